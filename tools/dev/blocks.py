#!/usr/bin/env python
"""Basic-block summary of a SASS listing produced by tools/dev/sass.sh (address opcode operands...).

    python tools/dev/blocks.py /tmp/sass/parity.sass [lo hi] [--cold a-b,c-d]

Prints, per basic block inside [lo, hi) (hex addresses), the number of FP64-pipe instructions (DADD/DMUL/DFMA/DSETP,
two issue cycles each on B200) and of all other instructions (one issue cycle), and the modelled issue cycles
2*fp64 + other of the blocks that are not marked cold.  Used to compare kernel variants without a GPU: the model
matched the measured cycles per warp-step within 1 % in round 1.
"""
import re
import sys


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    cold = []
    for a in sys.argv[1:]:
        if a.startswith("--cold="):
            for r in a[7:].split(","):
                lo, hi = r.split("-")
                cold.append((int(lo, 16), int(hi, 16)))
    path = args[0]
    lo = int(args[1], 16) if len(args) > 1 else 0
    hi = int(args[2], 16) if len(args) > 2 else 1 << 30
    ins = []
    for line in open(path):
        m = re.match(r"([0-9a-f]{4,}) (.*)", line.strip())
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    targets = set()
    for _, t in ins:
        for m in re.finditer(r"0x([0-9a-f]+) ;", t):
            if re.search(r"\b(BRA|BSSY|CALL|BRA\.U|BRA\.DIV)", t):
                targets.add(int(m.group(1), 16))
    blocks, cur = [], []
    for a, t in ins:
        if a in targets and cur:
            blocks.append(cur); cur = []
        cur.append((a, t))
        if re.search(r"\b(BRA|EXIT|RET|CALL)", t):
            blocks.append(cur); cur = []
    if cur:
        blocks.append(cur)
    tot_f = tot_o = 0
    for b in blocks:
        a0 = b[0][0]
        if a0 < lo or a0 >= hi:
            continue
        f = sum(1 for _, t in b if re.match(r"(@!?U?P\d+ )?D(ADD|MUL|FMA|SETP)", t))
        o = len(b) - f
        is_cold = any(c0 <= a0 < c1 for c0, c1 in cold)
        if not is_cold:
            tot_f += f; tot_o += o
        print("%05x..%05x fp64=%3d other=%3d %s %s" % (a0, b[-1][0], f, o, "COLD" if is_cold else "    ", b[-1][1][:60]))
    print("hot: fp64=%d other=%d model_cycles=%d" % (tot_f, tot_o, 2 * tot_f + tot_o))


if __name__ == "__main__":
    main()
