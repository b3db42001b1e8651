"""Regenerates diffeq_b200/csrc/deq_sincos_table.h from the running glibc's libm (locates __sincostab by the bit
pattern of sin(1/128))."""
import struct, sys
import mpmath
mpmath.mp.prec = 200
b = open('/lib/x86_64-linux-gnu/libm.so.6', 'rb').read()
s = float(mpmath.sin(mpmath.mpf(1) / 128))
off = b.find(struct.pack('<d', s)) - 32
t = struct.unpack('<440d', b[off:off + 440 * 8])
assert t[0] == 0.0 and t[2] == 1.0
rows = [", ".join(x.hex() for x in t[i:i + 4]) for i in range(0, 440, 4)]
hdr = open(sys.argv[1]).read().split("#define DEQ_SINCOSTAB_INIT")[0]
open(sys.argv[1], 'w').write(hdr + "#define DEQ_SINCOSTAB_INIT \\\n" + ", \\\n".join("    " + r for r in rows) + "\n")
