"""Builds libdiffeq_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m diffeq_b200.build [--force]

One object per Butcher tableau (kernels_inst.cu with -DDEQ_TB=<id>) so the 13 translation units compile in
parallel; every device TU uses -fmad=false (parity contract: the reference never fuses a*b+c, see stepper.cuh).
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libdiffeq_b200.so")
NVCC = os.environ.get("DEQ_NVCC", "/usr/local/cuda/bin/nvcc")
HOSTCXX = os.environ.get("DEQ_HOSTCXX", "/usr/bin/g++")
N_TB = 13

NVCC_FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false",
              "--expt-relaxed-constexpr", "-ccbin", HOSTCXX, "-Xcompiler", "-fPIC,-ffp-contract=off"]


def _sources():
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    deps.append(os.path.join(HERE, "..", "include", "diffeq_b200.h"))
    return deps


def _fingerprint():
    h = hashlib.sha256()
    for p in _sources():
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stderr


def build(force=False, verbose=False, extra_flags=(), lib_path=None, obj_dir=None):
    """extra_flags / lib_path / obj_dir build an experiment variant next to the product library."""
    global NVCC_FLAGS
    base_flags = NVCC_FLAGS
    OBJ_ = obj_dir or OBJ
    LIB_ = lib_path or LIB
    os.makedirs(OBJ_, exist_ok=True)
    stamp = os.path.join(OBJ_, "fingerprint")
    fp = _fingerprint() + " ".join(extra_flags)
    if not force and os.path.exists(LIB_) and os.path.exists(stamp) and open(stamp).read() == fp:
        return LIB_
    NVCC_FLAGS = list(base_flags) + list(extra_flags)
    try:
        return _build(OBJ_, LIB_, stamp, fp, verbose)
    finally:
        NVCC_FLAGS = base_flags


def _build(OBJ, LIB, stamp, fp, verbose):
    jobs = []
    for tb in range(N_TB):
        o = os.path.join(OBJ, "kernels_tb%d.o" % tb)
        jobs.append((o, [NVCC] + NVCC_FLAGS + ["-DDEQ_TB=%d" % tb, "-c", os.path.join(CSRC, "kernels_inst.cu"), "-o", o]))
    for name in ("kernels_common.cu", "capi.cu"):
        o = os.path.join(OBJ, name.replace(".cu", ".o"))
        jobs.append((o, [NVCC] + NVCC_FLAGS + ["-c", os.path.join(CSRC, name), "-o", o]))
    o = os.path.join(OBJ, "nvrtc_rhs.o")
    jobs.append((o, [NVCC] + NVCC_FLAGS + ["-x", "cu", "-c", os.path.join(CSRC, "nvrtc_rhs.cpp"), "-o", o]))
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for msg in ex.map(lambda j: _run(j[1]), jobs):
            if verbose and msg:
                print(msg, file=sys.stderr)
    objs = [j[0] for j in jobs]
    _run([NVCC, "-shared", "-o", LIB, "-ccbin", HOSTCXX] + objs + ["-lcudart", "-lnvrtc", "-lcuda"])
    with open(stamp, "w") as f:
        f.write(fp)
    return LIB


if __name__ == "__main__":
    extra = [a[len("--flag="):] for a in sys.argv if a.startswith("--flag=")]
    name = [a[len("--variant="):] for a in sys.argv if a.startswith("--variant=")]
    if name:
        print(build(force=True, verbose=False, extra_flags=extra,
                    lib_path=os.path.join(HERE, "libdiffeq_b200_%s.so" % name[0]), obj_dir=os.path.join(HERE, "_build_" + name[0])))
    else:
        print(build(force="--force" in sys.argv, verbose=True, extra_flags=extra))
