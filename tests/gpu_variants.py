"""Experiment driver (not a test): time config 2 with each libdiffeq_b200_<variant>.so present."""
import glob, os, subprocess, sys
libs = sorted(glob.glob("diffeq_b200/libdiffeq_b200_*.so"))
for l in libs:
    env = dict(os.environ, DEQ_LIB=os.path.abspath(l))
    r = subprocess.run([sys.executable, "tests/gpu_quick.py"] + sys.argv[1:], env=env, capture_output=True, text=True)
    lines = [x for x in r.stdout.splitlines() if x.startswith("refill=1")]
    print(os.path.basename(l), lines[-1] if lines else r.stderr[-300:])
