"""Probe (not a test): contiguous vs block-cyclic sharding of parameter-sorted sweeps over the GPUs of one box, through the
one-shot C-ABI call deq_solve_host (host buffers, one host thread + stream per device).  usage: python tests/gpu_shard_balance.py [log2_ntraj]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import diffeq_b200 as D
from diffeq_b200 import synth

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
T = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
n = 1 << lg
ndev = C.c_int()
C.CDLL("libcudart.so.12").cudaGetDeviceCount(C.byref(ndev))
devs = list(range(ndev.value))
vd = D.Rhs.builtin(D.System.VanDerPol)


def run(name, y0, mu, ts, ode, base, tset=D.TableauSet.Reference):
    out = None
    for chunk in (0, 1 << 16):
        o = D.OdeOptionMap().insert(*base, D.ShardChunk(chunk))
        best = None
        for rep in range(3):
            t0 = time.perf_counter()
            out = D.solve_host_multi(vd, y0, mu, ts, ode, o, devices=devs, tableau_set=tset, out=out)
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        steps = int(out["accepted"].sum(dtype=np.int64) + out["rejected"].sum(dtype=np.int64))
        per_dev = None
        if chunk == 0:
            b = np.array_split(out["accepted"].astype(np.int64) + out["rejected"], len(devs))
            per_dev = [int(x.sum()) for x in b]
        print(json.dumps({"workload": name, "gpus": len(devs), "ntraj": len(y0), "shard_chunk": chunk, "wall_ms": best * 1e3,
                          "steps": steps, "steps_per_s_e2e": steps / best, "steps_per_contiguous_shard": per_dev}), flush=True)


y0, mu = synth.vdp_sweep(0, n, n)
run(f"config-3 sweep (mu = 0.1..10 sorted), rk45 rtol 1e-6, T = {T}, final states", y0, mu, [0.0, T], D.Ode.Ode45fe,
    (D.Reltol(1e-6), D.Abstol(1e-8)))
mus = np.geomspace(0.1, 1000.0, n)[:, None]
run(f"stiffness sweep (mu = 0.1..1000 geometric), ode23s TEXTBOOK rtol 1e-5, T = {T}, final states", np.tile([2.0, 0.0], (n, 1)), mus,
    [0.0, T], D.Ode.Ode23s, (D.Reltol(1e-5), D.Abstol(1e-8), D.MaxSteps(200000)), tset=D.TableauSet.Textbook)
