#!/usr/bin/env python
"""bench.py — headline benchmark of diffeq_b200: f64 trajectory-steps/s on BASELINE.json's config[1]
(Lorenz-63 ensemble of 2^20 random initial conditions per GPU, Dormand–Prince 5(4), rtol = atol = 1e-8, tspan [0,10]).

    python bench.py [--gpus N] [--steps K] [--warmup W]          one JSON line, our CUDA path
    python bench.py --impl reference ...                          the reference's CPU path (oracle restatement,
                                                                  all host threads) on a bounded sample

A "step" is one pass of the hot path over the whole per-GPU ensemble: hinit pre-pass + the persistent adaptive
kernel (`deq_solve`).  `value` counts accepted + rejected step attempts of all ranks divided by the device time of
the timed steps (CUDA events on the launching stream, max over ranks), inputs resident in HBM.  `e2e` is the same
metric through the host-buffer API (`diffeq_b200.solve_host`): pinned host y0 -> H2D -> solve -> D2H of final
states and step counters, every step.  Multi-GPU (torchrun, one rank per GPU): rank r integrates the global
trajectory indices [r*2^20, (r+1)*2^20) — weak scaling, no data-path collective (trajectories are independent).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NTRAJ_PER_GPU = 1 << 20
TSPAN = [0.0, 10.0]
RTOL = ATOL = 1e-8
FLOP_PER_STEP = 314.0  # dopri5 (nnz a=20, b=5+6, S=7, FSAL) on Lorenz (n=3, F_rhs=8)
METRIC = "f64 trajectory-steps/s (Dopri5 Lorenz ensemble)"
UNIT = "trajectory-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ntraj", type=int, default=NTRAJ_PER_GPU, help="trajectories per GPU (default 2^20)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=65536)
    ap.add_argument("--no-config3", action="store_true", help="skip the dense-output end-to-end section")
    ap.add_argument("--strong-log2", type=int, nargs="*", default=[20, 23, 26],
                    help="BASELINE config 5: total trajectory counts (log2) of the strong-scaling sweep")
    return ap.parse_args()


def config_dict(args, world):
    return {"workload": "configs[1]: Lorenz-63 (sigma=10, rho=28, beta=8/3) ensemble, x,y~U[-20,20], z~U[0,50] "
                        "(splitmix64 counter RNG), Ode45 = Dormand-Prince 5(4) reference tableau, rtol=atol=1e-8, "
                        "tspan=[0,10], Points::All final state + step counters",
            "trajectories_per_gpu": args.ntraj, "trajectories_total": args.ntraj * world, "tspan": TSPAN,
            "rtol": RTOL, "atol": ATOL, "method": "Ode45/dopri5", "parallelism": f"ensemble-sharded x{world}",
            "l2": "flushed between timed steps (256 MiB memset outside the event pairs)"}


KERNEL_MODE = ("headline = Mode.Parity, the library default: no FMA contraction, glibc-exact pow, bit-identical to the CPU oracle "
               "(tests/test_gpu_parity.py); fast_mode = opt-in Mode.Fast (FMA, zero-skipping, one pow per step)")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, so ask the OS, not OpenMP)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_run(ntraj, nthreads=0, offset=0, alloc_faithful=False):
    """The reference's CPU path (C++ restatement, OpenMP over trajectories) on trajectories [offset, offset+ntraj).
    alloc_faithful: the variant that also reproduces the reference's heap-allocation pattern (`Y = Vec<f64>`, ~20 allocations
    per step attempt); same bits, slower — the default (stack arrays) is the conservative baseline."""
    from oracle import oracle as O
    from diffeq_b200 import synth
    nthreads = nthreads or host_threads()
    y0 = synth.lorenz_ics(offset, offset + ntraj)
    fn = O.ensemble_adaptive_vec if alloc_faithful else O.ensemble_adaptive
    t = time.perf_counter()
    r = fn(O.ODE45, O.LORENZ, synth.LORENZ_PARAMS, y0, TSPAN, O.opts(RTOL, ATOL), nthreads=nthreads)
    dt = time.perf_counter() - t
    steps = int(r["accepted"].sum()) + int(r["rejected"].sum())
    return steps, dt, nthreads


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; Rust cannot be built in
    this image) with all host threads, each step a bounded sample of the workload.  Rank 0 only."""
    if rank != 0:
        return
    sample = 65536   # at least 65 536 trajectories per timed step
    for _ in range(args.warmup):
        cpu_reference_run(2048)
    steps_total, t_total, cores = 0, 0.0, 1
    for k in range(args.steps):
        s, dt, cores = cpu_reference_run(sample, offset=k * sample)
        steps_total += s; t_total += dt
    v = steps_total / t_total
    desc = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample} trajectories of the workload per step (C++ restatement of the Rust solver, OpenMP "
                      f"schedule(dynamic,256), -O2 -ffp-contract=off)"}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(args.steps, 1),
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                      "data": "synthetic", "config": config_dict(args, world), "kernel_mode": "reference CPU path (oracle port)",
                      "cpu_baseline": desc,
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import diffeq_b200 as D
    from diffeq_b200 import sharding, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (diffeq_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # the version banner goes to stdout, which carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    D.lib()
    D.set_device(local_rank)
    stream = torch.cuda.current_stream()
    D.set_stream(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    ntraj = args.ntraj
    y0 = synth.lorenz_ics(rank * ntraj, (rank + 1) * ntraj)   # contiguous shard of the global index range
    rhs = D.Rhs.builtin(D.System.Lorenz)
    prob = D.OdeProblem.builder().fun(rhs).params(synth.LORENZ_PARAMS).init(y0).tspan(TSPAN).build()  # resident in HBM
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    fp64_peak, _ = D.measure_fp64_peak()

    opts_fast = D.OdeOptionMap().insert(D.Reltol(RTOL), D.Abstol(ATOL), D.Mode.Fast)
    opts_par = D.OdeOptionMap().insert(D.Reltol(RTOL), D.Abstol(ATOL), D.Mode.Parity)

    def device_arm(opts, sampler=None):
        """K timed passes with the ensemble resident in HBM; returns (device ms, steps, kernel ms list, wall s)."""
        for _ in range(args.warmup):
            prob.solve(D.Ode.Ode45, opts, fetch=False)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        if sampler:
            sampler.start()
        barrier()
        wall0 = time.perf_counter()
        steps_done, kernel_ms, totals = 0, [], None
        for k in range(args.steps):
            flush.zero_()
            ev[k][0].record(stream)
            sol = prob.solve(D.Ode.Ode45, opts, fetch=False)
            ev[k][1].record(stream)
            steps_done += sol.totals["accepted"] + sol.totals["rejected"]
            kernel_ms.append(sol.kernel_ms)
            totals = sol.totals
        barrier()
        wall = time.perf_counter() - wall0
        return sum(a.elapsed_time(b) for a, b in ev), steps_done, kernel_ms, wall, totals

    # ---------------- device-resident arms: PARITY (headline, bit-exact) and FAST (opt-in) ----------------
    sampler = ClockSampler(local_rank)
    dev_ms, steps_done, kernel_ms, wall, tot_par = device_arm(opts_par, sampler if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None
    fast_ms, fast_steps, fast_kernel_ms, _, tot_fast = device_arm(opts_fast)
    opts_fold = D.OdeOptionMap().insert(D.Reltol(RTOL), D.Abstol(ATOL), D.Mode.Parity, D.Libm.LlvmFolded)
    fold_ms, fold_steps, fold_kernel_ms, _, tot_fold = device_arm(opts_fold)
    # full-size parity property: both arithmetic modes must take exactly the same accepted / rejected steps, and the
    # final states must agree within the north-star tolerance (10 x rtol)
    s_fast = prob.solve(D.Ode.Ode45, opts_fast)
    s_par = prob.solve(D.Ode.Ode45, opts_par)
    f_fast, f_par = s_fast.final, s_par.final
    same_counts = (s_fast.accepted == s_par.accepted) & (s_fast.rejected == s_par.rejected)
    dev_units = np.max(np.abs(f_fast - f_par) / (np.abs(f_par) * RTOL + ATOL), axis=1)
    mode_check = {"fraction_of_trajectories_with_identical_accepted_and_rejected_counts": float(same_counts.mean()),
                  "trajectories_with_different_counts": int((~same_counts).sum()),
                  "step_totals_identical": bool(tot_fast == tot_par), "accepted": tot_par["accepted"], "rejected": tot_par["rejected"],
                  "final_state_deviation_fast_vs_parity_in_units_of_(rtol*|y|+atol)": {
                      "median": float(np.median(dev_units)), "p99": float(np.quantile(dev_units, 0.99)),
                      "p99.99": float(np.quantile(dev_units, 0.9999)), "max": float(dev_units.max()),
                      "fraction_within_10": float((dev_units <= 10.0).mean())},
                  "note": "Lorenz trajectories that pass close to the saddle at the origin amplify ANY rounding difference by up to "
                          "~1e10 within T=10, so only bit-identical arithmetic (parity mode) reproduces them to 10 x rtol; against a "
                          "high-precision solution FAST is as accurate as PARITY (tests/test_gpu_parity.py::test_fast_mode_accuracy_against_truth)"}
    del s_fast, s_par

    # ---------------- end-to-end arms (host buffers through the public API) ----------------
    h_y0 = torch.from_numpy(y0).pin_memory()
    h_final = torch.empty((ntraj, 3), dtype=torch.float64).pin_memory()
    h_acc = torch.empty(ntraj, dtype=torch.int32).pin_memory()
    h_rej = torch.empty(ntraj, dtype=torch.int32).pin_memory()

    e2e_out = {"final": h_final.numpy(), "accepted": h_acc.numpy(), "rejected": h_rej.numpy()}

    def e2e_arm(opts):
        """K calls of the one-shot C-ABI entry point deq_solve_host on pinned HOST buffers: H2D of y0 / params / tspan, the solve,
        D2H of final states and accepted / rejected counters, all inside the timed region."""
        for _ in range(min(args.warmup, 2)):
            D.solve_host_multi(rhs, h_y0.numpy(), synth.LORENZ_PARAMS, TSPAN, D.Ode.Ode45, opts, out=e2e_out)
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            D.solve_host_multi(rhs, h_y0.numpy(), synth.LORENZ_PARAMS, TSPAN, D.Ode.Ode45, opts, out=e2e_out)
        barrier()
        dt = time.perf_counter() - t0
        per_call = int(h_acc.numpy().sum(dtype=np.int64) + h_rej.numpy().sum(dtype=np.int64))   # counted after the timed region
        return per_call * args.steps, dt

    e2e_steps, e2e_s = e2e_arm(opts_par)
    e2e_fast_steps, e2e_fast_s = e2e_arm(opts_fast)
    h2d = y0.nbytes + 8 * len(TSPAN) + 8 * 3
    d2h = h_final.numel() * 8 + h_acc.numel() * 4 + h_rej.numel() * 4

    # CPU-side rendezvous for the sections below in which some ranks must stay off their GPU (an NCCL barrier is a kernel that
    # spins on the device of every waiting rank)
    cpu_group = dist.new_group(backend="gloo") if distributed else None

    def cpu_barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier(group=cpu_group)

    # ---------------- final gather of the sharded final states to rank 0 (the only inter-GPU traffic the path needs) ----------------
    final_gather = None
    if distributed:
        loc = torch.from_numpy(f_par).cuda()
        bufs = [torch.empty_like(loc) for _ in range(world)] if rank == 0 else None
        for _ in range(3):   # warm the communicator: the first collective on it pays the lazy NCCL connection set-up
            dist.gather(loc, bufs, dst=0)
        barrier()
        times = []
        for _ in range(5):
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            dist.gather(loc, bufs, dst=0)
            g1.record(stream)
            torch.cuda.synchronize()
            times.append(g0.elapsed_time(g1))
        tg = torch.tensor([min(times)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        if rank == 0:
            assert torch.equal(bufs[0], loc) and all(b.shape == loc.shape for b in bufs)
            moved = loc.numel() * 8 * (world - 1)
            final_gather = {"ms": float(tg.item()), "bytes_into_rank0": moved, "GB_per_s": moved / (float(tg.item()) * 1e-3) / 1e9,
                            "what": f"dist.gather (NCCL send/recv) of {world} x {loc.numel() * 8 / 2**20:.0f} MiB final states to rank 0, "
                                    "communicator warmed, best of 5, max over ranks"}
        del loc, bufs

    # ---------------- BASELINE config 5: strong scaling, 2^20 .. 2^26 trajectories IN TOTAL over the N ranks ----------------
    def timed_solve(problem, opts, reps):
        problem.solve(D.Ode.Ode45, opts, fetch=False)
        best, steps_ = None, 0
        for _ in range(reps):
            flush.zero_()
            sol = problem.solve(D.Ode.Ode45, opts, fetch=False)
            steps_ = sol.totals["accepted"] + sol.totals["rejected"]
            best = sol.kernel_ms if best is None else min(best, sol.kernel_ms)
        return best, steps_

    strong = []
    for lg in args.strong_log2:
        total = 1 << lg
        b, e = sharding.shard_range(total, rank, world)
        ys = synth.lorenz_ics(b, e)
        ps = D.OdeProblem.builder().fun(rhs).params(synth.LORENZ_PARAMS).init(ys).tspan(TSPAN).build()
        barrier()
        ms, st = timed_solve(ps, opts_par, 2 if lg <= 23 else 1)
        msf, stf = timed_solve(ps, opts_fast, 2 if lg <= 23 else 1)
        # end to end: pinned host shard -> deq_solve_host -> host results
        hy = torch.from_numpy(ys).pin_memory()
        ho = {"final": torch.empty((e - b, 3), dtype=torch.float64).pin_memory().numpy(),
              "accepted": torch.empty(e - b, dtype=torch.int32).pin_memory().numpy().view(np.uint32),
              "rejected": torch.empty(e - b, dtype=torch.int32).pin_memory().numpy().view(np.uint32)}
        D.solve_host_multi(rhs, hy.numpy(), synth.LORENZ_PARAMS, TSPAN, D.Ode.Ode45, opts_par, out=ho)
        barrier()
        t0 = time.perf_counter()
        D.solve_host_multi(rhs, hy.numpy(), synth.LORENZ_PARAMS, TSPAN, D.Ode.Ode45, opts_par, out=ho)
        torch.cuda.synchronize()
        e2e_local = time.perf_counter() - t0
        del ps, hy, ho
        tt = torch.tensor([ms, msf, e2e_local * 1e3], dtype=torch.float64, device="cuda")
        cc = torch.tensor([st, stf], dtype=torch.float64, device="cuda")
        if distributed:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        entry = {"trajectories_total": total, "trajectories_per_gpu": e - b, "n_gpus": world,
                 "parity_steps_per_s": cc[0].item() / (tt[0].item() * 1e-3), "parity_ms": tt[0].item(),
                 "fast_steps_per_s": cc[1].item() / (tt[1].item() * 1e-3), "fast_ms": tt[1].item(),
                 "e2e_parity_steps_per_s": cc[0].item() / (tt[2].item() * 1e-3), "e2e_parity_ms": tt[2].item()}
        if distributed:
            # the same total on ONE GPU (rank 0 alone; the others wait on the CPU): efficiency = t(1) / (N t(N))
            cpu_barrier()
            if rank == 0:
                y1 = synth.lorenz_ics(0, total)
                p1 = D.OdeProblem.builder().fun(rhs).params(synth.LORENZ_PARAMS).init(y1).tspan(TSPAN).build()
                del y1
                ms1, _ = timed_solve(p1, opts_par, 1)
                ms1f, _ = timed_solve(p1, opts_fast, 1)
                del p1
                entry.update({"one_gpu_parity_ms": ms1, "one_gpu_fast_ms": ms1f,
                              "parity_efficiency_vs_1_gpu": ms1 / (world * entry["parity_ms"]),
                              "fast_efficiency_vs_1_gpu": ms1f / (world * entry["fast_ms"])})
            cpu_barrier()
        strong.append(entry)

    # ---------------- the in-library multi-GPU path (what a Rust host calls): ONE process, deq_solve_host over all N devices ----------------
    in_library = None
    if distributed:
        cpu_barrier()
        if rank == 0:
            n_all = ntraj * world
            hy = torch.from_numpy(synth.lorenz_ics(0, n_all)).pin_memory()
            ho = {"final": torch.empty((n_all, 3), dtype=torch.float64).pin_memory().numpy(),
                  "accepted": torch.empty(n_all, dtype=torch.int32).pin_memory().numpy().view(np.uint32),
                  "rejected": torch.empty(n_all, dtype=torch.int32).pin_memory().numpy().view(np.uint32)}
            devs = list(range(world))
            for _ in range(2):
                D.solve_host_multi(rhs, hy.numpy(), synth.LORENZ_PARAMS, TSPAN, D.Ode.Ode45, opts_par, devices=devs, out=ho)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                D.solve_host_multi(rhs, hy.numpy(), synth.LORENZ_PARAMS, TSPAN, D.Ode.Ode45, opts_par, devices=devs, out=ho)
            dt_lib = (time.perf_counter() - t0) / reps
            st_lib = int(ho["accepted"].sum(dtype=np.int64) + ho["rejected"].sum(dtype=np.int64))
            ok = bool(np.array_equal(ho["final"][:ntraj], f_par))   # rank 0's shard: same bits as the rank-per-GPU solve
            in_library = {"steps_per_s": st_lib / dt_lib, "ms_per_call": dt_lib * 1e3, "trajectories": n_all, "devices": devs,
                          "first_shard_bit_identical_to_rank0_solve": ok,
                          "what": "deq_solve_host(devices = 0..N-1) called by ONE process on pinned host buffers (H2D + solve + D2H of "
                                  "every shard inside the call, one host thread and stream per device); compare with e2e.value"}
            del hy, ho
        cpu_barrier()

    # ---------------- BASELINE config 3 end to end: dense rows to HOST buffers through deq_solve_host_dense ----------------
    # (2^18 of the 2^22 trajectories: 4.2 GB of rows; the full 67 GB take the same per-byte time and need a 67 GB pinned buffer)
    config3_e2e = None
    if distributed:
        cpu_barrier()
    if rank == 0 and not args.no_config3:
        n3, nt3 = 1 << 18, 1000
        y3, mu3 = synth.vdp_sweep(0, n3, 1 << 22)
        ts3 = D.linspace(0.0, 20.0, nt3)
        vdp = D.Rhs.builtin(D.System.VanDerPol)
        o3 = D.OdeOptionMap().insert(D.Reltol(1e-6), D.Abstol(1e-8), D.Output.Tspan, D.DenseLayout.TrajMajor)
        h3 = {"final": torch.empty((n3, 2), dtype=torch.float64).pin_memory().numpy(),
              "accepted": torch.empty(n3, dtype=torch.int32).pin_memory().numpy().view(np.uint32),
              "rejected": torch.empty(n3, dtype=torch.int32).pin_memory().numpy().view(np.uint32),
              "nvalid": torch.empty(n3, dtype=torch.int32).pin_memory().numpy().view(np.uint32),
              "dense": torch.empty((n3, nt3, 2), dtype=torch.float64).pin_memory().numpy()}
        hy3, hm3 = torch.from_numpy(y3).pin_memory().numpy(), torch.from_numpy(mu3).pin_memory().numpy()
        nbytes = h3["dense"].nbytes
        # the plain pinned device-to-host copy of the same bytes: the ceiling of this path
        dsrc = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        hdst = torch.from_numpy(h3["dense"].view(np.uint8).reshape(-1))
        hdst.copy_(dsrc); torch.cuda.synchronize()
        t0 = time.perf_counter(); hdst.copy_(dsrc, non_blocking=True); torch.cuda.synchronize(); t_copy = time.perf_counter() - t0
        del dsrc
        # kernel alone (device-resident rows, trajectory-major like the host result)
        p3 = D.OdeProblem.builder().fun(vdp).params(mu3).init(y3).tspan(ts3).build()
        p3.solve(D.Ode.Ode45fe, o3, fetch=False)
        k3 = p3.solve(D.Ode.Ode45fe, o3, fetch=False)
        del p3
        D.solve_host_multi(vdp, hy3, hm3, ts3, D.Ode.Ode45fe, o3, out=h3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        D.solve_host_multi(vdp, hy3, hm3, ts3, D.Ode.Ode45fe, o3, out=h3)
        t_e2e = time.perf_counter() - t0
        steps3 = int(h3["accepted"].sum(dtype=np.int64) + h3["rejected"].sum(dtype=np.int64))
        config3_e2e = {"workload": "configs[2] at 2^18 of its 2^22 trajectories: Van der Pol mu sweep, Ode45fe, rtol 1e-6, atol 1e-8, 1000 tspan rows "
                                   "per trajectory to pinned HOST memory, trajectory-major [ntraj][1000][2]",
                       "rows_bytes": nbytes, "e2e_ms": t_e2e * 1e3, "GB_per_s_to_host": nbytes / t_e2e / 1e9,
                       "plain_pinned_d2h_ms": t_copy * 1e3, "plain_pinned_d2h_GB_per_s": nbytes / t_copy / 1e9,
                       "kernel_only_ms": k3.kernel_ms, "steps_per_s_e2e": steps3 / t_e2e,
                       "e2e_over_max_of_kernel_and_copy": t_e2e / max(t_copy, k3.kernel_ms * 1e-3),
                       "what": "deq_solve_host_dense: chunks of ~128 MiB of rows, chunk i+1 integrates while chunk i travels to the host"}
        del h3, hy3, hm3
    if distributed:
        cpu_barrier()

    # ---------------- BASELINE config 4 (quarter size) and dense rows of a sweep handed over in random order, informational ----------------
    config4 = None
    dense_permuted = None
    if rank == 0 and not args.no_config3:
        n4 = 1 << 22
        p4 = D.OdeProblem.builder().fun(D.Rhs.builtin(D.System.Pendulum)).params(synth.PENDULUM_PARAMS).init(synth.pendulum_ics(0, n4)).tspan([0.0, 10.0]).build()
        config4 = {"workload": "configs[3] at 2^22 of its 2^24 trajectories: damped driven pendulum, Ode78 = Fehlberg 7(8) TEXTBOOK coefficients, rtol = atol = 1e-12, "
                               "tspan [0, 10], final states; PARITY evaluates glibc's sin / cos bit for bit (13 sin + 9 cos per step attempt)"}
        for name, mode in (("parity", D.Mode.Parity), ("fast", D.Mode.Fast)):
            o4 = D.OdeOptionMap().insert(D.Reltol(1e-12), D.Abstol(1e-12), mode)
            p4.solve(D.Ode.Ode78, o4, tableau_set=D.TableauSet.Textbook, fetch=False)
            s4 = p4.solve(D.Ode.Ode78, o4, tableau_set=D.TableauSet.Textbook, fetch=False)
            st4 = s4.totals["accepted"] + s4.totals["rejected"]
            config4[name] = {"kernel_ms": s4.kernel_ms, "step_attempts": st4, "steps_per_s": st4 / (s4.kernel_ms * 1e-3),
                             "TFLOPs_at_495_flop_per_step": st4 * 495 / (s4.kernel_ms * 1e-3) / 1e12}
            del s4
        del p4
        nd, ntd = 1 << 18, 1000
        yd, mud = synth.vdp_sweep(0, nd, nd)
        tsd = D.linspace(0.0, 20.0, ntd)
        permd = np.random.default_rng(0).permutation(nd)
        dense_permuted = {"workload": "Van der Pol mu sweep of 2^18 trajectories (a quarter of the probe in profiles/r2_dense_sorted_vs_permuted.jsonl), Ode45fe, rtol 1e-6, "
                                      "atol 1e-8, 1000 tspan rows per trajectory (4.2 GB) kept on the device; kernel_ms of the second of two solves; "
                                      "'permuted' = the same sweep in random order (deq_solve sorts the work items by first step size; SoA: + in-place column permutation)"}
        for order, m in (("sorted", mud), ("permuted", mud[permd].copy())):
            pd = D.OdeProblem.builder().fun(D.Rhs.builtin(D.System.VanDerPol)).params(m).init(yd).tspan(tsd).build()
            for lay in (D.DenseLayout.SoA, D.DenseLayout.TrajMajor):
                od = D.OdeOptionMap().insert(D.Reltol(1e-6), D.Abstol(1e-8), D.Output.Tspan, lay)
                pd.solve(D.Ode.Ode45fe, od, fetch=False)
                sd = pd.solve(D.Ode.Ode45fe, od, fetch=False)
                dense_permuted[order + "_" + lay.name + "_ms"] = sd.kernel_ms
                del sd
            del pd

    # ---------------- stiff path, informational: ode23s on a Van der Pol stiffness sweep ----------------
    stiff = None
    if rank == 0:
        ns = 1 << 18
        mu = np.geomspace(0.1, 1000.0, ns)[:, None]
        sp = D.OdeProblem.builder().fun(D.Rhs.builtin(D.System.VanDerPol)).params(mu).init(np.tile([2.0, 0.0], (ns, 1))).tspan([0.0, 20.0]).build()
        so = D.OdeOptionMap().insert(D.Reltol(1e-5), D.Abstol(1e-8), D.MaxSteps(20000))
        stiff = {"workload": "2^18 Van der Pol, mu = geomspace(0.1, 1000), y0 = [2, 0], tspan [0, 20], ode23s rtol 1e-5 atol 1e-8, max_steps 20000"}
        for name, tset, extra in (("reference", 0, ()), ("textbook", 1, ()), ("textbook_fast_mode", 1, (D.Mode.Fast,))):
            so = D.OdeOptionMap().insert(D.Reltol(1e-5), D.Abstol(1e-8), D.MaxSteps(20000), *extra)
            sp.solve(D.Ode.Ode23s, so, tableau_set=tset, fetch=False)
            r = sp.solve(D.Ode.Ode23s, so, tableau_set=tset, fetch=False)
            att = r.totals["accepted"] + r.totals["rejected"]
            stiff[name] = {"steps_per_s": att / (r.kernel_ms * 1e-3), "kernel_ms": r.kernel_ms, "step_attempts": att}
        del sp
    if distributed:
        cpu_barrier()   # (the other ranks wait here on the CPU, not inside an NCCL kernel, while rank 0 runs the informational sections)

    # ---------------- reduce over ranks ----------------
    t = torch.tensor([dev_ms, e2e_s, wall, fast_ms, e2e_fast_s, fold_ms], dtype=torch.float64, device="cuda")
    c = torch.tensor([steps_done, e2e_steps, fast_steps, e2e_fast_steps, fold_steps], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_s_max, wall_max, fast_ms_max, e2e_fast_s_max, fold_ms_max = t.tolist()
    steps_all, e2e_steps_all, fast_steps_all, e2e_fast_steps_all, fold_steps_all = c.tolist()

    def ncu_summary(key):
        """dram__bytes_read.sum + dram__bytes_write.sum and the FP64-pipe utilisation of ONE launch of the kernel at 2^20
        trajectories, read from the committed `ncu --set full` summary of this round (tools/ncu_summary.py); traffic scales
        with the trajectory count, not with the step count.  None if the summary is absent."""
        path = os.path.join(ROOT, "profiles", "r2_%s_kernel_ncu_full_summary.csv" % key)
        out = {"traffic": None, "fp64_pipe_active_pct_ncu": None, "ncu_summary": None}
        if not os.path.exists(path):
            return out
        vals, scale = {}, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for line in open(path):
            parts = [c.strip().strip('"') for c in line.rstrip("\n").split('","')]
            if len(parts) == 3:
                try:
                    vals[parts[0]] = float(parts[2]) * scale.get(parts[1], 1.0)
                except ValueError:
                    pass
        rd, wr = vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum")
        out["ncu_summary"] = os.path.relpath(path, ROOT)
        if rd is not None and wr is not None:
            out["traffic"] = (rd + wr) * (ntraj / float(1 << 20))
        out["fp64_pipe_active_pct_ncu"] = vals.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed")
        return out

    def roofline(kernel_ms_list, steps_local, kernel_name, mode_note, traffic_key):
        kms = sum(kernel_ms_list) / len(kernel_ms_list)
        achieved = (steps_local / args.steps) * FLOP_PER_STEP / (kms * 1e-3) / 1e12
        return {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak if fp64_peak else None,
                **ncu_summary(traffic_key), "traffic_unit": "bytes per launch (DRAM, from the committed ncu summary named in ncu_summary)",
                "kernel": kernel_name,
                "kernel_ms": kms, "flop_per_step": FLOP_PER_STEP, "arithmetic": mode_note,
                "peak_source": "DFMA-stream microbenchmark measured in this run (deq_measure_fp64_peak); MEASURED_PEAKS.json "
                               "has no FP64 figure; nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2 TFLOP/s",
                "note": "state lives in registers: HBM traffic is ~100 B per trajectory per ~1200 steps (ncu: 78 MB per launch), "
                        "so the FP64 pipe — not HBM, not tensor cores — bounds this kernel"}

    if rank == 0:
        cfg = config_dict(args, world)
        par_note = ("no FMA allowed (the reference never fuses a*b+c): every multiply-add costs two FP64 issue slots, so at most 50 % "
                    "of the DFMA peak is reachable by construction in this mode")
        out = {
            "metric": METRIC, "value": steps_all / (dev_ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "kernel_mode": KERNEL_MODE,
            "e2e": {"value": e2e_steps_all / e2e_s_max, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s_max / args.steps},
            "gpu_launches": 3 * args.steps,   # per timed step: hinit_kernel, adapt_kernel, sum_counts_kernel (the step totals)
            "accepted_only_steps_per_s": (tot_par["accepted"] * world) / (dev_ms_max / args.steps * 1e-3),
            "roofline": roofline(kernel_ms, steps_done, "deqk::adapt_kernel<Dopri5<false>, Lorenz, false, MODE_FINAL, FAST=false>", par_note, "parity"),
            "fast_mode": {"value": fast_steps_all / (fast_ms_max * 1e-3), "unit": UNIT, "ms_per_step": fast_ms_max / args.steps,
                          "e2e": {"value": e2e_fast_steps_all / e2e_fast_s_max, "unit": UNIT, "ms_per_step": 1e3 * e2e_fast_s_max / args.steps},
                          "roofline": roofline(fast_kernel_ms, fast_steps, "deqk::adapt_kernel<Dopri5<false>, Lorenz, false, MODE_FINAL, FAST=true>",
                                               "FMA counted as 2 flop; algorithmic flop per step attempt from SURVEY.md 8(d)", "fast")},
            "parity_llvm_folded_libm": {
                "value": fold_steps_all / (fold_ms_max * 1e-3), "unit": UNIT, "ms_per_step": fold_ms_max / args.steps,
                "step_totals_identical_to_headline": bool(tot_fold == tot_par),
                "roofline": roofline(fold_kernel_ms, fold_steps, "deqk::adapt_kernel<Dopri5<false>, Lorenz, false, MODE_FINAL, FAST=false> (libm_folded)",
                                     par_note, "parity"),
                "note": "bit-exact PARITY arithmetic with the libm calls an LLVM-optimised build of the crate makes (sqrt for the error norm, "
                        "exp10 in hinit; DESIGN.md 4.1) — one pow per step instead of two"},
            "mode_check": mode_check,
            "stiff_ode23s": stiff,
            "config3_e2e": config3_e2e, "config4": config4, "dense_permuted_sweep": dense_permuted, "strong_scaling": strong, "solve_host_in_library": in_library, "final_gather": final_gather,
            "clocks": clocks, "wall_s_timed_region": wall_max,
        }
        if not args.no_cpu_baseline:
            s, dt, cores = cpu_reference_run(args.cpu_sample)
            out["cpu_baseline"] = {"value": s / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"first {args.cpu_sample} trajectories of the workload ({s} step attempts, "
                                             f"{dt:.1f} s): C++ restatement of the Rust solver (oracle/), OpenMP over trajectories"}
            s2, dt2, _ = cpu_reference_run(max(args.cpu_sample // 4, 1024), alloc_faithful=True)
            out["cpu_baseline"]["with_reference_allocation_pattern"] = {
                "value": s2 / dt2, "unit": UNIT,
                "note": "same restatement with the reference's memory behaviour (every state a heap Vec, ~20 allocations per step "
                        "attempt, growing Vec<Vec<f64>> output); bit-identical results; `value` above (stack arrays) is the conservative one"}
        print(json.dumps(out), flush=True)
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
