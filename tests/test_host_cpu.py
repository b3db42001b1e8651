"""CPU: host-side logic, the C ABI surface and the libm replicas (no GPU needed, no compute calls)."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_abi_exports_every_declared_symbol(deq):
    """libdiffeq_b200.so loads and exports exactly the functions include/diffeq_b200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "diffeq_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(deq_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(deq.ABI_SYMBOLS)
    L = ctypes.CDLL(os.path.join(ROOT, "diffeq_b200", "libdiffeq_b200.so"))
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in deq.lib().deq_version()


def test_product_does_not_touch_the_oracle():
    """The product path must not import / link the oracle (it is test infrastructure)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "diffeq_b200")):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"(import\s+oracle|from\s+oracle|#include[^\n]*oracle|liboracle|oracle_[a-z]+\s*\()", txt), \
                    f"{f} uses the oracle"
    out = subprocess.run(["ldd", os.path.join(ROOT, "diffeq_b200", "libdiffeq_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_product_tableaus_match_reference_source(deq):
    tabs = json.load(open(os.path.join(HERE, "golden", "tableaus.json")))
    name2tab = {"Feuler": "feuler", "Heun": "heun", "Midpoint": "midpoint", "Ode4": "rk4", "Ode21": "rk21", "Ode23": "rk23",
                "Ode45": "dopri5", "Ode45fe": "rk45", "Ode78": "feh78"}
    fh = lambda v: [fh(x) for x in v] if isinstance(v, list) else float.fromhex(v)  # noqa: E731
    for tset, key in ((deq.TableauSet.Reference, "reference"), (deq.TableauSet.Textbook, "textbook")):
        for m, tn in name2tab.items():
            g = tabs[key][tn]
            t = deq.tableau(getattr(deq.Ode, m), tset)
            assert (t["S"], t["adaptive"], t["order_min"], t["fsal"]) == (g["S"], g["adaptive"], g["order_min"], g["fsal"]), (m, key)
            assert np.array_equal(t["a"], np.array(fh(g["a"]))) and np.array_equal(t["c"], np.array(fh(g["c"])))
            gb = np.array(fh(g["b"]))
            assert np.array_equal(t["b"] if g["adaptive"] else t["b"][:, 0], gb if g["adaptive"] else gb[:, 0])


def test_product_rosenbrock_coefficients_match_reference_source(deq):
    """The kernels' kr4 / s4 sets (stiff.cuh) against the values parsed from rosenbrock.rs (tools/gen_golden.py stiff)."""
    gold = json.load(open(os.path.join(HERE, "golden", "stiff_cases.json")))["rosenbrock_coeffs"]
    fh = lambda v: [fh(x) for x in v] if isinstance(v, list) else float.fromhex(v)  # noqa: E731
    for which, m in (("kr4", deq.Ode.Ode4skr), ("s4", deq.Ode.Ode4ss)):
        co, g = deq.rosenbrock_coeffs(m), gold[which]
        assert co["gamma"] == fh(g["gamma"]) and co["a"].tolist() == fh(g["a"]) and co["b"].tolist() == fh(g["b"]) and co["c"].tolist() == fh(g["c"])
    with pytest.raises(deq.OdeError):
        deq.rosenbrock_coeffs(deq.Ode.Ode45)
    with pytest.raises(deq.OdeError) as e:
        deq.tableau(deq.Ode.Ode23s)          # the stiff methods have no Butcher tableau
    assert e.value.kind == "Unsupported"
    assert deq._traits(deq.Ode.Ode23s) == (True, True, True) and deq._traits(deq.Ode.Ode4ss) == (False, True, True)
    assert deq._traits(deq.Ode.Ode45) == (True, True, False) and deq._traits(deq.Ode.Ode4) == (False, False, False)


def test_ode_enum_and_from_str(deq):
    """mod.rs:14-46: variant order and the FromStr table (no name for Ode45fe; "ode4s" -> Ode4ss)."""
    assert [e.name for e in deq.Ode][:11] == ["Feuler", "Heun", "Midpoint", "Ode23", "Ode23s", "Ode4", "Ode45", "Ode45fe",
                                               "Ode4skr", "Ode4ss", "Ode78"]
    assert deq.Ode.from_str("ode45") is deq.Ode.Ode45 and deq.Ode.from_str("ode4s") is deq.Ode.Ode4ss
    for bad in ("ode45fe", "rk4", ""):
        with pytest.raises(ValueError):
            deq.Ode.from_str(bad)


def test_option_defaults_and_map(deq):
    """options.rs:283-299 defaults; unknown reference options (Norm, StepTimeout, Retries) are ignored."""
    o = deq.lib().deq_options_default()
    assert (o.reltol, o.abstol, o.initstep, o.has_minstep, o.has_maxstep, o.points) == (1e-5, 1e-8, 0.0, 0, 0, 0)
    m = deq.OdeOptionMap().insert(deq.Reltol(1e-3), deq.Maxstep(0.5), deq.Points.All)
    m["StepTimeout"] = 7
    c = m.to_c()
    assert (c.reltol, c.abstol, c.has_maxstep, c.maxstep, c.has_minstep) == (1e-3, 1e-8, 1, 0.5, 0)


def test_builder_uninitialized_errors(deq):
    """problem.rs:92-104: build() fails with Uninitialized for every missing field (checked before any CUDA call)."""
    rhs = deq.Rhs.builtin(deq.System.Lorenz)
    assert rhs.dim == 3 and rhs.nparams == 3
    for b in (deq.OdeProblem.builder().init([1, 2, 3]).tspan([0, 1]),
              deq.OdeProblem.builder().fun(rhs).params([10, 28, 8 / 3]).tspan([0, 1]),
              deq.OdeProblem.builder().fun(rhs).params([10, 28, 8 / 3]).init([1, 2, 3]),
              deq.OdeProblem.builder().fun(rhs).init([1, 2, 3]).tspan([0, 1])):
        with pytest.raises(deq.OdeError) as e:
            b.build()
        assert e.value.kind == "Uninitialized"
    with pytest.raises(deq.OdeError):
        deq.Rhs.builtin(99)


def test_linspace_and_synth_are_deterministic(deq, oracle):
    ts = deq.linspace(0.0, 100.0, 100001)
    assert np.array_equal(ts, oracle.linspace(0.0, 100.0, 100001))
    assert ts[0] == 0.0 and len(set(np.diff(ts).tolist())) == 16   # 16 distinct dt values on this grid
    from diffeq_b200 import synth
    a = synth.lorenz_ics(0, 1000)
    b = np.concatenate([synth.lorenz_ics(0, 333), synth.lorenz_ics(333, 1000)])
    assert np.array_equal(a, b)                      # counter RNG: shards generate the same inputs
    assert a[:, :2].min() >= -20 and a[:, :2].max() < 20 and a[:, 2].min() >= 0 and a[:, 2].max() < 50
    assert synth.splitmix64(np.array([0], dtype=np.uint64))[0] == np.uint64(0xE220A8397B1DCDAF)


def test_device_math_header_matches_libm_on_host():
    """deq_math.cuh compiled for the host (explicit FMAs, -ffp-contract=off) == the running glibc, bit for bit."""
    exe = "/tmp/deq_math_check"
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma",
                           os.path.join(HERE, "cpp", "math_check.cpp"), "-o", exe, "-lm"])
    out = subprocess.run([exe, "1000000"], capture_output=True, text=True)
    assert out.returncode == 0 and "MISMATCHES 0" in out.stdout, out.stdout[-500:]


def test_shard_ranges():
    from diffeq_b200 import sharding
    for total, world in ((10, 3), (1 << 20, 8), (7, 8), (0, 2)):
        r = [sharding.shard_range(total, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == total and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
    assert sharding.weak_range(100, 3) == (300, 400)


def test_cyclic_shards_partition_the_index_range():
    """Block-cyclic sharding (deq_options.shard_chunk): every trajectory belongs to exactly one rank."""
    from diffeq_b200 import sharding
    for ntraj, world, chunk in ((1000, 3, 64), (10, 4, 3), (5, 8, 2), (4096, 8, 512), (7, 2, 100)):
        parts = [sharding.cyclic_indices(ntraj, r, world, chunk) for r in range(world)]
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(ntraj)), (ntraj, world, chunk)
        assert all(np.all(np.diff(p) > 0) for p in parts if len(p) > 1)
        assert parts[0][0] == 0 and (len(parts[1]) == 0 or parts[1][0] == chunk)


def test_bench_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300)
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "trajectory-steps/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def _build_hpp_smoke():
    exe = "/tmp/deq_hpp_smoke"
    libdir = os.path.join(ROOT, "diffeq_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", os.path.join(HERE, "cpp", "hpp_smoke.cpp"), "-o", exe,
                           "-L" + libdir, "-ldiffeq_b200", "-Wl,-rpath," + libdir])
    return exe


def test_cpp_host_mirror_compiles_and_fails_loudly_without_gpu(deq):
    """include/diffeq_b200.hpp (C++ mirror of the reference API) builds against the C ABI; without a CUDA device the
    solve must fail with a CUDA error — never fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    out = subprocess.run([_build_hpp_smoke()], capture_output=True, text=True)
    assert out.returncode == 1 and "HPP_FAIL 7" in out.stdout


def test_c_abi_is_c99_clean_and_links_from_c(deq):
    """include/diffeq_b200.h compiles as strict C99 and the library links from a C program (no C++ types in the ABI)."""
    exe = "/tmp/deq_abi_smoke"
    libdir = os.path.join(ROOT, "diffeq_b200")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", os.path.join(HERE, "cpp", "abi_smoke.c"),
                           "-L" + libdir, "-ldiffeq_b200", "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "ABI_VERSION" in out.stdout and ("ABI_OK" in out.stdout or "ABI_NO_GPU" in out.stdout), out.stdout


def test_missing_extension_fails_loudly():
    """No CPU fallback: with the shared library absent the package refuses to work (ImportError), it does not fall back."""
    code = "import diffeq_b200 as D\ntry:\n    D.lib()\nexcept ImportError as e:\n    print('IMPORT_ERROR', 'no CPU fallback' in str(e))\n"
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT,
                         env=dict(os.environ, DEQ_LIB="/nonexistent/libdiffeq_b200.so"))
    assert "IMPORT_ERROR True" in out.stdout, out.stdout + out.stderr


def test_binding_struct_layout_matches_library(deq):
    import ctypes as C
    assert deq.lib().deq_abi_options_size() == C.sizeof(deq._Options)
    o = deq.lib().deq_options_default()
    assert (o.max_steps, o.refill, o.mode, o.precision, o.libm, o.dense_layout, o.output) == (10_000_000, 2, 0, 0, 0, 0, 0)


# ---- the Rust -sys crate cannot be compiled in this image: keep it in step with the header mechanically ----
def _c_type_to_rust(t):
    t = " ".join(t.replace("*", " * ").split())
    toks = t.split()
    base_map = {"double": "c_double", "int": "c_int", "float": "c_float", "char": "c_char", "void": "c_void", "size_t": "usize",
                "uint64_t": "u64", "uint32_t": "u32", "uint8_t": "u8", "deq_status": "deq_status", "deq_options": "deq_options",
                "deq_rhs": "deq_rhs", "deq_ensemble": "deq_ensemble", "deq_result": "deq_result"}
    const = False
    if toks[0] == "const":
        const, toks = True, toks[1:]
    base, stars = base_map[toks[0]], toks[1:]
    assert all(x == "*" for x in stars), t
    out = base
    for i, _ in enumerate(stars):
        out = ("*const " if (const and i == 0) else "*mut ") + out
    return out


def _parse_header(text):
    import re
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    fields = re.search(r"typedef struct deq_options \{(.*?)\} deq_options;", text, flags=re.S).group(1)
    opts = []
    for decl in fields.split(";"):
        decl = decl.strip()
        if decl:
            ty, name = decl.rsplit(None, 1)
            opts.append((name, _c_type_to_rust(ty)))
    status = {m.group(1): int(m.group(2)) for m in re.finditer(r"(DEQ_(?:OK|ERR_\w+))\s*=\s*(\d+)", text)}
    funcs = {}
    for m in re.finditer(r"(?m)^\s*((?:const\s+)?\w+\s*\**)\s*(deq_\w+)\s*\(([^;{}]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        alist = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                ty = re.sub(r"\b\w+$", "", a).strip() if not a.endswith("*") else a   # drop the parameter name
                alist.append(_c_type_to_rust(ty))
        funcs[name] = (None if ret.strip() == "void" else _c_type_to_rust(ret), alist)
    return opts, status, funcs


def _parse_rust_sys(text):
    import re
    body = re.search(r"pub struct deq_options \{(.*?)\n\}", text, flags=re.S).group(1)
    opts = [(m.group(1).replace("r#", ""), m.group(2).strip()) for m in re.finditer(r"pub\s+([\w#]+)\s*:\s*([^,\n]+),", body)]
    status = {m.group(1): int(m.group(2)) for m in re.finditer(r"pub const (DEQ_\w+): deq_status = (\d+);", text)}
    funcs = {}
    ext = re.search(r'extern "C" \{(.*)\n\}', text, flags=re.S).group(1)
    for m in re.finditer(r"pub fn (deq_\w+)\s*\((.*?)\)\s*(?:->\s*([^;]+))?;", ext, flags=re.S):
        name, args, ret = m.group(1), " ".join(m.group(2).split()), m.group(3)
        alist = [a.split(":", 1)[1].strip() for a in args.split(",") if a.strip()]
        funcs[name] = (ret.strip() if ret else None, alist)
    return opts, status, funcs


def test_rust_sys_crate_matches_the_header():
    """rust/diffeq-b200-sys/src/lib.rs declares exactly the functions of include/diffeq_b200.h with the same argument and
    return types, `#[repr(C)] deq_options` has the header's fields in the header's order, and the status codes agree — so the
    crate a maintainer compiles (there is no Rust toolchain in this image) binds the ABI the tests exercise."""
    hopts, hstatus, hfuncs = _parse_header(open(os.path.join(ROOT, "include", "diffeq_b200.h")).read())
    ropts, rstatus, rfuncs = _parse_rust_sys(open(os.path.join(ROOT, "rust", "diffeq-b200-sys", "src", "lib.rs")).read())
    assert len(hfuncs) >= 38 and len(hopts) == 17
    assert ropts == hopts
    assert rstatus == hstatus
    assert sorted(rfuncs) == sorted(hfuncs), (sorted(set(hfuncs) - set(rfuncs)), sorted(set(rfuncs) - set(hfuncs)))
    for name, sig in hfuncs.items():
        assert rfuncs[name] == sig, (name, rfuncs[name], sig)
    # the safe wrapper exposes the reference's per-method entry points (problem.rs:150-190) and the option kinds
    wrapper = open(os.path.join(ROOT, "rust", "diffeq-b200", "src", "lib.rs")).read()
    for fn in ("feuler", "heun", "midpoint", "ode4", "ode21", "ode23", "ode45", "ode45_dp", "ode45_fe", "ode78", "ode23s", "ode4s_kr", "ode4s_s",
               "solve_host", "tspan_linspace"):
        assert ("pub fn %s(" % fn) in wrapper, fn
    for ty in ("enum Points", "enum Output", "enum Mode", "enum TableauSet", "enum DenseLayout", "fn flatten", "fn ptr_or_null"):
        assert ty in wrapper, ty
    # every sys function the wrapper calls exists in the header
    import re
    for used in set(re.findall(r"sys::(deq_\w+)\(", wrapper)):
        assert used in hfuncs, used


def test_two_stage_column_permutation_plan_logic():
    """The factorisation behind `permute_staged_kernel` (kernels_common.cu), restated in numpy: with m(s) the stable rank of slot s
    under the key order[s] // BK, destination bucket b owns exactly the intermediate positions [BK b, BK (b + 1)), so stage 2 is a
    shuffle inside each BK-block, and stage 1 writes, per chunk of CH consecutive slots, runs of consecutive intermediate
    positions in ascending order.  Checked on ragged sizes (last chunk and last bucket partial), with the library's CH / BK."""
    CH, BK = 8192, 4096
    rng = np.random.default_rng(3)
    for n in (1, 5, 4096, 8192, 8193, 70001):
        order = rng.permutation(n).astype(np.uint32)                       # slot s -> destination column
        s1 = np.argsort(order // BK, kind="stable").astype(np.uint32)      # intermediate position m -> slot
        d2 = (order[s1] % BK).astype(np.uint16)                            # stage 2: position inside the block
        m1 = np.argsort(s1 // CH, kind="stable").astype(np.uint32)         # stage-1 list: positions grouped by source chunk, ascending
        l1 = (s1[m1] % CH).astype(np.uint16)                               # ... and the slot inside the chunk each one reads
        # bucket property: position m lies in the block of its destination
        assert np.array_equal(np.arange(n) // BK, order[s1] // BK)
        # the stage-1 list of chunk c covers exactly the slots of chunk c, positions ascending
        for c in range((n + CH - 1) // CH):
            lo, hi = c * CH, min(n, (c + 1) * CH)
            assert np.array_equal(np.sort(l1[lo:hi].astype(np.int64)), np.arange(hi - lo))
            assert np.all(np.diff(m1[lo:hi].astype(np.int64)) > 0)
        row = rng.standard_normal(n)                                       # one row in slot order
        mid = np.empty(n)
        for c in range((n + CH - 1) // CH):                                # stage 1, chunk by chunk through a "tile"
            lo, hi = c * CH, min(n, (c + 1) * CH)
            tile = row[lo:hi]
            mid[m1[lo:hi]] = tile[l1[lo:hi]]
        out = np.empty(n)
        for b in range((n + BK - 1) // BK):                                # stage 2, block by block
            lo, hi = b * BK, min(n, (b + 1) * BK)
            tile = np.empty(hi - lo)
            tile[d2[lo:hi]] = mid[lo:hi]
            out[lo:hi] = tile
        want = np.empty(n)
        want[order] = row                                                  # column order[s] <- slot s
        assert np.array_equal(out, want)
