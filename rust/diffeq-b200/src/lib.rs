//! Safe wrapper mirroring `diffeq::ode::{problem::OdeProblem, Ode, options::*, solution::OdeSolution}` for ensembles.
//!
//! ```ignore
//! use diffeq_b200::{Abstol, Ode, OdeOptionMap, OdeProblem, Reltol, Rhs, System};
//! let problem = OdeProblem::builder()
//!     .fun(Rhs::builtin(System::Lorenz)?)          // device functor instead of a closure
//!     .params(vec![10.0, 28.0, 8.0 / 3.0])         // what the closure used to capture
//!     .init(vec![vec![0.1, 0.0, 0.0]; 1 << 20])    // one `Y` per trajectory
//!     .tspan(vec![0.0, 10.0])
//!     .build()?;
//! let mut opts = OdeOptionMap::default();
//! opts.insert(Reltol(1e-8)); opts.insert(Abstol(1e-8));
//! let solution = problem.solve(Ode::Ode45, opts)?;  // solution.yout[i] = final state of trajectory i
//! ```
//! NOT compiled in the build image (no Rust toolchain); every call maps 1:1 onto the verified C ABI
//! (`include/diffeq_b200.h`; `tests/test_host_cpu.py::test_rust_sys_crate_matches_the_header` keeps the `-sys` crate in step
//! with it).  Every buffer handed to the C ABI is length-checked here first: the library reads `ntraj * n` state and
//! `nparams` (or `ntraj * nparams`) parameter doubles through raw pointers.
use diffeq_b200_sys as sys;
use std::ffi::{CStr, CString};
use std::ptr;

/// Same variants and order as `diffeq::ode::Ode` (src/ode/mod.rs:14-26); `Ode21` is `.ode21()` (problem.rs:164-166).
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum Ode { Feuler = 0, Heun, Midpoint, Ode23, Ode23s, Ode4, Ode45, Ode45fe, Ode4skr, Ode4ss, Ode78, Ode21 }

impl Ode {
    fn is_fixed_step(self) -> bool { matches!(self, Ode::Feuler | Ode::Heun | Ode::Midpoint | Ode::Ode4) }
    /// per-trajectory counters exist for the adaptive methods and for the Rosenbrock ones (a step can stop a trajectory)
    fn has_counters(self) -> bool { !self.is_fixed_step() }
}

impl std::str::FromStr for Ode {
    type Err = String;
    fn from_str(s: &str) -> Result<Self, String> {  // src/ode/mod.rs:28-46
        Ok(match s {
            "feuler" => Ode::Feuler, "heun" => Ode::Heun, "midpoint" => Ode::Midpoint, "ode23" => Ode::Ode23,
            "ode23s" => Ode::Ode23s, "ode4" => Ode::Ode4, "ode45" => Ode::Ode45, "ode4skr" => Ode::Ode4skr,
            "ode4s" => Ode::Ode4ss, "ode78" => Ode::Ode78,
            _ => return Err(format!("{} is not a valid Ode identifier", s)),
        })
    }
}

#[derive(Clone, Copy, Debug)]
#[repr(i32)]
pub enum System { Lorenz = 0, VanDerPol = 1, Pendulum = 2 }

/// `deq_tableau_set`: the crate's coefficients bug for bug, or the corrected textbook ones.
#[derive(Clone, Copy, Debug, Default)]
#[repr(i32)]
pub enum TableauSet { #[default] Reference = 0, Textbook = 1 }

/// `Points` (src/ode/options.rs:120-137).
#[derive(Clone, Copy, Debug)]
#[repr(i32)]
pub enum Points { All = 0, Specified = 1 }

/// `deq_output`: what a solve materialises.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum Output { Final = 0, Tspan = 1, All = 2 }

/// `deq_mode`: bit-identical to the reference (default) or FMA / zero-skipping arithmetic.
#[derive(Clone, Copy, Debug)]
#[repr(i32)]
pub enum Mode { Parity = 0, Fast = 1 }

/// `deq_libm`: libm calls as written in the source, or as LLVM folds them in an optimised build.
#[derive(Clone, Copy, Debug)]
#[repr(i32)]
pub enum Libm { AsWritten = 0, LlvmFolded = 1 }

/// `deq_dense_layout` of `Output::Tspan`.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum DenseLayout { SoA = 0, TrajMajor = 1 }

/// Mirrors `diffeq::error::OdeError` (src/error.rs:4-23) plus backend variants.
#[derive(Debug)]
pub enum OdeError {
    Uninitialized { msg: String }, InvalidButcherTableauWeightType, NAN, ZeroTimeSpan, InvalidInitstep, InvalidMatrix,
    Backend { status: i32, msg: String },
}

fn invalid(msg: String) -> OdeError { OdeError::Backend { status: sys::DEQ_ERR_INVALID_ARG, msg } }

fn check(rc: sys::deq_status) -> Result<(), OdeError> {
    if rc == sys::DEQ_OK { return Ok(()); }
    let msg = unsafe { CStr::from_ptr(sys::deq_last_error_message()) }.to_string_lossy().into_owned();
    Err(match rc {
        sys::DEQ_ERR_UNINITIALIZED => OdeError::Uninitialized { msg },
        sys::DEQ_ERR_WEIGHT_TYPE => OdeError::InvalidButcherTableauWeightType,
        sys::DEQ_ERR_NAN => OdeError::NAN,
        sys::DEQ_ERR_ZERO_TIMESPAN => OdeError::ZeroTimeSpan,
        sys::DEQ_ERR_INVALID_INITSTEP => OdeError::InvalidInitstep,
        sys::DEQ_ERR_INVALID_MATRIX => OdeError::InvalidMatrix,
        status => OdeError::Backend { status, msg },
    })
}

pub struct Rhs(*mut sys::deq_rhs);
impl Rhs {
    pub fn builtin(s: System) -> Result<Self, OdeError> {
        let mut h = ptr::null_mut();
        check(unsafe { sys::deq_rhs_builtin(s as i32, &mut h) })?;
        Ok(Rhs(h))
    }
    /// `__device__ void rhs(double t, const double* y, double* dy, const double* p)` compiled with NVRTC.
    pub fn from_source(src: &str, n: usize, nparams: usize) -> Result<Self, OdeError> {
        let c = CString::new(src).map_err(|e| invalid(e.to_string()))?;
        let mut h = ptr::null_mut();
        check(unsafe { sys::deq_rhs_from_source(c.as_ptr(), n as i32, nparams as i32, &mut h) })?;
        Ok(Rhs(h))
    }
    pub fn dim(&self) -> usize { unsafe { sys::deq_rhs_dim(self.0) as usize } }
    pub fn nparams(&self) -> usize { unsafe { sys::deq_rhs_nparams(self.0) as usize } }
}
impl Drop for Rhs { fn drop(&mut self) { unsafe { sys::deq_rhs_destroy(self.0) } } }

/// What the reference's closure captures: one parameter set for the whole ensemble, or one per trajectory.
#[derive(Clone, Debug)]
pub enum Params { None, Shared(Vec<f64>), PerTrajectory(Vec<Vec<f64>>) }

/// Flattens and LENGTH-CHECKS the buffers the C ABI reads through raw pointers.
struct Flat { y0: Vec<f64>, params: Vec<f64>, per_traj: i32, ntraj: usize }
fn flatten(f: &Rhs, y0: &[Vec<f64>], params: &Params) -> Result<Flat, OdeError> {
    let (n, np, ntraj) = (f.dim(), f.nparams(), y0.len());
    if let Some((i, row)) = y0.iter().enumerate().find(|(_, r)| r.len() != n) {
        return Err(invalid(format!("initial state {} has {} components, the right-hand side expects {}", i, row.len(), n)));
    }
    let (flat_p, per_traj) = match params {
        Params::None if np == 0 => (Vec::new(), 0),
        Params::None => return Err(OdeError::Uninitialized { msg: "Required problem parameters must be initialized".into() }),
        Params::Shared(p) if p.len() == np => (p.clone(), 0),
        Params::Shared(p) => return Err(invalid(format!("{} shared parameters given, the right-hand side expects {}", p.len(), np))),
        Params::PerTrajectory(rows) => {
            if rows.len() != ntraj { return Err(invalid(format!("{} parameter rows for {} trajectories", rows.len(), ntraj))); }
            if let Some((i, r)) = rows.iter().enumerate().find(|(_, r)| r.len() != np) {
                return Err(invalid(format!("parameter row {} has {} entries, the right-hand side expects {}", i, r.len(), np)));
            }
            (rows.iter().flatten().copied().collect(), 1)
        }
    };
    Ok(Flat { y0: y0.iter().flatten().copied().collect(), params: flat_p, per_traj, ntraj })
}
/// NULL for an empty buffer (never the dangling pointer of an empty Vec).
fn ptr_or_null(v: &[f64]) -> *const f64 { if v.is_empty() { ptr::null() } else { v.as_ptr() } }

// Option newtypes, src/ode/options.rs:258-281 (+ the extensions of the C ABI)
pub struct Reltol(pub f64); pub struct Abstol(pub f64); pub struct Minstep(pub f64); pub struct Maxstep(pub f64);
pub struct Initstep(pub f64); pub struct MaxSteps(pub u64); pub struct ShardChunk(pub u64);
pub trait IntoOption { fn apply(self, o: &mut sys::deq_options); }
impl IntoOption for Reltol { fn apply(self, o: &mut sys::deq_options) { o.reltol = self.0 } }
impl IntoOption for Abstol { fn apply(self, o: &mut sys::deq_options) { o.abstol = self.0 } }
impl IntoOption for Minstep { fn apply(self, o: &mut sys::deq_options) { o.has_minstep = 1; o.minstep = self.0 } }
impl IntoOption for Maxstep { fn apply(self, o: &mut sys::deq_options) { o.has_maxstep = 1; o.maxstep = self.0 } }
impl IntoOption for Initstep { fn apply(self, o: &mut sys::deq_options) { o.initstep = self.0 } }
impl IntoOption for MaxSteps { fn apply(self, o: &mut sys::deq_options) { o.max_steps = self.0 } }
impl IntoOption for ShardChunk { fn apply(self, o: &mut sys::deq_options) { o.shard_chunk = self.0 } }
impl IntoOption for Points { fn apply(self, o: &mut sys::deq_options) { o.points = self as i32 } }
impl IntoOption for Output { fn apply(self, o: &mut sys::deq_options) { o.output = self as i32 } }
impl IntoOption for Mode { fn apply(self, o: &mut sys::deq_options) { o.mode = self as i32 } }
impl IntoOption for Libm { fn apply(self, o: &mut sys::deq_options) { o.libm = self as i32 } }
impl IntoOption for DenseLayout { fn apply(self, o: &mut sys::deq_options) { o.dense_layout = self as i32 } }

/// `OdeOptionMap` (src/ode/options.rs:8-25); defaults as options.rs:283-299.
#[derive(Clone, Copy)]
pub struct OdeOptionMap(sys::deq_options);
impl Default for OdeOptionMap {
    fn default() -> Self {
        assert_eq!(unsafe { sys::deq_abi_options_size() }, std::mem::size_of::<sys::deq_options>(), "deq_options layout mismatch");
        OdeOptionMap(unsafe { sys::deq_options_default() })
    }
}
impl OdeOptionMap { pub fn insert<T: IntoOption>(&mut self, o: T) { o.apply(&mut self.0) } }

/// `OdeSolution { tout, yout }` (src/ode/solution.rs:22-29) per trajectory + the diagnostics the reference counts and drops.
/// `Output::Final`: `tout = [t_reached]`, `yout = [final state]`.  `Output::Tspan`: the tspan rows the reference would have
/// emitted (`nvalid` leading rows, and the end point).  `Output::All`: the complete `Points::All` stream.
pub struct OdeSolution { pub tout: Vec<f64>, pub yout: Vec<Vec<f64>> }
pub struct EnsembleSolution {
    pub solutions: Vec<OdeSolution>,
    pub accepted: Vec<u32>, pub rejected: Vec<u32>, pub nfe: Vec<u32>,
    /// `deq_traj_status`: 0 done, 1 minstep break, 2 max_steps, 3 invalid matrix (the reference's `Err(InvalidMatrix)`)
    pub status: Vec<u8>,
}

#[derive(Default)]
pub struct OdeBuilder { f: Option<Rhs>, y0: Option<Vec<Vec<f64>>>, tspan: Option<Vec<f64>>, params: Option<Params> }
impl OdeBuilder {
    pub fn fun(mut self, f: Rhs) -> Self { self.f = Some(f); self }
    pub fn params(mut self, p: Vec<f64>) -> Self { self.params = Some(Params::Shared(p)); self }
    pub fn params_per_trajectory(mut self, p: Vec<Vec<f64>>) -> Self { self.params = Some(Params::PerTrajectory(p)); self }
    pub fn init(mut self, y0: Vec<Vec<f64>>) -> Self { self.y0 = Some(y0); self }
    pub fn tspan(mut self, t: Vec<f64>) -> Self { self.tspan = Some(t); self }
    pub fn tspan_linspace(mut self, from: f64, to: f64, n: usize) -> Self {
        let mut t = vec![0.0; n];
        unsafe { sys::deq_tspan_linspace(from, to, n, t.as_mut_ptr()) };
        self.tspan = Some(t); self
    }
    /// src/ode/problem.rs:92-104
    pub fn build(self) -> Result<OdeProblem, OdeError> {
        let f = self.f.ok_or(OdeError::Uninitialized { msg: "Required problem must be initialized".into() })?;
        let y0 = self.y0.ok_or(OdeError::Uninitialized { msg: "Initial starting point must be initialized".into() })?;
        let tspan = self.tspan.ok_or(OdeError::Uninitialized { msg: "Time span must be initialized".into() })?;
        let params = self.params.unwrap_or(Params::None);
        let flat = flatten(&f, &y0, &params)?;
        let mut h = ptr::null_mut();
        check(unsafe { sys::deq_ensemble_create(f.0, flat.ntraj, ptr_or_null(&flat.y0), ptr_or_null(&flat.params), flat.per_traj,
                                                tspan.as_ptr(), tspan.len(), &mut h) })?;
        Ok(OdeProblem { h, n: f.dim(), ntraj: flat.ntraj, tspan, _f: f })
    }
}

pub struct OdeProblem { h: *mut sys::deq_ensemble, n: usize, ntraj: usize, tspan: Vec<f64>, _f: Rhs }
impl OdeProblem {
    pub fn builder() -> OdeBuilder { OdeBuilder::default() }

    /// src/ode/problem.rs:133-147
    pub fn solve(&self, ode: Ode, opts: OdeOptionMap) -> Result<EnsembleSolution, OdeError> { self.solve_with(ode, TableauSet::Reference, opts) }

    pub fn solve_with(&self, ode: Ode, set: TableauSet, opts: OdeOptionMap) -> Result<EnsembleSolution, OdeError> {
        struct Res(*mut sys::deq_result);
        impl Drop for Res { fn drop(&mut self) { unsafe { sys::deq_result_destroy(self.0) } } }
        let mut r = ptr::null_mut();
        check(unsafe { sys::deq_solve(self.h, ode as i32, set as i32, &opts.0, &mut r) })?;
        let r = Res(r);
        let (n, ntraj, nt) = (self.n, self.ntraj, self.tspan.len());
        let mut out = EnsembleSolution { solutions: Vec::with_capacity(ntraj), accepted: vec![0; ntraj], rejected: vec![0; ntraj],
                                         nfe: vec![0; ntraj], status: vec![0; ntraj] };
        if nt == 0 || ntraj == 0 {   // problem.rs:211-214: nothing to solve, empty solutions
            out.solutions = (0..ntraj).map(|_| OdeSolution { tout: Vec::new(), yout: Vec::new() }).collect();
            return Ok(out);
        }
        let mut t_reached = vec![0.0; ntraj];
        if ode.has_counters() {
            check(unsafe { sys::deq_result_counts(r.0, out.accepted.as_mut_ptr(), out.rejected.as_mut_ptr(), out.nfe.as_mut_ptr(),
                                                  out.status.as_mut_ptr(), t_reached.as_mut_ptr()) })?;
        } else {
            t_reached.iter_mut().for_each(|t| *t = self.tspan[nt - 1]);
        }
        let output = opts.0.output;
        if output == Output::All as i32 && !ode.is_fixed_step() {
            let mut offsets = vec![0u64; ntraj + 1];
            check(unsafe { sys::deq_result_all_offsets(r.0, offsets.as_mut_ptr()) })?;
            let rows = offsets[ntraj] as usize;
            let (mut tout, mut yout) = (vec![0.0; rows], vec![0.0; rows * n]);
            if rows > 0 { check(unsafe { sys::deq_result_all_copy(r.0, 0, ntraj, tout.as_mut_ptr(), yout.as_mut_ptr()) })?; }
            for i in 0..ntraj {
                let (a, b) = (offsets[i] as usize, offsets[i + 1] as usize);
                out.solutions.push(OdeSolution { tout: tout[a..b].to_vec(), yout: yout[a * n..b * n].chunks(n).map(|c| c.to_vec()).collect() });
            }
        } else if output == Output::Tspan as i32 {
            // one trajectory-range at a time keeps the host copy bounded; layout as the library reports it
            let tm = opts.0.dense_layout == DenseLayout::TrajMajor as i32 && !ode.is_fixed_step()
                && !matches!(ode, Ode::Ode23s | Ode::Ode4skr | Ode::Ode4ss);
            let mut dense = vec![0.0; n * nt * ntraj];
            let mut nvalid = vec![0u32; ntraj];
            check(unsafe { sys::deq_result_dense(r.0, 0, ntraj, dense.as_mut_ptr(), nvalid.as_mut_ptr()) })?;
            for i in 0..ntraj {
                let rows = (nvalid[i] as usize).min(nt);
                let mut idx: Vec<usize> = (0..rows).collect();
                if !ode.is_fixed_step() && !matches!(ode, Ode::Ode23s) && rows < nt { idx.push(nt - 1); }   // the end point is always pushed (problem.rs:320-322)
                let yout = idx.iter().map(|&k| (0..n).map(|d| if tm { dense[(i * nt + k) * n + d] } else { dense[(d * nt + k) * ntraj + i] }).collect()).collect();
                out.solutions.push(OdeSolution { tout: idx.iter().map(|&k| self.tspan[k]).collect(), yout });
            }
        } else {
            let mut flat = vec![0.0; ntraj * n];
            check(unsafe { sys::deq_result_final(r.0, flat.as_mut_ptr()) })?;
            for (i, c) in flat.chunks(n).enumerate() {
                out.solutions.push(OdeSolution { tout: vec![t_reached[i]], yout: vec![c.to_vec()] });
            }
        }
        Ok(out)
    }

    // the per-method entry points of problem.rs:150-190 (the fixed-step ones take no options there either)
    pub fn feuler(&self) -> Result<EnsembleSolution, OdeError> { self.solve(Ode::Feuler, OdeOptionMap::default()) }
    pub fn heun(&self) -> Result<EnsembleSolution, OdeError> { self.solve(Ode::Heun, OdeOptionMap::default()) }
    pub fn midpoint(&self) -> Result<EnsembleSolution, OdeError> { self.solve(Ode::Midpoint, OdeOptionMap::default()) }
    pub fn ode4(&self) -> Result<EnsembleSolution, OdeError> { self.solve(Ode::Ode4, OdeOptionMap::default()) }
    pub fn ode21(&self, opts: OdeOptionMap) -> Result<EnsembleSolution, OdeError> { self.solve(Ode::Ode21, opts) }
    pub fn ode23(&self, opts: OdeOptionMap) -> Result<EnsembleSolution, OdeError> { self.solve(Ode::Ode23, opts) }
    pub fn ode45(&self, opts: OdeOptionMap) -> Result<EnsembleSolution, OdeError> { self.solve(Ode::Ode45, opts) }
    pub fn ode45_dp(&self, opts: OdeOptionMap) -> Result<EnsembleSolution, OdeError> { self.solve(Ode::Ode45, opts) }
    pub fn ode45_fe(&self, opts: OdeOptionMap) -> Result<EnsembleSolution, OdeError> { self.solve(Ode::Ode45fe, opts) }
    pub fn ode78(&self, opts: OdeOptionMap) -> Result<EnsembleSolution, OdeError> { self.solve(Ode::Ode78, opts) }
    /// problem.rs:409 — per-trajectory status 3 is the reference's `Err(OdeError::InvalidMatrix)` of a single problem
    pub fn ode23s(&self, opts: OdeOptionMap) -> Result<EnsembleSolution, OdeError> { self.solve(Ode::Ode23s, opts) }
    pub fn ode4s_kr(&self) -> Result<EnsembleSolution, OdeError> { self.solve(Ode::Ode4skr, OdeOptionMap::default()) }   // problem.rs:643
    pub fn ode4s_s(&self) -> Result<EnsembleSolution, OdeError> { self.solve(Ode::Ode4ss, OdeOptionMap::default()) }     // problem.rs:648
}
impl Drop for OdeProblem { fn drop(&mut self) { unsafe { sys::deq_ensemble_destroy(self.h) } } }

/// Final states and counters of `solve_host`.
pub struct HostSolution { pub yfinal: Vec<Vec<f64>>, pub accepted: Vec<u32>, pub rejected: Vec<u32>, pub nfe: Vec<u32>, pub status: Vec<u8>, pub t_reached: Vec<f64> }

/// `deq_solve_host`: one call on host buffers, sharded over `devices` (CUDA ordinals; empty = the current device) with one
/// host thread and stream per GPU inside the library — the multi-GPU form of `OdeProblem::solve` for a process that owns
/// all GPUs of the box.  The result of trajectory i does not depend on the device list.
pub fn solve_host(f: &Rhs, y0: &[Vec<f64>], params: &Params, tspan: &[f64], ode: Ode, set: TableauSet, opts: OdeOptionMap,
                  devices: &[i32]) -> Result<HostSolution, OdeError> {
    if opts.0.output != Output::Final as i32 { return Err(invalid("solve_host returns final states and counters (Output::Final)".into())); }
    let flat = flatten(f, y0, params)?;
    let (n, ntraj) = (f.dim(), flat.ntraj);
    let mut yf = vec![0.0; ntraj * n];
    let mut s = HostSolution { yfinal: Vec::new(), accepted: vec![0; ntraj], rejected: vec![0; ntraj], nfe: vec![0; ntraj],
                               status: vec![0; ntraj], t_reached: vec![0.0; ntraj] };
    check(unsafe { sys::deq_solve_host(f.0, ntraj, ptr_or_null(&flat.y0), ptr_or_null(&flat.params), flat.per_traj, tspan.as_ptr(), tspan.len(),
                                       ode as i32, set as i32, &opts.0, if devices.is_empty() { ptr::null() } else { devices.as_ptr() },
                                       devices.len() as i32, yf.as_mut_ptr(), s.accepted.as_mut_ptr(), s.rejected.as_mut_ptr(),
                                       s.nfe.as_mut_ptr(), s.status.as_mut_ptr(), s.t_reached.as_mut_ptr()) })?;
    s.yfinal = yf.chunks(n.max(1)).map(|c| c.to_vec()).collect();
    Ok(s)
}
