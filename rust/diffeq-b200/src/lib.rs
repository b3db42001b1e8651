//! Safe wrapper mirroring `diffeq::ode::{problem::OdeProblem, Ode, options::*, solution::OdeSolution}` for ensembles.
//!
//! ```ignore
//! use diffeq_b200::{Ode, OdeOptionMap, OdeProblem, Rhs, System, Reltol, Abstol};
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
//! NOT compiled in the build image (no Rust toolchain); every call maps 1:1 onto the verified C ABI.
use diffeq_b200_sys as sys;
use std::ffi::{CStr, CString};
use std::ptr;

/// Same variants and order as `diffeq::ode::Ode` (src/ode/mod.rs:14-26).
#[derive(Clone, Copy, Debug)]
#[repr(i32)]
pub enum Ode { Feuler = 0, Heun, Midpoint, Ode23, Ode23s, Ode4, Ode45, Ode45fe, Ode4skr, Ode4ss, Ode78, Ode21 }

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

/// Mirrors `diffeq::error::OdeError` (src/error.rs:4-23) plus backend variants.
#[derive(Debug)]
pub enum OdeError {
    Uninitialized { msg: String }, InvalidButcherTableauWeightType, NAN, ZeroTimeSpan, InvalidInitstep, InvalidMatrix,
    Backend { status: i32, msg: String },
}

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
        let c = CString::new(src).map_err(|e| OdeError::Backend { status: 10, msg: e.to_string() })?;
        let mut h = ptr::null_mut();
        check(unsafe { sys::deq_rhs_from_source(c.as_ptr(), n as i32, nparams as i32, &mut h) })?;
        Ok(Rhs(h))
    }
    pub fn dim(&self) -> usize { unsafe { sys::deq_rhs_dim(self.0) as usize } }
}
impl Drop for Rhs { fn drop(&mut self) { unsafe { sys::deq_rhs_destroy(self.0) } } }

// Option newtypes, src/ode/options.rs:258-281
pub struct Reltol(pub f64); pub struct Abstol(pub f64); pub struct Minstep(pub f64); pub struct Maxstep(pub f64);
pub struct Initstep(pub f64);
pub trait IntoOption { fn apply(self, o: &mut sys::deq_options); }
impl IntoOption for Reltol { fn apply(self, o: &mut sys::deq_options) { o.reltol = self.0 } }
impl IntoOption for Abstol { fn apply(self, o: &mut sys::deq_options) { o.abstol = self.0 } }
impl IntoOption for Minstep { fn apply(self, o: &mut sys::deq_options) { o.has_minstep = 1; o.minstep = self.0 } }
impl IntoOption for Maxstep { fn apply(self, o: &mut sys::deq_options) { o.has_maxstep = 1; o.maxstep = self.0 } }
impl IntoOption for Initstep { fn apply(self, o: &mut sys::deq_options) { o.initstep = self.0 } }

/// `OdeOptionMap` (src/ode/options.rs:8-25); defaults as options.rs:283-299.
pub struct OdeOptionMap(sys::deq_options);
impl Default for OdeOptionMap { fn default() -> Self { OdeOptionMap(unsafe { sys::deq_options_default() }) } }
impl OdeOptionMap { pub fn insert<T: IntoOption>(&mut self, o: T) { o.apply(&mut self.0) } }

/// `OdeSolution { tout, yout }` (src/ode/solution.rs:22-29) + the diagnostics the reference drops.
pub struct OdeSolution { pub tout: Vec<f64>, pub yout: Vec<Vec<f64>>, pub accepted: Vec<u32>, pub rejected: Vec<u32> }

#[derive(Default)]
pub struct OdeBuilder { f: Option<Rhs>, y0: Option<Vec<Vec<f64>>>, tspan: Option<Vec<f64>>, params: Vec<f64> }
impl OdeBuilder {
    pub fn fun(mut self, f: Rhs) -> Self { self.f = Some(f); self }
    pub fn params(mut self, p: Vec<f64>) -> Self { self.params = p; self }
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
        let flat: Vec<f64> = y0.iter().flatten().copied().collect();
        let mut h = ptr::null_mut();
        check(unsafe { sys::deq_ensemble_create(f.0, y0.len(), flat.as_ptr(), self.params.as_ptr(), 0, tspan.as_ptr(),
                                                tspan.len(), &mut h) })?;
        Ok(OdeProblem { h, n: f.dim(), ntraj: y0.len(), tspan, _f: f })
    }
}

pub struct OdeProblem { h: *mut sys::deq_ensemble, n: usize, ntraj: usize, tspan: Vec<f64>, _f: Rhs }
impl OdeProblem {
    pub fn builder() -> OdeBuilder { OdeBuilder::default() }
    /// src/ode/problem.rs:133-147
    pub fn solve(&self, ode: Ode, opts: OdeOptionMap) -> Result<OdeSolution, OdeError> {
        let mut r = ptr::null_mut();
        check(unsafe { sys::deq_solve(self.h, ode as i32, 0, &opts.0, &mut r) })?;
        let mut flat = vec![0.0; self.ntraj * self.n];
        let (mut acc, mut rej) = (vec![0u32; self.ntraj], vec![0u32; self.ntraj]);
        let rc = unsafe { sys::deq_result_final(r, flat.as_mut_ptr()) };
        let adaptive = !matches!(ode, Ode::Feuler | Ode::Heun | Ode::Midpoint | Ode::Ode4);
        let rc2 = if adaptive { unsafe { sys::deq_result_counts(r, acc.as_mut_ptr(), rej.as_mut_ptr(), ptr::null_mut(), ptr::null_mut(), ptr::null_mut()) } } else { 0 };
        unsafe { sys::deq_result_destroy(r) };
        check(rc)?; check(rc2)?;
        Ok(OdeSolution { tout: vec![*self.tspan.last().unwrap_or(&0.0)], yout: flat.chunks(self.n).map(|c| c.to_vec()).collect(), accepted: acc, rejected: rej })
    }
    pub fn ode45(&self, opts: OdeOptionMap) -> Result<OdeSolution, OdeError> { self.solve(Ode::Ode45, opts) }
    pub fn ode45_fe(&self, opts: OdeOptionMap) -> Result<OdeSolution, OdeError> { self.solve(Ode::Ode45fe, opts) }
    pub fn ode78(&self, opts: OdeOptionMap) -> Result<OdeSolution, OdeError> { self.solve(Ode::Ode78, opts) }
    pub fn ode4(&self) -> Result<OdeSolution, OdeError> { self.solve(Ode::Ode4, OdeOptionMap::default()) }
    /// problem.rs:409 — per-trajectory status 3 maps to the reference's Err(OdeError::InvalidMatrix) for a single problem
    pub fn ode23s(&self, opts: OdeOptionMap) -> Result<OdeSolution, OdeError> { self.solve(Ode::Ode23s, opts) }
    pub fn ode4s_kr(&self) -> Result<OdeSolution, OdeError> { self.solve(Ode::Ode4skr, OdeOptionMap::default()) }   // problem.rs:643
    pub fn ode4s_s(&self) -> Result<OdeSolution, OdeError> { self.solve(Ode::Ode4ss, OdeOptionMap::default()) }     // problem.rs:648
}
impl Drop for OdeProblem { fn drop(&mut self) { unsafe { sys::deq_ensemble_destroy(self.h) } } }
