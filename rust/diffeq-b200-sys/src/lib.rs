//! Raw bindings: one `extern "C"` item per declaration in `include/diffeq_b200.h`.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_float, c_int, c_void};

pub type deq_status = c_int;
pub const DEQ_OK: deq_status = 0;
pub const DEQ_ERR_UNINITIALIZED: deq_status = 1;
pub const DEQ_ERR_WEIGHT_TYPE: deq_status = 2;
pub const DEQ_ERR_NAN: deq_status = 3;
pub const DEQ_ERR_ZERO_TIMESPAN: deq_status = 4;
pub const DEQ_ERR_INVALID_INITSTEP: deq_status = 5;
pub const DEQ_ERR_INVALID_MATRIX: deq_status = 6;
pub const DEQ_ERR_CUDA: deq_status = 7;
pub const DEQ_ERR_NVRTC: deq_status = 8;
pub const DEQ_ERR_UNSUPPORTED: deq_status = 9;
pub const DEQ_ERR_INVALID_ARG: deq_status = 10;

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct deq_options {
    pub reltol: c_double,
    pub abstol: c_double,
    pub has_minstep: c_int,
    pub minstep: c_double,
    pub has_maxstep: c_int,
    pub maxstep: c_double,
    pub initstep: c_double,
    pub points: c_int,
    pub output: c_int,
    pub max_steps: u64,
    pub refill: c_int,
    pub r#async: c_int,
    pub mode: c_int,
    pub precision: c_int,
    pub libm: c_int,
    pub dense_layout: c_int,
    pub shard_chunk: u64,
}

#[repr(C)] pub struct deq_rhs { _p: [u8; 0] }
#[repr(C)] pub struct deq_ensemble { _p: [u8; 0] }
#[repr(C)] pub struct deq_result { _p: [u8; 0] }

extern "C" {
    pub fn deq_last_error_message() -> *const c_char;
    pub fn deq_version() -> *const c_char;
    pub fn deq_abi_options_size() -> usize;
    pub fn deq_rhs_builtin(system: c_int, out: *mut *mut deq_rhs) -> deq_status;
    pub fn deq_rhs_from_source(cuda_src: *const c_char, n: c_int, nparams: c_int, out: *mut *mut deq_rhs) -> deq_status;
    pub fn deq_rhs_dim(rhs: *const deq_rhs) -> c_int;
    pub fn deq_rhs_nparams(rhs: *const deq_rhs) -> c_int;
    pub fn deq_rhs_destroy(rhs: *mut deq_rhs);
    pub fn deq_tspan_linspace(from: c_double, to: c_double, n: usize, out: *mut c_double) -> deq_status;
    pub fn deq_ensemble_create(rhs: *const deq_rhs, ntraj: usize, y0: *const c_double, params: *const c_double,
                               params_per_traj: c_int, tspan: *const c_double, ntspan: usize,
                               out: *mut *mut deq_ensemble) -> deq_status;
    pub fn deq_ensemble_create_device(rhs: *const deq_rhs, ntraj: usize, d_y0_soa: *const c_double,
                                      params: *const c_double, params_per_traj: c_int, tspan: *const c_double,
                                      ntspan: usize, out: *mut *mut deq_ensemble) -> deq_status;
    pub fn deq_ensemble_destroy(ens: *mut deq_ensemble);
    pub fn deq_ensemble_ntraj(ens: *const deq_ensemble) -> usize;
    pub fn deq_set_device(device: c_int) -> deq_status;
    pub fn deq_get_device(device: *mut c_int) -> deq_status;
    pub fn deq_set_stream(cuda_stream: *mut c_void) -> deq_status;
    pub fn deq_options_default() -> deq_options;
    pub fn deq_solve(ens: *mut deq_ensemble, method: c_int, tableau_set: c_int, opts: *const deq_options,
                     out: *mut *mut deq_result) -> deq_status;
    pub fn deq_solve_host(rhs: *const deq_rhs, ntraj: usize, y0: *const c_double, params: *const c_double, params_per_traj: c_int,
                          tspan: *const c_double, ntspan: usize, method: c_int, tableau_set: c_int, opts: *const deq_options,
                          devices: *const c_int, ndevices: c_int, yfinal: *mut c_double, accepted: *mut u32, rejected: *mut u32,
                          nfe: *mut u32, status: *mut u8, t_reached: *mut c_double) -> deq_status;
    pub fn deq_solve_host_dense(rhs: *const deq_rhs, ntraj: usize, y0: *const c_double, params: *const c_double, params_per_traj: c_int,
                                tspan: *const c_double, ntspan: usize, method: c_int, tableau_set: c_int, opts: *const deq_options,
                                devices: *const c_int, ndevices: c_int, yfinal: *mut c_double, accepted: *mut u32, rejected: *mut u32,
                                nfe: *mut u32, status: *mut u8, t_reached: *mut c_double, dense: *mut c_double, nvalid: *mut u32) -> deq_status;
    pub fn deq_result_final(res: *mut deq_result, y: *mut c_double) -> deq_status;
    pub fn deq_result_counts(res: *mut deq_result, accepted: *mut u32, rejected: *mut u32, nfe: *mut u32,
                             status: *mut u8, t_reached: *mut c_double) -> deq_status;
    pub fn deq_result_totals(res: *mut deq_result, accepted: *mut u64, rejected: *mut u64, nfe: *mut u64) -> deq_status;
    pub fn deq_result_dense(res: *mut deq_result, traj_begin: usize, traj_end: usize, out: *mut c_double,
                            nvalid: *mut u32) -> deq_status;
    pub fn deq_result_all_offsets(res: *mut deq_result, offsets: *mut u64) -> deq_status;
    pub fn deq_result_all_copy(res: *mut deq_result, traj_begin: usize, traj_end: usize, tout: *mut c_double,
                               yout: *mut c_double) -> deq_status;
    pub fn deq_result_device_ptrs(res: *mut deq_result, d_yfinal_soa: *mut *const c_double,
                                  d_dense_soa: *mut *const c_double) -> deq_status;
    pub fn deq_result_trajectory(res: *mut deq_result, traj: usize, yout: *mut c_double) -> deq_status;
    pub fn deq_result_kernel_ms(res: *mut deq_result, ms: *mut c_float) -> deq_status;
    pub fn deq_result_destroy(res: *mut deq_result);
    pub fn deq_tableau_get(method: c_int, tableau_set: c_int, stages: *mut c_int, a: *mut c_double, b: *mut c_double,
                           c: *mut c_double, adaptive: *mut c_int, order_min: *mut c_int, fsal: *mut c_int) -> deq_status;
    pub fn deq_rosenbrock_get(method: c_int, gamma: *mut c_double, a: *mut c_double, b: *mut c_double, c: *mut c_double) -> deq_status;
    pub fn deq_host_alloc(bytes: usize, out: *mut *mut c_void) -> deq_status;
    pub fn deq_host_free(p: *mut c_void);
    pub fn deq_test_pow(x: *const c_double, y: *const c_double, out: *mut c_double, n: usize) -> deq_status;
    pub fn deq_test_log10(x: *const c_double, out: *mut c_double, n: usize) -> deq_status;
    pub fn deq_test_div(seed: u64, count: u64, mismatches: *mut u64, checked: *mut u64) -> deq_status;
    pub fn deq_measure_fp64_peak(tflops: *mut c_double, sm_clock_mhz_hint: *mut c_double) -> deq_status;
}
