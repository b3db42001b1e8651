// params.h — plain-old-data launch descriptors shared by the kernels (stepper.cuh), the per-tableau launchers
// (kernels_inst.cu) and the C ABI implementation (capi.cu).  All pointers are DEVICE pointers; ensemble arrays are
// structure-of-arrays over trajectories ([component][trajectory]) so that a warp's loads and stores coalesce.
#pragma once
#include "deq_rtc_compat.h"

namespace deqk {

enum : int { MODE_FINAL = 0, MODE_DENSE = 1, MODE_DENSE_TM = 2, MODE_REC = 3 };   // kernel template parameter
// launcher-level id of the Points::Specified flavour: adapt_kernel<..., MODE_REC, FAST = false, SPEC = true>
enum : int { MODE_REC_SPEC = 4 };
// launcher-level id of the row-synchronous SoA dense kernel (adapt_rows_kernel): same output as MODE_DENSE, any trajectory order
enum : int { MODE_DENSE_ROWS = 5 };
// the two passes of a MODE_REC kernel (the reference's full ragged Points::All stream): count rows, then write them
enum : int { REC_COUNT = 1, REC_WRITE = 2 };
enum : uint8_t { ST_DONE = 0, ST_MINSTEP_BREAK = 1, ST_MAX_STEPS = 2,
                 ST_INVALID_MATRIX = 3 };   // stiff path: try_inverse() failed (the reference's OdeError::InvalidMatrix)

#ifndef DEQ_BLOCK
#define DEQ_BLOCK 128
#endif
constexpr int BLOCK = DEQ_BLOCK;
constexpr int MAX_NP = 8;
constexpr int MAX_N = 64;   // state dimension of user (NVRTC) systems: stage vectors live in registers up to 16, in local memory beyond

struct SharedParams { double v[MAX_NP]; };

struct AdaptParams {
    // inputs (device, SoA over trajectories)
    const double* y0;      // [N][ntraj]
    const double* f0;      // [N][ntraj]   f(t0, y0) from hinit_kernel (k0 of the first step)
    const double* h0;      // [ntraj]      first step from hinit_kernel (ignored if use_initstep)
    const double* params;  // [NP][ntraj]  when per-trajectory, else unused
    SharedParams shared;   // parameters shared by the whole ensemble
    const double* tspan;   // [nt]
    unsigned long long nt;
    unsigned long long ntraj;
    double t0, tend, tdir, minstep, maxstep, reltol, abstol, initstep;
    double tdir_tend;      // tdir * tend (one IEEE product, problem.rs:337), formed once on the host
    int use_initstep;
    int libm_folded;       // 0: pow(s, 0.5) as written (types.rs:115); 1: sqrt(s), what LLVM folds pow(x, 0.5) to
    int refill;            // 1: persistent lanes re-filled individually; 0: a warp takes 32 trajectories at a time
    unsigned long long max_steps;  // guard (extension; the reference loops forever); 0 = none
    // outputs
    double* yfinal;        // [N][ntraj]
    uint32_t* accepted; uint32_t* rejected; uint32_t* nfe; uint8_t* status; double* t_reached;
    double* dense;         // MODE_DENSE: [N][nt][ntraj]; MODE_DENSE_TM: [ntraj][nt][N]
    uint32_t* nvalid;      // MODE_DENSE: number of leading tspan rows written (reference quirk)
    int rec_mode;          // MODE_REC: REC_COUNT / REC_WRITE
    uint32_t* rec_counts;  // REC_COUNT: rows of (tout, yout) the reference would produce, per trajectory
    const unsigned long long* rec_offsets;  // REC_WRITE: first row of each trajectory in rec_rows
    double* rec_rows;      // REC_WRITE: rows [t, y_0 .. y_{N-1}] of all trajectories, ragged, trajectory-major (staged writes)
    unsigned long long* work_counter;  // zeroed before launch
    unsigned long long* totals;        // [3] += accepted, rejected, nfe over the ensemble
    // ode23s only (stiff.cuh)
    int true_inverse;      // TEXTBOOK set: invert W = I - h d J itself, not its packed LU factors (problem.rs:469,481)
    int points_specified;  // Points::Specified: MODE_REC rows are the tspan points only (problem.rs:536)
    // MODE_DENSE / MODE_DENSE_TM (lane-refill kernels): work item s integrates trajectory order[s] — inputs are read from, and
    // the per-trajectory results written to, index order[s].  SoA: its dense rows go to COLUMN s of `dense` (capi.cu permutes
    // the columns back afterwards); trajectory-major: they go straight to the stream of trajectory order[s].  nullptr: identity.
    // Used to run an ensemble in the order of its first step sizes, so that neighbouring lanes advance alike (see deq_solve).
    const uint32_t* order;
};

struct HinitParams {
    const double* y0; const double* params; SharedParams shared;
    unsigned long long ntraj;
    double t0, tend, reltol, abstol;
    int order;
    int libm_folded;       // 0: pow(10, p) as written (problem.rs:893); 1: exp10(p), what LLVM folds pow(10.0, x) to
    double* f0; double* h0;
};

struct FixedParams {
    const double* y0; const double* params; SharedParams shared;
    const double* tspan; unsigned long long nt; unsigned long long ntraj;
    double* yfinal;   // [N][ntraj]
    double* yall;     // [N][nt][ntraj] or nullptr
    // oderosenbrock only (stiff.cuh): per-trajectory bookkeeping, a step can stop a trajectory (InvalidMatrix)
    uint32_t* accepted; uint32_t* rejected; uint32_t* nfe; uint8_t* status; double* t_reached; uint32_t* nvalid;
    unsigned long long* totals;
};

}  // namespace deqk
