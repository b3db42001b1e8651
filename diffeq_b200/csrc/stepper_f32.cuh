// stepper_f32.cuh — optional single-precision variant of the adaptive kernel (BASELINE.json north_star: "f64 is the
// reference precision, with an optional f32 mode").
//
// The reference cannot instantiate an f32 state (its impls are commented out, /root/reference/src/ode/types.rs:262-278),
// so there is no reference result to be identical to: this mode is validated against the f64 path at f32-appropriate
// tolerances (tests/test_gpu_parity.py::test_f32_mode_tracks_f64).  State, stages, error norm and the step factor are
// float; time and the step size stay double like the reference's `f64` time axis (problem.rs:46) so that t += dt does
// not lose the horizon; HBM arrays stay double (converted on load/store), so the C ABI is unchanged.  FMA arithmetic,
// zero coefficients skipped, one powf per step — the FAST formulas of stepper.cuh.  Same persistent lane-refill
// scheduling.  FP32 instructions do not hold the issue port for two cycles, which is where the speed-up comes from.
#pragma once
#include "stepper.cuh"

namespace deqk {

template <class TB, class F>
__device__ __forceinline__ void stages_f32(double t, const float (&y)[F::N], double dt, float (&k)[TB::S][F::N], const float* prm) {
    constexpr auto tbc = TB::data();
    const float dtf = (float)dt;
#pragma unroll
    for (int row = 1; row < TB::S; ++row) {
        float yi[F::N];
#pragma unroll
        for (int d = 0; d < F::N; ++d) yi[d] = y[d];
#pragma unroll
        for (int col = 0; col < row; ++col) {
            if (tbc.a[row][col] != 0.0) {
                const float w = (float)tbc.a[row][col] * dtf;
#pragma unroll
                for (int d = 0; d < F::N; ++d) yi[d] = __fmaf_rn(k[col][d], w, yi[d]);
            }
        }
        F::eval_r((float)(t + tbc.c[row] * dt), yi, k[row], prm);
    }
}

template <class TB, class F, bool PT, int MODE>
__global__ void __launch_bounds__(BLOCK, DEQ_MIN_BLOCKS) adapt_kernel_f32(const AdaptParams P) {
    constexpr int N = F::N, S = TB::S, NPR = F::NP ? F::NP : 1;
    constexpr auto tbc = TB::data();
    constexpr bool FSAL = deqt::fsal_of(TB::data());
    constexpr float PWH = (float)(-0.5 / (double)(TB::ORDER_MIN + 1));
    const unsigned lane = threadIdx.x & 31u;
    const double tdir = P.tdir, tend = P.tend;
    const float reltol = (float)P.reltol, abstol = (float)P.abstol;
    const uint32_t max_steps = P.max_steps == 0ull || P.max_steps > 0xffffffffull ? 0xffffffffu : (uint32_t)P.max_steps;

    bool active = false, pool_empty = false, last_step = false;
    unsigned long long traj = 0, it = 1;
    double t = 0.0, dt = 0.0;
    float cy[N], ck[N], prm[NPR];
    unsigned timeout = 0;
    uint32_t acc = 0, rej = 0, attempts = 0;
    unsigned long long tot_acc = 0, tot_rej = 0, tot_nfe = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) { cy[d] = 0.f; ck[d] = 0.f; }
#pragma unroll
    for (int j = 0; j < NPR; ++j) prm[j] = 0.f;

    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (idle != 0u && !pool_empty && (P.refill || idle == 0xffffffffu)) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(P.work_counter, (unsigned long long)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + (unsigned long long)__popc(idle) >= P.ntraj) pool_empty = true;
            if (!active) {
                const unsigned long long my = base + (unsigned long long)__popc(idle & ((1u << lane) - 1u));
                if (my < P.ntraj) {
                    traj = my;
                    active = true;
#pragma unroll
                    for (int j = 0; j < F::NP; ++j) prm[j] = (float)(PT ? P.params[(unsigned long long)j * P.ntraj + traj] : P.shared.v[j]);
#pragma unroll
                    for (int d = 0; d < N; ++d) {
                        cy[d] = (float)P.y0[(unsigned long long)d * P.ntraj + traj];
                        ck[d] = (float)P.f0[(unsigned long long)d * P.ntraj + traj];
                    }
                    t = P.t0;
                    dt = P.use_initstep ? P.initstep : P.h0[traj];
                    timeout = 0;
                    last_step = fabs((t + dt) - tend) <= DBL_EPSILON;
                    acc = 0; rej = 0; attempts = 0; it = 1;
                    if (MODE == MODE_DENSE) {
#pragma unroll
                        for (int d = 0; d < N; ++d) P.dense[((unsigned long long)d * P.nt) * P.ntraj + traj] = (double)cy[d];
                    }
                }
            }
        }
        if (__ballot_sync(0xffffffffu, active) == 0u) break;
#pragma unroll 1
        for (int burst = 0; burst < BURST_ATTEMPTS && active; ++burst) {
            uint8_t fin = 0xff;
            if (attempts >= max_steps) {
                fin = ST_MAX_STEPS;
            } else {
                ++attempts;
                float k[S][N];
#pragma unroll
                for (int d = 0; d < N; ++d) k[0][d] = ck[d];
                stages_f32<TB, F>(t, cy, dt, k, prm);
                const float dtf = (float)dt;
                float ytrial[N];
                float ssum = 0.f;
#pragma unroll
                for (int d = 0; d < N; ++d) {
                    float p = 0.f, e = 0.f;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        if (tbc.b[s][0] != 0.0) p = __fmaf_rn(k[s][d], (float)tbc.b[s][0], p);
                        if (tbc.b[s][0] - tbc.b[s][1] != 0.0) e = __fmaf_rn(k[s][d], (float)(tbc.b[s][0] - tbc.b[s][1]), e);
                    }
                    ytrial[d] = __fmaf_rn(p, dtf, cy[d]);
                    const float sc = __fmaf_rn(fmaxf(fabsf(cy[d]), fabsf(ytrial[d])), reltol, abstol);
                    const float q = __fdividef(e * dtf, sc);
                    ssum = __fmaf_rn(q, q, ssum);
                }
                double ndt;
                bool accept;
                if (ssum != ssum) {
                    accept = false; ndt = 0.2 * dt; timeout = 5;
                } else {
                    const float g = __powf(ssum, PWH) * 0.8f;    // (1/err)^(1/(p+1)) = ssum^(-1/(2(p+1)))
                    double nd = sel_min(P.maxstep, (sel_max(0.2, (double)g) * tdir) * dt);
                    if (timeout > 0) { nd = sel_min(nd, dt); timeout -= 1; }
                    ndt = tdir * nd;
                    accept = ssum < 1.f;
                }
                if (accept) {
                    acc += 1;
                    float f1[N];
                    if (FSAL) {
#pragma unroll
                        for (int d = 0; d < N; ++d) f1[d] = k[S - 1][d];
                    } else {
                        F::eval_r((float)(t + dt), ytrial, f1, prm);
                    }
                    if (MODE == MODE_DENSE) {
                        const double tn = t + dt;
                        while (it < P.nt) {
                            const double tq = __ldg(P.tspan + it);
                            if (!(tdir * t < tdir * tq && tdir * tq < tdir * tn)) break;
                            const float th = (float)((tq - t) / dt);
#pragma unroll
                            for (int d = 0; d < N; ++d) {
                                const float v = (cy[d] * (1.f - th) + ytrial[d] * th) +
                                                ((ytrial[d] - cy[d]) * (1.f - 2.f * th) + k[0][d] * (th - 1.f) * dtf + f1[d] * th * dtf) * th * (th - 1.f);
                                P.dense[((unsigned long long)d * P.nt + it) * P.ntraj + traj] = (double)v;
                            }
                            ++it;
                        }
                    }
#pragma unroll
                    for (int d = 0; d < N; ++d) { ck[d] = f1[d]; cy[d] = ytrial[d]; }
                    t = t + dt;
                    if (last_step) {
                        fin = ST_DONE;
                    } else {
                        dt = ndt;
                        if (tdir * ((t + dt) + dt * 0.01) >= tdir * tend) { dt = tend - t; last_step = true; }
                    }
                } else if (fabs(ndt) < P.minstep) {
                    fin = ST_MINSTEP_BREAK;
                } else {
                    rej += 1; last_step = false; dt = ndt; timeout = 5;
                }
            }
            if (fin != 0xff) {
                const uint32_t nfe = 2u + attempts * (uint32_t)(S - 1) + (FSAL ? 0u : acc);
#pragma unroll
                for (int d = 0; d < N; ++d) P.yfinal[(unsigned long long)d * P.ntraj + traj] = (double)cy[d];
                P.accepted[traj] = acc; P.rejected[traj] = rej; P.nfe[traj] = nfe; P.status[traj] = fin; P.t_reached[traj] = t;
                if (MODE == MODE_DENSE) {
#pragma unroll
                    for (int d = 0; d < N; ++d) P.dense[((unsigned long long)d * P.nt + (P.nt - 1)) * P.ntraj + traj] = (double)cy[d];
                    P.nvalid[traj] = (uint32_t)it;
                }
                tot_acc += acc; tot_rej += rej; tot_nfe += nfe;
                active = false;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tot_acc += __shfl_down_sync(0xffffffffu, tot_acc, o);
        tot_rej += __shfl_down_sync(0xffffffffu, tot_rej, o);
        tot_nfe += __shfl_down_sync(0xffffffffu, tot_nfe, o);
    }
    if (lane == 0) { atomicAdd(P.totals + 0, tot_acc); atomicAdd(P.totals + 1, tot_rej); atomicAdd(P.totals + 2, tot_nfe); }
}

}  // namespace deqk
