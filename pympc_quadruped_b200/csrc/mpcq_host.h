// Host-side helpers shared by the C-ABI (mpcq_api.cu) and the test-only warp emulator:
// public mpcq_config -> kernel constants, size classes.
#pragma once

#include <math.h>
#include <string>

#include "../../include/mpcq.h"
#include "mpcq_core.cuh"

namespace mpcq {

// size classes by number of stance foot-steps (slot capacity = 3 * stance, rounded to 32 rows/lane)
struct SizeClass { int ncap, ns_lo, ns_hi, nw; };      // nw = warps of the team that owns one environment
// team sizes are compile-time constants of the kernels (tunable with -DMPCQ_NWx=... for experiments).
// Measured on B200 (profiles/r01_team_size_sweep.txt): one warp is best for n <= 60 (the CTA barriers and the lower
// FMA density cost more than the extra warps hide), 2-3 warps give 1.3-1.5x for n = 120 / 180, larger teams lose.
#ifndef MPCQ_NW0
#define MPCQ_NW0 1
#endif
#ifndef MPCQ_NW1
#define MPCQ_NW1 2
#endif
#ifndef MPCQ_NW2
#define MPCQ_NW2 3
#endif
#ifndef MPCQ_NW3
#define MPCQ_NW3 12
#endif
static_assert(MPCQ_NW0 <= 16 && MPCQ_NW1 <= 16 && MPCQ_NW2 <= 16 && MPCQ_NW3 <= 16, "team scratch (Ctx::red / redi) holds one slot per warp, 16 at most");
static const SizeClass kClasses[4] = {{64, 0, class_ns_hi(64), MPCQ_NW0}, {128, class_ns_hi(64) + 1, class_ns_hi(128), MPCQ_NW1},
                                      {192, class_ns_hi(128) + 1, class_ns_hi(192), MPCQ_NW2}, {384, class_ns_hi(192) + 1, class_ns_hi(384), MPCQ_NW3}};
inline int class_nmax(const SizeClass& c) { return (3 * c.ns_hi + 3) & ~3; }   // rows of the largest system in the class

inline int num_classes(int horizon) {
    int n = 1;
    while (n < 4 && 4 * horizon >= kClasses[n].ns_lo) ++n;
    return n;
}

inline bool consts_from_config(const mpcq_config& c, Consts& k, std::string& err) {
    if (c.horizon < 1 || c.horizon > 32) { err = "horizon must be in 1..32"; return false; }
    if (c.dtype != MPCQ_F32 && c.dtype != MPCQ_F64) { err = "dtype must be MPCQ_F32 or MPCQ_F64"; return false; }
    if (!(c.dt > 0) || !(c.mu > 0) || !(c.fz_max >= 0) || !(c.mass > 0)) { err = "dt, mu, mass must be > 0 and fz_max >= 0"; return false; }
    for (int i = 0; i < 12; ++i)
        if (!(c.r_diag[i] > 0)) { err = "r_diag must be > 0 (strict convexity)"; return false; }
    for (int i = 0; i < 13; ++i)
        if (!(c.q_diag[i] >= 0)) { err = "q_diag must be >= 0"; return false; }
    const bool f64 = c.dtype == MPCQ_F64;
    k.horizon = c.horizon;
    k.pdas_cap = c.max_pdas_rounds > 0 ? c.max_pdas_rounds : 14;
    k.as_cap = c.max_as_iter > 0 ? c.max_as_iter : 12 * c.horizon + 30;
    k.refine_max = c.max_refine > 0 ? c.max_refine : 20;
    // the reference builds the friction pyramid in float32 (mpc.py:239-245): its solver sees float32(mu); fz_max * gait likewise
    k.dt = c.dt; k.mu = (double)(float)c.mu; k.fz_max = c.fz_max;
    k.inv_mass = (double)(float)(1.0 / c.mass);            // Bc[9:12] = I/m is stored in float32 (mpc.py:190)
    for (int i = 0; i < 9; ++i) k.inertia[i] = c.inertia[i];
    for (int i = 0; i < 13; ++i) k.q[i] = c.q_diag[i];
    for (int i = 0; i < 12; ++i) k.r[i] = c.r_diag[i];
    k.tol_p = c.tol_primal > 0 ? c.tol_primal : (f64 ? 1e-9 : 1e-7);
    k.tol_d = c.tol_dual > 0 ? c.tol_dual : (f64 ? 1e-9 : 1e-7);
    k.tol_pc = f64 ? k.tol_p : (k.tol_p > 1e-5 ? k.tol_p : 1e-5);
    k.tol_dc = f64 ? k.tol_d : (k.tol_d > 1e-5 ? k.tol_d : 1e-5);
    k.tol_r_tight = c.tol_residual > 0 ? c.tol_residual : (f64 ? 1e-12 : 1e-9);
    k.tol_r_loose = c.tol_residual_loose > 0 ? c.tol_residual_loose : (f64 ? 1e-9 : 1e-6);
    if (k.tol_r_loose < k.tol_r_tight) k.tol_r_loose = k.tol_r_tight;
    // u0 = -H^-1 g: the rounds build on it (u = u0 - P A' lam), so its residual is the starting residual of the fp64 finish;
    // stopping after ONE application of the inverse (1e-4) saves a step here and costs it again there (measured: 3.96 vs 3.97
    // operator applications per robot)
    k.tol_r_first = k.tol_r_loose;
    k.tol_active = c.tol_active > 0 ? c.tol_active : 1e-6;
    double rmin = c.r_diag[0];
    for (int i = 1; i < 12; ++i) rmin = c.r_diag[i] < rmin ? c.r_diag[i] : rmin;
    k.tol_r_abs = 2.0 * rmin * (f64 ? 2e-6 : 2e-4);            // forces to 2e-4 N (f32 mode) / 2e-6 N (f64 mode) in the weakest direction
    double det = c.inertia[0] * (c.inertia[4] * c.inertia[8] - c.inertia[5] * c.inertia[7]) -
                 c.inertia[1] * (c.inertia[3] * c.inertia[8] - c.inertia[5] * c.inertia[6]) +
                 c.inertia[2] * (c.inertia[3] * c.inertia[7] - c.inertia[4] * c.inertia[6]);
    if (!(det > 0)) { err = "inertia must be positive definite"; return false; }
    return true;
}

}  // namespace mpcq
