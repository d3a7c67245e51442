// Device code of the per-leg layer, one call per (environment, leg).  Included by mpcq_api.cu (the
// sm_100a kernels) and, with MPCQ_HOST_EMU, by the test-only host emulation (tests/emu) so that the same arithmetic is
// checked against the reference fixtures in the CPU-only container.
#pragma once

#include <math.h>
#include <stdint.h>

#include "mpcq_warp.cuh"

namespace mpcq {
// Per-leg layer that consumes the MPC forces: swing-foot targets (reference
// linear_mpc/swing_foot_trajectory_generator.py:38-129) and the joint torque map (linear_mpc/leg_controller.py:38-91).
// One thread per (environment, leg); elementwise, fp64 like the reference (float32 where the reference rounds: base position
// and velocity, trajectory break points, torque output).  HBM stream bound: ~350 B in / 100 B out per leg.
struct SwingArgs {
    const double *pos_base, *vel_base, *R_base, *thighs, *pos_feet, *swing_state, *v_des, *yaw_rate, *swing_time, *stance_time;
    uint8_t* active;
    double *remaining, *foot_init, *foot_final, *pos_t, *vel_t;
    double swing_height, dt_control, gravity, foot_z;
    int B;
};

MPCQ_HD void swing_leg(const SwingArgs& a, int idx) {
    const int b = idx >> 2;
    const size_t o = (size_t)idx * 3;
    const double ss = a.swing_state[idx];
    if (!(ss > 0.0)) {                                             // scripts/isaacgym_a1.py:143-147: stance legs keep zero targets
MPCQ_UNROLL
        for (int i = 0; i < 3; ++i) { a.pos_t[o + i] = 0.0; a.vel_t[o + i] = 0.0; }
        return;
    }
    double pb[3], vb[3], R[9], vd[3], th[3];
MPCQ_UNROLL
    for (int i = 0; i < 3; ++i) {
        pb[i] = (double)(float)a.pos_base[3 * b + i];              // np.array(..., dtype=np.float32) (:83-84)
        vb[i] = (double)(float)a.vel_base[3 * b + i];
        vd[i] = a.v_des[3 * b + i];
        th[i] = a.thighs[o + i];
    }
MPCQ_UNROLL
    for (int i = 0; i < 9; ++i) R[i] = a.R_base[9 * b + i];
    const double Tsw = a.swing_time[b], Tst = a.stance_time[b], yr = a.yaw_rate[b];
    const bool started = a.active[idx] != 0;                       // !is_first_swing
    const double rem = started ? a.remaining[idx] - a.dt_control : Tsw;          // :96-100
    double sn, cs;
    sincos(yr * 0.5 * Tst, &sn, &cs);                              // :103
    const double in[3] = {cs * th[0] - sn * th[1] + vd[0] * rem, sn * th[0] + cs * th[1] + vd[1] * rem, th[2] + vd[2] * rem};
    double fin[3], ini[3];
MPCQ_UNROLL
    for (int i = 0; i < 3; ++i) {
        const double vdw = R[3 * i] * vd[0] + R[3 * i + 1] * vd[1] + R[3 * i + 2] * vd[2];       // R_base @ base_vel_base_des (:94)
        fin[i] = pb[i] + (R[3 * i] * in[0] + R[3 * i + 1] * in[1] + R[3 * i + 2] * in[2]) + 0.5 * Tst * vb[i] + 0.03 * (vb[i] - vdw);
    }
    const double kz = 0.5 * pb[2] / a.gravity;                     // :113-117
    fin[0] += kz * (vb[1] * yr);
    fin[1] += kz * (-vb[0] * yr);
    fin[2] = a.foot_z;
MPCQ_UNROLL
    for (int i = 0; i < 3; ++i) ini[i] = started ? a.foot_init[o + i] : a.pos_feet[o + i];        // :121-123
    a.active[idx] = ss >= 1.0 ? 0 : 1;                             // :125-126 swing finished -> next swing starts afresh
    a.remaining[idx] = rem;
MPCQ_UNROLL
    for (int i = 0; i < 3; ++i) {
        if (!started) a.foot_init[o + i] = ini[i];                 // unchanged during a swing: not written back (96 B per robot)
        a.foot_final[o + i] = fin[i];
    }
    // three-point zero-velocity cubic Hermite (:38-63; Drake PiecewisePolynomial.CubicHermite, float32 break points)
    const double t1 = (double)(float)(Tsw / 2.0), t2 = (double)(float)Tsw;
    double t = Tsw - rem;
    t = t < 0.0 ? 0.0 : (t > t2 ? t2 : t);
    const bool second = !(t < t1);
    const double h = second ? t2 - t1 : t1, s = second ? t - t1 : t;
    // one division per leg: with x = s / h,  p = ya + dy x^2 (3 - 2 x),  p' = 6 dy x (1 - x) / h  (the same cubic as Drake's
    // c2 = 3 dy / h^2, c3 = -2 dy / h^3; fp64 divisions were a third of this kernel's FP64 instructions)
    const double ih = 1.0 / h, x = s * ih, wp = x * x * (3.0 - 2.0 * x), wv = 6.0 * x * (1.0 - x) * ih;
    double p[3], v[3];
MPCQ_UNROLL
    for (int i = 0; i < 3; ++i) {
        const double mid = i == 2 ? a.swing_height : (ini[i] + fin[i]) * 0.5;
        const double ya = second ? mid : ini[i], yb = second ? fin[i] : mid;
        const double dy = yb - ya;
        p[i] = ya + dy * wp - pb[i];
        v[i] = dy * wv - vb[i];
    }
MPCQ_UNROLL
    for (int i = 0; i < 3; ++i) {                                  // base_R_world @ (. - pos_base) (:76-78)
        a.pos_t[o + i] = R[i] * p[0] + R[3 + i] * p[1] + R[6 + i] * p[2];
        a.vel_t[o + i] = R[i] * v[0] + R[3 + i] * v[1] + R[6 + i] * v[2];
    }
}

struct TorqueArgs {
    const double *Jv, *R_base, *bpf, *bvf, *swing_state, *pos_t, *vel_t;
    const void* forces;
    float* tau;
    double kp[9], kd[9];
    int B, ncol;
};

template <typename T>
MPCQ_HD void torque_leg(const TorqueArgs& a, int idx) {
    const int b = idx >> 2, leg = idx & 3;
    const size_t o = (size_t)idx * 3;
    double e[3];
    if (a.swing_state[idx] != 0.0) {                               // `if swing_states[leg_idx]:` (leg_controller.py:76; NaN is truthy)
        double R[9], dp[3], dv[3];
MPCQ_UNROLL
        for (int i = 0; i < 9; ++i) R[i] = a.R_base[9 * b + i];
MPCQ_UNROLL
        for (int i = 0; i < 3; ++i) {
            const double *pt = a.pos_t + o, *vt = a.vel_t + o, *pf = a.bpf + o, *vf = a.bvf + o;
            dp[i] = (R[3 * i] * pt[0] + R[3 * i + 1] * pt[1] + R[3 * i + 2] * pt[2]) - (R[3 * i] * pf[0] + R[3 * i + 1] * pf[1] + R[3 * i + 2] * pf[2]);
            dv[i] = (R[3 * i] * vt[0] + R[3 * i + 1] * vt[1] + R[3 * i + 2] * vt[2]) - (R[3 * i] * vf[0] + R[3 * i + 1] * vf[1] + R[3 * i + 2] * vf[2]);
        }
MPCQ_UNROLL
        for (int i = 0; i < 3; ++i)
            e[i] = (a.kp[3 * i] * dp[0] + a.kp[3 * i + 1] * dp[1] + a.kp[3 * i + 2] * dp[2]) +
                   (a.kd[3 * i] * dv[0] + a.kd[3 * i + 1] * dv[1] + a.kd[3 * i + 2] * dv[2]);
    } else {
        const T* f = static_cast<const T*>(a.forces) + o;
MPCQ_UNROLL
        for (int i = 0; i < 3; ++i) e[i] = -(double)f[i];          // tau = Jv.T @ -f (:87)
    }
    // the leg's own joint columns of the 3 x ncol foot Jacobian (:84,88): 6 + 3 leg in the reference layout, 0 for bare blocks
    const double* J = a.Jv + (size_t)idx * 3 * a.ncol + (a.ncol == 18 ? 6 + 3 * leg : 0);
MPCQ_UNROLL
    for (int c = 0; c < 3; ++c)
        a.tau[o + c] = (float)(J[c] * e[0] + J[a.ncol + c] * e[1] + J[2 * a.ncol + c] * e[2]);
}
}  // namespace mpcq
