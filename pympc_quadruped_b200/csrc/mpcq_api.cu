// libmpcq.so: sm_100a kernels + the C ABI declared in include/mpcq.h.
//
// One warp (a 32-thread CTA) owns one environment; environments are split into size classes by
// their number of stance foot-steps so that each class gets exactly the shared memory its dense
// factor needs (occupancy follows the problem size, the hardware block scheduler balances the
// very uneven per-env work).  A class kernel is launched over the whole batch; CTAs whose env
// belongs to another class exit after reading its contact table.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <new>
#include <string>
#include <unordered_map>

#include "mpcq_host.h"
#include "mpcq_legs.cuh"

namespace {

using mpcq::Consts;
using mpcq::IO;

constexpr size_t kMaxSmem = 227 * 1024;
constexpr int kHostStreams = 4;             // chunks / streams of mpcq_solve_host
constexpr int kGlobalCtas = 148 * 2;       // resident CTAs of a class whose factor lives in global memory

// warps of the team that owns one environment, per slot capacity (must match mpcq::kClasses)
template <int NCAP> struct TeamWarps { static constexpr int value = NCAP <= 64 ? MPCQ_NW0 : (NCAP <= 128 ? MPCQ_NW1 : (NCAP <= 192 ? MPCQ_NW2 : MPCQ_NW3)); };

// One-warp teams (the small class) are launched several to a CTA, each warp with its own robot and its own slice of the
// CTA's shared memory: the warps of a CTA start together and run the same instruction stream through the uniform part of the
// path (model, factorisation, inversion - no data-dependent control flow there), so the SM's instruction caches serve one
// copy of it instead of one per resident warp (measured: the kernel stalls on instruction fetch, profiles/r02_*).
#ifndef MPCQ_EPC0
#define MPCQ_EPC0 9
#endif
#ifndef MPCQ_EPC1
#define MPCQ_EPC1 1
#endif
template <int NCAP> struct EnvsPerCta { static constexpr int value = NCAP <= 64 ? MPCQ_EPC0 : (NCAP <= 128 ? MPCQ_EPC1 : 1); };

template <class T, int NCAP, bool LGLOBAL>
__global__ void __launch_bounds__(32 * TeamWarps<NCAP>::value * EnvsPerCta<NCAP>::value, NCAP <= 64 ? (sizeof(T) == 4 && EnvsPerCta<NCAP>::value <= 4 ? 8 / EnvsPerCta<NCAP>::value : 1) : 1)
mpcq_solve_kernel(const __grid_constant__ Consts cs, const __grid_constant__ IO<T> io, T* gws, size_t gws_stride,
                  int ns_lo, int ns_hi, unsigned env_bytes) {
    extern __shared__ __align__(32) char smem[];
    constexpr int EPC = EnvsPerCta<NCAP>::value;
    if constexpr (EPC > 1) {
        constexpr int NT = 32 * TeamWarps<NCAP>::value;         // threads of one team (1 warp by default)
        const int wq = threadIdx.x / NT;
        const int i = blockIdx.x * (blockDim.x / NT) + wq;      // fewer teams than EPC when the workspace is large (long horizons)
        if (i < io.B) {
            const int b = io.perm ? io.perm[i] : i;
            mpcq::solve_env<T, NCAP, TeamWarps<NCAP>::value>(cs, io, b, smem + (size_t)wq * env_bytes, nullptr, ns_lo, ns_hi, 1 + wq);
        }
    } else {
        T* lg = LGLOBAL ? gws + (size_t)blockIdx.x * gws_stride : nullptr;
        for (int i = blockIdx.x; i < io.B; i += gridDim.x) {
            const int b = io.perm ? io.perm[i] : i;
            mpcq::solve_env<T, NCAP, TeamWarps<NCAP>::value>(cs, io, b, smem, lg, ns_lo, ns_hi);
            __syncthreads();
        }
    }
}

// Expected-work-first schedule.  The number of active-set rounds an environment needs grows with how far it is from
// its reference: the Q-weighted initial tracking error sqrt(sum_c q_c (x0_c - xref_0,c)^2) has a rank correlation of
// 0.7-0.8 with the measured factorisation count.  Environments are bucketed by that score (64 buckets) and launched
// hardest-first, so the long ones start at t = 0 instead of trailing behind the batch (list scheduling, LPT rule);
// measured -21 % batch time on the 50/50 mix, within 2 % of the oracle order.  One CTA: histogram, prefix, scatter.
template <class T>
__global__ void __launch_bounds__(1024)
mpcq_schedule_kernel(const __grid_constant__ Consts cs, const __grid_constant__ IO<T> io, int32_t* perm, uint8_t* bucket) {
    __shared__ int hist[64];
    __shared__ int start[64];
    const int tid = threadIdx.x, H = cs.horizon;
    if (tid < 64) hist[tid] = 0;
    __syncthreads();
    for (int b = tid; b < io.B; b += blockDim.x) {
        const T* x0 = io.x0 + (size_t)b * 13;
        const T* xr = io.x_ref + (size_t)b * 13 * H;
        float s = 0.f;
        for (int c = 0; c < 12; ++c) {
            if (c == 2) continue;                              // yaw: reference starts at the current yaw (and wraps)
            const float e = (float)x0[c] - (float)xr[c];
            s += (float)cs.q[c] * e * e;
        }
        s = sqrtf(s);
        int k = (int)(64.f * s / (1.f + s));
        k = !(s == s) ? 63 : (k > 63 ? 63 : (k < 0 ? 0 : k));  // non-finite states first: they exit at once
        bucket[b] = (uint8_t)k;
        atomicAdd(&hist[k], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int k = 63; k >= 0; --k) { start[k] = acc; acc += hist[k]; }     // highest score first
    }
    __syncthreads();
    for (int b = tid; b < io.B; b += blockDim.x) perm[atomicAdd(&start[bucket[b]], 1)] = b;
}

// The same schedule as two small multi-CTA launches (used by the device entry point): with one CTA the 4 096-robot pre-pass
// is a chain of four dependent DRAM round trips per thread plus two passes over the bucket array (16 us on the critical
// path of a 0.68 ms step); here every robot has its own thread in the scoring pass, each CTA keeps its own histogram, and the
// scatter pass derives each CTA's offsets from the per-CTA histograms - no atomics on global memory, no grid barrier.
constexpr int kSchedThreads = 256, kSchedMaxCtas = 64;

template <class T>
MPCQ_DEV int schedule_bucket(const Consts& cs, const IO<T>& io, int b) {
    const T* x0 = io.x0 + (size_t)b * 13;
    const T* xr = io.x_ref + (size_t)b * 13 * cs.horizon;
    float s = 0.f;
    for (int c = 0; c < 12; ++c) {
        if (c == 2) continue;
        const float e = (float)x0[c] - (float)xr[c];
        s += (float)cs.q[c] * e * e;
    }
    s = sqrtf(s);
    const int k = (int)(64.f * s / (1.f + s));
    return !(s == s) ? 63 : (k > 63 ? 63 : (k < 0 ? 0 : k));
}

template <class T>
__global__ void __launch_bounds__(kSchedThreads)
mpcq_score_kernel(const __grid_constant__ Consts cs, const __grid_constant__ IO<T> io, uint8_t* bucket, int* cta_hist) {
    __shared__ int hist[64];
    const int tid = threadIdx.x;
    if (tid < 64) hist[tid] = 0;
    __syncthreads();
    for (int b = blockIdx.x * kSchedThreads + tid; b < io.B; b += gridDim.x * kSchedThreads) {
        const int k = schedule_bucket<T>(cs, io, b);
        bucket[b] = (uint8_t)k;
        atomicAdd(&hist[k], 1);
    }
    __syncthreads();
    if (tid < 64) cta_hist[blockIdx.x * 64 + tid] = hist[tid];
}

__global__ void __launch_bounds__(kSchedThreads)
mpcq_scatter_kernel(int B, const uint8_t* bucket, const int* cta_hist, int32_t* perm) {
    __shared__ int total[64];
    __shared__ int start[64];
    const int tid = threadIdx.x;
    int before = 0;
    if (tid < 64) {
        int tot = 0;
        for (int c = 0; c < (int)gridDim.x; ++c) {
            const int v = cta_hist[c * 64 + tid];
            before += c < (int)blockIdx.x ? v : 0;
            tot += v;
        }
        total[tid] = tot;
    }
    __syncthreads();
    if (tid < 64) {
        int higher = 0;
        for (int k = tid + 1; k < 64; ++k) higher += total[k];            // highest score first
        start[tid] = higher + before;
    }
    __syncthreads();
    for (int b = blockIdx.x * kSchedThreads + tid; b < B; b += gridDim.x * kSchedThreads)
        perm[atomicAdd(&start[bucket[b]], 1)] = b;
}

// stage kernel for parity tests: dense (H, g, ub) exactly as the reference hands them to its solver
template <class T>
__global__ void __launch_bounds__(32)
mpcq_build_qp_kernel(const __grid_constant__ Consts cs, const __grid_constant__ IO<T> io, double* Hout, double* gout,
                     double* ubout) {
    extern __shared__ __align__(32) char smem[];
    const int b = blockIdx.x, lane = threadIdx.x, H = cs.horizon, n = 12 * H;
    mpcq::Work<T> w;
    mpcq::carve<T>(w, smem, reinterpret_cast<T*>(smem), H, 384, true);   // no factor needed: vectors only
    // every foot-step counts as stance here so that g comes out in the full [H][12] layout
    for (int k = lane; k < 4 * H; k += 32) { w.fk[k] = (uint8_t)k; w.cidx[k] = (uint8_t)k; }
    w.ns = 4 * H;
    w.nv = 12 * H;
    w.n = 12 * H;
    w.t.tid = lane; w.t.nt = 32; w.t.wid = 0;
    __syncwarp();
    const double yaw = io.yaw ? (double)io.yaw[b] : (double)io.x0[(size_t)b * 13 + 2];
    mpcq::setup_model(cs, w, io.x0 + (size_t)b * 13, yaw, io.r_feet + (size_t)b * 12, io.x_ref + (size_t)b * 13 * H);
    double* Hb = Hout + (size_t)b * n * n;
    for (int row = 0; row < n; ++row) {
        const int i = row / 12, r = row - 12 * i;
        for (int col = lane; col < n; col += 32) {
            const int j = col / 12, c = col - 12 * j;
            const int m = i > j ? i : j;
            double v = 2.0 * (double)(H - m) * w.Md[12 * r + c] + (double)w.NS2[2 * (i * H + j) + 1] * w.Md[144 + 12 * r + c];
            if (row == col) v += 2.0 * cs.r[r];
            Hb[(size_t)row * n + col] = v;
        }
    }
    for (int idx = lane; idx < n; idx += 32) gout[(size_t)b * n + idx] = w.g[idx];
    const float* gait = io.gait + (size_t)b * 4 * H;
    for (int k = lane; k < 4 * H; k += 32) {
        double* ub = ubout + (size_t)b * 20 * H + 5 * k;
        ub[0] = ub[1] = ub[2] = ub[3] = (double)INFINITY;
        ub[4] = (double)(float)((double)gait[k] * cs.fz_max);
    }
}

// Per-leg layer: bodies in mpcq_legs.cuh, one thread per (environment, leg).
using mpcq::SwingArgs;
using mpcq::TorqueArgs;

// 8 blocks per SM (64 registers): 0.164 ms per 2^20 robots; 7 blocks (72 registers, unconstrained) 0.176 ms, 12 (40, spills) 0.226 ms
__global__ void __launch_bounds__(128, 8)
mpcq_swing_kernel(SwingArgs a) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 4 * a.B) mpcq::swing_leg(a, idx);
}

template <typename T>
__global__ void __launch_bounds__(128)
mpcq_torque_kernel(TorqueArgs a) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 4 * a.B) mpcq::torque_leg<T>(a, idx);
}

// Contact schedule on the device (reference linear_mpc/gait.py:76-135): one thread per environment writes
// its 4H-entry contact table and, optionally, the swing / stance phase states the leg controller reads.  Integer modular
// arithmetic for the table; the phase states follow the reference's float32 phase / float64 arithmetic mix.
struct GaitArgs {
    const int32_t *offsets, *durations, *num_segment, *cur_iteration;
    int iterations_between_mpc, B, H;
};

__global__ void __launch_bounds__(128)
mpcq_gait_kernel(GaitArgs a, float* table, double* swing_state, double* stance_state) {
    __shared__ __align__(16) float stage[128 * 64];
    const int b0 = blockIdx.x * blockDim.x;
    const bool valid = b0 + (int)threadIdx.x < a.B;
    const int b = valid ? b0 + (int)threadIdx.x : a.B - 1;         // idle threads shadow the last robot, their stores are masked
    const int nvalid = a.B - b0 < (int)blockDim.x ? a.B - b0 : (int)blockDim.x;
    const int seg = a.num_segment[b], ibm = a.iterations_between_mpc;
    const int cur = a.cur_iteration[b];
    int off[4], dur[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { off[j] = a.offsets[4 * b + j]; dur[j] = a.durations[4 * b + j]; }
    // set_iteration (gait.py:76-79): iteration = floor(cur / ibm) % seg, phase = (cur % (ibm seg)) / (ibm seg)
    const int iteration = (cur / ibm) % seg;
    int offm[4];                                                 // offsets reduced once (Python's non-negative modulo, gait.py:93-97)
#pragma unroll
    for (int j = 0; j < 4; ++j) { offm[j] = off[j] % seg; offm[j] += offm[j] < 0 ? seg : 0; }
    int ph = (iteration + 1) % seg;                             // (i + 1 + iteration) % seg, advanced step by step
    // the table rows are staged in shared memory (up to 16 horizon steps at a time) and written by the whole block in runs of
    // 4 x steps consecutive floats per robot: one thread writing its own row touches 32 sectors per store instruction
    for (int i0 = 0; i0 < a.H; i0 += 16) {
        const int hc = a.H - i0 < 16 ? a.H - i0 : 16;
        if (valid) {
            float4* st4 = reinterpret_cast<float4*>(stage) + threadIdx.x * hc;
            for (int i = 0; i < hc; ++i) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int c = ph - offm[j];                          // both in [0, seg): one conditional add is the modulo
                    c += c < 0 ? seg : 0;
                    v[j] = c < dur[j] ? 1.0f : 0.0f;
                }
                st4[i] = make_float4(v[0], v[1], v[2], v[3]);
                ph = ph + 1 == seg ? 0 : ph + 1;
            }
        }
        __syncthreads();
        const int run = 4 * hc;
        const unsigned inv = 0xFFFFFFFFu / (unsigned)run + 1u;   // idx / run = umulhi(idx, inv) for idx < 2^16
        for (int idx = threadIdx.x; idx < nvalid * run; idx += blockDim.x) {
            const int rb = (int)__umulhi((unsigned)idx, inv), c = idx - rb * run;
            table[(size_t)(b0 + rb) * 4 * a.H + 4 * i0 + c] = stage[idx];
        }
        __syncthreads();
    }
    if (!valid) return;
    if (!swing_state && !stance_state) return;
    const int period = ibm * seg;
    const double phase = (double)(float)((double)(cur % period) / (double)period);   // np.full(4, phase, dtype=np.float32)
    double offn[4], durn[4], swo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { offn[j] = (double)off[j] / seg; durn[j] = (double)dur[j] / seg; swo[j] = offn[j] + durn[j]; }
    // the reference subtracts 1 from the WHOLE vector each time one entry exceeds 1 (gait.py:104-106); reproduced as written
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (swo[i] > 1.0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) swo[j] -= 1.0;
        }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (swing_state) {
            const double sd = 1.0 - durn[j];
            double s = phase - swo[j];
            if (s < 0.0) s += 1.0;
            swing_state[4 * b + j] = s > sd ? 0.0 : s / sd;
        }
        if (stance_state) {
            double s = phase - offn[j];
            if (s < 0.0) s += 1.0;
            stance_state[4 * b + j] = s > durn[j] ? 0.0 : s / durn[j];
        }
    }
}

// state assembly + reference trajectory, one thread per environment (HBM-stream bound: ~34 doubles in, 13 + 13H reals out)
struct AssembleArgs {
    const double *quat, *pos, *omega, *vel, *R, *vdes, *yawrate;
    double *xy_des, *yaw_des, *rp_init;
    int first_run, do_mpc, B, H;
    double dt_control, dt, com_height, gravity;
};

template <class T>
__global__ void __launch_bounds__(128)
mpcq_assemble_kernel(const __grid_constant__ AssembleArgs a, T* x0, T* yaw_out, T* x_ref) {
    // x_ref rows are staged per horizon step in shared memory and written by the whole block in runs of 13 consecutive values
    // per robot (one thread per robot writing its own 13H values touches 32 cache lines per store instruction)
    __shared__ T stage[128 * 13 * (sizeof(T) == 4 ? 7 : 3)];
    const int b0 = blockIdx.x * blockDim.x;
    const bool valid = b0 + (int)threadIdx.x < a.B;
    const int b = valid ? b0 + (int)threadIdx.x : a.B - 1;         // idle threads shadow the last robot, their stores are masked
    const int nvalid = a.B - b0 < (int)blockDim.x ? a.B - b0 : (int)blockDim.x;
    const double qw = a.quat[4 * b], qx = a.quat[4 * b + 1], qy = a.quat[4 * b + 2], qz = a.quat[4 * b + 3];
    // quat2ZYXangle (utils/kinematics.py:40-49), float64
    const double roll = atan2(2 * (qw * qx + qy * qz), 1 - 2 * (qx * qx + qy * qy));
    const double pitch = asin(2 * (qw * qy - qz * qx));
    const double yaw = atan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz));
    // current_state is a float32 array (mpc.py:57,70-76)
    float st[13];
    st[0] = (float)roll; st[1] = (float)pitch; st[2] = (float)yaw;
    for (int i = 0; i < 3; ++i) {
        st[3 + i] = (float)a.pos[3 * b + i];
        st[6 + i] = (float)a.omega[3 * b + i];
        st[9 + i] = (float)a.vel[3 * b + i];
    }
    st[12] = (float)(-a.gravity);
    for (int i = 0; i < 13; ++i) stage[threadIdx.x * 13 + i] = (T)st[i];
    __syncthreads();
    for (int idx = threadIdx.x; idx < nvalid * 13; idx += blockDim.x) x0[(size_t)b0 * 13 + idx] = stage[idx];   // contiguous
    __syncthreads();
    if (valid) yaw_out[b] = (T)yaw;
    // vel_base_des = R_base @ base_vel_base_des (mpc.py:83)
    double R[9];
    if (a.R) {
        for (int i = 0; i < 9; ++i) R[i] = a.R[9 * b + i];
    } else {                                                    // quat2matrix (utils/kinematics.py:51-71)
        R[0] = qw * qw + qx * qx - qy * qy - qz * qz; R[1] = 2 * (qx * qy - qw * qz); R[2] = 2 * (qw * qy + qx * qz);
        R[3] = 2 * (qw * qz + qx * qy); R[4] = qw * qw - qx * qx + qy * qy - qz * qz; R[5] = 2 * (qy * qz - qw * qx);
        R[6] = 2 * (qx * qz - qw * qy); R[7] = 2 * (qw * qx + qy * qz); R[8] = qw * qw - qx * qx - qy * qy + qz * qz;
    }
    const double vb0 = a.vdes[3 * b], vb1 = a.vdes[3 * b + 1], vb2 = a.vdes[3 * b + 2];
    const double vx = (R[0] * vb0 + R[1] * vb1) + R[2] * vb2, vy = (R[3] * vb0 + R[4] * vb1) + R[5] * vb2;
    const double rate = a.yawrate[b];
    double xd, yd, yawd;
    if (a.first_run == 2) { xd = (double)st[3]; yd = (double)st[4]; yawd = yaw; }   // respawn: desired pose = current pose
    else if (a.first_run) { xd = 0.0; yd = 0.0; yawd = yaw; }  // mpc.py:84-88
    else {                                                      // mpc.py:89-92
        xd = a.xy_des[2 * b] + a.dt_control * vx;
        yd = a.xy_des[2 * b + 1] + a.dt_control * vy;
        yawd = yaw + a.dt_control * rate;
    }
    if (a.do_mpc) {                                             // generate_reference_trajectory, mpc.py:110-170
        const double x3 = st[3], x4 = st[4];
        const double lim = 0.1;
        if (xd - x3 > lim) xd = x3 + lim;
        if (x3 - xd > lim) xd = x3 - lim;
        if (yd - x4 > lim) yd = x4 + lim;
        if (x4 - yd > lim) yd = x4 - lim;
        double roll_init = a.first_run == 2 ? 0.0 : a.rp_init[2 * b], pitch_init = a.first_run == 2 ? 0.0 : a.rp_init[2 * b + 1];
        if (fabs((double)st[9]) > 0.2) pitch_init += a.dt * (0.0 - (double)st[1]) / (double)st[9];
        if (fabs((double)st[10]) > 0.1) roll_init += a.dt * (0.0 - (double)st[0]) / (double)st[10];
        roll_init = fmin(fmax(roll_init, -0.25), 0.25);
        pitch_init = fmin(fmax(pitch_init, -0.25), 0.25);
        if (valid) { a.rp_init[2 * b] = roll_init; a.rp_init[2 * b + 1] = pitch_init; }
        const float rc = (float)((double)st[10] * roll_init), pc = (float)((double)st[9] * pitch_init);
        float ry = (float)yawd, rx = (float)xd, rY = (float)yd;
        // HC horizon steps per staging pass: the block then writes runs of 13 HC consecutive values per robot
        constexpr int HC = sizeof(T) == 4 ? 7 : 3;
        for (int i0 = 0; i0 < a.H; i0 += HC) {
            const int hc = a.H - i0 < HC ? a.H - i0 : HC;
            for (int ii = 0; ii < hc; ++ii) {
                if (i0 + ii > 0) {                                  // float32 storage, float64 increments (mpc.py:163-166)
                    ry = (float)((double)ry + a.dt * rate);
                    rx = (float)((double)rx + a.dt * vx);
                    rY = (float)((double)rY + a.dt * vy);
                }
                T* r = stage + (threadIdx.x * hc + ii) * 13;
                r[0] = (T)rc; r[1] = (T)pc; r[2] = (T)ry; r[3] = (T)rx; r[4] = (T)rY; r[5] = (T)(float)a.com_height;
                r[6] = (T)0; r[7] = (T)0; r[8] = (T)(float)rate; r[9] = (T)(float)vx; r[10] = (T)(float)vy; r[11] = (T)0;
                r[12] = (T)(float)(-a.gravity);
            }
            __syncthreads();
            const int run = 13 * hc;
            const unsigned inv = 0xFFFFFFFFu / (unsigned)run + 1u;  // idx / run = umulhi(idx, inv) for idx < 2^16
            for (int idx = threadIdx.x; idx < nvalid * run; idx += blockDim.x) {
                const int rb = (int)__umulhi((unsigned)idx, inv), c = idx - run * rb;
                x_ref[((size_t)(b0 + rb) * a.H + i0) * 13 + c] = stage[idx];
            }
            __syncthreads();
        }
    }
    if (valid) { a.xy_des[2 * b] = xd; a.xy_des[2 * b + 1] = yd; a.yaw_des[b] = yawd; }
}

// mpcq_tick_host: the packed per-robot rows a host loop hands over, split into the arrays the gait / assembly / solve kernels
// take (one thread per robot; 272 B in per robot)
struct TickArrays {
    double *quat, *pos, *omega, *vel, *vdes, *yawrate;          // [B,4] [B,3] [B,3] [B,3] [B,3] [B]
    int32_t *offs, *durs, *seg, *iter;                          // [B,4] [B,4] [B] [B]
};
template <typename T>
__global__ void __launch_bounds__(128)
mpcq_tick_unpack_kernel(int B, const double* __restrict__ state_cmd, const int32_t* __restrict__ gait_params, TickArrays a, T* feet) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double* s = state_cmd + (size_t)b * 29;
    for (int i = 0; i < 4; ++i) a.quat[4 * b + i] = s[i];
    for (int i = 0; i < 3; ++i) { a.pos[3 * b + i] = s[4 + i]; a.omega[3 * b + i] = s[7 + i]; a.vel[3 * b + i] = s[10 + i]; a.vdes[3 * b + i] = s[25 + i]; }
    for (int i = 0; i < 12; ++i) feet[(size_t)12 * b + i] = (T)s[13 + i];
    a.yawrate[b] = s[28];
    const int32_t* g = gait_params + (size_t)b * 10;
    for (int i = 0; i < 4; ++i) { a.offs[4 * b + i] = g[i]; a.durs[4 * b + i] = g[4 + i]; }
    a.seg[b] = g[8];
    a.iter[b] = g[9];
}

}  // namespace

struct mpcq_handle {
    mpcq_config cfg;
    Consts cs;
    int ncls = 0;
    size_t smem[4] = {0, 0, 0, 0};
    bool lglobal[4] = {false, false, false, false};
    void* gws = nullptr;                  // global-memory factors of the largest class
    size_t gws_stride = 0;                // elements per CTA
    size_t real_size = 4;
    int last_launches = 0;
    std::string err;
    // staging for mpcq_solve_host
    char* pin = nullptr;
    char* dev = nullptr;
    size_t stage_cap = 0;
    cudaStream_t streams[4] = {nullptr, nullptr, nullptr, nullptr};
    std::unordered_map<const void*, size_t> pinned_cache; // caller buffer -> bytes verified page-locked from that address (0 = pageable)
    // mpcq_tick_host: device arrays for `tick_cap` robots; the controller state (desired xy / yaw, roll / pitch compensation)
    // persists between calls
    char* tick_dev[2] = {nullptr, nullptr};
    double* tick_state[2] = {nullptr, nullptr};   // [cap,5]: xy_des 2 | yaw_des | rp_init 2, per pipeline slot
    size_t tick_cap[2] = {0, 0};
    bool tick_inflight[2] = {false, false};        // submitted, not yet waited for
    // measurement hooks
    bool profiling = false;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int ev_launches = 0;
    // launch-order buffers of the expected-work-first schedule: [0] for mpcq_solve, [1] for the chunks of mpcq_solve_host
    int32_t* perm[2] = {nullptr, nullptr};
    uint8_t* bucket[2] = {nullptr, nullptr};
    size_t perm_cap[2] = {0, 0};
    bool schedule = true;
    bool direct_results = true;           // mpcq_solve_host: kernels write into page-locked result buffers (MPCQ_HOST_DIRECT=0 turns it off)
    int* cta_hist = nullptr;              // per-CTA histograms of the two-launch schedule: [1 + kHostStreams][kSchedMaxCtas][64]
    // warm start of the next mpcq_solve calls (mpcq_set_warm_start)
    const uint8_t* face_in = nullptr;
    uint8_t* face_out = nullptr;
    // mpcq_solve runs its size classes side by side: classes 1.. are forked onto these streams and joined back (launch_all)
    cudaStream_t aux[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t fork_ev = nullptr, join_ev[3] = {nullptr, nullptr, nullptr};
};

namespace {
thread_local std::string g_create_err;

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

bool cuda_ok(mpcq_handle* h, cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    if (h) h->err = buf; else g_create_err = buf;
    return false;
}

template <class T, int NCAP, bool LG>
cudaError_t launch_class(mpcq_handle* h, const IO<T>& io, int ci, cudaStream_t st) {
    auto kern = mpcq_solve_kernel<T, NCAP, LG>;
    const mpcq::SizeClass& sc = mpcq::kClasses[ci];
    int epc = EnvsPerCta<NCAP>::value;
    while (epc > 1 && h->smem[ci] * epc > kMaxSmem) --epc;
    const int grid = LG ? (io.B < kGlobalCtas ? io.B : kGlobalCtas) : (io.B + epc - 1) / epc;
    kern<<<grid, 32 * TeamWarps<NCAP>::value * epc, h->smem[ci] * epc, st>>>(h->cs, io, static_cast<T*>(h->gws), h->gws_stride, sc.ns_lo,
                                                                          sc.ns_hi, (unsigned)h->smem[ci]);
    return cudaGetLastError();
}

// which (capacity, factor-in-global) combinations exist: the two small classes always fit in shared
// memory, the largest never does (fp32 348 KB); only the 192-slot class can go either way (fp64, long horizons)
template <class T, int NCAP>
cudaError_t launch_class_lg(mpcq_handle* h, const IO<T>& io, int ci, cudaStream_t st) {
    if constexpr (NCAP <= 128) {
        return launch_class<T, NCAP, false>(h, io, ci, st);
    } else if constexpr (NCAP >= 384) {
        return launch_class<T, NCAP, true>(h, io, ci, st);
    } else {
        return h->lglobal[ci] ? launch_class<T, NCAP, true>(h, io, ci, st) : launch_class<T, NCAP, false>(h, io, ci, st);
    }
}

template <class T>
cudaError_t launch_one(mpcq_handle* h, const IO<T>& io, int ci, cudaStream_t st) {
    switch (ci) {
        case 0: return launch_class_lg<T, 64>(h, io, ci, st);
        case 1: return launch_class_lg<T, 128>(h, io, ci, st);
        case 2: return launch_class_lg<T, 192>(h, io, ci, st);
        default: return launch_class_lg<T, 384>(h, io, ci, st);
    }
}

template <typename T>
cudaError_t launch_all(mpcq_handle* h, IO<T> io, int32_t* perm, uint8_t* bucket, cudaStream_t st, int* cta_hist) {
    const bool overlap = cta_hist != nullptr;
    cudaError_t e = cudaSuccess;
    h->last_launches = 0;
    h->ev_launches = 0;
    if (perm && io.B >= 512) {                                 // below ~one wave the order cannot matter
        if (overlap) {
            int g = (io.B + kSchedThreads - 1) / kSchedThreads;
            g = g > kSchedMaxCtas ? kSchedMaxCtas : g;
            mpcq_score_kernel<T><<<g, kSchedThreads, 0, st>>>(h->cs, io, bucket, cta_hist);
            mpcq_scatter_kernel<<<g, kSchedThreads, 0, st>>>(io.B, bucket, cta_hist, perm);
            h->last_launches += 2;
        } else {
            mpcq_schedule_kernel<T><<<1, 1024, 0, st>>>(h->cs, io, perm, bucket);
            ++h->last_launches;
        }
        e = cudaGetLastError();
        io.perm = perm;
    }
    if (overlap && !h->profiling && h->ncls > 1 && h->fork_ev && e == cudaSuccess) {
        // every environment belongs to exactly one class and the launches share nothing but read-only inputs, so the classes
        // run side by side: the larger ones (more work per environment) start first on their own streams, class 0 follows
        // on the caller's stream, which then waits for the others.  A class that is empty for this batch (trot at H = 10:
        // everything is in class 0) costs nothing on the critical path instead of a grid of CTAs that exit at once.
        cudaEventRecord(h->fork_ev, st);
        for (int ci = h->ncls - 1; ci >= 1 && e == cudaSuccess; --ci) {
            cudaStreamWaitEvent(h->aux[ci - 1], h->fork_ev, 0);
            e = launch_one<T>(h, io, ci, h->aux[ci - 1]);
            cudaEventRecord(h->join_ev[ci - 1], h->aux[ci - 1]);
            ++h->last_launches;
        }
        if (e == cudaSuccess) { e = launch_one<T>(h, io, 0, st); ++h->last_launches; }
        for (int ci = 1; ci < h->ncls; ++ci) cudaStreamWaitEvent(st, h->join_ev[ci - 1], 0);
        return e;
    }
    for (int ci = 0; ci < h->ncls && e == cudaSuccess; ++ci) {
        if (h->profiling) cudaEventRecord(h->ev[2 * ci], st);
        e = launch_one<T>(h, io, ci, st);
        if (h->profiling) { cudaEventRecord(h->ev[2 * ci + 1], st); ++h->ev_launches; }
        ++h->last_launches;
    }
    return e;
}

// The dynamic shared-memory limit of a kernel is per-device state shared by every handle of the process (a handle for a
// short horizon must not lower what a long-horizon handle needs), so it is always raised to the hardware maximum.
template <class T, int NCAP>
cudaError_t set_attr(mpcq_handle* h, int ci) {
    if constexpr (NCAP <= 128) {
        if (h->lglobal[ci]) return cudaErrorInvalidValue;
        return cudaFuncSetAttribute(mpcq_solve_kernel<T, NCAP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
    } else if constexpr (NCAP >= 384) {
        if (!h->lglobal[ci]) return cudaErrorInvalidValue;
        return cudaFuncSetAttribute(mpcq_solve_kernel<T, NCAP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
    } else {
        if (h->lglobal[ci])
            return cudaFuncSetAttribute(mpcq_solve_kernel<T, NCAP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
        return cudaFuncSetAttribute(mpcq_solve_kernel<T, NCAP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
    }
}

template <class T>
cudaError_t configure(mpcq_handle* h) {
    const int H = h->cs.horizon;
    h->ncls = mpcq::num_classes(H);
    size_t gws_elems = 0;
    for (int ci = 0; ci < h->ncls; ++ci) {
        const int ncap = mpcq::kClasses[ci].ncap;
        const int nmax = mpcq::class_nmax(mpcq::kClasses[ci]);
        size_t s = mpcq::work_bytes<T>(H, ncap, true, false, nmax, mpcq::kClasses[ci].nw);
        h->lglobal[ci] = s > kMaxSmem || ncap >= 384;
        if (h->lglobal[ci]) {
            s = mpcq::work_bytes<T>(H, ncap, false, false, nmax, mpcq::kClasses[ci].nw);
            if (s > kMaxSmem) return cudaErrorInvalidValue;
            size_t need = (size_t)mpcq::ps_elems(nmax);
            need = (need + 31) / 32 * 32;
            if (need > gws_elems) gws_elems = need;
        }
        if (const char* ov = getenv("MPCQ_CTAS_PER_SM")) {       // experiments only: fewer resident teams per SM (class 0)
            const int k = atoi(ov);
            if (ci == 0 && k >= 1 && k <= 12 && !h->lglobal[ci]) {
                const size_t want = ((size_t)233472 / k - 1024) / 128 * 128;
                if (want > s && want <= kMaxSmem) s = want;
            }
        }
        h->smem[ci] = (s + 127) / 128 * 128;
    }
    cudaError_t e = cudaSuccess;
    for (int ci = 0; ci < h->ncls && e == cudaSuccess; ++ci) {
        switch (ci) {
            case 0: e = set_attr<T, 64>(h, ci); break;
            case 1: e = set_attr<T, 128>(h, ci); break;
            case 2: e = set_attr<T, 192>(h, ci); break;
            default: e = set_attr<T, 384>(h, ci); break;
        }
    }
    if (e != cudaSuccess) return e;
    if (mpcq::work_bytes<T>(H, 384, false, true) > kMaxSmem) return cudaErrorInvalidValue;
    e = cudaFuncSetAttribute(mpcq_build_qp_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
    if (e != cudaSuccess) return e;
    if (gws_elems) {
        h->gws_stride = gws_elems;
        e = cudaMalloc(&h->gws, (gws_elems * kGlobalCtas + 4 * 512) * sizeof(T));   // slack: invert_factor reads up to 3 columns past a region
    }
    return e;
}

template <class T>
IO<T> make_io(int B, const void* x0, const void* yaw, const void* r_feet, const float* gait, const void* x_ref, void* f_out,
              void* u_full, int32_t* iters, double* resid, int32_t* status, uint8_t* active) {
    IO<T> io;
    io.x0 = static_cast<const T*>(x0);
    io.yaw = static_cast<const T*>(yaw);
    io.r_feet = static_cast<const T*>(r_feet);
    io.gait = gait;
    io.x_ref = static_cast<const T*>(x_ref);
    io.f_out = static_cast<T*>(f_out);
    io.u_full = static_cast<T*>(u_full);
    io.iters = iters;
    io.resid = resid;
    io.status = status;
    io.active = active;
    io.perm = nullptr;
    io.B = B;
    io.face_in = nullptr;
    io.face_out = nullptr;
    return io;
}

}  // namespace

extern "C" {

int mpcq_version(void) { return MPCQ_VERSION; }

const char* mpcq_last_error(const mpcq_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int mpcq_create(const mpcq_config* cfg, mpcq_handle** out) {
    if (!cfg || !out) { g_create_err = "null argument"; return MPCQ_ERR_INVALID; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        g_create_err = "no CUDA device: libmpcq has no CPU path";
        return MPCQ_ERR_NO_DEVICE;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_err = "device ordinal out of range"; return MPCQ_ERR_INVALID; }
    mpcq_handle* h = new (std::nothrow) mpcq_handle;
    if (!h) { g_create_err = "out of host memory"; return MPCQ_ERR_INVALID; }
    h->cfg = *cfg;
    std::string err;
    if (!mpcq::consts_from_config(*cfg, h->cs, err)) { g_create_err = err; delete h; return MPCQ_ERR_INVALID; }
    h->real_size = cfg->dtype == MPCQ_F64 ? 8 : 4;
    h->schedule = cfg->schedule >= 0;
    if (const char* ov = getenv("MPCQ_HOST_DIRECT")) h->direct_results = atoi(ov) != 0;
    DeviceGuard guard(cfg->device);
    cudaError_t e = cfg->dtype == MPCQ_F64 ? configure<double>(h) : configure<float>(h);
    if (e == cudaErrorInvalidValue && !h->smem[0]) { g_create_err = "horizon needs more shared memory than one SM has"; delete h; return MPCQ_ERR_UNSUPPORTED; }
    if (!cuda_ok(nullptr, e, "configure kernels")) { if (h->gws) cudaFree(h->gws); delete h; return MPCQ_ERR_CUDA; }
    for (int i = 0; i < kHostStreams; ++i)
        if (!cuda_ok(nullptr, cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking), "cudaStreamCreate")) {
            if (h->gws) cudaFree(h->gws);
            delete h;
            return MPCQ_ERR_CUDA;
        }
    // side streams of mpcq_solve's class overlap; without them (creation failed) the classes simply run in sequence
    bool side = cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 3 && side; ++i)
        side = cudaStreamCreateWithFlags(&h->aux[i], cudaStreamNonBlocking) == cudaSuccess &&
               cudaEventCreateWithFlags(&h->join_ev[i], cudaEventDisableTiming) == cudaSuccess;
    if (!side) { cudaGetLastError(); if (h->fork_ev) cudaEventDestroy(h->fork_ev); h->fork_ev = nullptr; }
    *out = h;
    return MPCQ_OK;
}

void mpcq_destroy(mpcq_handle* h) {
    if (!h) return;
    DeviceGuard guard(h->cfg.device);
    for (int i = 0; i < 3; ++i) {
        if (h->aux[i]) { cudaStreamSynchronize(h->aux[i]); cudaStreamDestroy(h->aux[i]); }
        if (h->join_ev[i]) cudaEventDestroy(h->join_ev[i]);
    }
    if (h->fork_ev) cudaEventDestroy(h->fork_ev);
    for (int i = 0; i < kHostStreams; ++i)
        if (h->streams[i]) { cudaStreamSynchronize(h->streams[i]); cudaStreamDestroy(h->streams[i]); }
    for (int i = 0; i < 8; ++i)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->gws) cudaFree(h->gws);
    for (int i = 0; i < 2; ++i) {
        if (h->perm[i]) cudaFree(h->perm[i]);
        if (h->bucket[i]) cudaFree(h->bucket[i]);
    }
    if (h->cta_hist) cudaFree(h->cta_hist);
    if (h->dev) cudaFree(h->dev);
    if (h->pin) cudaFreeHost(h->pin);
    for (int sl = 0; sl < 2; ++sl) {
        if (h->tick_dev[sl]) cudaFree(h->tick_dev[sl]);
        if (h->tick_state[sl]) cudaFree(h->tick_state[sl]);
    }
    delete h;
}

// launch-order buffers grow to the largest batch seen (slot 0: mpcq_solve, slot 1: chunks of mpcq_solve_host)
static int ensure_perm(mpcq_handle* h, int slot, size_t envs) {
    if (!h->schedule || envs <= h->perm_cap[slot]) return MPCQ_OK;
    cudaDeviceSynchronize();                                   // an earlier call may still read the old buffers
    if (h->perm[slot]) cudaFree(h->perm[slot]);
    if (h->bucket[slot]) cudaFree(h->bucket[slot]);
    h->perm[slot] = nullptr; h->bucket[slot] = nullptr; h->perm_cap[slot] = 0;
    const size_t cap = envs < 4096 ? 4096 : envs;
    if (!cuda_ok(h, cudaMalloc(&h->perm[slot], cap * sizeof(int32_t)), "cudaMalloc schedule")) return MPCQ_ERR_CUDA;
    if (!cuda_ok(h, cudaMalloc(&h->bucket[slot], cap), "cudaMalloc schedule")) return MPCQ_ERR_CUDA;
    if (!h->cta_hist &&
        !cuda_ok(h, cudaMalloc(&h->cta_hist, (size_t)(1 + kHostStreams) * kSchedMaxCtas * 64 * sizeof(int)), "cudaMalloc schedule")) return MPCQ_ERR_CUDA;
    h->perm_cap[slot] = cap;
    return MPCQ_OK;
}

// the synchronous host entry points share staging, launch-order buffers and streams with the asynchronous tick slots:
// whatever was submitted and not yet waited for is finished first (its results are then complete in the caller's buffers)
static void drain_ticks(mpcq_handle* h) {
    for (int sl = 0; sl < 2; ++sl)
        if (h->tick_inflight[sl]) { cudaStreamSynchronize(h->streams[sl]); h->tick_inflight[sl] = false; }
}

// solve envs [0,B) of the given arrays; `slot`/`off` select the region of the launch-order buffers this call may use
static int solve_impl(mpcq_handle* h, int32_t B, const void* x0, const void* yaw, const void* r_feet, const float* gait,
                      const void* x_ref, void* f_out, void* u_full, int32_t* iters, double* resid, int32_t* status,
                      uint8_t* active, int slot, size_t off, cudaStream_t st, int chunk = 0) {
    int32_t* perm = (h->schedule && h->perm[slot]) ? h->perm[slot] + off : nullptr;
    uint8_t* bucket = (h->schedule && h->bucket[slot]) ? h->bucket[slot] + off : nullptr;
    cudaError_t e;
    const bool warm = slot == 0;                               // the device entry point honours mpcq_set_warm_start
    // per-CTA histograms of the two-launch schedule: region 0 for mpcq_solve, 1 + chunk for the chunks of mpcq_solve_host
    int* hist = h->cta_hist ? h->cta_hist + (size_t)(slot == 0 ? 0 : 1 + chunk) * kSchedMaxCtas * 64 : nullptr;
    if (h->cfg.dtype == MPCQ_F64) {
        IO<double> io = make_io<double>(B, x0, yaw, r_feet, gait, x_ref, f_out, u_full, iters, resid, status, active);
        if (warm) { io.face_in = h->face_in; io.face_out = h->face_out; }
        e = launch_all<double>(h, io, perm, bucket, st, hist);
    } else {
        IO<float> io = make_io<float>(B, x0, yaw, r_feet, gait, x_ref, f_out, u_full, iters, resid, status, active);
        if (warm) { io.face_in = h->face_in; io.face_out = h->face_out; }
        e = launch_all<float>(h, io, perm, bucket, st, hist);
    }
    return cuda_ok(h, e, "mpcq_solve launch") ? MPCQ_OK : MPCQ_ERR_CUDA;
}

int mpcq_solve(mpcq_handle* h, int32_t B, const void* x0, const void* yaw, const void* r_feet, const float* gait,
               const void* x_ref, void* f_out, void* u_full, int32_t* iters, double* resid, int32_t* status, uint8_t* active,
               void* stream) {
    if (!h) return MPCQ_ERR_INVALID;
    if (B < 0 || (B > 0 && (!x0 || !r_feet || !gait || !x_ref || !f_out))) { h->err = "null input/output pointer"; return MPCQ_ERR_INVALID; }
    h->last_launches = 0;
    if (B == 0) return MPCQ_OK;
    DeviceGuard guard(h->cfg.device);
    const int rc = ensure_perm(h, 0, (size_t)B);
    if (rc != MPCQ_OK) return rc;
    return solve_impl(h, B, x0, yaw, r_feet, gait, x_ref, f_out, u_full, iters, resid, status, active, 0, 0, static_cast<cudaStream_t>(stream));
}

int mpcq_set_warm_start(mpcq_handle* h, const uint8_t* faces_in, uint8_t* faces_out) {
    if (!h) return MPCQ_ERR_INVALID;
    h->face_in = faces_in;
    h->face_out = faces_out;
    return MPCQ_OK;
}

int mpcq_build_qp(mpcq_handle* h, int32_t B, const void* x0, const void* yaw, const void* r_feet, const float* gait,
                  const void* x_ref, double* H_out, double* g_out, double* ub_out, void* stream) {
    if (!h) return MPCQ_ERR_INVALID;
    if (B < 0 || (B > 0 && (!x0 || !r_feet || !gait || !x_ref || !H_out || !g_out || !ub_out))) { h->err = "null input/output pointer"; return MPCQ_ERR_INVALID; }
    h->last_launches = 0;
    if (B == 0) return MPCQ_OK;
    DeviceGuard guard(h->cfg.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int H = h->cs.horizon;
    if (h->cfg.dtype == MPCQ_F64) {
        IO<double> io = make_io<double>(B, x0, yaw, r_feet, gait, x_ref, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        mpcq_build_qp_kernel<double><<<B, 32, mpcq::work_bytes<double>(H, 384, false, true), st>>>(h->cs, io, H_out, g_out, ub_out);
    } else {
        IO<float> io = make_io<float>(B, x0, yaw, r_feet, gait, x_ref, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        mpcq_build_qp_kernel<float><<<B, 32, mpcq::work_bytes<float>(H, 384, false, true), st>>>(h->cs, io, H_out, g_out, ub_out);
    }
    h->last_launches = 1;
    return cuda_ok(h, cudaGetLastError(), "mpcq_build_qp launch") ? MPCQ_OK : MPCQ_ERR_CUDA;
}

int mpcq_assemble(mpcq_handle* h, int32_t B, const double* quat, const double* pos, const double* omega, const double* vel,
                  const double* R_base, const double* v_des_body, const double* yaw_rate_des, double* xy_des, double* yaw_des,
                  double* rp_init, int32_t first_run, int32_t do_mpc, void* x0, void* yaw, void* x_ref, void* stream) {
    if (!h) return MPCQ_ERR_INVALID;
    if (B < 0 || (B > 0 && (!quat || !pos || !omega || !vel || !v_des_body || !yaw_rate_des || !xy_des || !yaw_des || !rp_init ||
                            !x0 || !yaw || (do_mpc && !x_ref)))) { h->err = "null input/output pointer"; return MPCQ_ERR_INVALID; }
    h->last_launches = 0;
    if (B == 0) return MPCQ_OK;
    DeviceGuard guard(h->cfg.device);
    AssembleArgs a{quat, pos, omega, vel, R_base, v_des_body, yaw_rate_des, xy_des, yaw_des, rp_init, first_run, do_mpc, B,
                   h->cs.horizon, h->cfg.dt_control, h->cfg.dt, h->cfg.com_height_des, h->cfg.gravity};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = (B + 127) / 128;
    if (h->cfg.dtype == MPCQ_F64)
        mpcq_assemble_kernel<double><<<grid, 128, 0, st>>>(a, static_cast<double*>(x0), static_cast<double*>(yaw), static_cast<double*>(x_ref));
    else
        mpcq_assemble_kernel<float><<<grid, 128, 0, st>>>(a, static_cast<float*>(x0), static_cast<float*>(yaw), static_cast<float*>(x_ref));
    h->last_launches = 1;
    return cuda_ok(h, cudaGetLastError(), "mpcq_assemble launch") ? MPCQ_OK : MPCQ_ERR_CUDA;
}

int mpcq_gait_tables(mpcq_handle* h, int32_t B, const int32_t* stance_offsets, const int32_t* stance_durations,
                     const int32_t* num_segment, const int32_t* cur_iteration, int32_t iterations_between_mpc, float* table,
                     double* swing_state, double* stance_state, void* stream) {
    if (!h) return MPCQ_ERR_INVALID;
    if (B < 0 || iterations_between_mpc < 1 ||
        (B > 0 && (!stance_offsets || !stance_durations || !num_segment || !cur_iteration || !table))) {
        h->err = "mpcq_gait_tables: null pointer or iterations_between_mpc < 1";
        return MPCQ_ERR_INVALID;
    }
    h->last_launches = 0;
    if (B == 0) return MPCQ_OK;
    DeviceGuard guard(h->cfg.device);
    GaitArgs a{stance_offsets, stance_durations, num_segment, cur_iteration, iterations_between_mpc, B, h->cs.horizon};
    mpcq_gait_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(a, table, swing_state, stance_state);
    h->last_launches = 1;
    return cuda_ok(h, cudaGetLastError(), "mpcq_gait_tables launch") ? MPCQ_OK : MPCQ_ERR_CUDA;
}

int mpcq_swing_targets(mpcq_handle* h, int32_t B, const mpcq_leg_params* lp, const double* pos_base, const double* lin_vel_base,
                       const double* R_base, const double* base_pos_base_thighs, const double* pos_feet, const double* swing_state,
                       const double* v_des_body, const double* yaw_rate_des, const double* swing_time, const double* stance_time,
                       uint8_t* swing_active, double* remaining_swing_time, double* footpos_init, double* footpos_final,
                       double* pos_targets, double* vel_targets, void* stream) {
    if (!h) return MPCQ_ERR_INVALID;
    if (B < 0 || !lp || !(lp->gravity > 0.0) ||
        (B > 0 && (!pos_base || !lin_vel_base || !R_base || !base_pos_base_thighs || !pos_feet || !swing_state || !v_des_body ||
                   !yaw_rate_des || !swing_time || !stance_time || !swing_active || !remaining_swing_time || !footpos_init ||
                   !footpos_final || !pos_targets || !vel_targets))) {
        h->err = "mpcq_swing_targets: null pointer or gravity <= 0";
        return MPCQ_ERR_INVALID;
    }
    h->last_launches = 0;
    if (B == 0) return MPCQ_OK;
    DeviceGuard guard(h->cfg.device);
    SwingArgs a{pos_base, lin_vel_base, R_base, base_pos_base_thighs, pos_feet, swing_state, v_des_body, yaw_rate_des, swing_time,
                stance_time, swing_active, remaining_swing_time, footpos_init, footpos_final, pos_targets, vel_targets,
                lp->swing_height, lp->dt_control, lp->gravity, lp->foot_z_final, B};
    mpcq_swing_kernel<<<(4 * B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
    h->last_launches = 1;
    return cuda_ok(h, cudaGetLastError(), "mpcq_swing_targets launch") ? MPCQ_OK : MPCQ_ERR_CUDA;
}

int mpcq_leg_torques(mpcq_handle* h, int32_t B, const mpcq_leg_params* lp, const double* Jv_feet, int32_t ncol, const double* R_base,
                     const double* base_pos_base_feet, const double* base_vel_base_feet, const void* contact_forces,
                     const double* swing_state, const double* pos_targets, const double* vel_targets, float* torque_cmds,
                     void* stream) {
    if (!h) return MPCQ_ERR_INVALID;
    if (B < 0 || !lp || (ncol != 18 && ncol != 3) ||
        (B > 0 && (!Jv_feet || !R_base || !base_pos_base_feet || !base_vel_base_feet || !contact_forces || !swing_state ||
                   !pos_targets || !vel_targets || !torque_cmds))) {
        h->err = "mpcq_leg_torques: null pointer, or ncol is neither 18 (reference Jacobian layout) nor 3 (joint blocks)";
        return MPCQ_ERR_INVALID;
    }
    h->last_launches = 0;
    if (B == 0) return MPCQ_OK;
    DeviceGuard guard(h->cfg.device);
    TorqueArgs a{Jv_feet, R_base, base_pos_base_feet, base_vel_base_feet, swing_state, pos_targets, vel_targets, contact_forces,
                 torque_cmds, {}, {}, B, ncol};
    for (int i = 0; i < 9; ++i) { a.kp[i] = lp->kp_swing[i]; a.kd[i] = lp->kd_swing[i]; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (h->cfg.dtype == MPCQ_F64) mpcq_torque_kernel<double><<<(4 * B + 127) / 128, 128, 0, st>>>(a);
    else mpcq_torque_kernel<float><<<(4 * B + 127) / 128, 128, 0, st>>>(a);
    h->last_launches = 1;
    return cuda_ok(h, cudaGetLastError(), "mpcq_leg_torques launch") ? MPCQ_OK : MPCQ_ERR_CUDA;
}

// Is [p, p + bytes) page-locked host memory?  The driver query costs 1-2 us per pointer, so the verified extent is remembered
// per address; first and last byte are both checked (a buffer registered only in part must go through the staging copy:
// the kernels would otherwise write through a device alias beyond the mapped range).
static bool host_range_pinned(mpcq_handle* h, const void* p, size_t bytes) {
    if (!p || !bytes) return false;
    auto it = h->pinned_cache.find(p);
    if (it != h->pinned_cache.end() && (it->second == 0 || it->second >= bytes)) return it->second != 0;
    cudaPointerAttributes a0, a1;
    bool pin = cudaPointerGetAttributes(&a0, p) == cudaSuccess && a0.type == cudaMemoryTypeHost;
    if (pin) pin = cudaPointerGetAttributes(&a1, static_cast<const char*>(p) + bytes - 1) == cudaSuccess && a1.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (it != h->pinned_cache.end()) it->second = pin ? bytes : 0;
    else if (h->pinned_cache.size() < 4096) h->pinned_cache.emplace(p, pin ? bytes : 0);
    return pin;
}

int mpcq_solve_host(mpcq_handle* h, int32_t B, const void* x0, const void* yaw, const void* r_feet, const float* gait,
                    const void* x_ref, void* f_out, void* u_full, int32_t* iters, double* resid, int32_t* status,
                    uint8_t* active) {
    if (!h) return MPCQ_ERR_INVALID;
    if (B < 0 || (B > 0 && (!x0 || !r_feet || !gait || !x_ref || !f_out))) { h->err = "null input/output pointer"; return MPCQ_ERR_INVALID; }
    h->last_launches = 0;
    if (B == 0) return MPCQ_OK;
    DeviceGuard guard(h->cfg.device);
    drain_ticks(h);
    const size_t rs = h->real_size, H = (size_t)h->cs.horizon, b = (size_t)B;
    // per-environment byte widths of the 5 inputs and 6 outputs
    const size_t width[11] = {13 * rs, yaw ? rs : 0, 12 * rs, 4 * H * 4, 13 * H * rs,
                              12 * rs, u_full ? 12 * H * rs : 0, iters ? (size_t)8 : 0, resid ? (size_t)16 : 0,
                              status ? (size_t)4 : 0, active ? 4 * H : 0};
    const void* src[5] = {x0, yaw, r_feet, gait, x_ref};
    void* dst[6] = {f_out, u_full, iters, resid, status, active};
    size_t off[12], cur = 0;
    for (int i = 0; i < 11; ++i) { off[i] = cur; cur += (b * width[i] + 255) / 256 * 256; }
    off[11] = cur;
    const size_t total = cur;
    if (total > h->stage_cap) {
        for (int i = 0; i < kHostStreams; ++i) cudaStreamSynchronize(h->streams[i]);
        if (h->dev) cudaFree(h->dev);
        if (h->pin) cudaFreeHost(h->pin);
        h->dev = h->pin = nullptr;
        h->stage_cap = 0;
        if (!cuda_ok(h, cudaMalloc(&h->dev, total), "cudaMalloc staging")) return MPCQ_ERR_CUDA;
        if (!cuda_ok(h, cudaMallocHost(&h->pin, total), "cudaMallocHost staging")) return MPCQ_ERR_CUDA;
        h->stage_cap = total;
    }
    // caller buffers that are already page-locked are used for DMA directly; pageable ones go through the pinned staging
    bool pinned[11];
    for (int i = 0; i < 11; ++i) {
        const void* p = i < 5 ? src[i] : dst[i - 5];
        pinned[i] = (p && width[i]) ? host_range_pinned(h, p, b * width[i]) : false;
    }
    // the batch is cut into chunks, each on its own stream: H2D(c+1) overlaps solve(c) overlaps D2H(c-1), and the
    // kernels of neighbouring chunks fill each other's tails.  A class whose factor lives in the shared global
    // workspace cannot run twice concurrently, so such handles use one chunk.
    bool any_global = false;
    for (int ci = 0; ci < h->ncls; ++ci) any_global = any_global || h->lglobal[ci];
    // measured at B = 4096 (v13): 1 / 2 / 3 / 4 chunks = 0.80-0.82 / 0.81 / 0.81-0.84 / 0.83 ms (smaller chunks hide more of the copies
    // but give the expected-work-first schedule less to work with); large batches take all four streams.  Reading page-locked
    // inputs in place (zero copy, one chunk) was tried and is slower (0.85 ms): the kernels read their inputs twice and the
    // first wave of CTAs queues 2.8 MB of PCIe reads in front of the hardest environments.
    // v15: with the results written in place one chunk wins at 4 096 (0.78-0.79 vs 0.79-0.81 ms): the launch order sees the whole batch
    int nchunk = any_global ? 1 : (B >= 16384 ? kHostStreams : (B >= 8192 ? 2 : 1));
    if (const char* ov = getenv("MPCQ_HOST_CHUNKS")) {          // experiments only
        const int v = atoi(ov);
        if (v >= 1 && v <= kHostStreams && !any_global) nchunk = v;
    }
    {
        const int rcp = ensure_perm(h, 1, b);
        if (rcp != MPCQ_OK) return rcp;
    }
    char* d = h->dev;
    int launches = 0;
    // on a failure in mid-loop the copies and kernels already queued still use the caller's buffers: wait for them before returning
    auto fail = [&](int code) { for (int i = 0; i < kHostStreams; ++i) cudaStreamSynchronize(h->streams[i]); return code; };
    // Results that go to page-locked caller buffers are written there by the kernels themselves (mapped host memory: a few
    // posted PCIe writes per environment when it finishes) instead of being staged on the device and copied behind the kernel.
    char* zdst[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int i = 5; i < 11; ++i) {
        if (!width[i] || !pinned[i] || !h->direct_results) continue;
        void* dptr = nullptr;
        if (cudaHostGetDevicePointer(&dptr, dst[i - 5], 0) == cudaSuccess) zdst[i - 5] = static_cast<char*>(dptr);
        else cudaGetLastError();
    }
    for (int c = 0; c < nchunk; ++c) {
        const size_t lo = b * c / nchunk, hi = b * (c + 1) / nchunk, nb = hi - lo;
        if (nb == 0) continue;
        cudaStream_t st = h->streams[c % kHostStreams];
        for (int i = 0; i < 5; ++i) {
            if (!width[i]) continue;
            const char* from = static_cast<const char*>(src[i]) + lo * width[i];
            if (!pinned[i]) {
                memcpy(h->pin + off[i] + lo * width[i], from, nb * width[i]);
                from = h->pin + off[i] + lo * width[i];
            }
            if (!cuda_ok(h, cudaMemcpyAsync(d + off[i] + lo * width[i], from, nb * width[i], cudaMemcpyHostToDevice, st), "H2D")) return fail(MPCQ_ERR_CUDA);
        }
        auto dp = [&](int i) -> char* {
            if (!width[i]) return nullptr;
            return (i >= 5 && zdst[i - 5]) ? zdst[i - 5] + lo * width[i] : d + off[i] + lo * width[i];
        };
        const int rc = solve_impl(h, (int32_t)nb, dp(0), dp(1), dp(2), reinterpret_cast<float*>(dp(3)), dp(4), dp(5), dp(6),
                                  reinterpret_cast<int32_t*>(dp(7)), reinterpret_cast<double*>(dp(8)),
                                  reinterpret_cast<int32_t*>(dp(9)), reinterpret_cast<uint8_t*>(dp(10)), 1, lo, st, c % kHostStreams);
        if (rc != MPCQ_OK) return fail(rc);
        launches += h->last_launches;
        for (int i = 5; i < 11; ++i) {
            if (!width[i] || zdst[i - 5]) continue;
            char* to = pinned[i] ? static_cast<char*>(dst[i - 5]) + lo * width[i] : h->pin + off[i] + lo * width[i];
            if (!cuda_ok(h, cudaMemcpyAsync(to, d + off[i] + lo * width[i], nb * width[i], cudaMemcpyDeviceToHost, st), "D2H")) return fail(MPCQ_ERR_CUDA);
        }
    }
    for (int c = 0; c < nchunk && c < kHostStreams; ++c)
        if (!cuda_ok(h, cudaStreamSynchronize(h->streams[c]), "mpcq_solve_host sync")) return MPCQ_ERR_CUDA;
    for (int i = 5; i < 11; ++i)
        if (width[i] && !pinned[i]) memcpy(dst[i - 5], h->pin + off[i], b * width[i]);
    h->last_launches = launches;
    return MPCQ_OK;
}

// slot: which of the two independent tick pipelines (device buffers, controller state, stream) the call uses; wait: synchronise
// and finish staged copies before returning (the asynchronous entry points need page-locked buffers and skip it)
static int tick_impl(mpcq_handle* h, int slot, bool wait, int32_t B, const double* state_cmd, const int32_t* gait_params,
                     int32_t iterations_between_mpc, int32_t first_run, void* f_out, int32_t* status) {
    if (!h) return MPCQ_ERR_INVALID;
    if (B < 0 || iterations_between_mpc < 1 || (B > 0 && (!state_cmd || !gait_params || !f_out))) {
        h->err = "mpcq_tick_host: null pointer or iterations_between_mpc < 1";
        return MPCQ_ERR_INVALID;
    }
    h->last_launches = 0;
    if (B == 0) return MPCQ_OK;
    DeviceGuard guard(h->cfg.device);
    if (wait) drain_ticks(h);                                   // the synchronous call may use several streams
    else if (h->tick_inflight[slot]) { cudaStreamSynchronize(h->streams[slot]); h->tick_inflight[slot] = false; }   // submitted twice
    const size_t rs = h->real_size, H = (size_t)h->cs.horizon, b = (size_t)B;
    cudaStream_t st = h->streams[slot];
    // device layout (256-byte aligned sections): packed inputs | unpacked arrays | x0, yaw, x_ref, gait table, feet | results
    const size_t width[16] = {29 * 8, 10 * 4,                              // state_cmd, gait_params
                              4 * 8, 3 * 8, 3 * 8, 3 * 8, 3 * 8, 8,        // quat pos omega vel vdes yawrate
                              16, 16, 4, 4,                                // offs durs seg iter
                              (13 + 1 + 12 + 13 * H) * rs + 16 * H,        // x0 | yaw | feet | x_ref | table (sub-sections below)
                              12 * rs, 4, 0};                              // f_out, status
    size_t off[17], cur = 0;
    for (int i = 0; i < 16; ++i) { off[i] = cur; cur += (b * width[i] + 255) / 256 * 256 + (i == 12 ? 5 * 256 : 0); }
    off[16] = cur;
    if (b > h->tick_cap[slot]) {
        cudaStreamSynchronize(st);
        if (h->tick_dev[slot]) cudaFree(h->tick_dev[slot]);
        if (h->tick_state[slot]) cudaFree(h->tick_state[slot]);
        h->tick_dev[slot] = nullptr; h->tick_state[slot] = nullptr; h->tick_cap[slot] = 0;
        if (!cuda_ok(h, cudaMalloc(&h->tick_dev[slot], cur), "cudaMalloc tick")) return MPCQ_ERR_CUDA;
        if (!cuda_ok(h, cudaMalloc(&h->tick_state[slot], b * 5 * sizeof(double)), "cudaMalloc tick state")) return MPCQ_ERR_CUDA;
        if (!cuda_ok(h, cudaMemsetAsync(h->tick_state[slot], 0, b * 5 * sizeof(double), st), "memset tick state")) return MPCQ_ERR_CUDA;
        cudaStreamSynchronize(st);                              // the chunks below run on several streams
        h->tick_cap[slot] = b;
    }
    {
        const int rcp = ensure_perm(h, 1, 2 * b);                // each slot owns one half of the launch-order buffers
        if (rcp != MPCQ_OK) return rcp;
    }
    const size_t perm_off = slot ? h->perm_cap[1] / 2 : 0;
    char* d = h->tick_dev[slot];
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    // page-locked caller buffers are DMA sources / are written in place by the kernels; pageable ones go through staging
    const size_t need_pin = al(b * width[0]) + al(b * width[1]) + al(b * width[13]) + al(b * width[14]);
    if (need_pin > h->stage_cap) {
        for (int i = 0; i < kHostStreams; ++i) cudaStreamSynchronize(h->streams[i]);
        if (h->dev) cudaFree(h->dev);
        if (h->pin) cudaFreeHost(h->pin);
        h->dev = h->pin = nullptr; h->stage_cap = 0;
        if (!cuda_ok(h, cudaMalloc(&h->dev, need_pin), "cudaMalloc staging")) return MPCQ_ERR_CUDA;
        if (!cuda_ok(h, cudaMallocHost(&h->pin, need_pin), "cudaMallocHost staging")) return MPCQ_ERR_CUDA;
        h->stage_cap = need_pin;
    }
    // the batch is cut into chunks on separate streams (like mpcq_solve_host): the copies and the small kernels of one chunk
    // hide behind the solve of the other, and the solve kernels of neighbouring chunks fill each other's tails
    bool any_global = false;
    for (int ci = 0; ci < h->ncls; ++ci) any_global = any_global || h->lglobal[ci];
    int nchunk = (any_global || !wait) ? 1 : (B >= 16384 ? kHostStreams : (B >= 8192 ? 2 : 1));   // measured at 4 096 robots: 1 / 2 / 3 chunks = 585 / 608 / 642 us
    if (const char* ov = getenv("MPCQ_HOST_CHUNKS")) {          // experiments only
        const int v = atoi(ov);
        if (v >= 1 && v <= kHostStreams && !any_global && wait) nchunk = v;
    }
    const void* src[2] = {state_cmd, gait_params};
    bool src_pinned[2];
    size_t poffs[2] = {0, al(b * width[0])};
    for (int i = 0; i < 2; ++i) src_pinned[i] = host_range_pinned(h, src[i], b * width[i]);
    if (!wait && (!src_pinned[0] || !src_pinned[1])) { h->err = "mpcq_tick_host_submit needs page-locked buffers"; return MPCQ_ERR_INVALID; }
    // results: in place into page-locked caller buffers, else staged
    void* dst[2] = {f_out, status};
    char* dres[2] = {d + off[13], d + off[14]};
    bool direct[2] = {false, false}, dst_pinned[2] = {false, false};
    for (int i = 0; i < 2; ++i) {
        if (!dst[i]) continue;
        dst_pinned[i] = host_range_pinned(h, dst[i], b * width[13 + i]);
        if (!wait && !dst_pinned[i]) { h->err = "mpcq_tick_host_submit needs page-locked buffers"; return MPCQ_ERR_INVALID; }
        if (!h->direct_results || !dst_pinned[i]) continue;
        void* dptr = nullptr;
        if (cudaHostGetDevicePointer(&dptr, dst[i], 0) == cudaSuccess) { dres[i] = static_cast<char*>(dptr); direct[i] = true; }
        else cudaGetLastError();
    }
    char* pres = h->pin + al(b * width[0]) + al(b * width[1]);
    auto fail = [&](int code) { for (int i = 0; i < kHostStreams; ++i) cudaStreamSynchronize(h->streams[i]); h->tick_inflight[0] = h->tick_inflight[1] = false; return code; };
    TickArrays ta{reinterpret_cast<double*>(d + off[2]), reinterpret_cast<double*>(d + off[3]), reinterpret_cast<double*>(d + off[4]),
                  reinterpret_cast<double*>(d + off[5]), reinterpret_cast<double*>(d + off[6]), reinterpret_cast<double*>(d + off[7]),
                  reinterpret_cast<int32_t*>(d + off[8]), reinterpret_cast<int32_t*>(d + off[9]), reinterpret_cast<int32_t*>(d + off[10]),
                  reinterpret_cast<int32_t*>(d + off[11])};
    char* x0 = d + off[12];
    char* yaw = x0 + al(b * 13 * rs);
    char* feet = yaw + al(b * rs);
    char* xref = feet + al(b * 12 * rs);
    char* table = xref + al(b * 13 * H * rs);
    double* cst = h->tick_state[slot];
    int launches = 0;
    for (int c = 0; c < nchunk; ++c) {
        const size_t lo = b * c / nchunk, hi = b * (c + 1) / nchunk, nb = hi - lo;
        if (nb == 0) continue;
        cudaStream_t st = h->streams[(slot + c) % kHostStreams];
        for (int i = 0; i < 2; ++i) {
            const char* from = static_cast<const char*>(src[i]) + lo * width[i];
            if (!src_pinned[i]) { memcpy(h->pin + poffs[i] + lo * width[i], from, nb * width[i]); from = h->pin + poffs[i] + lo * width[i]; }
            if (!cuda_ok(h, cudaMemcpyAsync(d + off[i] + lo * width[i], from, nb * width[i], cudaMemcpyHostToDevice, st), "H2D")) return fail(MPCQ_ERR_CUDA);
        }
        TickArrays tc{ta.quat + 4 * lo, ta.pos + 3 * lo, ta.omega + 3 * lo, ta.vel + 3 * lo, ta.vdes + 3 * lo, ta.yawrate + lo,
                      ta.offs + 4 * lo, ta.durs + 4 * lo, ta.seg + lo, ta.iter + lo};
        const int nbi = (int)nb, grid = (nbi + 127) / 128;
        const double* sc = reinterpret_cast<const double*>(d + off[0]) + 29 * lo;
        const int32_t* gp = reinterpret_cast<const int32_t*>(d + off[1]) + 10 * lo;
        char* cx0 = x0 + lo * 13 * rs; char* cyaw = yaw + lo * rs; char* cfeet = feet + lo * 12 * rs; char* cxref = xref + lo * 13 * H * rs;
        float* ctab = reinterpret_cast<float*>(table) + lo * 4 * H;
        if (h->cfg.dtype == MPCQ_F64) mpcq_tick_unpack_kernel<double><<<grid, 128, 0, st>>>(nbi, sc, gp, tc, reinterpret_cast<double*>(cfeet));
        else mpcq_tick_unpack_kernel<float><<<grid, 128, 0, st>>>(nbi, sc, gp, tc, reinterpret_cast<float*>(cfeet));
        GaitArgs ga{tc.offs, tc.durs, tc.seg, tc.iter, iterations_between_mpc, nbi, h->cs.horizon};
        mpcq_gait_kernel<<<grid, 128, 0, st>>>(ga, ctab, nullptr, nullptr);
        AssembleArgs aa{tc.quat, tc.pos, tc.omega, tc.vel, nullptr, tc.vdes, tc.yawrate, cst + 2 * lo, cst + 2 * b + lo, cst + 3 * b + 2 * lo,
                        first_run, 1, nbi, h->cs.horizon, h->cfg.dt_control, h->cfg.dt, h->cfg.com_height_des, h->cfg.gravity};
        if (h->cfg.dtype == MPCQ_F64)
            mpcq_assemble_kernel<double><<<grid, 128, 0, st>>>(aa, reinterpret_cast<double*>(cx0), reinterpret_cast<double*>(cyaw), reinterpret_cast<double*>(cxref));
        else
            mpcq_assemble_kernel<float><<<grid, 128, 0, st>>>(aa, reinterpret_cast<float*>(cx0), reinterpret_cast<float*>(cyaw), reinterpret_cast<float*>(cxref));
        if (!cuda_ok(h, cudaGetLastError(), "mpcq_tick_host launch")) return fail(MPCQ_ERR_CUDA);
        const int rc = solve_impl(h, nbi, cx0, cyaw, cfeet, ctab, cxref, dres[0] + lo * width[13], nullptr, nullptr, nullptr,
                                  status ? reinterpret_cast<int32_t*>(dres[1] + lo * width[14]) : nullptr, nullptr, 1, perm_off + lo, st, (slot + c) % kHostStreams);
        if (rc != MPCQ_OK) return fail(rc);
        launches += h->last_launches + 3;
        for (int i = 0; i < 2; ++i) {
            if (!dst[i] || direct[i]) continue;
            char* to = (dst_pinned[i] ? static_cast<char*>(dst[i]) : pres + (i ? al(b * width[13]) : 0)) + lo * width[13 + i];
            if (!cuda_ok(h, cudaMemcpyAsync(to, dres[i] + lo * width[13 + i], nb * width[13 + i], cudaMemcpyDeviceToHost, st), "D2H")) return fail(MPCQ_ERR_CUDA);
        }
    }
    h->last_launches = launches;
    if (!wait) { h->tick_inflight[slot] = true; return MPCQ_OK; }
    for (int c = 0; c < nchunk && c < kHostStreams; ++c)
        if (!cuda_ok(h, cudaStreamSynchronize(h->streams[(slot + c) % kHostStreams]), "mpcq_tick_host sync")) return fail(MPCQ_ERR_CUDA);
    for (int i = 0; i < 2; ++i)
        if (dst[i] && !direct[i] && !dst_pinned[i]) memcpy(dst[i], pres + (i ? al(b * width[13]) : 0), b * width[13 + i]);
    h->last_launches = launches;
    return MPCQ_OK;
}

int mpcq_tick_host(mpcq_handle* h, int32_t B, const double* state_cmd, const int32_t* gait_params, int32_t iterations_between_mpc,
                   int32_t first_run, void* f_out, int32_t* status) {
    return tick_impl(h, 0, true, B, state_cmd, gait_params, iterations_between_mpc, first_run, f_out, status);
}

int mpcq_tick_host_submit(mpcq_handle* h, int32_t slot, int32_t B, const double* state_cmd, const int32_t* gait_params,
                          int32_t iterations_between_mpc, int32_t first_run, void* f_out, int32_t* status) {
    if (!h) return MPCQ_ERR_INVALID;
    if (slot < 0 || slot > 1) { h->err = "mpcq_tick_host_submit: slot must be 0 or 1"; return MPCQ_ERR_INVALID; }
    return tick_impl(h, slot, false, B, state_cmd, gait_params, iterations_between_mpc, first_run, f_out, status);
}

int mpcq_tick_host_wait(mpcq_handle* h, int32_t slot) {
    if (!h) return MPCQ_ERR_INVALID;
    if (slot < 0 || slot > 1) { h->err = "mpcq_tick_host_wait: slot must be 0 or 1"; return MPCQ_ERR_INVALID; }
    DeviceGuard guard(h->cfg.device);
    h->tick_inflight[slot] = false;
    return cuda_ok(h, cudaStreamSynchronize(h->streams[slot]), "mpcq_tick_host_wait") ? MPCQ_OK : MPCQ_ERR_CUDA;
}

int mpcq_tick_reset(mpcq_handle* h) {
    if (!h) return MPCQ_ERR_INVALID;
    DeviceGuard guard(h->cfg.device);
    for (int i = 0; i < kHostStreams; ++i) cudaStreamSynchronize(h->streams[i]);
    h->tick_inflight[0] = h->tick_inflight[1] = false;
    for (int sl = 0; sl < 2; ++sl)
        if (h->tick_state[sl] && !cuda_ok(h, cudaMemset(h->tick_state[sl], 0, h->tick_cap[sl] * 5 * sizeof(double)), "mpcq_tick_reset")) return MPCQ_ERR_CUDA;
    return MPCQ_OK;
}

int mpcq_last_launch_count(const mpcq_handle* h) { return h ? h->last_launches : 0; }

#ifdef MPCQ_PHASE_CLOCKS
// development builds only (not declared in mpcq.h): read and clear the per-phase cycle table
int mpcq_debug_phase_cycles(unsigned long long* out16) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out16, mpcq::g_phase_cycles, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    unsigned long long zero[16] = {};
    return cudaMemcpyToSymbol(mpcq::g_phase_cycles, zero, sizeof zero) == cudaSuccess ? 0 : -1;
}
#endif

int mpcq_set_profiling(mpcq_handle* h, int32_t enable) {
    if (!h) return MPCQ_ERR_INVALID;
    DeviceGuard guard(h->cfg.device);
    if (enable)
        for (int i = 0; i < 8; ++i)
            if (!h->ev[i] && !cuda_ok(h, cudaEventCreate(&h->ev[i]), "cudaEventCreate")) return MPCQ_ERR_CUDA;
    h->profiling = enable != 0;
    h->ev_launches = 0;
    return MPCQ_OK;
}

int mpcq_last_kernel_ms(mpcq_handle* h, float* ms, int32_t cap) {
    if (!h || !ms) return MPCQ_ERR_INVALID;
    DeviceGuard guard(h->cfg.device);
    const int n = h->ev_launches;
    for (int i = 0; i < n && i < cap; ++i) {
        if (!cuda_ok(h, cudaEventSynchronize(h->ev[2 * i + 1]), "cudaEventSynchronize")) return MPCQ_ERR_CUDA;
        if (!cuda_ok(h, cudaEventElapsedTime(&ms[i], h->ev[2 * i], h->ev[2 * i + 1]), "cudaEventElapsedTime")) return MPCQ_ERR_CUDA;
    }
    return n;
}

}  // extern "C"
