// Per-environment convex-MPC QP engine: one warp builds and solves one robot's QP.
//
// Replaces, for a batch, the body of ModelPredictiveController._solve_mpc
// (reference linear_mpc/mpc.py:262-290, drake branch):
//   _generate_state_space_model  :173-192   -> setup_model()      (K1: Rz, world inertia, foot levers)
//   _discretize_continuous_model :194-208   -> closed form, Ac^3 = 0 (no expm)
//   _generate_QP_cost            :211-235   -> setup_model() + chol_factor() assembly (K2)
//   _generate_QP_constraints     :237-260   -> stance list + friction-pyramid faces (K3)
//   Solve()                      :277-286   -> pdas / active-set drivers over chol_factor + tri_solve (K4)
//
// Math (oracle/structured.py states the same in numpy and is pinned against the reference):
//   H = 2 (N (x) M00 + S (x) M11 + I (x) R),  M00 = B0'QB0, M11 = B1'QB1 (12x12),
//   N_ij = H - max(i,j),  S_ij = sum_{k>=max(i,j)} (k-i+1/2)(k-j+1/2),
//   g_j = 2 (B0' E0_j + B1' E1_j) from the suffix sums of Q (A^(k+1) x0 - xref_k).
// The constraints 0 <= C u <= ub are, per foot-step, membership of f = (fx,fy,fz) in the truncated
// friction pyramid K = {|fx|,|fy| <= mu fz, 0 <= fz <= fmax}.  A *face* of K is coded (sx,sy,sz):
//   sx,sy in {-1,0,+1}: fx = sx mu fz / free;  sz in {-1,0,+1}: apex f = 0 / fz free / fz = fmax.
// On a choice of faces the QP is an unconstrained strictly convex problem in the free
// parameters (f = c + Z w); it is solved by a dense Cholesky in shared memory (precision T) with
// fp64 residual refinement through the structured operator.  Faces are updated by primal-dual
// active-set rounds, with a feasible primal active-set method as the anti-cycling fallback.
// Swing foot-steps (ub_fz = 0 pins f = 0) never enter the factorisation.
#pragma once

#include "mpcq_warp.cuh"
#if defined(MPCQ_HOST_EMU) && defined(MPCQ_TRACE)
#include <stdio.h>
#endif

namespace mpcq {

// Optional per-phase cycle accounting (-DMPCQ_PHASE_CLOCKS, development builds only): lane 0 of every team adds the
// clock64() deltas of the phases it walks through to a global table read back by mpcq_debug_phase_cycles().
#if defined(MPCQ_PHASE_CLOCKS) && !defined(MPCQ_HOST_EMU)
__device__ unsigned long long g_phase_cycles[16];
struct PhaseClock {
    long long t0;
    int id;
    __device__ __forceinline__ explicit PhaseClock(int i) : t0(0), id(i) {
#ifdef __CUDA_ARCH__
        t0 = clock64();
#endif
    }
    __device__ __forceinline__ ~PhaseClock() {
#ifdef __CUDA_ARCH__
        if ((threadIdx.x & 31) == 0) atomicAdd(&g_phase_cycles[id], (unsigned long long)(clock64() - t0));
#endif
    }
};
#define MPCQ_PHASE(id) PhaseClock _pc_##id(id)
#elif defined(MPCQ_HOST_EMU)
// host emulation (tests only): count how often thread 0 of a team enters each phase
static long g_emu_phase_calls[16];
#define MPCQ_PHASE(id) do { if (::mpcq_emu::thread_id() == 0) ++g_emu_phase_calls[id]; } while (0)
#else
#define MPCQ_PHASE(id)
#endif

// ---------------------------------------------------------------------------------------------
struct Consts {
    int horizon;
    int pdas_cap;          // primal-dual active-set rounds before the fallback
    int as_cap;            // primal active-set iterations in the fallback
    int refine_max;        // residual-refinement solves per factorisation
    double dt, mu, fz_max, inv_mass;
    double inertia[9];
    double q[13];
    double r[12];
    double tol_p, tol_d;   // relative primal / dual tolerances of the face tests
    double tol_r_loose, tol_r_tight;   // reduced-gradient tolerances (relative to 1 + |g|_inf)
    double tol_active;     // slack tolerance of the reported constraint activity
    double tol_r_abs;      // absolute cap of the tight reduced-gradient tolerance: 2 min(R) * (force accuracy in N)
};

template <class T> struct IO {
    const T* x0;           // [B,13]
    const T* yaw;          // [B] or null
    const T* r_feet;       // [B,12]
    const float* gait;     // [B,4H]
    const T* x_ref;        // [B,13H]
    T* f_out;              // [B,12]
    T* u_full;             // [B,12H] or null
    int32_t* iters;        // [B,2] or null
    double* resid;         // [B,2] or null
    int32_t* status;       // [B] or null
    uint8_t* active;       // [B,4H] or null
    const int32_t* perm;   // [B] launch order (expected-work-first schedule) or null = natural order
    int B;
    // warm start (mpcq_set_warm_start): one face code per foot-step, (sx & 3) | (sy & 3) << 2 | (sz & 3) << 4 with -1 stored
    // as 3, so that a zeroed buffer means "every face free" = cold start
    const uint8_t* face_in;   // [B,4H] or null
    uint8_t* face_out;        // [B,4H] or null
};

MPCQ_DEV int face_decode(int c) { c &= 3; return c == 3 ? -1 : (c == 2 ? 0 : c); }
MPCQ_DEV int face_encode(int sx, int sy, int sz) { return (sx & 3) | ((sy & 3) << 2) | ((sz & 3) << 4); }

enum : int { ST_VERIFIED = 1, ST_FALLBACK = 2, ST_MAXITER = 4, ST_NUMERIC = 8, ST_NO_STANCE = 32 };

// ---------------------------------------------------------------------------------------------
// per-warp workspace
MPCQ_HD constexpr size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
MPCQ_HD constexpr int l_elems(int n) { return n * n / 2 + 2 * n + 32; }   // +32: unpredicated row sweeps may overrun

template <class T> struct Work {
    // fp64
    double *Md, *GW, *g, *u, *gam, *P0, *P1, *ucur, *utrial, *hd, *fmax, *zero3;
    double *hs, *hq, *rd;  // hess_apply tables: stage-1 scales [9], stage-2 scales [12], diag(R) [12]
    // precision T
    T *L, *dblk, *vec, *cw, *zt, *Mf, *r2;   // r2 = 2 * diag(R) in precision T
    T *NS2;                // [H][H][2]: (2 N_ij, 2 S_ij) pairs for the Hessian assembly
    int32_t* sinf;         // per slot: step | leg << 8 | foot << 16 | dead << 30
    // bytes
    uint8_t *fk;           // stance list: full foot-step index k = 4*step + leg
    uint8_t *fo;           // stance index (position in the ORIGINAL stance list) of the foot now at position p;
                           // the fp64 vectors g, u, gam, ucur, utrial are stored in that fixed compact order
    uint8_t *cidx;         // full foot-step k -> stance index, 255 = swing
    int8_t* face;          // 3 per stance foot-step
    int8_t* face2;         // trial faces of the fallback
    int8_t* facef;         // faces the current factor was built for
    int n, ns, H, nv;      // nv = 3 * ns: live length of the compact vectors
    team::Ctx t;             // the team of warps that owns this environment
};

// Shared-memory layout of one team's workspace.  nvc = capacity of everything indexed by slot (the largest system of the
// size class, `nmax`, or `ncap` when not given); cw is double-buffered only for multi-warp teams.  The small class must
// stay within 18 432 B (fp32) so that 12 one-warp teams fit in the 228 KB of an SM (12 x (18 432 + 1 024 reserved)).
struct Layout { size_t nd, nt, nb; int nvc, cw; };
MPCQ_HD constexpr Layout layout(int H, int ncap, bool l_in_smem, bool with_md, int nmax, int nw) {
    Layout l{};
    l.nvc = nmax > 0 ? nmax : ncap;
    l.cw = nw > 1 ? 256 : 128;
    const size_t nv = (size_t)l.nvc;
    l.nd = (with_md ? 288 : 0) + 72 + 9 * (size_t)H + 12 * (size_t)H + 6 * nv + nv / 3 + 1 + 12 + 40;
    l.nt = (l_in_smem ? (size_t)l_elems(l.nvc) : 0) + 3 * nv + nv + (size_t)l.cw + 3 * nv + 288 + 12 + 2 * (size_t)H * H;
    l.nb = (nv / 3 + 1) * 11 + 4 * (size_t)H + 16 + 4 * nv;
    return l;
}

template <class T>
MPCQ_HD constexpr size_t work_bytes(int H, int ncap, bool l_in_smem, bool with_md = false, int nmax = 0, int nw = 2) {
    const Layout l = layout(H, ncap, l_in_smem, with_md, nmax, nw);
    return align_up(l.nd * 8, 16) + align_up(l.nt * sizeof(T), 16) + align_up(l.nb, 16);
}

template <class T>
MPCQ_DEV void carve(Work<T>& w, char* base, T* l_global, int H, int ncap, bool with_md = false, int nmax = 0, int nw = 2) {
    const Layout l = layout(H, ncap, l_global == nullptr, with_md, nmax, nw);
    const int nv = l.nvc;
    double* d = reinterpret_cast<double*>(base);
    w.Md = with_md ? d : nullptr; d += with_md ? 288 : 0;
    w.GW = d; d += 72;
    w.g = d; d += nv;
    w.u = d; d += nv;
    w.gam = d; d += nv;
    w.P0 = d; d += 9 * H;        // hess_apply stage 1
    w.P1 = d; d += 12 * H;       // hess_apply stage 2; setup_model keeps its suffix sums here
    w.ucur = d; d += nv;
    w.utrial = d; d += nv;
    w.hd = d; d += nv;
    w.fmax = d; d += nv / 3 + 1;
    w.t.red = d; d += 8;
    w.t.redi = reinterpret_cast<int*>(d); d += 4;
    w.zero3 = d; d += 4;
    w.hs = d; d += 12;
    w.hq = d; d += 12;
    w.rd = d; d += 12;
    T* t = reinterpret_cast<T*>(base + align_up(l.nd * 8, 16));
    if (l_global) {
        w.L = l_global;
    } else {
        w.L = t; t += l_elems(nv);
    }
    w.dblk = t; t += 3 * nv;     // 12 values per block of 4 columns
    w.vec = t; t += nv;
    w.cw = t; t += l.cw;
    w.zt = t; t += 3 * nv;
    w.Mf = t; t += 288;
    w.r2 = t; t += 12;
    w.NS2 = t; t += 2 * H * H;
    uint8_t* b = reinterpret_cast<uint8_t*>(base + align_up(l.nd * 8, 16) + align_up(l.nt * sizeof(T), 16));
    w.fk = b; b += nv / 3 + 1;
    w.fo = b; b += nv / 3 + 1;
    w.cidx = b; b += 4 * H;
    w.face = reinterpret_cast<int8_t*>(b); b += 3 * (nv / 3 + 1);
    w.face2 = reinterpret_cast<int8_t*>(b); b += 3 * (nv / 3 + 1);
    w.facef = reinterpret_cast<int8_t*>(b); b += 3 * (nv / 3 + 1);
    w.sinf = reinterpret_cast<int32_t*>(reinterpret_cast<uintptr_t>(b + 15) & ~uintptr_t(15));
    w.H = H;
}

// ---------------------------------------------------------------------------------------------
// packed lower-triangular storage, column-major, columns grouped by 4 and starting at row 4*(j/4)
// so that (a) a lane-per-row sweep down a column is conflict-free and (b) the 4 panel rows
// L[k0..k0+3, j] are one aligned 4-vector.  element (r, j) lives at colbase(j, n) + r.
MPCQ_DEV int colbase(int j, int n) {
    int g = j >> 2, t = j & 3;
    return 4 * g * n - 8 * g * (g - 1) + t * (n - 4 * g) - 4 * g;
}

MPCQ_DEV void load4(const float* p, float& a, float& b, float& c, float& d) {
#ifdef MPCQ_HOST_EMU
    a = p[0]; b = p[1]; c = p[2]; d = p[3];
#else
    float4 v = *reinterpret_cast<const float4*>(p);
    a = v.x; b = v.y; c = v.z; d = v.w;
#endif
}
MPCQ_DEV void load4(const double* p, double& a, double& b, double& c, double& d) {
#ifdef MPCQ_HOST_EMU
    a = p[0]; b = p[1]; c = p[2]; d = p[3];
#else
    double2 v0 = *reinterpret_cast<const double2*>(p);
    double2 v1 = *reinterpret_cast<const double2*>(p + 2);
    a = v0.x; b = v0.y; c = v1.x; d = v1.y;
#endif
}

MPCQ_DEV void load2(const float* p, float& a, float& b) {
#ifdef MPCQ_HOST_EMU
    a = p[0]; b = p[1];
#else
    float2 v = *reinterpret_cast<const float2*>(p);
    a = v.x; b = v.y;
#endif
}
MPCQ_DEV void load2(const double* p, double& a, double& b) {
#ifdef MPCQ_HOST_EMU
    a = p[0]; b = p[1];
#else
    double2 v = *reinterpret_cast<const double2*>(p);
    a = v.x; b = v.y;
#endif
}

template <int V> struct IntC { static constexpr int value = V; };

template <class T, int NS> MPCQ_DEV T pick(const T (&a)[NS], int m) {
    T v = a[0];
    MPCQ_UNROLL
    for (int i = 1; i < NS; ++i) v = (i == m) ? a[i] : v;
    return v;
}

MPCQ_DEV double dmax(double a, double b) { return (a > b || a != a) ? a : b; }      // NaN propagates
MPCQ_DEV double dmin(double a, double b) { return a < b ? a : b; }
MPCQ_DEV double dabs(double a) { return a < 0 ? -a : a; }

// ---------------------------------------------------------------------------------------------
// K1 + K2 (per-env part): model matrices M00, M11, horizon table S and the linear term g.
template <class T>
MPCQ_DEV void setup_model(const Consts& cs, Work<T>& w, const T* x0p, double yaw, const T* feetp, const T* xrefp) {
    MPCQ_PHASE(0);
    const int lane = w.t.tid;
    const int H = cs.horizon;
    // --- Rz (float32-rounded like the reference), world inertia, its inverse: every lane, redundantly
    const double c = (double)(float)cos(yaw), s = (double)(float)sin(yaw);
    const double Rz[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
    double RI[9], WI[9];
    MPCQ_UNROLL
    for (int i = 0; i < 3; ++i)
        MPCQ_UNROLL
        for (int j = 0; j < 3; ++j) {
            double a = 0;
            MPCQ_UNROLL
            for (int k = 0; k < 3; ++k) a += Rz[3 * i + k] * cs.inertia[3 * k + j];
            RI[3 * i + j] = a;
        }
    MPCQ_UNROLL
    for (int i = 0; i < 3; ++i)
        MPCQ_UNROLL
        for (int j = 0; j < 3; ++j) {
            double a = 0;
            MPCQ_UNROLL
            for (int k = 0; k < 3; ++k) a += RI[3 * i + k] * Rz[3 * j + k];
            WI[3 * i + j] = (double)(float)a;           // the reference holds world_I in float32
        }
    double inv[9];
    {
        const double a = WI[0], b = WI[1], cc = WI[2], d = WI[3], e = WI[4], f = WI[5], g = WI[6], h = WI[7], i = WI[8];
        const double A = e * i - f * h, Bc = -(d * i - f * g), C = d * h - e * g;
        const double det = a * A + b * Bc + cc * C;
        const double id = 1.0 / det;
        inv[0] = A * id; inv[1] = -(b * i - cc * h) * id; inv[2] = (b * f - cc * e) * id;
        inv[3] = Bc * id; inv[4] = (a * i - cc * g) * id; inv[5] = -(a * f - cc * d) * id;
        inv[6] = C * id; inv[7] = -(a * h - b * g) * id; inv[8] = (a * e - b * d) * id;
    }
    // --- G_a = inv(world_I) [r_a]x (float32 on store, mpc.py:188) and W_a = Rz' G_a  -> GW[leg][G|W][k][y]
    if (lane < 4) {
        const int a = lane;
        const double rx = (double)feetp[3 * a], ry = (double)feetp[3 * a + 1], rz = (double)feetp[3 * a + 2];
        const double sk[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
        double G[9];
        MPCQ_UNROLL
        for (int k = 0; k < 3; ++k)
            MPCQ_UNROLL
            for (int y = 0; y < 3; ++y) {
                double acc = 0;
                MPCQ_UNROLL
                for (int l = 0; l < 3; ++l) acc += inv[3 * k + l] * sk[3 * l + y];
                G[3 * k + y] = (double)(float)acc;
            }
        MPCQ_UNROLL
        for (int k = 0; k < 3; ++k)
            MPCQ_UNROLL
            for (int y = 0; y < 3; ++y) {
                double acc = 0;
                MPCQ_UNROLL
                for (int l = 0; l < 3; ++l) acc += Rz[3 * l + k] * G[3 * l + y];   // Rz' G
                w.GW[18 * a + 3 * k + y] = G[3 * k + y];
                w.GW[18 * a + 9 + 3 * k + y] = acc;
            }
    }
    if (lane < 12) {
        w.r2[lane] = (T)(2.0 * cs.r[lane]);
        w.rd[lane] = cs.r[lane];
        // hess_apply tables (shared memory instead of per-lane indexed constant-bank reads and branches)
        w.hs[lane] = lane < 3 ? cs.q[6 + lane] : (lane < 6 ? cs.q[lane - 3] : cs.inv_mass);
        w.hq[lane] = 0.5 * ((lane >= 3 && lane < 6) ? cs.q[6 + lane] : (lane >= 9 ? cs.q[lane - 6] : 1.0));   // 1/2: NS2 holds 2N, 2S
    }
    if (lane < 4) w.zero3[lane] = 0.0;
    // --- horizon table S
    for (int idx = lane; idx < H * H; idx += w.t.nt) {
        const int i = idx / H, j = idx - i * H;
        const int m = i > j ? i : j;
        const double a = m - i + 0.5, b = m - j + 0.5, Ln = H - m;
        const double sv = Ln * a * b + (a + b) * Ln * (Ln - 1) * 0.5 + (Ln - 1) * Ln * (2 * Ln - 1) / 6.0;
        w.NS2[2 * idx] = (T)(2.0 * Ln);                     // 2 N_ij
        w.NS2[2 * idx + 1] = (T)(2.0 * sv);                 // 2 S_ij: multiples of 1/2 below 2^18, exact in float
    }
    team::sync(w.t);
    // --- M00 = B0'QB0, M11 = B1'QB1
    const double dt = cs.dt, dt2 = dt * dt, dt4 = dt2 * dt2, im2 = cs.inv_mass * cs.inv_mass;
    for (int idx = lane; idx < 144; idx += w.t.nt) {
        const int r = idx / 12, cc2 = idx - r * 12;
        const int a = r / 3, x = r - 3 * a, b = cc2 / 3, y = cc2 - 3 * b;
        double m0 = 0, m1 = 0;
        MPCQ_UNROLL
        for (int k = 0; k < 3; ++k) {
            m0 += w.GW[18 * a + 3 * k + x] * cs.q[6 + k] * w.GW[18 * b + 3 * k + y];
            m1 += w.GW[18 * a + 9 + 3 * k + x] * cs.q[k] * w.GW[18 * b + 9 + 3 * k + y];
        }
        if (x == y) { m0 += cs.q[9 + x] * im2; m1 += cs.q[3 + x] * im2; }
        m0 *= dt2; m1 *= dt4;
        if (w.Md) { w.Md[idx] = m0; w.Md[144 + idx] = m1; }
        w.Mf[idx] = (T)m0; w.Mf[144 + idx] = (T)m1;
    }
    // --- suffix sums E0, E1 of Q e_k (free response minus reference), one lane per state component
    if (lane < 12) {
        const int cidx = lane;
        const double xc = (double)x0p[cidx];
        double acx = 0, ac2x = 0;
        if (cidx < 3) {
            const double w0 = (double)x0p[6], w1 = (double)x0p[7], w2 = (double)x0p[8];
            acx = Rz[cidx] * w0 + Rz[3 + cidx] * w1 + Rz[6 + cidx] * w2;      // (Rz' omega)[c]
        } else if (cidx < 6) {
            acx = (double)x0p[6 + cidx];                                       // v
        } else if (cidx == 11) {
            acx = (double)x0p[12];                                             // -g
        }
        if (cidx == 5) ac2x = (double)x0p[12];
        double E0 = 0, E1 = 0;
        for (int j = H - 1; j >= 0; --j) {
            const double t = (j + 1) * dt;
            const double qe = cs.q[cidx] * (xc + t * acx + 0.5 * t * t * ac2x - (double)xrefp[13 * j + cidx]);
            E1 = E1 + E0 + 0.5 * qe;
            E0 = E0 + qe;
            w.P1[12 * j + cidx] = cidx < 6 ? E1 : E0;           // only E1[0:6] and E0[6:12] are used below
        }
    }
    team::sync(w.t);
    for (int v = lane; v < w.nv; v += w.t.nt) {
        const int p = v / 3, y = v - 3 * p, j = w.fk[p] >> 2, a = w.fk[p] & 3;
        const double* E0 = w.P1 + 12 * j;                       // components 6..11
        const double* E1 = w.P1 + 12 * j;                       // components 0..5
        double t0 = cs.inv_mass * E0[9 + y], t1 = cs.inv_mass * E1[3 + y];
        MPCQ_UNROLL
        for (int k = 0; k < 3; ++k) {
            t0 += w.GW[18 * a + 3 * k + y] * E0[6 + k];
            t1 += w.GW[18 * a + 9 + 3 * k + y] * E1[k];
        }
        w.g[3 * w.fo[p] + y] = 2.0 * (dt * t0 + dt2 * t1);
    }
    team::sync(w.t);
}

// gam = H u + g in fp64 through the factored structure H = 2 (N (x) B0'QB0 + S (x) B1'QB1 + R):
//   1. y_i = (G u_i, W u_i, sum_legs u_i / m)  per step      2. mix over steps with N and S
//   3. project back with G', W'.   u, gam in full [H][12] layout; P0 / P1 are scratch.
template <class T>
MPCQ_DEV void hess_apply(const Consts& cs, Work<T>& w, const double* uin, double* out, bool add_g) {
    MPCQ_PHASE(1);
    const int lane = w.t.tid;
    const int H = cs.horizon;
    for (int idx = lane; idx < 9 * H; idx += w.t.nt) {
        const int i = idx / 9, k = idx - 9 * i;
        // y = scale * sum_legs coef . u_leg with coef = row k of G (k < 3) / of W (k < 6) / the unit vector e_{k-6};
        // written with selects instead of per-lane branches (k differs from lane to lane)
        const bool rot = k < 6;
        const double* gw = w.GW + wp::sel(k < 3, 3 * k, wp::sel(rot, 9 + 3 * (k - 3), 0));
        const double e0 = wp::sel(k == 6, 1.0, 0.0), e1 = wp::sel(k == 7, 1.0, 0.0), e2 = wp::sel(k == 8, 1.0, 0.0);
        double y = 0;
        MPCQ_UNROLL
        for (int a = 0; a < 4; ++a) {
            const int s = w.cidx[4 * i + a];
            const double* ua = s != 255 ? uin + 3 * s : w.zero3;
            const double c0 = wp::sel(rot, gw[18 * a], e0), c1 = wp::sel(rot, gw[18 * a + 1], e1), c2 = wp::sel(rot, gw[18 * a + 2], e2);
            y += c0 * ua[0] + c1 * ua[1] + c2 * ua[2];
        }
        w.P0[idx] = y * w.hs[k];
    }
    team::sync(w.t);
    for (int idx = lane; idx < 12 * H; idx += w.t.nt) {
        const int j = idx / 12, c = idx - 12 * j;
        const bool useN = c < 6;
        const int src = wp::sel(c < 3, c, wp::sel(c < 6, 3 + c, c - 3));     // y0r | fs | y1r | fs
        const T* ns = w.NS2 + 2 * j + (useN ? 0 : 1);          // column j of 2N (c < 6) or of 2S
        const double* p0 = w.P0 + src;
        double acc0 = 0, acc1 = 0;
        int i = 0;
        MPCQ_NOUNROLL                                          // hess_apply is inlined a dozen times: keep each copy small (I-cache)
        for (; i + 1 < H; i += 2) {
            acc0 += (double)ns[2 * H * i] * p0[9 * i];
            acc1 += (double)ns[2 * H * (i + 1)] * p0[9 * (i + 1)];
        }
        if (i < H) acc0 += (double)ns[2 * H * i] * p0[9 * i];
        w.P1[idx] = (acc0 + acc1) * w.hq[c];
    }
    team::sync(w.t);
    const double dt2 = cs.dt * cs.dt, dt4 = dt2 * dt2;
    for (int v = lane; v < w.nv; v += w.t.nt) {
        const int p = v / 3, y = v - 3 * p, j = w.fk[p] >> 2, a = w.fk[p] & 3;
        const int o = 3 * w.fo[p] + y;
        const double* Y = w.P1 + 12 * j;
        const double* gw = w.GW + 18 * a + y;
        const double t0 = cs.inv_mass * Y[3 + y] + gw[0] * Y[0] + gw[3] * Y[1] + gw[6] * Y[2];
        const double t1 = cs.inv_mass * Y[9 + y] + gw[9] * Y[6] + gw[12] * Y[7] + gw[15] * Y[8];
        out[o] = (add_g ? w.g[o] : 0.0) + 2.0 * (w.rd[3 * a + y] * uin[o] + dt2 * t0 + dt4 * t1);
    }
    team::sync(w.t);
}

template <class T>
MPCQ_DEV void hess_apply(const Consts& cs, Work<T>& w) { hess_apply(cs, w, w.u, w.gam, true); }

// ---------------------------------------------------------------------------------------------
// K3: faces -> slot vectors z (precision T) and the face constants c written into u.
// slot v = 3p + comp of stance foot-step p; dead slots (z = 0) become identity rows.
template <class T>
MPCQ_DEV bool build_slots(const Consts& cs, Work<T>& w) {
    MPCQ_PHASE(2);
    const int lane = w.t.tid;
    const T mu = (T)cs.mu;
    bool nonzero_c = false;
    for (int idx = lane; idx < w.nv; idx += w.t.nt) w.u[idx] = 0.0;
    team::sync(w.t);
    for (int p = lane; p < w.n / 3 + 1; p += w.t.nt) {
        T z[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (p < w.ns) {
            const int sx = w.face[3 * p], sy = w.face[3 * p + 1], sz = w.face[3 * p + 2];
            if (sz >= 0) {
                if (sx == 0) z[0] = 1;
                if (sy == 0) z[4] = 1;
                if (sz == 0) { z[6] = sx * mu; z[7] = sy * mu; z[8] = 1; }
                else {
                    const double fm = w.fmax[p];
                    double* up = w.u + 3 * w.fo[p];
                    up[0] = sx * cs.mu * fm; up[1] = sy * cs.mu * fm; up[2] = fm;
                    nonzero_c = true;
                }
            }
        }
        MPCQ_UNROLL
        for (int c = 0; c < 9; ++c)
            if (3 * p + c / 3 < w.n) w.zt[9 * p + c] = z[c];
        MPCQ_UNROLL
        for (int c = 0; c < 3; ++c)
            if (3 * p + c < w.n) {
                const bool dead = z[3 * c] == 0 && z[3 * c + 1] == 0 && z[3 * c + 2] == 0;
                const int k = p < w.ns ? w.fk[p] : 0;
                w.sinf[3 * p + c] = (k >> 2) | ((k & 3) << 8) | (p << 16) | (dead ? (1 << 30) : 0);
            }
    }
    team::sync(w.t);
    return team::any(w.t, nonzero_c);
}

// ---------------------------------------------------------------------------------------------
// K2 + K4a: assemble K = Z'HZ column panel by column panel (never materialised) and factor it.
// Left-looking, 4-column panels, executed by the whole team (NW warps).  Rows are dealt to threads in blocks of
// 4, round-robin over the warps:   row = 4 * (NW * (8 * slot + lane / 4) + warp) + lane % 4,
// so that every warp keeps a share of the rows that are still active as the panel index grows (for NW = 1 this
// is row = lane + 32 * slot).  Returns false on a bad pivot.  The inner sweeps are unpredicated: rows above the
// panel / beyond n compute garbage that is never stored.
template <class T, int NCAP, int NW>
MPCQ_DEV bool chol_factor(const Consts& cs, Work<T>& w, int k_start) {
    MPCQ_PHASE(3);
    constexpr int NSLOT = NCAP / (32 * NW);
    constexpr int RSTEP = 32 * NW;                 // rows between two slots of a thread
    const int lane = wp::lane(), wid = w.t.wid;
    const int n = w.n, H = cs.horizon;
    T* L = w.L;
    bool ok = true;
    const int row0 = 4 * (NW * (lane >> 2) + wid) + (lane & 3);
    // row data of this thread, fixed for the whole factorisation
    T zr[NSLOT][3], zq[NSLOT][3];                  // z of the row and 2 R z (the R term of a same-foot entry)
    int ri[NSLOT];
    MPCQ_UNROLL
    for (int m = 0; m < NSLOT; ++m) {
        const int v = row0 + RSTEP * m;
        const bool in = v < n;
        ri[m] = in ? w.sinf[v] : (1 << 30);
        zr[m][0] = in ? w.zt[3 * v] : (T)0;
        zr[m][1] = in ? w.zt[3 * v + 1] : (T)0;
        zr[m][2] = in ? w.zt[3 * v + 2] : (T)0;
        const T* r2 = w.r2 + 3 * ((ri[m] >> 8) & 3);
        zq[m][0] = zr[m][0] * r2[0]; zq[m][1] = zr[m][1] * r2[1]; zq[m][2] = zr[m][2] * r2[2];
    }
    // cw[c][mat][leg][x] = sum_y M_mat[3 leg + x][3 b_c + y] z_c[y] for the 4 panel columns: one (c,mat,leg) per lane of
    // warp 0, double-buffered so the next panel's table can be written while other warps still read this one
    auto write_cw = [&](int k0, T* dstbuf) {
        const int c = lane >> 3, mat = (lane >> 2) & 1, leg = lane & 3;
        const int info = w.sinf[k0 + c];
        const int b = (info >> 8) & 3;
        const T* z = w.zt + 3 * (k0 + c);
        const T z0 = z[0], z1 = z[1], z2 = z[2];
        const T* mrow = w.Mf + mat * 144 + (3 * leg) * 12 + 3 * b;
        T* dst = dstbuf + 4 * lane;
        dst[0] = mrow[0] * z0 + mrow[1] * z1 + mrow[2] * z2;
        dst[1] = mrow[12] * z0 + mrow[13] * z1 + mrow[14] * z2;
        dst[2] = mrow[24] * z0 + mrow[25] * z1 + mrow[26] * z2;
        dst[3] = 0;
    };
    if (wid == 0) write_cw(0, w.cw);
    team::sync(w.t);
    // Rows < k_start keep their factor (leading block unchanged since the previous factorisation, see
    // reorder_feet): for panels k0 < k_start only the rows >= k_start are recomputed (L21 = K21 L11^-T, using the
    // stored diagonal blocks); from k_start on it is the plain left-looking factorisation.
    // One panel.  M0 >= 0 fixes the first live row slot at compile time (slots below it hold only rows above the panel):
    // with predication instead, a one-warp team would still ISSUE the dead slot's multiply-adds - half of all
    // instructions of the later panels, and the kernel is issue-bound under load.  M0 = -1: decided at run time.
    auto panel = [&](int k0, auto m0c) {
        constexpr int M0 = decltype(m0c)::value;
        const bool keep_diag = k0 < k_start;
        const int row_lo = keep_diag ? k_start : k0;
        const T* cw = w.cw + (NW > 1 ? 128 * ((k0 >> 2) & 1) : 0);   // double-buffered for multi-warp teams only
        int ci[4];
        MPCQ_UNROLL
        for (int c = 0; c < 4; ++c) ci[c] = w.sinf[k0 + c];
        // ---- initial entries of the panel
        T acc[NSLOT][4];
        const int m0 = M0 >= 0 ? M0 : row_lo / RSTEP;   // first live row slot (compile-time in the specialised bodies)
        {
        MPCQ_PHASE(10);
        MPCQ_UNROLL
        for (int m = 0; m < NSLOT; ++m) {
            if (m < m0) continue;
            const int iv = ri[m] & 0xff, av = (ri[m] >> 8) & 3;
            const bool dead = (ri[m] >> 30) != 0;
            const int v = row0 + RSTEP * m;
            MPCQ_UNROLL
            for (int c = 0; c < 4; ++c) {
                const int jw = ci[c] & 0xff;
                T q0, q1, q2, q3, s0, s1, s2, s3;
                load4(cw + 4 * (8 * c + av), q0, q1, q2, q3);
                load4(cw + 4 * (8 * c + 4 + av), s0, s1, s2, s3);
                const T t0 = zr[m][0] * q0 + zr[m][1] * q1 + zr[m][2] * q2;
                const T t1 = zr[m][0] * s0 + zr[m][1] * s1 + zr[m][2] * s2;
                T n2w, s2w;
                load2(w.NS2 + 2 * (iv * H + jw), n2w, s2w);        // 2 N_ij, 2 S_ij
                T e = n2w * t0 + s2w * t1;
                // same foot-step (and same dead flag): R term / unit diagonal - branch-free (every lane executes a
                // divergent branch body anyway as soon as one lane takes it)
                const bool same = ((ri[m] ^ ci[c]) >> 16) == 0;
                const T* zw = w.zt + 3 * (k0 + c);
                const T rterm = zq[m][0] * zw[0] + zq[m][1] * zw[1] + zq[m][2] * zw[2];
                e += wp::sel(same, rterm, (T)0);
                e = wp::sel(same && dead && v == k0 + c, (T)1, e);
                acc[m][c] = e;
            }
        }
        }
        if (NW == 1) {                                          // one-warp team: the table of the NEXT panel is written now, off
            wp::sync();                                         // the serial end of this panel (every lane is done reading cw)
            if (k0 + 4 < n) write_cw(k0 + 4, w.cw);
        }
        // ---- left-looking update with all previous columns
        const T* colg;
        {
            MPCQ_PHASE(11);
            const T* col = L;
            int stride = n;
            // software-pipelined over groups of 4 columns: the loads of group g + 1 are issued before the multiply-adds
            // of group g (a lone warp otherwise waits out the shared-memory latency once per group)
            const int ng = k0 >> 2;
            if constexpr (NSLOT > 2) {                        // many rows per thread: enough independent work, keep registers
                for (int g = 0; g < ng; ++g) {
                    MPCQ_UNROLL
                    for (int t = 0; t < 4; ++t) {
                        T p[4];
                        load4(col + k0, p[0], p[1], p[2], p[3]);
                        MPCQ_UNROLL
                        for (int m = 0; m < NSLOT; ++m) {
                            if (m < m0) continue;
                            wp::fma4_sub(acc[m], col[row0 + RSTEP * m], p);
                        }
                        col += stride;
                    }
                    col -= 4;
                    stride -= 4;
                }
            } else {
            T pa[4][4], la[4][NSLOT], pb[4][4], lb[4][NSLOT];   // two register sets, used alternately (no copies)
            auto fetch = [&](T (&p)[4][4], T (&l)[4][NSLOT], const T* cc, int st) {
                MPCQ_UNROLL
                for (int t = 0; t < 4; ++t) {
                    load4(cc + k0, p[t][0], p[t][1], p[t][2], p[t][3]);
                    MPCQ_UNROLL
                    for (int m = 0; m < NSLOT; ++m)
                        if (m >= m0) l[t][m] = cc[row0 + RSTEP * m];
                    cc += st;
                }
            };
            auto apply = [&](const T (&p)[4][4], const T (&l)[4][NSLOT]) {
                MPCQ_UNROLL
                for (int t = 0; t < 4; ++t) {
                    MPCQ_UNROLL
                    for (int m = 0; m < NSLOT; ++m) {
                        if (m < m0) continue;
                        wp::fma4_sub(acc[m], l[t][m], p[t]);       // acc[m][c] -= l * p[c] (packed FFMA2 pairs in fp32)
                    }
                }
            };
            if (ng > 0) fetch(pa, la, col, stride);
            int g = 0;
            for (; g + 1 < ng; g += 2) {
                col += 4 * stride - 4; stride -= 4;
                fetch(pb, lb, col, stride);
                apply(pa, la);
                col += 4 * stride - 4; stride -= 4;
                if (g + 2 < ng) fetch(pa, la, col, stride);
                apply(pb, lb);
            }
            if (g < ng) {
                apply(pa, la);
                col += 4 * stride - 4; stride -= 4;
            }
            }
            colg = col;                                       // == L + colbase(k0, n)
        }
        // ---- 4x4 diagonal block: kept, or factored by the warp that owns its rows and published through dblk
        T m10 = 0, m20 = 0, m21 = 0, m30 = 0, m31 = 0, m32 = 0, m00 = 0, m11 = 0, m22 = 0, m33 = 0;
        if (!keep_diag) {
            MPCQ_PHASE(12);
            const int kb = k0 >> 2, q = kb / NW;
            if (wid == kb - q * NW) {
                const int ld = 4 * (q & 7), md = q >> 3;
                T a0 = acc[0][0], a1 = acc[0][1], a2 = acc[0][2], a3 = acc[0][3];
                MPCQ_UNROLL
                for (int m = 1; m < NSLOT; ++m)
                    if (m == md) { a0 = acc[m][0]; a1 = acc[m][1]; a2 = acc[m][2]; a3 = acc[m][3]; }
                const T d00 = wp::shfl(a0, ld);
                const T d10 = wp::shfl(a0, ld + 1), d11 = wp::shfl(a1, ld + 1);
                const T d20 = wp::shfl(a0, ld + 2), d21 = wp::shfl(a1, ld + 2), d22 = wp::shfl(a2, ld + 2);
                const T d30 = wp::shfl(a0, ld + 3), d31 = wp::shfl(a1, ld + 3), d32 = wp::shfl(a2, ld + 3), d33 = wp::shfl(a3, ld + 3);
                T piv = d00;
                ok = ok && (piv > (T)0);
                const T i0 = wp::rsqrt_(piv > (T)0 ? piv : (T)1);
                const T l10 = d10 * i0, l20 = d20 * i0, l30 = d30 * i0;
                piv = d11 - l10 * l10;
                ok = ok && (piv > (T)0);
                const T i1 = wp::rsqrt_(piv > (T)0 ? piv : (T)1);
                const T l21 = (d21 - l20 * l10) * i1, l31 = (d31 - l30 * l10) * i1;
                piv = d22 - l20 * l20 - l21 * l21;
                ok = ok && (piv > (T)0);
                const T i2 = wp::rsqrt_(piv > (T)0 ? piv : (T)1);
                const T l32 = (d32 - l30 * l20 - l31 * l21) * i2;
                piv = d33 - l30 * l30 - l31 * l31 - l32 * l32;
                ok = ok && (piv > (T)0);
                const T i3 = wp::rsqrt_(piv > (T)0 ? piv : (T)1);
                // dblk holds the INVERSE M of the 4x4 diagonal block (lower triangular): the block solves of the
                // panel and of the triangular sweeps become 4 independent dot products instead of a 4-deep chain.
                // Every lane computes it (the pivots were broadcast by the shuffles), lane 0 stores it for later sweeps;
                // a one-warp team goes on with the register copy instead of a store / barrier / load round trip.
                m10 = -(l10 * i0) * i1; m21 = -(l21 * i1) * i2; m32 = -(l32 * i2) * i3;
                m20 = -(l20 * i0 + l21 * m10) * i2; m31 = -(l31 * i1 + l32 * m21) * i3;
                m30 = -(l30 * i0 + l31 * m10 + l32 * m20) * i3;
                m00 = i0; m11 = i1; m22 = i2; m33 = i3;
                if (lane == 0) {
                    T* db = w.dblk + 3 * k0;          // 12 values per block of 4 columns
                    db[0] = m10; db[1] = m20; db[2] = m21; db[3] = m30;
                    db[4] = m31; db[5] = m32; db[6] = i0; db[7] = i1;
                    db[8] = i2; db[9] = i3; db[10] = 0; db[11] = 0;
                }
            }
            if (NW > 1) team::sync(w.t);                        // dblk visible to the team
        }
        MPCQ_PHASE(13);
        if (NW > 1 || keep_diag) {
            T pad0, pad1;
            const T* db = w.dblk + 3 * k0;
            load4(db, m10, m20, m21, m30);
            load4(db + 4, m31, m32, m00, m11);
            load4(db + 8, m22, m33, pad0, pad1);
        }
        // ---- panel rows: x = acc * inv(Ld)' = (M acc')'
        {
            const int stride = n - k0;
            T* c0 = const_cast<T*>(colg);
            MPCQ_UNROLL
            for (int m = 0; m < NSLOT; ++m) {
                const int v = row0 + RSTEP * m;
                if (m < m0) continue;
                const int dv = v - k0;
                const T x0 = m00 * acc[m][0];
                const T x1 = wp::sel(dv >= 1, m10 * acc[m][0] + m11 * acc[m][1], (T)0);
                const T x2 = wp::sel(dv >= 2, (m20 * acc[m][0] + m21 * acc[m][1]) + m22 * acc[m][2], (T)0);
                const T x3 = wp::sel(dv >= 3, (m30 * acc[m][0] + m31 * acc[m][1]) + (m32 * acc[m][2] + m33 * acc[m][3]), (T)0);
                if (v >= row_lo && v < n) {
                    c0[v] = x0;
                    c0[stride + v] = x1;
                    c0[2 * stride + v] = x2;
                    c0[3 * stride + v] = x3;
                }
            }
        }
        if (NW > 1 && wid == 0 && k0 + 4 < n) write_cw(k0 + 4, w.cw + 128 * (((k0 >> 2) + 1) & 1));
        team::sync(w.t);                                        // panel columns + next cw visible
    };
    for (int k0 = 0; k0 < n; k0 += 4) {
        const int row_lo = k0 < k_start ? k_start : k0;
        if constexpr (NSLOT == 2) {
            if (row_lo >= RSTEP) panel(k0, IntC<1>{}); else panel(k0, IntC<0>{});
        } else {
            panel(k0, IntC<-1>{});
        }
    }
    return team::any(w.t, !ok) == false;
}

// K4b: solve L L' x = vec in place (vec in shared memory, precision T).  One warp; lane owns rows lane + 32 m.
// The row updates are written branch-free (selects and zero-masked loads): per-lane if/else chains cost a divergent
// branch region per slot and block (measured 400 cycles per 4-column block before, see profiles/r01_phase_cycles_*).
template <class T, int NSLOT>
MPCQ_DEV void tri_solve(Work<T>& w) {
    MPCQ_PHASE(4);
    const int lane = wp::lane();
    const int n = w.n;
    const T* L = w.L;
    T bv[NSLOT];
    int cbv[NSLOT];
    MPCQ_UNROLL
    for (int m = 0; m < NSLOT; ++m) {
        const int v = lane + 32 * m;
        bv[m] = v < n ? w.vec[v] : (T)0;
        cbv[m] = v < n ? colbase(v, n) : 0;
    }
    // One 4-column block of the forward / backward sweep.  MD = the slot that holds the block's rows when known at compile
    // time (NSLOT == 2: slots below MD in the forward sweep / above MD in the backward sweep hold no row that still
    // changes, and are not even issued); MD = -1: generic.
    auto fwd = [&](int k0, auto mdc) {
        constexpr int MD = decltype(mdc)::value;
        const int md = MD >= 0 ? MD : (k0 >> 5), ld = k0 & 31;
        const T mine = MD >= 0 ? bv[MD >= 0 ? MD : 0] : pick<T, NSLOT>(bv, md);
        const T b0 = wp::shfl(mine, ld), b1 = wp::shfl(mine, ld + 1), b2 = wp::shfl(mine, ld + 2), b3 = wp::shfl(mine, ld + 3);
        T m10, m20, m21, m30, m31, m32, m00, m11, m22, m33, pad0, pad1;
        const T* db = w.dblk + 3 * k0;
        load4(db, m10, m20, m21, m30);
        load4(db + 4, m31, m32, m00, m11);
        load4(db + 8, m22, m33, pad0, pad1);
        const T y0 = m00 * b0;                                  // y = M b (M = inverse of the diagonal block)
        const T y1 = m10 * b0 + m11 * b1;
        const T y2 = (m20 * b0 + m21 * b1) + m22 * b2;
        const T y3 = (m30 * b0 + m31 * b1) + (m32 * b2 + m33 * b3);
        const int stride = n - k0;
        const T* c0 = L + colbase(k0, n);
        MPCQ_UNROLL
        for (int m = (MD >= 0 ? MD : 0); m < NSLOT; ++m) {
            const int v = lane + 32 * m;
            const int dv = v - k0;
            const bool upd = dv >= 4 && v < n;
            const T l0 = upd ? c0[v] : (T)0, l1 = upd ? c0[stride + v] : (T)0;
            const T l2 = upd ? c0[2 * stride + v] : (T)0, l3 = upd ? c0[3 * stride + v] : (T)0;
            const T contrib = (l0 * y0 + l1 * y1) + (l2 * y2 + l3 * y3);
            const T yv = wp::sel(dv < 2, wp::sel(dv == 0, y0, y1), wp::sel(dv == 2, y2, y3));
            bv[m] = wp::sel((unsigned)dv < 4u, yv, bv[m] - contrib);
        }
    };
    auto bwd = [&](int k0, auto mdc) {
        constexpr int MD = decltype(mdc)::value;
        const int md = MD >= 0 ? MD : (k0 >> 5), ld = k0 & 31;
        const T mine = MD >= 0 ? bv[MD >= 0 ? MD : 0] : pick<T, NSLOT>(bv, md);
        const T b0 = wp::shfl(mine, ld), b1 = wp::shfl(mine, ld + 1), b2 = wp::shfl(mine, ld + 2), b3 = wp::shfl(mine, ld + 3);
        T m10, m20, m21, m30, m31, m32, m00, m11, m22, m33, pad0, pad1;
        const T* db = w.dblk + 3 * k0;
        load4(db, m10, m20, m21, m30);
        load4(db + 4, m31, m32, m00, m11);
        load4(db + 8, m22, m33, pad0, pad1);
        const T x3 = m33 * b3;                                  // x = M' b
        const T x2 = m22 * b2 + m32 * b3;
        const T x1 = (m11 * b1 + m21 * b2) + m31 * b3;
        const T x0 = (m00 * b0 + m10 * b1) + (m20 * b2 + m30 * b3);
        MPCQ_UNROLL
        for (int m = 0; m < (MD >= 0 ? MD + 1 : NSLOT); ++m) {
            const int v = lane + 32 * m;
            const int dv = v - k0;
            T a0, a1, a2, a3;
            load4(L + cbv[m] + k0, a0, a1, a2, a3);              // L[k0..k0+3, v]; any row >= k0 reads in-bounds garbage, masked below
            const T contrib = (a0 * x0 + a1 * x1) + (a2 * x2 + a3 * x3);
            const T xv = wp::sel(dv < 2, wp::sel(dv == 0, x0, x1), wp::sel(dv == 2, x2, x3));
            bv[m] = wp::sel(dv < 0, bv[m] - contrib, wp::sel(dv < 4, xv, bv[m]));
        }
    };
    if constexpr (NSLOT == 2) {
        const int nlo = n < 32 ? n : 32;
        for (int k0 = 0; k0 < nlo; k0 += 4) fwd(k0, IntC<0>{});
        for (int k0 = 32; k0 < n; k0 += 4) fwd(k0, IntC<1>{});
        for (int k0 = n - 4; k0 >= 32; k0 -= 4) bwd(k0, IntC<1>{});
        for (int k0 = nlo - 4; k0 >= 0; k0 -= 4) bwd(k0, IntC<0>{});
    } else {
        for (int k0 = 0; k0 < n; k0 += 4) fwd(k0, IntC<-1>{});
        for (int k0 = n - 4; k0 >= 0; k0 -= 4) bwd(k0, IntC<-1>{});
    }
    MPCQ_UNROLL
    for (int m = 0; m < NSLOT; ++m) {
        const int v = lane + 32 * m;
        if (v < n) w.vec[v] = bv[m];
    }
    wp::sync();
}

// ---------------------------------------------------------------------------------------------
// -z_v' gam_p for slot v = 3p + c with z rebuilt from the face code in fp64 (the float z of the factor carries a
// float-rounded mu: good enough for a preconditioner, not for the stationarity test)
MPCQ_DEV double slot_residual(int sx, int sy, int sz, int c, const double* gp, double mu) {
    if (sz < 0) return 0.0;
    if (c == 0) return sx == 0 ? -gp[0] : 0.0;
    if (c == 1) return sy == 0 ? -gp[1] : 0.0;
    return sz == 0 ? -(sx * mu * gp[0] + sy * mu * gp[1] + gp[2]) : 0.0;
}

// the three slot residuals -Z_p' gam_p of one foot-step, branch-free (faces differ from lane to lane)
MPCQ_DEV void foot_residual(int sx, int sy, int sz, const double* gp, double mu, double& r0, double& r1, double& r2) {
    const bool live = sz >= 0;
    r0 = wp::sel(live && sx == 0, -gp[0], 0.0);
    r1 = wp::sel(live && sy == 0, -gp[1], 0.0);
    r2 = wp::sel(sz == 0, -((double)sx * mu * gp[0] + (double)sy * mu * gp[1] + gp[2]), 0.0);
}

// reduced gradient r = -Z' gam into vec (precision T); returns |r|_inf (fp64).  One lane per foot-step.
template <class T>
MPCQ_DEV double reduced_gradient(const Consts& cs, Work<T>& w) {
    MPCQ_PHASE(5);
    const int lane = w.t.tid;
    double rmax = 0;
    for (int p = lane; p < w.n / 3 + 1; p += w.t.nt) {
        double r0 = 0.0, r1 = 0.0, r2 = 0.0;
        if (p < w.ns) foot_residual(w.face[3 * p], w.face[3 * p + 1], w.face[3 * p + 2], w.gam + 3 * w.fo[p], cs.mu, r0, r1, r2);
        if (3 * p < w.n) w.vec[3 * p] = (T)r0;                   // the slots beyond 3 ns (padding of n to a multiple of 4) get 0
        if (3 * p + 1 < w.n) w.vec[3 * p + 1] = (T)r1;
        if (3 * p + 2 < w.n) w.vec[3 * p + 2] = (T)r2;
        rmax = dmax(rmax, dmax(dabs(r0), dmax(dabs(r1), dabs(r2))));
    }
    rmax = team::reduce_max(w.t, rmax);
    team::sync(w.t);
    return rmax;
}

// u += Z w  (w = vec); Z is rebuilt from the face codes in fp64 so the face equalities hold
// to double precision whatever T is
template <class T>
MPCQ_DEV void apply_step(const Consts& cs, Work<T>& w) {
    MPCQ_PHASE(6);
    const int lane = w.t.tid;
    for (int p = lane; p < w.ns; p += w.t.nt) {
        const int sx = w.face[3 * p], sy = w.face[3 * p + 1], sz = w.face[3 * p + 2];
        if (sz < 0) continue;
        double* up = w.u + 3 * w.fo[p];
        const double w0 = (double)w.vec[3 * p], w1 = (double)w.vec[3 * p + 1], w2 = (double)w.vec[3 * p + 2];
        if (sz == 0) {
            up[2] += w2;
            up[0] = sx != 0 ? sx * cs.mu * up[2] : up[0] + w0;
            up[1] = sy != 0 ? sy * cs.mu * up[2] : up[1] + w1;
        } else {
            if (sx == 0) up[0] += w0;
            if (sy == 0) up[1] += w1;
        }
    }
    team::sync(w.t);
}

// Improve u on the current faces until the reduced gradient is below tol; leaves gam = Hu+g.
// Conjugate gradients on the reduced system Z'HZ w = -Z'(Hu+g), preconditioned by the (precision-T) Cholesky
// factor, with every residual and the operator in fp64.  With an accurate factor this is one step of
// iterative refinement per iteration; with a factor that is only a rough inverse (fp32, n = 180, cond 1e6)
// plain refinement stagnates while CG still converges in a few steps.  The search direction lives in the
// full space (d = Z p, kept in utrial), so no reduced-space fp64 vectors are needed.
template <class T, int NSLOT>
MPCQ_DEV double refine(const Consts& cs, Work<T>& w, double tol_abs, bool u_is_zero, bool use_cg) {
    const int lane = w.t.tid;
    // ONE loop (one call site each for tri_solve and reduced_gradient - the kernel stalls on instruction fetch, code size
    // matters) with two phases:
    //  plain  u += Z M^-1 r, gam recomputed: with an accurate factor it gains 2-3 digits per step and is the cheapest.
    //         Intermediate active-set rounds (use_cg = false) stop there - they only need the faces right.
    //  cg     entered when the plain steps converge too slowly or run out: preconditioned CG, gam updated incrementally.
    double rmax = 0.0, prev = 0.0, rz = 0.0;
    double* d = w.utrial;
    const int cap = use_cg ? cs.refine_max : 3;
    bool cg = false;
    int it = 0, itcg = 0;
    for (;;) {
        if (!cg) {
            if (it == 0 && u_is_zero) {
                for (int idx = lane; idx < w.nv; idx += w.t.nt) w.gam[idx] = w.g[idx];
                team::sync(w.t);
            } else {
                hess_apply(cs, w);
            }
        }
        rmax = reduced_gradient(cs, w);                              // r = -Z' gam -> vec
        if (!cg) {
            const bool stop = !(rmax > tol_abs && it < cap);
            if (stop || (it > 0 && rmax > 0.2 * prev)) {         // done, out of steps, or converging too slowly: hand over to CG
                if (!use_cg || rmax <= tol_abs) return rmax;
                cg = true;
            } else {
                prev = rmax;
            }
        }
        if (cg && !(rmax > tol_abs && itcg < cs.refine_max)) break;
        if (w.t.wid == 0) tri_solve<T, NSLOT>(w);               // z = M^-1 r; the serial chain runs on one warp
        team::sync(w.t);
        if (!cg) {
            apply_step(cs, w);
            ++it;
            continue;
        }
        // rz = r'z (r recomputed from gam in fp64), beta, d = Z z + beta d
        double part = 0.0;
        for (int v = lane; v < w.n; v += w.t.nt) {
            const int p = v / 3;
            if (p < w.ns)
                part += slot_residual(w.face[3 * p], w.face[3 * p + 1], w.face[3 * p + 2], v - 3 * p, w.gam + 3 * w.fo[p], cs.mu) *
                        (double)w.vec[v];
        }
        const double rz_new = team::reduce_sum(w.t, part);
        const double beta = itcg == 0 ? 0.0 : rz_new / rz;
        rz = rz_new;
        if (!(rz > 0.0)) break;                                  // converged to rounding (or a broken factor)
        for (int p = lane; p < w.ns; p += w.t.nt) {
            const int sx = w.face[3 * p], sy = w.face[3 * p + 1], sz = w.face[3 * p + 2];
            double* dp = d + 3 * w.fo[p];
            double z0 = 0.0, z1 = 0.0, z2 = 0.0;
            if (sz >= 0) {
                const double w0 = (double)w.vec[3 * p], w1 = (double)w.vec[3 * p + 1], w2 = (double)w.vec[3 * p + 2];
                if (sz == 0) { z2 = w2; z0 = sx != 0 ? sx * cs.mu * w2 : w0; z1 = sy != 0 ? sy * cs.mu * w2 : w1; }
                else { z0 = sx == 0 ? w0 : 0.0; z1 = sy == 0 ? w1 : 0.0; }
            }
            if (itcg == 0) { dp[0] = z0; dp[1] = z1; dp[2] = z2; }
            else { dp[0] = z0 + beta * dp[0]; dp[1] = z1 + beta * dp[1]; dp[2] = z2 + beta * dp[2]; }
        }
        team::sync(w.t);
        hess_apply(cs, w, d, w.hd, false);                       // hd = H d
        part = 0.0;
        for (int idx = lane; idx < w.nv; idx += w.t.nt) part += d[idx] * w.hd[idx];
        const double dHd = team::reduce_sum(w.t, part);
        if (!(dHd > 0.0)) break;
        const double alpha = rz / dHd;
        for (int idx = lane; idx < w.nv; idx += w.t.nt) {
            w.u[idx] += alpha * d[idx];
            w.gam[idx] += alpha * w.hd[idx];
        }
        team::sync(w.t);
        ++itcg;
    }
    // the face equalities hold to rounding after the updates; make them exact again (changes u by ~1 ulp)
    for (int p = lane; p < w.ns; p += w.t.nt) {
        const int sx = w.face[3 * p], sy = w.face[3 * p + 1], sz = w.face[3 * p + 2];
        double* up = w.u + 3 * w.fo[p];
        if (sz < 0) { up[0] = up[1] = up[2] = 0.0; continue; }
        if (sz > 0) up[2] = w.fmax[p];
        if (sx != 0) up[0] = sx * cs.mu * up[2];
        if (sy != 0) up[1] = sy * cs.mu * up[2];
    }
    team::sync(w.t);
    return rmax;
}

// ---------------------------------------------------------------------------------------------
// face tests of a primal-dual round: the faces every offending foot should move to are written to face2 (all feet, so
// that commit_faces() can adopt them without evaluating the tests a second time); returns the violation counts.
struct FaceCheck { int n_primal, n_dual; };

template <class T>
MPCQ_DEV void commit_faces(Work<T>& w) {
    for (int idx = w.t.tid; idx < 3 * w.ns; idx += w.t.nt) w.face[idx] = w.face2[idx];
    team::sync(w.t);
}

template <class T>
MPCQ_DEV FaceCheck pdas_update(const Consts& cs, Work<T>& w) {
    MPCQ_PHASE(7);
    const int lane = w.t.tid;
    const double mu = cs.mu;
    int npv = 0, ndv = 0;
    for (int p = lane; p < w.ns; p += w.t.nt) {
        const double* f = w.u + 3 * w.fo[p];
        const double* ga = w.gam + 3 * w.fo[p];
        const double fm = w.fmax[p];
        int sx = w.face[3 * p], sy = w.face[3 * p + 1], sz = w.face[3 * p + 2];
        const double gs = 1.0 + dmax(dabs(ga[0]), dmax(dabs(ga[1]), dabs(ga[2])));
        if (sz < 0) {
            if (ga[2] - mu * (dabs(ga[0]) + dabs(ga[1])) < -cs.tol_d * gs) {
                ++ndv;
                sx = ga[0] > 0 ? -1 : (ga[0] < 0 ? 1 : 0);
                sy = ga[1] > 0 ? -1 : (ga[1] < 0 ? 1 : 0);
                sz = 0;
            }
        } else {
            const double sc = 1.0 + dmax(dabs(f[0]), dmax(dabs(f[1]), dabs(f[2])));
            const double tp = cs.tol_p * sc;
            bool pv = false;
            int nsx = sx, nsy = sy, nsz = sz;
            if (f[2] < -tp) { nsz = -1; pv = true; }
            else if (f[2] > fm + tp) { nsz = 1; pv = true; }
            if (sx == 0) {
                if (f[0] > mu * f[2] + tp) { nsx = 1; pv = true; }
                else if (f[0] < -mu * f[2] - tp) { nsx = -1; pv = true; }
            }
            if (sy == 0) {
                if (f[1] > mu * f[2] + tp) { nsy = 1; pv = true; }
                else if (f[1] < -mu * f[2] - tp) { nsy = -1; pv = true; }
            }
            if (pv) {
                ++npv;
            } else {
                const double lx = sx != 0 ? -sx * ga[0] : 0.0;
                const double ly = sy != 0 ? -sy * ga[1] : 0.0;
                bool dv = false;
                if (sx != 0 && lx < -cs.tol_d * gs) { nsx = 0; dv = true; }
                if (sy != 0 && ly < -cs.tol_d * gs) { nsy = 0; dv = true; }
                if (sz > 0 && (-ga[2] + mu * (lx + ly)) < -cs.tol_d * gs) { nsz = 0; dv = true; }
                if (dv) ++ndv;
            }
            sx = nsx; sy = nsy; sz = nsz;
        }
        w.face2[3 * p] = (int8_t)sx; w.face2[3 * p + 1] = (int8_t)sy; w.face2[3 * p + 2] = (int8_t)sz;
    }
    FaceCheck fc;
    fc.n_primal = team::reduce_sum(w.t, npv);
    fc.n_dual = team::reduce_sum(w.t, ndv);
    team::sync(w.t);
    return fc;
}

// Stable partition of the stance list: foot-steps whose face is the one the current factor was built for
// stay in front (same relative order), the changed ones move to the back.  The leading principal block of
// K = Z'HZ - and therefore the leading columns of its Cholesky factor - is then unchanged, so the
// factorisation restarts at the first changed column instead of column 0.  Returns that column (multiple of 4).
template <class T, int NFS>
MPCQ_DEV int reorder_feet(Work<T>& w) {
    MPCQ_PHASE(8);
    const int lane = wp::lane();
    const int ns = w.ns;
    unsigned balc[NFS];
    int fc[NFS], ff[NFS], fkv[NFS], fov[NFS];
    double fm[NFS];
    int n_changed = 0, first_changed = ns;
    MPCQ_UNROLL
    for (int t = 0; t < NFS; ++t) {
        const int p = lane + 32 * t;
        const bool valid = p < ns;
        fc[t] = valid ? ((w.face[3 * p] & 0xff) | ((w.face[3 * p + 1] & 0xff) << 8) | ((w.face[3 * p + 2] & 0xff) << 16)) : 0;
        ff[t] = valid ? ((w.facef[3 * p] & 0xff) | ((w.facef[3 * p + 1] & 0xff) << 8) | ((w.facef[3 * p + 2] & 0xff) << 16)) : 0;
        fkv[t] = valid ? w.fk[p] : 0;
        fov[t] = valid ? w.fo[p] : 0;
        fm[t] = valid ? w.fmax[p] : 0.0;
        balc[t] = wp::ballot(valid && fc[t] != ff[t]);
        if (balc[t] != 0u && first_changed == ns) {
            int b = 0;
            while (!((balc[t] >> b) & 1u)) ++b;
            first_changed = 32 * t + b;
        }
        n_changed += wp::popc(balc[t]);
    }
#if defined(MPCQ_HOST_EMU) && defined(MPCQ_TRACE)
    if (lane == 0) { printf("   changed steps:"); for (int t = 0; t < NFS; ++t) for (int bb = 0; bb < 32; ++bb) if ((balc[t] >> bb) & 1u) printf(" %d", w.fk[32 * t + bb] >> 2); printf("\n"); }
#endif
    if (n_changed == 0) return w.n;
    const int n_unchanged = ns - n_changed;
    wp::sync();
    int cb_base = 0;
    MPCQ_UNROLL
    for (int t = 0; t < NFS; ++t) {
        const int p = lane + 32 * t;
        if (p < ns) {
            const int cb = cb_base + wp::popc(balc[t] & ((1u << lane) - 1u));
            const bool changed = (balc[t] >> lane) & 1u;
            const int q = changed ? n_unchanged + cb : p - cb;
            w.fk[q] = (uint8_t)fkv[t];
            w.fo[q] = (uint8_t)fov[t];
            w.fmax[q] = fm[t];
            w.face[3 * q] = (int8_t)(fc[t] & 0xff); w.face[3 * q + 1] = (int8_t)((fc[t] >> 8) & 0xff); w.face[3 * q + 2] = (int8_t)((fc[t] >> 16) & 0xff);
            w.facef[3 * q] = (int8_t)(fc[t] & 0xff); w.facef[3 * q + 1] = (int8_t)((fc[t] >> 8) & 0xff); w.facef[3 * q + 2] = (int8_t)((fc[t] >> 16) & 0xff);
        }
        cb_base += wp::popc(balc[t]);
    }
    wp::sync();
    return (3 * first_changed) & ~3;
}

// one factor-and-solve on the current faces: u = argmin on the faces (to tol), gam = Hu+g
template <class T, int NCAP, int NW>
MPCQ_DEV bool face_solve(const Consts& cs, Work<T>& w, double tol_abs, double& rmax, bool use_cg, bool same_factor) {
    constexpr int NFS = (NCAP / 3 + 31) / 32;
    bool ok = true, cnz = true;
    if (!same_factor) {
    int k_start = 0;
    if (w.t.wid == 0) k_start = reorder_feet<T, NFS>(w);        // one warp permutes the stance list ...
    k_start = team::bcast(w.t, k_start);                           // ... and the team learns the restart column
#if defined(MPCQ_HOST_EMU) && defined(MPCQ_TRACE)
    if (w.t.tid == 0) printf("   face_solve: k_start %d of n %d\n", k_start, w.n);
#endif
    cnz = build_slots(cs, w);
    ok = k_start < w.n ? chol_factor<T, NCAP, NW>(cs, w, k_start) : true;
    }
    // same_factor: faces and factor are those of the previous call and u is its solution - only improve it (to a tighter tol)
    rmax = refine<T, NCAP / 32>(cs, w, tol_abs, !cnz, use_cg);
    return ok;
}

// ---------------------------------------------------------------------------------------------
// fallback: feasible primal active-set method started from the clamped last iterate.
// The start point and its faces are derived from the point alone: every foot is moved into K and
// a row is taken active when the point sits on it (or beyond it) within the primal tolerance.
template <class T>
MPCQ_DEV void clamp_into(const Consts& cs, Work<T>& w, const double* src, double* dst, int8_t* fdst) {
    const int lane = w.t.tid;
    const double mu = cs.mu;
    for (int idx = lane; idx < w.nv; idx += w.t.nt) dst[idx] = 0.0;
    team::sync(w.t);
    for (int p = lane; p < w.ns; p += w.t.nt) {
        const double* f = src + 3 * w.fo[p];
        double* o = dst + 3 * w.fo[p];
        const double fm = w.fmax[p];
        const double tp = cs.tol_p * (1.0 + dmax(dabs(f[0]), dmax(dabs(f[1]), dabs(f[2]))));
        int sx = 0, sy = 0, sz = 0;
        double fz = f[2];
        if (fz >= fm - tp) { fz = fm; sz = 1; }
        if (!(fz > tp)) {
            sz = -1;
            o[0] = o[1] = o[2] = 0.0;
        } else {
            double fx = f[0], fy = f[1];
            if (fx >= mu * fz - tp) { fx = mu * fz; sx = 1; } else if (fx <= -mu * fz + tp) { fx = -mu * fz; sx = -1; }
            if (fy >= mu * fz - tp) { fy = mu * fz; sy = 1; } else if (fy <= -mu * fz + tp) { fy = -mu * fz; sy = -1; }
            o[0] = fx; o[1] = fy; o[2] = fz;
        }
        fdst[3 * p] = (int8_t)sx; fdst[3 * p + 1] = (int8_t)sy; fdst[3 * p + 2] = (int8_t)sz;
    }
    team::sync(w.t);
}

// phi(v) = 1/2 v'Hv + g'v = 1/2 v'(gam + g) for the vector v currently in w.u with w.gam = Hv + g
template <class T>
MPCQ_DEV double objective(const Consts& cs, Work<T>& w) {
    double s = 0.0;
    for (int idx = w.t.tid; idx < w.nv; idx += w.t.nt) s += w.u[idx] * (w.gam[idx] + w.g[idx]);
    return 0.5 * team::reduce_sum(w.t, s);
}

// objective of an arbitrary vector v: 1/2 v'(Hv + g + g); uses hd as scratch
template <class T>
MPCQ_DEV double objective_of(const Consts& cs, Work<T>& w, const double* v) {
    hess_apply(cs, w, v, w.hd, true);
    double s = 0.0;
    for (int idx = w.t.tid; idx < w.nv; idx += w.t.nt) s += v[idx] * (w.hd[idx] + w.g[idx]);
    return 0.5 * team::reduce_sum(w.t, s);
}

MPCQ_DEV void row_slacks(const double* f, double mu, double fm, double (&s)[6]) {
    s[0] = f[0] + mu * f[2]; s[1] = -f[0] + mu * f[2];
    s[2] = f[1] + mu * f[2]; s[3] = -f[1] + mu * f[2];
    s[4] = f[2]; s[5] = fm - f[2];
}

// ratio test from the feasible point ucur towards the face minimiser u: largest alpha in [0,1] keeping every
// inactive row satisfied, and the (foot, row) that blocks (tag = 8*p + row, 0x7fffffff if none)
template <class T>
MPCQ_DEV void ratio_test(const Consts& cs, Work<T>& w, double& alpha, int& tag) {
    const int lane = w.t.tid;
    const double mu = cs.mu;
    alpha = 1.0;
    tag = 0x7fffffff;
    for (int p = lane; p < w.ns; p += w.t.nt) {
        const int sx = w.face[3 * p], sy = w.face[3 * p + 1], sz = w.face[3 * p + 2];
        if (sz < 0) continue;
        const double* f0 = w.ucur + 3 * w.fo[p];
        const double* f1 = w.u + 3 * w.fo[p];
        double s0[6], s1[6];
        row_slacks(f0, mu, w.fmax[p], s0);
        row_slacks(f1, mu, w.fmax[p], s1);
        const double sc = 1.0 + dmax(dabs(f1[0]), dmax(dabs(f1[1]), dabs(f1[2])));
        const bool act[6] = {sx == -1, sx == 1, sy == -1, sy == 1, false, sz == 1};
        MPCQ_UNROLL
        for (int r = 0; r < 6; ++r) {
            const double ds = s1[r] - s0[r];
            if (!act[r] && s1[r] < -cs.tol_p * sc && ds < 0) {
                const double a = dmax(0.0, s0[r] / (-ds));
                const int tg = p * 8 + r;
                if (a < alpha || (a == alpha && tg < tag)) { alpha = a; tag = tg; }
            }
        }
    }
    team::reduce_argmin(w.t, alpha, tag);
    if (tag == 0x7fffffff) alpha = 1.0;
}

// degenerate case alpha = 0: ucur cannot move because rows it already sits on would be violated.  Make ALL of
// them part of their faces at once (instead of one factorisation per row).
template <class T>
MPCQ_DEV void block_all_at_zero(const Consts& cs, Work<T>& w) {
    const int lane = w.t.tid;
    const double mu = cs.mu;
    for (int p = lane; p < w.ns; p += w.t.nt) {
        int sx = w.face[3 * p], sy = w.face[3 * p + 1], sz = w.face[3 * p + 2];
        if (sz < 0) continue;
        double* f0 = w.ucur + 3 * w.fo[p];
        const double* f1 = w.u + 3 * w.fo[p];
        double s0[6], s1[6];
        row_slacks(f0, mu, w.fmax[p], s0);
        row_slacks(f1, mu, w.fmax[p], s1);
        const double sc = 1.0 + dmax(dabs(f1[0]), dmax(dabs(f1[1]), dabs(f1[2])));
        const bool act[6] = {sx == -1, sx == 1, sy == -1, sy == 1, false, sz == 1};
        bool hit[6];
        bool any_hit = false;
        MPCQ_UNROLL
        for (int r = 0; r < 6; ++r) {
            const double ds = s1[r] - s0[r];
            hit[r] = !act[r] && s1[r] < -cs.tol_p * sc && ds < 0 && s0[r] <= 1e-12 * sc;
            any_hit = any_hit || hit[r];
        }
        if (!any_hit) continue;
        const bool apex = hit[4] || (hit[0] && (sx == 1 || hit[1])) || (hit[1] && sx == -1) ||
                          (hit[2] && (sy == 1 || hit[3])) || (hit[3] && sy == -1);
        if (hit[0]) sx = -1; else if (hit[1]) sx = 1;
        if (hit[2]) sy = -1; else if (hit[3]) sy = 1;
        if (hit[5]) sz = 1;
        if (apex || !(f0[2] > 0.0)) { sz = -1; f0[0] = f0[1] = f0[2] = 0.0; }
        else {
            if (sz == 1) f0[2] = w.fmax[p];
            if (sx != 0) f0[0] = sx * mu * f0[2];
            if (sy != 0) f0[1] = sy * mu * f0[2];
        }
        w.face[3 * p] = (int8_t)sx; w.face[3 * p + 1] = (int8_t)sy; w.face[3 * p + 2] = (int8_t)sz;
    }
    team::sync(w.t);
}

// move ucur by alpha towards u and make the blocking row part of its foot's face
template <class T>
MPCQ_DEV void blocked_step(const Consts& cs, Work<T>& w, double alpha, int tag) {
    const int lane = w.t.tid;
    for (int idx = lane; idx < w.nv; idx += w.t.nt) w.ucur[idx] += alpha * (w.u[idx] - w.ucur[idx]);
    team::sync(w.t);
    if (lane == 0) {
        const int p = tag >> 3, r = tag & 7;
        int sx = w.face[3 * p], sy = w.face[3 * p + 1], sz = w.face[3 * p + 2];
        const bool apex = (r == 4) || (r == 0 && sx == 1) || (r == 1 && sx == -1) || (r == 2 && sy == 1) || (r == 3 && sy == -1);
        if (r == 0) sx = -1; else if (r == 1) sx = 1; else if (r == 2) sy = -1; else if (r == 3) sy = 1; else if (r == 5) sz = 1;
        double* f = w.ucur + 3 * w.fo[p];
        if (apex || !(f[2] > 0.0)) { sz = -1; f[0] = f[1] = f[2] = 0.0; }
        else {                                              // put the point exactly on the new face
            if (sz == 1) f[2] = w.fmax[p];
            if (sx != 0) f[0] = sx * cs.mu * f[2];
            if (sy != 0) f[1] = sy * cs.mu * f[2];
        }
        w.face[3 * p] = (int8_t)sx; w.face[3 * p + 1] = (int8_t)sy; w.face[3 * p + 2] = (int8_t)sz;
    }
    team::sync(w.t);
}

// ---------------------------------------------------------------------------------------------
// the whole path for environment b.  NCAP = slot capacity of this size class.
template <class T, int NCAP, int NW>
MPCQ_DEV void solve_env(const Consts& cs, const IO<T>& io, int b, char* smem, T* l_global, int ns_lo, int ns_hi) {
    MPCQ_PHASE(9);
    const int nmax = (3 * (ns_hi < NCAP / 3 ? ns_hi : NCAP / 3) + 3) & ~3;     // largest system of this size class
    const int H = cs.horizon;
    Work<T> w;
    carve(w, smem, l_global, H, NCAP, false, nmax, NW);
    w.t.tid = NW == 1 ? wp::lane() : wp::team_tid();
    w.t.nt = 32 * NW;
    w.t.wid = NW == 1 ? 0 : (w.t.tid >> 5);
    const int lane = w.t.tid;                                   // thread index within the team
    constexpr int NSLOT = NCAP / 32;                            // rows per lane when ONE warp sweeps (triangular solves)
    // ---- K3a: stance list from the contact table (ub_fz = gait * fz_max > 0); every warp builds it redundantly
    // (identical values, so the concurrent writes are benign) and therefore knows ns without a broadcast
    const float* gait = io.gait + (size_t)b * 4 * H;
    int ns = 0;
    {
        const int wl = wp::lane();
        // The list is built LAST horizon step first: face changes concentrate on the early steps (measured), so the
        // volatile foot-steps end up at the back of the factor and re-factorisations restart late (reorder_feet).
        for (int k0 = 0; k0 < 4 * H; k0 += 32) {
            const int k = 4 * H - 1 - (k0 + wl);
            const double fm = k >= 0 ? (double)gait[k] * cs.fz_max : 0.0;
            const bool st = fm > 0.0;
            const unsigned bal = wp::ballot(st);
            const int pos = ns + wp::popc(bal & ((1u << wl) - 1u));
            if (w.t.wid == 0) {
                if (k >= 0) w.cidx[k] = (uint8_t)((st && pos < NCAP / 3) ? pos : 255);
                if (st && pos < NCAP / 3) {
                    w.fk[pos] = (uint8_t)k; w.fo[pos] = (uint8_t)pos; w.fmax[pos] = fm;
                    const int code = io.face_in ? io.face_in[(size_t)b * 4 * H + k] : 0;      // 0 = all free (cold start)
                    w.face[3 * pos] = (int8_t)face_decode(code); w.face[3 * pos + 1] = (int8_t)face_decode(code >> 2);
                    w.face[3 * pos + 2] = (int8_t)face_decode(code >> 4);
                    w.facef[3 * pos] = 99; w.facef[3 * pos + 1] = 99; w.facef[3 * pos + 2] = 99;
                }
            }
            ns += wp::popc(bal);
        }
    }
    if (ns < ns_lo || ns > ns_hi) return;                       // another size class owns this env
    w.ns = ns;
    w.nv = 3 * ns;
    w.n = (3 * ns + 3) & ~3;
    team::sync(w.t);
    int status = 0, nfac = 0, nas = 0;
    double rmax = 0.0, pviol = 0.0;
    // ---- non-finite inputs are flagged before any arithmetic: u = 0, MPCQ_ST_NUMERIC
    bool finite_in = true;
    {
        const T* x0p = io.x0 + (size_t)b * 13;
        const T* ftp = io.r_feet + (size_t)b * 12;
        const T* xrp = io.x_ref + (size_t)b * 13 * H;
        for (int idx = lane; idx < 13 * H; idx += w.t.nt) { const double v = (double)xrp[idx]; finite_in = finite_in && (v - v == 0.0); }
        if (lane < 13) { const double v = (double)x0p[lane]; finite_in = finite_in && (v - v == 0.0); }
        if (lane < 12) { const double v = (double)ftp[lane]; finite_in = finite_in && (v - v == 0.0); }
        for (int idx = lane; idx < 4 * H; idx += w.t.nt) { const double v = (double)gait[idx]; finite_in = finite_in && (v - v == 0.0); }
        if (io.yaw && lane == 0) { const double v = (double)io.yaw[b]; finite_in = finite_in && (v - v == 0.0); }
        finite_in = !team::any(w.t, !finite_in);
    }
    if (!finite_in) {
        status = ST_NUMERIC;
        for (int idx = lane; idx < w.nv; idx += w.t.nt) w.u[idx] = 0.0;
        team::sync(w.t);
    } else if (ns == 0) {
        status = ST_NO_STANCE | ST_VERIFIED;
        team::sync(w.t);
    } else {
        const double yaw = io.yaw ? (double)io.yaw[b] : (double)io.x0[(size_t)b * 13 + 2];
        setup_model(cs, w, io.x0 + (size_t)b * 13, yaw, io.r_feet + (size_t)b * 12, io.x_ref + (size_t)b * 13 * H);
        double gsc = 0.0;
        for (int idx = lane; idx < w.nv; idx += w.t.nt) gsc = dmax(gsc, dabs(w.g[idx]));
        gsc = 1.0 + team::reduce_max(w.t, gsc);
        // the weakest curvature of H is 2 min(R): a reduced gradient r can hide a force error of r / (2 min R)
        const double tol_tight = dmin(cs.tol_r_tight * gsc, cs.tol_r_abs);
        const double tol_loose = dmax(cs.tol_r_loose * gsc, tol_tight);
        bool numeric_ok = (gsc == gsc) && (gsc < 1e300);
        bool done = false;
        if (!numeric_ok) {                                     // nothing was solved: return zeros, flagged
            for (int idx = lane; idx < w.nv; idx += w.t.nt) w.u[idx] = 0.0;
            team::sync(w.t);
        }
        // ---- ONE loop drives both methods, so that the factor-and-solve code exists once in the kernel (it is the bulk of
        // the code and the kernel stalls on instruction fetch when warps run different copies of it):
        //   PDAS   primal-dual active-set rounds: solve on the current faces (loose tolerance), test every foot, move all
        //          offending feet at once;
        //   TIGHT  no violation left: refine on the same factor to the tight tolerance, re-test -> verified, or back to PDAS;
        //   AS     fallback after pdas_cap rounds - monotone projected active-set method.  Keeps a FEASIBLE iterate ucur
        //          whose objective strictly decreases: from the face minimiser u either (a) u is feasible -> take it and
        //          release every wrong-signed multiplier, or (b) the clamped point clamp(u) lowers the objective -> take it
        //          with the faces it lands on (many rows change at once), or (c) step to the first blocking row (ratio test).
        enum { M_PDAS = 0, M_TIGHT = 1, M_AS = 2 };
        int mode = M_PDAS, round = 0;
        double phi_cur = 0.0;
        bool need_phi = false;
        while (numeric_ok && !done) {
            if (mode == M_AS && need_phi) { phi_cur = objective_of(cs, w, w.ucur); need_phi = false; }
            const bool tight = mode != M_PDAS;
            numeric_ok = face_solve<T, NCAP, NW>(cs, w, tight ? tol_tight : tol_loose, rmax, tight, mode == M_TIGHT) && numeric_ok;
            if (mode != M_TIGHT) ++nfac;
            if (mode != M_AS) {
                const FaceCheck fc = pdas_update(cs, w);
                const bool clean = fc.n_primal == 0 && fc.n_dual == 0;
                if (clean && mode == M_PDAS) { mode = M_TIGHT; continue; }          // tighten on the same factor (CG), re-test
                if (clean) {
                    // verified only with the stationarity residual actually at tolerance (a NaN fails this test)
                    if (rmax <= 10.0 * tol_tight) done = true; else numeric_ok = false;
                    break;
                }
#if defined(MPCQ_HOST_EMU) && defined(MPCQ_TRACE)
                if (lane == 0) printf(" PDAS round %d: primal %d dual %d rmax %.2e\n", round, fc.n_primal, fc.n_dual, rmax);
#endif
                mode = M_PDAS;
                if (!numeric_ok) break;
                if (round++ < cs.pdas_cap) { commit_faces(w); continue; }
                // the rounds cycle: hand over to the monotone method, started from the clamped last iterate (clamp_into
                // derives the faces from the point)
                status |= ST_FALLBACK;
                clamp_into(cs, w, w.u, w.ucur, w.face);
                mode = M_AS; need_phi = true; nas = 1;
                continue;
            }
            // ---- AS iteration nas
            if (!numeric_ok) break;
            double alpha;
            int tag;
            ratio_test(cs, w, alpha, tag);
#if defined(MPCQ_HOST_EMU) && defined(MPCQ_TRACE)
            if (lane == 0) printf("  AS it %d: alpha %.3e tag %d phi_cur %.10e rmax %.2e\n", nas, alpha, tag, phi_cur, rmax);
#endif
            if (tag == 0x7fffffff) {                    // (a)
                for (int idx = lane; idx < w.nv; idx += w.t.nt) w.ucur[idx] = w.u[idx];
                phi_cur = objective(cs, w);
                if (!(rmax <= 10.0 * tol_tight)) { numeric_ok = false; break; }   // the factor cannot deliver the residual
                const FaceCheck fc = pdas_update(cs, w);
                commit_faces(w);
#if defined(MPCQ_HOST_EMU) && defined(MPCQ_TRACE)
                if (lane == 0) printf("     feasible minimiser: primal %d dual %d\n", fc.n_primal, fc.n_dual);
#endif
                if (fc.n_primal == 0 && fc.n_dual == 0) { done = true; break; }
            } else {
                clamp_into(cs, w, w.u, w.utrial, w.face2);
                const double phi_t = objective_of(cs, w, w.utrial);
#if defined(MPCQ_HOST_EMU) && defined(MPCQ_TRACE)
                if (lane == 0) printf("     trial phi %.10e -> %s\n", phi_t, phi_t < phi_cur - 1e-12 * dabs(phi_cur) ? "accept" : "ratio step");
#endif
                if (phi_t < phi_cur - 1e-12 * dabs(phi_cur)) {   // (b)
                    for (int idx = lane; idx < w.nv; idx += w.t.nt) w.ucur[idx] = w.utrial[idx];
                    for (int idx = lane; idx < 3 * w.ns; idx += w.t.nt) w.face[idx] = w.face2[idx];
                    team::sync(w.t);
                    phi_cur = phi_t;
                } else if (alpha <= 1e-13) {                // (c0) degenerate: no move possible, fix every such row
                    block_all_at_zero(cs, w);
                } else {                                    // (c)
                    blocked_step(cs, w, alpha, tag);
                    need_phi = true;
                }
            }
            if (++nas > cs.as_cap) break;
        }
        if (mode == M_AS && !done && numeric_ok) {          // iteration cap: return the feasible iterate
            for (int idx = lane; idx < w.nv; idx += w.t.nt) w.u[idx] = w.ucur[idx];
            team::sync(w.t);
        }
        if (done) status |= ST_VERIFIED; else status |= numeric_ok ? ST_MAXITER : ST_NUMERIC;
        if (!done && !numeric_ok) {                            // never hand out the debris of a numerical breakdown
            for (int idx = lane; idx < w.nv; idx += w.t.nt) w.u[idx] = 0.0;
            team::sync(w.t);
        }
    }
    // ---- outputs: forces, activity (on primal slack, like the oracle's kkt_report), residuals
    const double mu = cs.mu;
    for (int k = lane; k < 4 * H; k += w.t.nt) {
        const int sidx = w.cidx[k];
        double f[3] = {0.0, 0.0, 0.0};
        if (sidx != 255) { f[0] = w.u[3 * sidx]; f[1] = w.u[3 * sidx + 1]; f[2] = w.u[3 * sidx + 2]; }
        const double fm = dmax((double)gait[k] * cs.fz_max, 0.0);
        const double sc = 1.0 + dmax(dabs(f[0]), dmax(dabs(f[1]), dabs(f[2])));
        double s[6];
        row_slacks(f, mu, fm, s);
        unsigned bits = 0;
        MPCQ_UNROLL
        for (int r = 0; r < 6; ++r) {
            if (s[r] <= cs.tol_active * sc) bits |= 1u << r;
            pviol = dmax(pviol, -s[r]);
        }
        if (io.active) io.active[(size_t)b * 4 * H + k] = (uint8_t)bits;
        if (io.u_full) {
            T* uo = io.u_full + (size_t)b * 12 * H + 3 * k;
            uo[0] = (T)f[0]; uo[1] = (T)f[1]; uo[2] = (T)f[2];
        }
        if (k < 4) {
            T* fo = io.f_out + (size_t)b * 12 + 3 * k;
            fo[0] = (T)f[0]; fo[1] = (T)f[1]; fo[2] = (T)f[2];
        }
    }
    if (io.face_out) {                                          // faces of the returned point, for the next update's warm start
        uint8_t* fo = io.face_out + (size_t)b * 4 * H;
        for (int k = lane; k < 4 * H; k += w.t.nt) fo[k] = 0;
        team::sync(w.t);
        if ((status & ST_VERIFIED) && ns > 0)
            for (int p = lane; p < w.ns; p += w.t.nt)
            {
                const int sz = w.face[3 * p + 2];                // apex: sx, sy carry no information - canonical code
                fo[w.fk[p]] = (uint8_t)face_encode(sz < 0 ? 0 : w.face[3 * p], sz < 0 ? 0 : w.face[3 * p + 1], sz);
            }
    }
    pviol = team::reduce_max(w.t, pviol);
    if (lane == 0) {
        if (io.iters) { io.iters[2 * b] = nfac; io.iters[2 * b + 1] = nas; }
        if (io.resid) { io.resid[2 * b] = rmax; io.resid[2 * b + 1] = dmax(pviol, 0.0); }
        if (io.status) io.status[b] = status;
    }
}

}  // namespace mpcq
