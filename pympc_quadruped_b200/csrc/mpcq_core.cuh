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
#if defined(MPCQ_HOST_EMU)
#include <stdio.h>
#include <stdlib.h>
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
    double tol_pc, tol_dc; // the same for the precision-T ("cheap") rounds: decisions only, the fp64 tests have the last word
    double tol_r_loose, tol_r_tight;   // reduced-gradient tolerances (relative to 1 + |g|_inf)
    double tol_active;     // slack tolerance of the reported constraint activity
    double tol_r_abs;      // absolute cap of the tight reduced-gradient tolerance: 2 min(R) * (force accuracy in N)
    double tol_r_first;    // reduced-gradient tolerance of u0 = -H^-1 g (it only has to decide the first faces)
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
// The factor region is reused for P = H^-1 (full symmetric storage, leading dimension n + 1) and, above P's diagonal
// and in one extra column, the Schur complement of the active rows: (n + 1)^2 elements.
// size classes by slot capacity: most stance foot-steps of the class, rows of its largest system, leading dimension of P
// (compile-time, so that every address in the O(n^3) loops is a base register plus an immediate)
MPCQ_HD constexpr int class_ns_hi(int ncap) { return ncap <= 64 ? 20 : (ncap <= 128 ? 42 : (ncap <= 192 ? 64 : 128)); }
MPCQ_HD constexpr int class_rows(int ncap) { return (3 * class_ns_hi(ncap) + 3) & ~3; }
MPCQ_HD constexpr int class_ld(int ncap) { return class_rows(ncap) + 1; }
MPCQ_HD constexpr int ps_elems(int n) { return (((n + 1) * (n + 1) > l_elems(n) ? (n + 1) * (n + 1) : l_elems(n)) + 3) & ~3; }

template <class T> struct Work {
    // fp64
    double *Md, *GW, *g, *u, *gam, *P0, *P1, *ucur, *utrial, *hd, *fmax, *zero3;
    double *hs, *hq, *rd;  // hess_apply tables: stage-1 scales [9], stage-2 scales [12], diag(R) [12]
    // precision T
    T *L, *dblk, *vec, *cw, *Mf, *r2;        // Mf = (M00, M11) pairs; r2 = 2 * diag(R) in precision T
    T *NS2;                // [H][H][2]: (2 N_ij, 2 S_ij) pairs for the Hessian assembly
    T *stage;              // [4 n] column block of L, transposed, while P = H^-1 is formed
    T *u0f;                // unconstrained minimiser -H^-1 g (precision T)
    T *tv;                 // scratch vector of dual_apply
    T *lam;                // right-hand side / multipliers of the active rows
    T *rcf;                // active row i:  u[rc1[i]] + rcf[i] * u[rc2[i]] = b_i
    uint16_t *rc1, *rc2;
    uint16_t *rbase;       // first active row of stance foot-step p
    int32_t* sinf;         // per slot: step | leg << 8 | foot << 16 | dead << 30
    // bytes
    uint8_t *fk;           // stance list: full foot-step index k = 4*step + leg
    uint8_t *fo;           // stance index of the foot at position p (identity; the fp64 vectors g, u, gam, ucur, utrial
                           // are stored in that compact order)
    uint8_t *cidx;         // full foot-step k -> stance index, 255 = swing
    int8_t* face;          // 3 per stance foot-step
    int8_t* face2;         // trial faces
    int n, ns, H, nv;      // nv = 3 * ns: live length of the compact vectors
    int q;                 // number of active rows
    team::Ctx t;             // the team of warps that owns this environment
};

// Shared-memory layout of one team's workspace.  nvc = capacity of everything indexed by slot (the largest system of the
// size class, `nmax`, or `ncap` when not given); cw is double-buffered only for multi-warp teams.
struct Layout { size_t nd, nt, nb; int nvc, cw; };
MPCQ_HD constexpr Layout layout(int H, int ncap, bool l_in_smem, bool with_md, int nmax, int nw) {
    Layout l{};
    l.nvc = nmax > 0 ? nmax : ncap;
    l.cw = nw > 1 ? 128 : 32;                                    // scratch of the team reductions (10 per warp) and of a 4x4 diagonal block
    const size_t nv = (size_t)l.nvc;
    l.nd = (with_md ? 288 : 0) + 72 + 9 * (size_t)H + 12 * (size_t)H + 6 * nv + nv / 3 + 1 + 24 + 40;   // 24 = team scratch: 16 doubles + 16 ints (one per warp, teams of up to 16 warps)
    // (M00, M11) [288] is only read by the factorisation; u0f, tv, lam, rcf [4 nv + 8] are first written after it: one region
    const size_t shared = 4 * nv + 8 > 288 ? 4 * nv + 8 : 288;
    l.nt = (l_in_smem ? (size_t)ps_elems(l.nvc) : 0) + 3 * nv + nv + (size_t)l.cw + shared + 12 + ((2 * (size_t)H * H + 3) & ~(size_t)3) + 4 * nv;
    l.nb = (nv / 3 + 1) * 8 + 4 * (size_t)H + 16 + 4 * nv + 2 * 2 * (nv + 4) + 2 * (nv / 3 + 2) + 16;
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
    w.t.red = d; d += 16;                                  // one slot per warp of the team (up to 16 warps; the largest class runs 12)
    w.t.redi = reinterpret_cast<int*>(d); d += 8;
    w.zero3 = d; d += 4;
    w.hs = d; d += 12;
    w.hq = d; d += 12;
    w.rd = d; d += 12;
    T* t = reinterpret_cast<T*>(base + align_up(l.nd * 8, 16));
    if (l_global) {
        w.L = l_global;
    } else {
        w.L = t; t += ps_elems(nv);
    }
    w.dblk = t; t += 3 * nv;     // 12 values per block of 4 columns
    w.vec = t; t += nv;
    w.cw = t; t += l.cw;
    w.Mf = t;                    // shares its space with u0f / tv / lam / rcf (see layout)
    w.u0f = t; w.tv = t + nv; w.lam = t + 2 * nv; w.rcf = t + 3 * nv + 4;
    t += 4 * nv + 8 > 288 ? 4 * nv + 8 : 288;
    w.r2 = t; t += 12;
    w.NS2 = t; t += (2 * H * H + 3) & ~3;
    w.stage = t; t += 4 * nv;
    uint8_t* b = reinterpret_cast<uint8_t*>(base + align_up(l.nd * 8, 16) + align_up(l.nt * sizeof(T), 16));
    w.fk = b; b += nv / 3 + 1;
    w.fo = b; b += nv / 3 + 1;
    w.cidx = b; b += 4 * H;
    w.face = reinterpret_cast<int8_t*>(b); b += 3 * (nv / 3 + 1);
    w.face2 = reinterpret_cast<int8_t*>(b); b += 3 * (nv / 3 + 1);
    b = reinterpret_cast<uint8_t*>(reinterpret_cast<uintptr_t>(b + 1) & ~uintptr_t(1));
    w.rc1 = reinterpret_cast<uint16_t*>(b); b += 2 * (nv + 4);
    w.rc2 = reinterpret_cast<uint16_t*>(b); b += 2 * (nv + 4);
    w.rbase = reinterpret_cast<uint16_t*>(b); b += 2 * (nv / 3 + 2);
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

#ifdef MPCQ_HOST_EMU
// the emulation checks what the vector loads of the device require
inline void emu_check_align(const void* p, size_t a) {
    if (reinterpret_cast<uintptr_t>(p) % a) { fprintf(stderr, "mpcq_emu: MISALIGNED %zu-byte vector load at %p\n", a, p); abort(); }
}
#endif
MPCQ_DEV void load4(const float* p, float& a, float& b, float& c, float& d) {
#ifdef MPCQ_HOST_EMU
    emu_check_align(p, 16);
    a = p[0]; b = p[1]; c = p[2]; d = p[3];
#else
    float4 v = *reinterpret_cast<const float4*>(p);
    a = v.x; b = v.y; c = v.z; d = v.w;
#endif
}
MPCQ_DEV void load4(const double* p, double& a, double& b, double& c, double& d) {
#ifdef MPCQ_HOST_EMU
    emu_check_align(p, 16);
    a = p[0]; b = p[1]; c = p[2]; d = p[3];
#else
    double2 v0 = *reinterpret_cast<const double2*>(p);
    double2 v1 = *reinterpret_cast<const double2*>(p + 2);
    a = v0.x; b = v0.y; c = v1.x; d = v1.y;
#endif
}

MPCQ_DEV void load2(const float* p, float& a, float& b) {
#ifdef MPCQ_HOST_EMU
    emu_check_align(p, 8);
    a = p[0]; b = p[1];
#else
    float2 v = *reinterpret_cast<const float2*>(p);
    a = v.x; b = v.y;
#endif
}
MPCQ_DEV void load2(const double* p, double& a, double& b) {
#ifdef MPCQ_HOST_EMU
    emu_check_align(p, 16);
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
    double sy_, cy_;
    sincos(yaw, &sy_, &cy_);
    const double c = (double)(float)cy_, s = (double)(float)sy_;
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
    // --- horizon table (2 N_ij, 2 S_ij): small integers and halves, exact in float; idx / H by multiply-shift (idx < 1024)
    {
        const int rcp = (65536 + H - 1) / H;
        for (int idx = lane; idx < H * H; idx += w.t.nt) {
            const int i = (idx * rcp) >> 16, j = idx - i * H;
            const int m = i > j ? i : j;
            const float a = (float)(m - i) + 0.5f, b = (float)(m - j) + 0.5f, Ln = (float)(H - m);
            // sum_{k=0}^{Ln-1} (a + k)(b + k) = Ln a b + (a + b) Ln (Ln - 1) / 2 + (Ln - 1) Ln (2 Ln - 1) / 6
            const float sv = Ln * a * b + (a + b) * (Ln * (Ln - 1.0f) * 0.5f) + (float)(((H - m - 1) * (H - m) * (2 * (H - m) - 1)) / 6);
            w.NS2[2 * idx] = (T)(2.0f * Ln);                    // 2 N_ij
            w.NS2[2 * idx + 1] = (T)(2.0f * sv);                // 2 S_ij: multiples of 1/2 below 2^18, exact in float
        }
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
        w.Mf[2 * idx] = (T)m0; w.Mf[2 * idx + 1] = (T)m1;       // (M00, M11) pairs, row-major 12 x 12
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
        w.g[3 * p + y] = 2.0 * (dt * t0 + dt2 * t1);
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
        const int o = 3 * p + y;
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
// K2 + K4a: assemble H (stance slots only) column panel by column panel - never materialised - and factor it.
// Left-looking, 4-column panels, executed by the whole team (NW warps).  Rows are dealt to threads in blocks of
// 4, round-robin over the warps:   row = 4 * (NW * (8 * slot + lane / 4) + warp) + lane % 4,
// so that every warp keeps a share of the rows that are still active as the panel index grows (for NW = 1 this
// is row = lane + 32 * slot).  Returns false on a bad pivot.  The inner sweeps are unpredicated: rows above the
// panel / beyond n compute garbage that is never stored.  Slot v = 3 p + comp of stance foot-step p; the slots beyond
// 3 ns (padding of n to a multiple of 4) are identity rows.
// An entry is  2 N_ij M00[ra][cb] + 2 S_ij M11[ra][cb] (+ 2 R on the diagonal):  one (2N, 2S) pair and one (M00, M11) pair.
template <class T, int NCAP, int NW>
MPCQ_DEV bool chol_factor(const Consts& cs, Work<T>& w) {
    MPCQ_PHASE(3);
    constexpr int NSLOT = NCAP / (32 * NW);
    constexpr int RSTEP = 32 * NW;                 // rows between two slots of a thread
    const int lane = wp::lane(), wid = w.t.wid;
    const int n = w.n, H = cs.horizon;
    T* L = w.L;
    bool ok = true;
    // per slot: horizon step | (3 leg + comp) << 8 | dead << 30
    for (int v = w.t.tid; v < n; v += w.t.nt) {
        const int p = v / 3, c = v - 3 * p;
        const int k = p < w.ns ? w.fk[p] : 0;
        w.sinf[v] = (k >> 2) | ((3 * (k & 3) + c) << 8) | (p < w.ns ? 0 : (1 << 30));
    }
    team::sync(w.t);
    const int row0 = 4 * (NW * (lane >> 2) + wid) + (lane & 3);
    // row data of this thread, fixed for the whole factorisation
    const T* nsr[NSLOT];                           // row of the (2N, 2S) table of the slot's horizon step
    const T* mr[NSLOT];                            // row of the (M00, M11) table of the slot's force component
    T rdiag[NSLOT];                                // 2 R of the slot (1 for a padding row)
    bool dead[NSLOT];
    MPCQ_UNROLL
    for (int m = 0; m < NSLOT; ++m) {
        const int v = row0 + RSTEP * m;
        const int info = v < n ? w.sinf[v] : (1 << 30);
        const int ra = (info >> 8) & 15;
        dead[m] = (info >> 30) != 0;
        nsr[m] = w.NS2 + 2 * (info & 0xff) * H;
        mr[m] = w.Mf + 24 * ra;
        rdiag[m] = dead[m] ? (T)1 : w.r2[ra];
    }
    // One panel.  M0 >= 0 fixes the first live row slot at compile time (slots below it hold only rows above the panel):
    // with predication instead, a one-warp team would still ISSUE the dead slot's multiply-adds - half of all
    // instructions of the later panels, and the kernel is issue-bound under load.  M0 = -1: decided at run time.
    auto panel = [&](int k0, auto m0c) {
        constexpr int M0 = decltype(m0c)::value;
        const int row_lo = k0;
        int cj[4], cb[4];
        bool cdead[4];
        MPCQ_UNROLL
        for (int c = 0; c < 4; ++c) {
            const int info = w.sinf[k0 + c];
            cj[c] = 2 * (info & 0xff); cb[c] = 2 * ((info >> 8) & 15); cdead[c] = (info >> 30) != 0;
        }
        // ---- initial entries of the panel
        T acc[NSLOT][4];
        const int m0 = M0 >= 0 ? M0 : row_lo / RSTEP;   // first live row slot (compile-time in the specialised bodies)
        MPCQ_UNROLL
        for (int m = 0; m < NSLOT; ++m) {
            if (m < m0) continue;
            const int v = row0 + RSTEP * m;
            MPCQ_UNROLL
            for (int c = 0; c < 4; ++c) {
                T n2w, s2w, t0, t1;
                load2(nsr[m] + cj[c], n2w, s2w);                   // 2 N_ij, 2 S_ij
                load2(mr[m] + cb[c], t0, t1);                      // M00, M11 entry of the two force components
                T e = n2w * t0 + s2w * t1;
                e = wp::sel(dead[m] || cdead[c], (T)0, e);         // padding rows / columns: identity
                e += wp::sel(v == k0 + c, rdiag[m], (T)0);
                acc[m][c] = e;
            }
        }
        // ---- left-looking update with all previous columns
        const T* colg;
        {
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
            // two running pointers (pivot rows, own rows) stepped by the column stride: one add each per column instead of
            // an index computation per load - the multiply-add pipe also executes the integer multiply-adds
            const T* pk = L + k0;
            const T* pr = L + row0;
            int st = n;
            auto fetch = [&](T (&p)[4][4], T (&l)[4][NSLOT]) {
                MPCQ_UNROLL
                for (int t = 0; t < 4; ++t) {
                    load4(pk, p[t][0], p[t][1], p[t][2], p[t][3]);
                    MPCQ_UNROLL
                    for (int m = 0; m < NSLOT; ++m)
                        if (m >= m0) l[t][m] = pr[RSTEP * m];
                    pk += st; pr += st;
                }
                pk -= 4; pr -= 4; st -= 4;                      // next group of 4 columns: 4 rows shorter
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
            if (ng > 0) fetch(pa, la);
            int g = 0;
            for (; g + 1 < ng; g += 2) {
                fetch(pb, lb);
                apply(pa, la);
                fetch(pa, la);                                // past the last group this reads the panel's own (unwritten) columns: unused, but the body stays one basic block
                apply(pb, lb);
            }
            if (g < ng) apply(pa, la);
            col = L + colbase(k0, n);
            }
            colg = col;                                       // == L + colbase(k0, n)
        }
        // ---- 4x4 diagonal block: kept, or factored by the warp that owns its rows and published through dblk
        T m10 = 0, m20 = 0, m21 = 0, m30 = 0, m31 = 0, m32 = 0, m00 = 0, m11 = 0, m22 = 0, m33 = 0;
        {
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
        if (NW > 1) {
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
        team::sync(w.t);                                        // panel columns + next cw visible
    };
    for (int k0 = 0; k0 < n; k0 += 4) {
        if constexpr (NSLOT == 2) {
            if (k0 >= RSTEP) panel(k0, IntC<1>{}); else panel(k0, IntC<0>{});
        } else {
            panel(k0, IntC<-1>{});
        }
    }
    return team::any(w.t, !ok) == false;
}

// ---------------------------------------------------------------------------------------------
// K4b: P = H^-1 from the Cholesky factor, written over it.
//
// The factor is used exactly once per environment: every later linear solve of the active-set rounds goes through P
// (a Schur complement on the active rows, see schur_factor), so a round costs O(q^3/6 + n q) instead of a new O(n^3/3)
// factorisation.  P is formed by the backward block recurrence of L'P = L^-1 (4 columns J at a time, trailing block
// "2" = rows/columns beyond J):
//     T   = P22 L21                       (the O(n^3/3) part: one symmetric block-matvec per column block)
//     P21 = -T M,   M = L11^-1           (the stored inverse diagonal block of the factor)
//     P11 = M'(I + L21'T) M = M'M - (L21 M)'(P21)     (the second form needs no 4x4 matrix product)
// Storage: full symmetric, element (i, k) at P[k * ld + i] with ld = n + 1 (odd: a lane-per-row sweep down a column
// and the transposed store along a row are both conflict-free).  P is written over the packed factor in place:
// column block J of P lands at addresses >= j0 * ld + j0, above every factor column < j0 that is still to be read.
template <class T, int NCAP, int NW>
MPCQ_DEV void invert_factor(Work<T>& w) {
    MPCQ_PHASE(4);
    constexpr int NT = 32 * NW;
    constexpr int NSLOT = NCAP / NT;
    constexpr int ld = class_ld(NCAP);
    const int tid = w.t.tid, n = w.n;
    T* P = w.L;
    T* stage = w.stage;
    for (int j0 = n - 4; j0 >= 0; j0 -= 4) {
        // ---- stage the column block of L (rows below the diagonal block), row-major, and fetch M
        {
            const T* c0 = w.L + colbase(j0, n);
            const int stride = n - j0;
            MPCQ_UNROLL
            for (int m = 0; m < NSLOT; ++m) {
                const int r = tid + NT * m;
                if (r >= j0 + 4 && r < n) {
                    T* d = stage + 4 * r;
                    d[0] = c0[r]; d[1] = c0[stride + r]; d[2] = c0[2 * stride + r]; d[3] = c0[3 * stride + r];
                }
            }
        }
        T m10, m20, m21, m30, m31, m32, m00, m11, m22, m33, pad0, pad1;
        {
            const T* db = w.dblk + 3 * j0;
            load4(db, m10, m20, m21, m30);
            load4(db + 4, m31, m32, m00, m11);
            load4(db + 8, m22, m33, pad0, pad1);
        }
        team::sync(w.t);
        // ---- acc = -T = -(P22 L21) for the rows of this thread (rows above the trailing block compute garbage, never stored)
        T acc[NSLOT][4];
        MPCQ_UNROLL
        for (int m = 0; m < NSLOT; ++m) { acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = (T)0; }
        const int m0 = (j0 + 4) / NT;                          // slots below hold no row of the trailing block
        if constexpr (NW > 1 && NSLOT == 2) {        // (measured: no gain for the one-warp class, whose 9 warps per SM hide the pipe for each other)
            // A team of lone warps (one per scheduler): nothing else hides the shared-memory pipe, which takes 4 cycles per
            // load instruction of a warp.  Two register sets, the loads of group g + 1 issued between the multiply-adds of
            // group g, no branch inside the body (the last fetch reads one group past the end: columns of the Schur block or
            // the arrays behind the factor region, never used).
            const T* pc = P + (size_t)(j0 + 4) * ld + tid;
            const T* sg = stage + 4 * (j0 + 4);
            T la[4][4], pa[4][2], lb[4][4], pb[4][2];
            auto fetch = [&](T (&l)[4][4], T (&p)[4][2], const T* pcc, const T* sgg) {
                MPCQ_UNROLL
                for (int t = 0; t < 4; ++t) {
                    load4(sgg + 4 * t, l[t][0], l[t][1], l[t][2], l[t][3]);
                    p[t][0] = pcc[t * ld]; p[t][1] = pcc[t * ld + NT];
                }
            };
            auto apply = [&](const T (&l)[4][4], const T (&p)[4][2]) {
                MPCQ_UNROLL
                for (int t = 0; t < 4; ++t) {
                    if (m0 == 0) wp::fma4_sub(acc[0], p[t][0], l[t]);
                    wp::fma4_sub(acc[1], p[t][1], l[t]);
                }
            };
            const int ng = (n - j0 - 4) >> 2;
            fetch(la, pa, pc, sg);
            int g = 0;
            MPCQ_NOUNROLL
            for (; g + 1 < ng; g += 2) {
                fetch(lb, pb, pc + 4 * ld, sg + 16);
                apply(la, pa);
                pc += 8 * ld; sg += 32;
                fetch(la, pa, pc, sg);
                apply(lb, pb);
            }
            if (g < ng) apply(la, pa);
        } else {
            const T* pc = P + (size_t)(j0 + 4) * ld + tid;
            const T* sg = stage + 4 * (j0 + 4);
            MPCQ_UNROLL2                                         // two groups in flight: the loads of one hide behind the other's multiply-adds
            for (int k = j0 + 4; k < n; k += 4) {
                MPCQ_UNROLL
                for (int t = 0; t < 4; ++t) {
                    T l[4];
                    load4(sg + 4 * t, l[0], l[1], l[2], l[3]);
                    MPCQ_UNROLL
                    for (int m = 0; m < NSLOT; ++m) {
                        if (m < m0) continue;
                        wp::fma4_sub(acc[m], pc[t * ld + NT * m], l);
                    }
                }
                pc += 4 * ld;
                sg += 16;
            }
        }
        // ---- P21 rows: x = (-T) M, stored down the columns of J and along its rows; and the 10 sums of Y'X over the trailing
        // rows (Y = L21 M), which give P11 = M'M - Y'X without any 4x4 matrix product: every thread adds its rows, then the
        // team reduces
        T gp[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        MPCQ_UNROLL
        for (int m = 0; m < NSLOT; ++m) {
            const int r = tid + NT * m;
            if (r >= j0 + 4 && r < n) {
                const T x0 = (acc[m][0] * m00 + acc[m][1] * m10) + (acc[m][2] * m20 + acc[m][3] * m30);
                const T x1 = acc[m][1] * m11 + acc[m][2] * m21 + acc[m][3] * m31;
                const T x2 = acc[m][2] * m22 + acc[m][3] * m32;
                const T x3 = acc[m][3] * m33;
                T* pc = P + (size_t)j0 * ld + r;
                pc[0] = x0; pc[ld] = x1; pc[2 * ld] = x2; pc[3 * ld] = x3;
                T* pr = P + (size_t)r * ld + j0;
                pr[0] = x0; pr[1] = x1; pr[2] = x2; pr[3] = x3;
                T l0, l1, l2, l3;
                load4(stage + 4 * r, l0, l1, l2, l3);
                const T y0 = (l0 * m00 + l1 * m10) + (l2 * m20 + l3 * m30);
                const T y1 = l1 * m11 + l2 * m21 + l3 * m31;
                const T y2 = l2 * m22 + l3 * m32;
                const T y3 = l3 * m33;
                gp[0] += y0 * x0;
                gp[1] += y1 * x0; gp[2] += y1 * x1;
                gp[3] += y2 * x0; gp[4] += y2 * x1; gp[5] += y2 * x2;
                gp[6] += y3 * x0; gp[7] += y3 * x1; gp[8] += y3 * x2; gp[9] += y3 * x3;
            }
        }
        MPCQ_UNROLL
        for (int e = 0; e < 10; ++e) {
            MPCQ_UNROLL
            for (int sft = 16; sft > 0; sft >>= 1) gp[e] += wp::shfl_xor(gp[e], sft);
        }
        if (NW > 1) {
            T* red = w.cw;                                       // free while P is formed
            if (wp::lane() == 0) {
                MPCQ_UNROLL
                for (int e = 0; e < 10; ++e) red[10 * w.t.wid + e] = gp[e];
            }
            team::sync(w.t);
            MPCQ_UNROLL
            for (int e = 0; e < 10; ++e) {
                T sacc = red[e];
                for (int ww = 1; ww < NW; ++ww) sacc += red[10 * ww + e];
                gp[e] = sacc;
            }
        }
        // ---- P11 = M'M - Y'X (symmetric; lower entries p00 | p10 p11 | p20 p21 p22 | p30 p31 p32 p33), one thread stores it
        if (tid == 0) {
            const T p00 = ((m00 * m00 + m10 * m10) + (m20 * m20 + m30 * m30)) - gp[0];
            const T p10 = (m11 * m10 + m21 * m20 + m31 * m30) - gp[1], p11 = (m11 * m11 + m21 * m21 + m31 * m31) - gp[2];
            const T p20 = (m22 * m20 + m32 * m30) - gp[3], p21 = (m22 * m21 + m32 * m31) - gp[4], p22 = (m22 * m22 + m32 * m32) - gp[5];
            const T p30 = m33 * m30 - gp[6], p31 = m33 * m31 - gp[7], p32 = m33 * m32 - gp[8], p33 = m33 * m33 - gp[9];
            T* pd = P + (size_t)j0 * ld + j0;
            pd[0] = p00; pd[1] = p10; pd[2] = p20; pd[3] = p30;
            pd[ld] = p10; pd[ld + 1] = p11; pd[ld + 2] = p21; pd[ld + 3] = p31;
            pd[2 * ld] = p20; pd[2 * ld + 1] = p21; pd[2 * ld + 2] = p22; pd[2 * ld + 3] = p32;
            pd[3 * ld] = p30; pd[3 * ld + 1] = p31; pd[3 * ld + 2] = p32; pd[3 * ld + 3] = p33;
        }
        team::sync(w.t);
    }
}

// element (i, k) of the symmetric P, read from its lower triangle (the upper one is given to the Schur complement)
template <class T> MPCQ_DEV T psym(const T* P, int ld, int i, int k) {
    const int lo = i < k ? i : k, hi = i < k ? k : i;
    return P[lo * ld + hi];
}

// ---------------------------------------------------------------------------------------------
// K3b: the active rows of the current faces.  Row i reads  u[rc1] + rcf u[rc2] = b:
//   fx = sx mu fz -> (x, z, -sx mu, 0);  fy likewise;  fz = fmax -> (z, z, 0, fmax);  apex -> x, y, z = 0 (three rows).
// Rows are numbered foot by foot (rbase[p] = first row of foot p); lam receives rho = A u0 - b.  One warp.
template <class T, int NFS>
MPCQ_DEV int dual_rows(const Consts& cs, Work<T>& w) {
    const int lane = wp::lane();
    const T mu = (T)cs.mu;
    int base = 0;
    MPCQ_UNROLL
    for (int t = 0; t < NFS; ++t) {
        const int p = lane + 32 * t;
        const bool valid = p < w.ns;
        const int sx = valid ? w.face[3 * p] : 0, sy = valid ? w.face[3 * p + 1] : 0, sz = valid ? w.face[3 * p + 2] : 0;
        const int cnt = !valid ? 0 : (sz < 0 ? 3 : (sx != 0) + (sy != 0) + (sz > 0));
        int incl = cnt;                                          // inclusive scan over the lanes
        MPCQ_UNROLL
        for (int d = 1; d < 32; d <<= 1) {
            const int o = wp::shfl(incl, (lane - d) & 31);
            incl += lane >= d ? o : 0;
        }
        int r = base + incl - cnt;
        if (valid) {
            w.rbase[p] = (uint16_t)r;
            const int vx = 3 * p, vy = 3 * p + 1, vz = 3 * p + 2;
            const T ux = w.u0f[vx], uy = w.u0f[vy], uz = w.u0f[vz];
            if (sz < 0) {
                w.rc1[r] = (uint16_t)vx; w.rc2[r] = (uint16_t)vx; w.rcf[r] = (T)0; w.lam[r] = ux; ++r;
                w.rc1[r] = (uint16_t)vy; w.rc2[r] = (uint16_t)vy; w.rcf[r] = (T)0; w.lam[r] = uy; ++r;
                w.rc1[r] = (uint16_t)vz; w.rc2[r] = (uint16_t)vz; w.rcf[r] = (T)0; w.lam[r] = uz;
            } else {
                if (sx != 0) { w.rc1[r] = (uint16_t)vx; w.rc2[r] = (uint16_t)vz; w.rcf[r] = -(T)sx * mu; w.lam[r] = ux - (T)sx * mu * uz; ++r; }
                if (sy != 0) { w.rc1[r] = (uint16_t)vy; w.rc2[r] = (uint16_t)vz; w.rcf[r] = -(T)sy * mu; w.lam[r] = uy - (T)sy * mu * uz; ++r; }
                if (sz > 0) { w.rc1[r] = (uint16_t)vz; w.rc2[r] = (uint16_t)vz; w.rcf[r] = (T)0; w.lam[r] = uz - (T)w.fmax[p]; }
            }
        }
        base += wp::shfl(incl, 31);
    }
    return base;
}

// element (i, j), j <= i, of the Schur complement / its Cholesky factor: kept above P's diagonal, at P[(i + 1) * ld + j]
// (column i + 1 <= n exists because the region has n + 1 columns), so it costs no shared memory of its own
#define MPCQ_SA(i, j) P[((i) + 1) * ld + (j)]

// K4c: S = A P A' on the q active rows (padded with identity rows to a multiple of 4) and its Cholesky factor.
// The entries are written first (every thread its rows: no dependencies), then the factor is formed column by column,
// left-looking, one row per thread and slot.  Every thread accumulates the pivot of column j itself (it reads row j of
// the factor anyway), so a column needs no exchange - only the barrier that publishes it.  The inverse 4x4 diagonal
// blocks go to dblk (free once P is formed) for the blocked substitutions of schur_solve.  Returns false on a bad pivot.
template <class T, int NCAP, int NW>
MPCQ_DEV bool schur_factor(Work<T>& w) {
    MPCQ_PHASE(10);
    constexpr int NT = 32 * NW;
    constexpr int NSLOT = NCAP / NT;
    constexpr int ld = class_ld(NCAP);
    const int tid = w.t.tid, q = w.q, q4 = (q + 3) & ~3;
    T* P = w.L;
    const int nsl = (q4 + NT - 1) / NT;                          // live row slots
    bool ok = true;
    T* rp[NSLOT];
    MPCQ_UNROLL
    for (int m = 0; m < NSLOT; ++m) {
        if (m >= nsl) continue;
        const int i = tid + NT * m;
        const int ie = i < q4 ? i : q4 - 1;                      // rows beyond the system read (and discard) the last row
        rp[m] = &MPCQ_SA(ie, 0);
        if (i < q) {
            const int c1 = w.rc1[i], c2 = w.rc2[i];
            const T cf = w.rcf[i];
            // 4 entries per pass: all loads first, then the stores (the compiler cannot move a load of P across a store to
            // the Schur block - same array - so entry by entry every pass would wait out the shared-memory latency)
            for (int j0 = 0; j0 <= i; j0 += 4) {
                T e[4];
                MPCQ_UNROLL
                for (int t = 0; t < 4; ++t) {
                    const int j = j0 + t < q ? j0 + t : q - 1;
                    const int c1j = w.rc1[j], c2j = w.rc2[j];
                    const T cfj = w.rcf[j];
                    e[t] = (psym(P, ld, c1, c1j) + cfj * psym(P, ld, c1, c2j)) + cf * (psym(P, ld, c2, c1j) + cfj * psym(P, ld, c2, c2j));
                }
                MPCQ_UNROLL
                for (int t = 0; t < 4; ++t)
                    if (j0 + t <= i) rp[m][j0 + t] = e[t];
            }
        } else if (i < q4) {
            for (int j = 0; j < i; ++j) rp[m][j] = (T)0;
            rp[m][i] = (T)1;
        }
    }
    team::sync(w.t);
    // ---- left-looking factorisation in panels of 4 columns
    for (int k0 = 0; k0 < q4; k0 += 4) {
        // rows k0 .. k0+3 of the factor so far, transposed: stage[k] = L[k0 .. k0+3, k]  (one 4-vector per earlier column)
        for (int k = tid; k < k0; k += NT) {
            const T* c = &MPCQ_SA(k0, k);
            T* d = w.stage + 4 * k;
            d[0] = c[0]; d[1] = c[ld]; d[2] = c[2 * ld]; d[3] = c[3 * ld];
        }
        team::sync(w.t);
        T acc[NSLOT][4];
        MPCQ_UNROLL
        for (int m = 0; m < NSLOT; ++m) {
            if (m >= nsl) continue;
            const T* r = rp[m] + k0;
            acc[m][0] = r[0]; acc[m][1] = r[1]; acc[m][2] = r[2]; acc[m][3] = r[3];
        }
        for (int k = 0; k < k0; k += 4) {
            MPCQ_UNROLL
            for (int t = 0; t < 4; ++t) {
                T l[4];
                load4(w.stage + 4 * (k + t), l[0], l[1], l[2], l[3]);
                MPCQ_UNROLL
                for (int m = 0; m < NSLOT; ++m) {
                    if (m >= nsl) continue;
                    wp::fma4_sub(acc[m], rp[m][k + t], l);
                }
            }
        }
        // the 4x4 diagonal block: its owners publish it, every thread factors it (inverse M of its Cholesky factor)
        T* dg = w.cw;
        MPCQ_UNROLL
        for (int m = 0; m < NSLOT; ++m) {
            if (m >= nsl) continue;
            const int dv = tid + NT * m - k0;
            if ((unsigned)dv < 4u) { T* d = dg + 4 * dv; d[0] = acc[m][0]; d[1] = acc[m][1]; d[2] = acc[m][2]; d[3] = acc[m][3]; }
        }
        team::sync(w.t);
        T d00, d10, d11, d20, d21, d22, d30, d31, d32, d33, pad0, pad1, pad2, pad3, pad4, pad5;
        load4(dg, d00, pad0, pad1, pad2);
        load4(dg + 4, d10, d11, pad3, pad4);
        load4(dg + 8, d20, d21, d22, pad5);
        load4(dg + 12, d30, d31, d32, d33);
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
        const T m10 = -(l10 * i0) * i1, m21 = -(l21 * i1) * i2, m32 = -(l32 * i2) * i3;
        const T m20 = -(l20 * i0 + l21 * m10) * i2, m31 = -(l31 * i1 + l32 * m21) * i3;
        const T m30 = -(l30 * i0 + l31 * m10 + l32 * m20) * i3;
        if (tid == 0) {
            T* db = w.dblk + 3 * k0;
            db[0] = m10; db[1] = m20; db[2] = m21; db[3] = m30;
            db[4] = m31; db[5] = m32; db[6] = i0; db[7] = i1;
            db[8] = i2; db[9] = i3; db[10] = 0; db[11] = 0;
        }
        // panel rows: x = acc inv(Ld)'
        MPCQ_UNROLL
        for (int m = 0; m < NSLOT; ++m) {
            if (m >= nsl) continue;
            const int i = tid + NT * m, dv = i - k0;
            if (dv >= 0 && i < q4) {
                T* r = rp[m] + k0;
                r[0] = i0 * acc[m][0];
                if (dv >= 1) r[1] = m10 * acc[m][0] + i1 * acc[m][1];
                if (dv >= 2) r[2] = (m20 * acc[m][0] + m21 * acc[m][1]) + i2 * acc[m][2];
                if (dv >= 3) r[3] = (m30 * acc[m][0] + m31 * acc[m][1]) + (m32 * acc[m][2] + i3 * acc[m][3]);
            }
        }
        team::sync(w.t);
    }
    return ok;
}

// lam <- S^-1 lam through the factor: forward and backward substitution in blocks of 4 columns (inverse diagonal blocks
// from dblk: 4 independent dot products per block instead of a 4-deep chain).  One warp; lane owns rows lane + 32 m.
// Branch-free row updates (selects and masked loads), as the per-lane if/else chains cost a divergent region each.
template <class T, int NSLOT, int ld>
MPCQ_DEV void schur_solve(Work<T>& w) {
    MPCQ_PHASE(13);
    const int lane = wp::lane(), q = w.q, q4 = (q + 3) & ~3;
    const T* P = w.L;
    T bv[NSLOT];
    const T* rp[NSLOT];
    MPCQ_UNROLL
    for (int m = 0; m < NSLOT; ++m) {
        const int v = lane + 32 * m;
        bv[m] = v < q ? w.lam[v] : (T)0;
        rp[m] = &MPCQ_SA(v < q4 ? v : q4 - 1, 0);
    }
    const int nsl = (q4 + 31) >> 5;
    for (int k0 = 0; k0 < q4; k0 += 4) {                        // L y = rho
        const T mine = pick<T, NSLOT>(bv, k0 >> 5);
        const int l0 = k0 & 31;
        const T b0 = wp::shfl(mine, l0), b1 = wp::shfl(mine, l0 + 1), b2 = wp::shfl(mine, l0 + 2), b3 = wp::shfl(mine, l0 + 3);
        T m10, m20, m21, m30, m31, m32, m00, m11, m22, m33, pad0, pad1;
        const T* db = w.dblk + 3 * k0;
        load4(db, m10, m20, m21, m30);
        load4(db + 4, m31, m32, m00, m11);
        load4(db + 8, m22, m33, pad0, pad1);
        const T y0 = m00 * b0;
        const T y1 = m10 * b0 + m11 * b1;
        const T y2 = (m20 * b0 + m21 * b1) + m22 * b2;
        const T y3 = (m30 * b0 + m31 * b1) + (m32 * b2 + m33 * b3);
        MPCQ_UNROLL
        for (int m = 0; m < NSLOT; ++m) {
            if (m >= nsl) continue;
            const int dv = lane + 32 * m - k0;
            const T* r = rp[m] + k0;
            const T contrib = (r[0] * y0 + r[1] * y1) + (r[2] * y2 + r[3] * y3);
            const T yv = wp::sel(dv < 2, wp::sel(dv == 0, y0, y1), wp::sel(dv == 2, y2, y3));
            bv[m] = wp::sel((unsigned)dv < 4u, yv, wp::sel(dv >= 4, bv[m] - contrib, bv[m]));
        }
    }
    for (int k0 = q4 - 4; k0 >= 0; k0 -= 4) {                   // L' lam = y
        const T mine = pick<T, NSLOT>(bv, k0 >> 5);
        const int l0 = k0 & 31;
        const T b0 = wp::shfl(mine, l0), b1 = wp::shfl(mine, l0 + 1), b2 = wp::shfl(mine, l0 + 2), b3 = wp::shfl(mine, l0 + 3);
        T m10, m20, m21, m30, m31, m32, m00, m11, m22, m33, pad0, pad1;
        const T* db = w.dblk + 3 * k0;
        load4(db, m10, m20, m21, m30);
        load4(db + 4, m31, m32, m00, m11);
        load4(db + 8, m22, m33, pad0, pad1);
        const T x3 = m33 * b3;
        const T x2 = m22 * b2 + m32 * b3;
        const T x1 = (m11 * b1 + m21 * b2) + m31 * b3;
        const T x0 = (m00 * b0 + m10 * b1) + (m20 * b2 + m30 * b3);
        const T* c0 = &MPCQ_SA(k0, 0);                           // rows k0 .. k0+3 of the factor, read along the row
        MPCQ_UNROLL
        for (int m = 0; m < NSLOT; ++m) {
            if (m >= nsl) continue;
            const int v = lane + 32 * m;
            const int dv = v - k0;
            const int ve = dv < 0 ? v : 0;                       // rows at or beyond the block read column 0 (masked below)
            const T contrib = (c0[ve] * x0 + c0[ld + ve] * x1) + (c0[2 * ld + ve] * x2 + c0[3 * ld + ve] * x3);
            const T xv = wp::sel(dv < 2, wp::sel(dv == 0, x0, x1), wp::sel(dv == 2, x2, x3));
            bv[m] = wp::sel(dv < 0, bv[m] - contrib, wp::sel(dv < 4, xv, bv[m]));
        }
    }
    MPCQ_UNROLL
    for (int m = 0; m < NSLOT; ++m) {
        const int i = lane + 32 * m;
        if (i < q) w.lam[i] = bv[m];
    }
    wp::sync();
}

// The two uses of the Schur complement, one body (the kernel stalls on instruction fetch: every large piece exists once):
//   round   the minimiser on the current faces, precision T: rows, S and its factor, multipliers lam = S^-1 (A u0 - b),
//           u = u0 - P A' lam (made exact on the faces, stored as fp64) and gam = -A' lam.
//   !round  vec <- (P - P A' S^-1 A P) vec: the inverse of H restricted to the null space of the active rows, applied to
//           the slot-form reduced gradient (the right-hand side of the reduced system, lifted to the full space).
// Returns false on a bad pivot.
template <class T, int NCAP, int NW>
MPCQ_DEV bool dual_op(const Consts& cs, Work<T>& w, bool round) {
    MPCQ_PHASE(11);
    constexpr int NFS = (NCAP / 3 + 31) / 32;
    constexpr int ld = class_ld(NCAP);
    const int n = w.n;
    const T* P = w.L;
    bool ok = true;
    const T* base = w.u0f;
    if (round) {
        int q = 0;
        if (w.t.wid == 0) q = dual_rows<T, NFS>(cs, w);
        w.q = team::bcast(w.t, q);
        if (w.q > 0) ok = schur_factor<T, NCAP, NW>(w);
    } else {
        for (int v = w.t.tid; v < w.nv; v += w.t.nt) {
            T a0 = 0, a1 = 0;
            for (int k = 0; k < n; k += 4) {
                T r0, r1, r2, r3;
                load4(w.vec + k, r0, r1, r2, r3);
                a0 += psym(P, ld, v, k) * r0 + psym(P, ld, v, k + 2) * r2;
                a1 += psym(P, ld, v, k + 1) * r1 + psym(P, ld, v, k + 3) * r3;
            }
            w.tv[v] = a0 + a1;
        }
        team::sync(w.t);
        for (int i = w.t.tid; i < w.q; i += w.t.nt) w.lam[i] = w.tv[w.rc1[i]] + w.rcf[i] * w.tv[w.rc2[i]];
        team::sync(w.t);
        base = w.tv;
    }
    if (w.q > 0) {
        if (w.t.wid == 0) schur_solve<T, NCAP / 32, class_ld(NCAP)>(w);
        team::sync(w.t);
    }
    // vec = base - P A' lam
    {
        MPCQ_PHASE(14);
        const int q = w.q;
        for (int v = w.t.tid; v < w.nv; v += w.t.nt) {
            T a = base[v];
            for (int i = 0; i < q; ++i) {
                const T li = w.lam[i], cfi = w.rcf[i];
                const int c1i = w.rc1[i], c2i = w.rc2[i];
                a -= li * (psym(P, ld, v, c1i) + cfi * psym(P, ld, v, c2i));
            }
            w.vec[v] = a;
        }
        team::sync(w.t);
    }
    if (round) {
        const double mu = cs.mu;
        for (int p = w.t.tid; p < w.ns; p += w.t.nt) {
            const int sx = w.face[3 * p], sy = w.face[3 * p + 1], sz = w.face[3 * p + 2];
            double* up = w.u + 3 * p;
            double* gp = w.gam + 3 * p;
            int r = w.rbase[p];
            if (sz < 0) {
                up[0] = up[1] = up[2] = 0.0;
                gp[0] = -(double)w.lam[r]; gp[1] = -(double)w.lam[r + 1]; gp[2] = -(double)w.lam[r + 2];
            } else {
                double lx = 0.0, ly = 0.0, lz = 0.0;
                if (sx != 0) lx = (double)w.lam[r++];
                if (sy != 0) ly = (double)w.lam[r++];
                if (sz > 0) lz = (double)w.lam[r];
                const double fz = sz > 0 ? w.fmax[p] : (double)w.vec[3 * p + 2];
                up[2] = fz;
                up[0] = sx != 0 ? sx * mu * fz : (double)w.vec[3 * p];
                up[1] = sy != 0 ? sy * mu * fz : (double)w.vec[3 * p + 1];
                gp[0] = -lx; gp[1] = -ly; gp[2] = mu * (sx * lx + sy * ly) - lz;
            }
        }
        team::sync(w.t);
    }
    return ok;
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
        if (p < w.ns) foot_residual(w.face[3 * p], w.face[3 * p + 1], w.face[3 * p + 2], w.gam + 3 * p, cs.mu, r0, r1, r2);
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
        double* up = w.u + 3 * p;
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

// ---------------------------------------------------------------------------------------------
// face tests of a primal-dual round: the faces every offending foot should move to are written to face2 (all feet, so
// that commit_faces() can adopt them without evaluating the tests a second time); returns the violation counts.
struct FaceCheck { int n_primal, n_dual; };

template <class T>
MPCQ_DEV void commit_faces(Work<T>& w) {
    for (int idx = w.t.tid; idx < 3 * w.ns; idx += w.t.nt) w.face[idx] = w.face2[idx];
    team::sync(w.t);
}

// tol_p, tol_d: relative tolerances; tol_d_abs caps the multiplier threshold (a wrong-signed multiplier eps hides a force
// error of eps / (2 min R): the final test must not accept more than the force accuracy allows)
template <class T>
MPCQ_DEV FaceCheck pdas_update(const Consts& cs, Work<T>& w, double tol_p, double tol_d, double tol_d_abs) {
    MPCQ_PHASE(7);
    const int lane = w.t.tid;
    const double mu = cs.mu;
    int npv = 0, ndv = 0;
    for (int p = lane; p < w.ns; p += w.t.nt) {
        const double* f = w.u + 3 * p;
        const double* ga = w.gam + 3 * p;
        const double fm = w.fmax[p];
        int sx = w.face[3 * p], sy = w.face[3 * p + 1], sz = w.face[3 * p + 2];
        const double td = dmin(tol_d * (1.0 + dmax(dabs(ga[0]), dmax(dabs(ga[1]), dabs(ga[2])))), tol_d_abs);
        if (sz < 0) {
            if (ga[2] - mu * (dabs(ga[0]) + dabs(ga[1])) < -td) {
                ++ndv;
                sx = ga[0] > 0 ? -1 : (ga[0] < 0 ? 1 : 0);
                sy = ga[1] > 0 ? -1 : (ga[1] < 0 ? 1 : 0);
                sz = 0;
            }
        } else {
            const double sc = 1.0 + dmax(dabs(f[0]), dmax(dabs(f[1]), dabs(f[2])));
            const double tp = tol_p * sc;
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
                if (sx != 0 && lx < -td) { nsx = 0; dv = true; }
                if (sy != 0 && ly < -td) { nsy = 0; dv = true; }
                if (sz > 0 && (-ga[2] + mu * (lx + ly)) < -td) { nsz = 0; dv = true; }
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
        const double* f = src + 3 * p;
        double* o = dst + 3 * p;
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
        const double* f0 = w.ucur + 3 * p;
        const double* f1 = w.u + 3 * p;
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
        double* f0 = w.ucur + 3 * p;
        const double* f1 = w.u + 3 * p;
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
        double* f = w.ucur + 3 * p;
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
MPCQ_DEV void solve_env(const Consts& cs, const IO<T>& io, int b, char* smem, T* l_global, int ns_lo, int ns_hi, int team_bar = 0) {
    MPCQ_PHASE(9);
    const int nmax = (3 * (ns_hi < NCAP / 3 ? ns_hi : NCAP / 3) + 3) & ~3;     // largest system of this size class
    const int H = cs.horizon;
    Work<T> w;
    carve(w, smem, l_global, H, NCAP, false, nmax, NW);
    w.t.tid = NW == 1 ? wp::lane() : wp::team_tid(32 * NW);
    w.t.nt = 32 * NW;
    w.t.bar = team_bar;
    w.t.wid = NW == 1 ? 0 : (w.t.tid >> 5);
    const int lane = w.t.tid;                                   // thread index within the team
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
            const double fm = k >= 0 ? (double)(float)((double)gait[k] * cs.fz_max) : 0.0;   // float32 like the reference's ub (mpc.py:248-258)
            const bool st = fm > 0.0;
            const unsigned bal = wp::ballot(st);
            const int pos = ns + wp::popc(bal & ((1u << wl) - 1u));
            if (w.t.wid == 0) {
                if (k >= 0) w.cidx[k] = (uint8_t)((st && pos < NCAP / 3) ? pos : 255);
                if (st && pos < NCAP / 3) {
                    w.fk[pos] = (uint8_t)k; w.fmax[pos] = fm;
                    const int code = io.face_in ? io.face_in[(size_t)b * 4 * H + k] : 0;      // 0 = all free (cold start)
                    w.face[3 * pos] = (int8_t)face_decode(code); w.face[3 * pos + 1] = (int8_t)face_decode(code >> 2);
                    w.face[3 * pos + 2] = (int8_t)face_decode(code >> 4);
                }
            }
            ns += wp::popc(bal);
        }
    }
    if (ns < ns_lo || ns > ns_hi) return;                       // another size class owns this env
    w.ns = ns;
    w.nv = 3 * ns;
    w.n = (3 * ns + 3) & ~3;
    w.q = 0;
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
        bool numeric_ok = (gsc == gsc) && (gsc < 1e300);
        bool done = false;
        if (!numeric_ok) {                                     // nothing was solved: return zeros, flagged
            for (int idx = lane; idx < w.nv; idx += w.t.nt) w.u[idx] = 0.0;
            team::sync(w.t);
        }
        if (numeric_ok) {
            // ---- ONE factorisation per environment: H on the all-free faces, then P = H^-1 over it.  Warm-start faces wait
            // in face2 meanwhile (the factor is always that of the full Hessian).
            for (int idx = lane; idx < 3 * w.ns; idx += w.t.nt) { w.face2[idx] = w.face[idx]; w.face[idx] = 0; }
            team::sync(w.t);
            numeric_ok = chol_factor<T, NCAP, NW>(cs, w);
            invert_factor<T, NCAP, NW>(w);
            // ---- u0 = -H^-1 g, refined in fp64 through the structured operator (q = 0: the preconditioner is P itself)
            w.q = 0;
            for (int idx = lane; idx < w.nv; idx += w.t.nt) w.u[idx] = 0.0;
            team::sync(w.t);
        }
        // Rounds.  A round = the minimiser on the current faces through the Schur complement of the active rows (precision T:
        // "cheap"), the face tests on it, and - once the cheap tests pass (or always, in EXACT mode) - fp64 refinement on
        // those faces and the face tests at the tight tolerances.  PDAS moves every offending foot at once; after pdas_cap
        // rounds (the rounds can cycle) the monotone primal active-set method takes over: feasible iterate ucur, step to
        // the first blocking row (ratio test), release wrong-signed multipliers only at a feasible face minimiser.
        //
        // The refinement (plain residual correction u += P_A r while it gains >= 5x per step, preconditioned conjugate
        // gradients on the reduced system when it stagnates; operator and residuals in fp64) is part of the same flat state
        // machine, so that dual_op, hess_apply, reduced_gradient and the face tests each exist ONCE in the kernel: it stalls
        // on instruction fetch as soon as warps run different copies of the same code.
        enum { M_PDAS = 0, M_AS = 1 };
        enum { S_DUAL = 0, S_HESS = 1, S_RFDONE = 2, S_TEST = 3, S_RATIO = 4 };
        int mode = M_PDAS, round = 0, st = S_HESS;
        bool exact = false, first = true, ratio_done = false, tight = true, dual_round = false;
        // refinement state
        int rf_it = 0, rf_itcg = 0;
        bool rf_cg = false, rf_skip = true, rf_half = false;     // rf_skip: u = 0, so gam = g; rf_half: the operator was applied to the CG direction
        double rf_prev = 0.0, rf_rz = 0.0;
        // u0 is only refined to the loose tolerance: it is exact enough to decide the first faces, and in nine cases out of ten
        // it is infeasible anyway; a clean first test tightens it before it is accepted
        double rf_tol = dmax(tol_tight, cs.tol_r_first * gsc);
        double* const dcg = w.utrial;
        const double* hin = w.u;
        double* hout = w.gam;
        while (numeric_ok && !done) {
            if (st == S_DUAL) {
                numeric_ok = dual_op<T, NCAP, NW>(cs, w, dual_round) && numeric_ok;
                if (dual_round) {
                    ++nfac;
#if defined(MPCQ_HOST_EMU) && defined(MPCQ_TRACE)
                    if (lane == 0) printf(" round %d mode %d exact %d: q %d ok %d\n", round, mode, (int)exact, w.q, (int)numeric_ok);
#endif
                    if (!numeric_ok) break;
                    if (exact) { rf_it = rf_itcg = 0; rf_cg = rf_skip = rf_half = false; rf_prev = rf_rz = 0.0; rf_tol = tol_tight; hin = w.u; hout = w.gam; st = S_HESS; }
                    else if (mode == M_AS) st = S_RATIO;
                    else { tight = false; st = S_TEST; }
                    continue;
                }
                // ---- inside the refinement: vec = z = P_A r
                if (!rf_cg) {
                    apply_step(cs, w);
                    ++rf_it;
                    hin = w.u; hout = w.gam; rf_half = false;
                    st = S_HESS;
                    continue;
                }
                // rz = r'z (r recomputed from gam in fp64), beta, d = Z z + beta d
                double part = 0.0;
                for (int v = lane; v < w.n; v += w.t.nt) {
                    const int p = v / 3;
                    if (p < w.ns)
                        part += slot_residual(w.face[3 * p], w.face[3 * p + 1], w.face[3 * p + 2], v - 3 * p, w.gam + 3 * p, cs.mu) *
                                (double)w.vec[v];
                }
                const double rz_new = team::reduce_sum(w.t, part);
                const double beta = rf_itcg == 0 ? 0.0 : rz_new / rf_rz;
                rf_rz = rz_new;
                if (!(rf_rz > 0.0)) { st = S_RFDONE; continue; }    // converged to rounding (or a broken preconditioner)
                for (int p = lane; p < w.ns; p += w.t.nt) {
                    const int sx = w.face[3 * p], sy = w.face[3 * p + 1], sz = w.face[3 * p + 2];
                    double* dp = dcg + 3 * p;
                    double z0 = 0.0, z1 = 0.0, z2 = 0.0;
                    if (sz >= 0) {
                        const double w0 = (double)w.vec[3 * p], w1 = (double)w.vec[3 * p + 1], w2 = (double)w.vec[3 * p + 2];
                        if (sz == 0) { z2 = w2; z0 = sx != 0 ? sx * cs.mu * w2 : w0; z1 = sy != 0 ? sy * cs.mu * w2 : w1; }
                        else { z0 = sx == 0 ? w0 : 0.0; z1 = sy == 0 ? w1 : 0.0; }
                    }
                    if (rf_itcg == 0) { dp[0] = z0; dp[1] = z1; dp[2] = z2; }
                    else { dp[0] = z0 + beta * dp[0]; dp[1] = z1 + beta * dp[1]; dp[2] = z2 + beta * dp[2]; }
                }
                team::sync(w.t);
                hin = dcg; hout = w.hd; rf_half = true;
                st = S_HESS;
                continue;
            }
            if (st == S_HESS) {
                if (rf_skip) {
                    for (int idx = lane; idx < w.nv; idx += w.t.nt) w.gam[idx] = w.g[idx];
                    team::sync(w.t);
                    rf_skip = false;
                } else {
                    hess_apply(cs, w, hin, hout, !rf_half);
                }
                if (rf_half) {                                   // hd = H d: finish the CG step
                    double part = 0.0;
                    for (int idx = lane; idx < w.nv; idx += w.t.nt) part += dcg[idx] * w.hd[idx];
                    const double dHd = team::reduce_sum(w.t, part);
                    if (!(dHd > 0.0)) { st = S_RFDONE; continue; }
                    const double alpha = rf_rz / dHd;
                    for (int idx = lane; idx < w.nv; idx += w.t.nt) {
                        w.u[idx] += alpha * dcg[idx];
                        w.gam[idx] += alpha * w.hd[idx];
                    }
                    team::sync(w.t);
                    ++rf_itcg;
                }
                rmax = reduced_gradient(cs, w);                      // r = -Z' gam -> vec
#if defined(MPCQ_HOST_EMU) && defined(MPCQ_TRACE)
                if (lane == 0) printf("     refine it %d cg %d(%d) rmax %.3e (tol %.3e) q %d\n", rf_it, (int)rf_cg, rf_itcg, rmax, rf_tol, w.q);
#endif
                bool fin = false;
                if (!rf_cg) {
                    const bool stop = !(rmax > rf_tol && rf_it < cs.refine_max);
                    if (stop || (rf_it > 0 && rmax > 0.2 * rf_prev)) {   // done, out of steps, or converging too slowly: hand over to CG
                        if (rmax <= rf_tol) fin = true; else rf_cg = true;
                    } else {
                        rf_prev = rmax;
                    }
                }
                if (rf_cg && !(rmax > rf_tol && rf_itcg < cs.refine_max)) fin = true;
                if (fin) { st = S_RFDONE; continue; }
                dual_round = false;
                st = S_DUAL;
                continue;
            }
            if (st == S_RFDONE) {
                if (rf_cg) {                                     // the face equalities hold to rounding after CG updates; make them exact again
                    for (int p = lane; p < w.ns; p += w.t.nt) {
                        const int sx = w.face[3 * p], sy = w.face[3 * p + 1], sz = w.face[3 * p + 2];
                        double* up = w.u + 3 * p;
                        if (sz < 0) { up[0] = up[1] = up[2] = 0.0; continue; }
                        if (sz > 0) up[2] = w.fmax[p];
                        if (sx != 0) up[0] = sx * cs.mu * up[2];
                        if (sy != 0) up[1] = sy * cs.mu * up[2];
                    }
                    team::sync(w.t);
                }
                if (first) {                                     // this was u0 = -H^-1 g
                    for (int idx = lane; idx < w.nv; idx += w.t.nt) w.u0f[idx] = (T)w.u[idx];
                    bool any_warm = false;
                    for (int idx = lane; idx < 3 * w.ns; idx += w.t.nt) any_warm = any_warm || w.face2[idx] != 0;
                    if (team::any(w.t, any_warm)) {              // start the rounds from the given faces
                        commit_faces(w);
                        first = false;
                        dual_round = true; st = S_DUAL;
                        continue;
                    }
                    if (nfac == 0) ++nfac;
                }
                if (mode == M_AS && !ratio_done) { st = S_RATIO; continue; }
                tight = true;
                st = S_TEST;
                continue;
            }
            if (st == S_RATIO) {
                double alpha;
                int tag;
                ratio_test(cs, w, alpha, tag);
                if (tag != 0x7fffffff) {
                    if (alpha <= 1e-13) block_all_at_zero(cs, w);      // degenerate: no move possible, fix every such row
                    else blocked_step(cs, w, alpha, tag);
                    if (++nas > cs.as_cap) break;
                    dual_round = true; st = S_DUAL;
                    continue;
                }
                for (int idx = lane; idx < w.nv; idx += w.t.nt) w.ucur[idx] = w.u[idx];   // feasible face minimiser
                team::sync(w.t);
                ratio_done = true;
                tight = exact;
                st = S_TEST;
                continue;
            }
            // ---- S_TEST
            const FaceCheck fc = pdas_update(cs, w, tight ? cs.tol_p : cs.tol_pc, tight ? cs.tol_d : cs.tol_dc, tight ? cs.tol_r_abs : 1e300);
            const bool clean = fc.n_primal == 0 && fc.n_dual == 0;
#if defined(MPCQ_HOST_EMU) && defined(MPCQ_TRACE)
            if (lane == 0) printf("   tests: primal %d dual %d tight %d rmax %.2e\n", fc.n_primal, fc.n_dual, (int)tight, rmax);
#endif
            if (clean) {
                if (!tight) {                                   // refine on these faces, then the tight tests
                    rf_it = rf_itcg = 0; rf_cg = rf_skip = rf_half = false; rf_prev = rf_rz = 0.0; rf_tol = tol_tight; hin = w.u; hout = w.gam;
                    st = S_HESS;
                    continue;
                }
                // verified only with the stationarity residual actually at tolerance (a NaN fails this test)
                if (rmax <= 10.0 * tol_tight) { done = true; break; }
                if (first && rf_tol > tol_tight) {               // the unconstrained minimiser is feasible: tighten it, test again
                    rf_it = rf_itcg = 0; rf_cg = rf_skip = rf_half = false; rf_prev = rf_rz = 0.0; rf_tol = tol_tight; hin = w.u; hout = w.gam;
                    st = S_HESS;
                    continue;
                }
                numeric_ok = false;
                break;
            }
            if (tight && !first) exact = true;                  // the cheap tests were too optimistic: decide in fp64 from now on
            first = false;
            commit_faces(w);
            if (mode == M_PDAS) {
                if (round++ >= cs.pdas_cap) {                   // the rounds cycle: hand over to the monotone method
                    status |= ST_FALLBACK;
                    clamp_into(cs, w, w.u, w.ucur, w.face);
                    mode = M_AS; nas = 1;
                }
            } else if (++nas > cs.as_cap) {
                break;
            }
            ratio_done = false;
            dual_round = true;
            st = S_DUAL;
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
        const double fm = dmax((double)(float)((double)gait[k] * cs.fz_max), 0.0);
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
