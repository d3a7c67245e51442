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
    int B;
};

enum : int { ST_VERIFIED = 1, ST_FALLBACK = 2, ST_MAXITER = 4, ST_NUMERIC = 8, ST_NO_STANCE = 32 };

// ---------------------------------------------------------------------------------------------
// per-warp workspace
MPCQ_HD constexpr size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
MPCQ_HD constexpr int l_elems(int n) { return n * n / 2 + 2 * n; }

template <class T> struct Work {
    // fp64
    double *Md, *Sd, *GW, *g, *u, *gam, *P0, *P1, *ucur, *fmax;
    // precision T
    T *L, *dblk, *vec, *cw, *zt, *Mf, *St;
    // bytes
    uint8_t *fk;           // stance list: full foot-step index k = 4*step + leg
    int8_t* face;          // 3 per stance foot-step
    int n, ns, H;
};

template <class T>
MPCQ_HD constexpr size_t work_bytes(int H, int ncap, bool l_in_smem) {
    size_t nd = 288 + (size_t)H * H + 72 + 6 * 12 * (size_t)H + ncap / 3 + 1;
    size_t nt = (l_in_smem ? l_elems(ncap) : 0) + 3 * (ncap / 4) * 4 + ncap + 96 + 3 * ncap + 288 + (size_t)H * H;
    size_t nb = (ncap / 3 + 1) * 4;
    return align_up(nd * 8, 16) + align_up(nt * sizeof(T), 16) + align_up(nb, 16);
}

template <class T>
MPCQ_DEV void carve(Work<T>& w, char* base, T* l_global, int H, int ncap) {
    double* d = reinterpret_cast<double*>(base);
    w.Md = d; d += 288;
    w.Sd = d; d += H * H;
    w.GW = d; d += 72;
    w.g = d; d += 12 * H;
    w.u = d; d += 12 * H;
    w.gam = d; d += 12 * H;
    w.P0 = d; d += 12 * H;
    w.P1 = d; d += 12 * H;
    w.ucur = d; d += 12 * H;
    w.fmax = d; d += ncap / 3 + 1;
    size_t nd = 288 + (size_t)H * H + 72 + 6 * 12 * (size_t)H + ncap / 3 + 1;
    T* t = reinterpret_cast<T*>(base + align_up(nd * 8, 16));
    size_t used = 0;
    if (l_global) {
        w.L = l_global;
    } else {
        w.L = t; t += l_elems(ncap); used += l_elems(ncap);
    }
    w.dblk = t; t += 3 * (ncap / 4) * 4; used += 3 * (ncap / 4) * 4;
    w.vec = t; t += ncap; used += ncap;
    w.cw = t; t += 96; used += 96;
    w.zt = t; t += 3 * ncap; used += 3 * ncap;
    w.Mf = t; t += 288; used += 288;
    w.St = t; t += H * H; used += (size_t)H * H;
    uint8_t* b = reinterpret_cast<uint8_t*>(base + align_up(nd * 8, 16) + align_up(used * sizeof(T), 16));
    w.fk = b; b += ncap / 3 + 1;
    w.face = reinterpret_cast<int8_t*>(b);
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

template <class T, int NS> MPCQ_DEV T pick(const T (&a)[NS], int m) {
    T v = a[0];
    MPCQ_UNROLL
    for (int i = 1; i < NS; ++i) v = (i == m) ? a[i] : v;
    return v;
}

MPCQ_DEV double dmax(double a, double b) { return a > b ? a : b; }
MPCQ_DEV double dmin(double a, double b) { return a < b ? a : b; }
MPCQ_DEV double dabs(double a) { return a < 0 ? -a : a; }

// ---------------------------------------------------------------------------------------------
// K1 + K2 (per-env part): model matrices M00, M11, horizon table S and the linear term g.
template <class T>
MPCQ_DEV void setup_model(const Consts& cs, Work<T>& w, const T* x0p, double yaw, const T* feetp, const T* xrefp) {
    const int lane = wp::lane();
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
    // --- horizon table S (fp64 and T)
    for (int idx = lane; idx < H * H; idx += 32) {
        const int i = idx / H, j = idx - i * H;
        const int m = i > j ? i : j;
        const double a = m - i + 0.5, b = m - j + 0.5, Ln = H - m;
        const double sv = Ln * a * b + (a + b) * Ln * (Ln - 1) * 0.5 + (Ln - 1) * Ln * (2 * Ln - 1) / 6.0;
        w.Sd[idx] = sv;
        w.St[idx] = (T)sv;
    }
    wp::sync();
    // --- M00 = B0'QB0, M11 = B1'QB1
    const double dt = cs.dt, dt2 = dt * dt, dt4 = dt2 * dt2, im2 = cs.inv_mass * cs.inv_mass;
    for (int idx = lane; idx < 144; idx += 32) {
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
        w.Md[idx] = m0; w.Md[144 + idx] = m1;
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
            w.P0[12 * j + cidx] = E0;
            w.P1[12 * j + cidx] = E1;
        }
    }
    wp::sync();
    for (int idx = lane; idx < 12 * H; idx += 32) {
        const int j = idx / 12, rr = idx - 12 * j, a = rr / 3, y = rr - 3 * a;
        const double* E0 = w.P0 + 12 * j;
        const double* E1 = w.P1 + 12 * j;
        double t0 = cs.inv_mass * E0[9 + y], t1 = cs.inv_mass * E1[3 + y];
        MPCQ_UNROLL
        for (int k = 0; k < 3; ++k) {
            t0 += w.GW[18 * a + 3 * k + y] * E0[6 + k];
            t1 += w.GW[18 * a + 9 + 3 * k + y] * E1[k];
        }
        w.g[idx] = 2.0 * (dt * t0 + dt2 * t1);
    }
    wp::sync();
}

// gam = H u + g, fp64, through the Kronecker structure (u, gam in full [H][12] layout)
template <class T>
MPCQ_DEV void hess_apply(const Consts& cs, Work<T>& w) {
    const int lane = wp::lane();
    const int H = cs.horizon;
    for (int idx = lane; idx < 12 * H; idx += 32) {
        const int i = idx / 12, r = idx - 12 * i;
        const double* ui = w.u + 12 * i;
        const double* m0 = w.Md + 12 * r;
        const double* m1 = w.Md + 144 + 12 * r;
        double p0 = 0, p1 = 0;
        MPCQ_UNROLL
        for (int c = 0; c < 12; ++c) { p0 += m0[c] * ui[c]; p1 += m1[c] * ui[c]; }
        w.P0[idx] = p0; w.P1[idx] = p1;
    }
    wp::sync();
    for (int idx = lane; idx < 12 * H; idx += 32) {
        const int j = idx / 12, r = idx - 12 * j;
        double acc = 0;
        for (int i = 0; i < H; ++i) {
            const double Nij = (double)(H - (i > j ? i : j));
            acc += Nij * w.P0[12 * i + r] + w.Sd[i * H + j] * w.P1[12 * i + r];
        }
        w.gam[idx] = w.g[idx] + 2.0 * (acc + cs.r[r] * w.u[idx]);
    }
    wp::sync();
}

// ---------------------------------------------------------------------------------------------
// K3: faces -> slot vectors z (precision T) and the face constants c written into u.
// slot v = 3p + comp of stance foot-step p; dead slots (z = 0) become identity rows.
template <class T>
MPCQ_DEV bool build_slots(const Consts& cs, Work<T>& w) {
    const int lane = wp::lane();
    const T mu = (T)cs.mu;
    bool nonzero_c = false;
    for (int idx = lane; idx < 12 * cs.horizon; idx += 32) w.u[idx] = 0.0;
    wp::sync();
    for (int p = lane; p < w.n / 3 + 1; p += 32) {
        T z[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (p < w.ns) {
            const int sx = w.face[3 * p], sy = w.face[3 * p + 1], sz = w.face[3 * p + 2];
            if (sz >= 0) {
                if (sx == 0) z[0] = 1;
                if (sy == 0) z[4] = 1;
                if (sz == 0) { z[6] = sx * mu; z[7] = sy * mu; z[8] = 1; }
                else {
                    const double fm = w.fmax[p];
                    double* up = w.u + 3 * w.fk[p];
                    up[0] = sx * cs.mu * fm; up[1] = sy * cs.mu * fm; up[2] = fm;
                    nonzero_c = true;
                }
            }
        }
        MPCQ_UNROLL
        for (int c = 0; c < 9; ++c)
            if (3 * p + c / 3 < w.n) w.zt[9 * p + c] = z[c];
    }
    wp::sync();
    return wp::any(nonzero_c);
}

// ---------------------------------------------------------------------------------------------
// K2 + K4a: assemble K = Z'HZ column panel by column panel (never materialised) and factor it.
// Left-looking, 4-column panels; lane owns rows lane, lane+32, ... ; returns false on a bad pivot.
template <class T, int NSLOT>
MPCQ_DEV bool chol_factor(const Consts& cs, Work<T>& w) {
    const int lane = wp::lane();
    const int n = w.n, H = cs.horizon;
    T* L = w.L;
    bool ok = true;
    for (int k0 = 0; k0 < n; k0 += 4) {
        // ---- cw[c][mat*12 + r] = sum_y M_mat[r][3b+y] z_w[y] for the 4 panel columns
        MPCQ_UNROLL
        for (int t = 0; t < 3; ++t) {
            const int idx = lane + 32 * t;
            const int c = idx / 24, rem = idx - 24 * c;
            const int wv = k0 + c, pw = wv / 3;
            T val = 0;
            if (pw < w.ns) {
                const int b = w.fk[pw] & 3;
                const T* z = w.zt + 3 * wv;
                const T* mrow = w.Mf + (rem / 12) * 144 + (rem % 12) * 12 + 3 * b;
                val = mrow[0] * z[0] + mrow[1] * z[1] + mrow[2] * z[2];
            }
            w.cw[idx] = val;
        }
        wp::sync();
        // ---- initial entries of the panel
        T acc[NSLOT][4];
        const int m0 = k0 >> 5;
        MPCQ_UNROLL
        for (int m = 0; m < NSLOT; ++m) {
            MPCQ_UNROLL
            for (int c = 0; c < 4; ++c) acc[m][c] = 0;
            const int v = lane + 32 * m;
            if (m >= m0 && v >= k0 && v < n) {
                const int pv = v / 3;
                const T z0 = w.zt[3 * v], z1 = w.zt[3 * v + 1], z2 = w.zt[3 * v + 2];
                const bool dead = (z0 == 0 && z1 == 0 && z2 == 0);
                int iv = 0, av = 0;
                if (pv < w.ns) { iv = w.fk[pv] >> 2; av = w.fk[pv] & 3; }
                MPCQ_UNROLL
                for (int c = 0; c < 4; ++c) {
                    const int wv = k0 + c;
                    if (wv > v) continue;
                    const int pw = wv / 3;
                    T e = 0;
                    if (!dead && pw < w.ns) {
                        const int jw = w.fk[pw] >> 2;
                        const T* cwc = w.cw + 24 * c + 3 * av;
                        const T t0 = z0 * cwc[0] + z1 * cwc[1] + z2 * cwc[2];
                        const T t1 = z0 * cwc[12] + z1 * cwc[13] + z2 * cwc[14];
                        const int mx = iv > jw ? iv : jw;
                        e = (T)2 * ((T)(H - mx) * t0 + w.St[iv * H + jw] * t1);
                        if (pw == pv) {
                            const T* zw = w.zt + 3 * wv;
                            e += (T)2 * (z0 * zw[0] * (T)cs.r[3 * av] + z1 * zw[1] * (T)cs.r[3 * av + 1] +
                                         z2 * zw[2] * (T)cs.r[3 * av + 2]);
                        }
                    }
                    if (dead && wv == v) e = 1;
                    acc[m][c] = e;
                }
            }
        }
        // ---- left-looking update with all previous columns
        for (int g = 0; g < (k0 >> 2); ++g) {
            const int stride = n - 4 * g;
            const int cb = 4 * g * n - 8 * g * (g - 1) - 4 * g;
            MPCQ_UNROLL
            for (int t = 0; t < 4; ++t) {
                const T* col = L + cb + t * stride;
                T p0, p1, p2, p3;
                load4(col + k0, p0, p1, p2, p3);
                MPCQ_UNROLL
                for (int m = 0; m < NSLOT; ++m) {
                    const int v = lane + 32 * m;
                    if (m >= m0 && v >= k0 && v < n) {
                        const T lr = col[v];
                        acc[m][0] -= lr * p0; acc[m][1] -= lr * p1; acc[m][2] -= lr * p2; acc[m][3] -= lr * p3;
                    }
                }
            }
        }
        // ---- 4x4 diagonal block: fetched from its owner lanes, factored redundantly by every lane
        const int ld = k0 & 31;
        T a0 = acc[0][0], a1 = acc[0][1], a2 = acc[0][2], a3 = acc[0][3];
        MPCQ_UNROLL
        for (int m = 1; m < NSLOT; ++m)
            if (m == m0) { a0 = acc[m][0]; a1 = acc[m][1]; a2 = acc[m][2]; a3 = acc[m][3]; }
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
        if (lane == 0) {
            T* db = w.dblk + 3 * k0;          // 12 values per block of 4 columns
            db[0] = l10; db[1] = l20; db[2] = l21; db[3] = l30;
            db[4] = l31; db[5] = l32; db[6] = i0; db[7] = i1;
            db[8] = i2; db[9] = i3; db[10] = 0; db[11] = 0;
        }
        // ---- panel rows: x = acc * inv(Ld)'
        {
            const int stride = n - k0;
            T* c0 = L + colbase(k0, n);
            MPCQ_UNROLL
            for (int m = 0; m < NSLOT; ++m) {
                const int v = lane + 32 * m;
                if (m >= m0 && v >= k0 && v < n) {
                    const T x0 = acc[m][0] * i0;
                    const T x1 = (acc[m][1] - x0 * l10) * i1;
                    const T x2 = (acc[m][2] - x0 * l20 - x1 * l21) * i2;
                    const T x3 = (acc[m][3] - x0 * l30 - x1 * l31 - x2 * l32) * i3;
                    const int dv = v - k0;
                    c0[v] = x0;
                    c0[stride + v] = dv >= 1 ? x1 : (T)0;
                    c0[2 * stride + v] = dv >= 2 ? x2 : (T)0;
                    c0[3 * stride + v] = dv >= 3 ? x3 : (T)0;
                }
            }
        }
        wp::sync();
    }
    return wp::all(ok);
}

// K4b: solve L L' x = vec in place (vec in shared memory, precision T)
template <class T, int NSLOT>
MPCQ_DEV void tri_solve(Work<T>& w) {
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
    // forward: L y = b
    for (int k0 = 0; k0 < n; k0 += 4) {
        const int md = k0 >> 5, ld = k0 & 31;
        const T mine = pick<T, NSLOT>(bv, md);
        const T b0 = wp::shfl(mine, ld), b1 = wp::shfl(mine, ld + 1), b2 = wp::shfl(mine, ld + 2), b3 = wp::shfl(mine, ld + 3);
        T l10, l20, l21, l30, l31, l32, i0, i1, i2, i3, pad0, pad1;
        const T* db = w.dblk + 3 * k0;
        load4(db, l10, l20, l21, l30);
        load4(db + 4, l31, l32, i0, i1);
        load4(db + 8, i2, i3, pad0, pad1);
        const T y0 = b0 * i0;
        const T y1 = (b1 - l10 * y0) * i1;
        const T y2 = (b2 - l20 * y0 - l21 * y1) * i2;
        const T y3 = (b3 - l30 * y0 - l31 * y1 - l32 * y2) * i3;
        const int stride = n - k0;
        const T* c0 = L + colbase(k0, n);
        MPCQ_UNROLL
        for (int m = 0; m < NSLOT; ++m) {
            const int v = lane + 32 * m;
            if (m >= md && v < n) {
                if (v >= k0 + 4) {
                    bv[m] -= (c0[v] * y0 + c0[stride + v] * y1) + (c0[2 * stride + v] * y2 + c0[3 * stride + v] * y3);
                } else if (v >= k0) {
                    const int dv = v - k0;
                    bv[m] = dv == 0 ? y0 : dv == 1 ? y1 : dv == 2 ? y2 : y3;
                }
            }
        }
    }
    // backward: L' x = y
    for (int k0 = n - 4; k0 >= 0; k0 -= 4) {
        const int md = k0 >> 5, ld = k0 & 31;
        const T mine = pick<T, NSLOT>(bv, md);
        const T b0 = wp::shfl(mine, ld), b1 = wp::shfl(mine, ld + 1), b2 = wp::shfl(mine, ld + 2), b3 = wp::shfl(mine, ld + 3);
        T l10, l20, l21, l30, l31, l32, i0, i1, i2, i3, pad0, pad1;
        const T* db = w.dblk + 3 * k0;
        load4(db, l10, l20, l21, l30);
        load4(db + 4, l31, l32, i0, i1);
        load4(db + 8, i2, i3, pad0, pad1);
        const T x3 = b3 * i3;
        const T x2 = (b2 - l32 * x3) * i2;
        const T x1 = (b1 - l21 * x2 - l31 * x3) * i1;
        const T x0 = (b0 - l10 * x1 - l20 * x2 - l30 * x3) * i0;
        MPCQ_UNROLL
        for (int m = 0; m < NSLOT; ++m) {
            const int v = lane + 32 * m;
            if (m <= md && v < n) {
                if (v < k0) {
                    T a0, a1, a2, a3;
                    load4(L + cbv[m] + k0, a0, a1, a2, a3);      // L[k0..k0+3, v]
                    bv[m] -= (a0 * x0 + a1 * x1) + (a2 * x2 + a3 * x3);
                } else if (v < k0 + 4) {
                    const int dv = v - k0;
                    bv[m] = dv == 0 ? x0 : dv == 1 ? x1 : dv == 2 ? x2 : x3;
                }
            }
        }
    }
    MPCQ_UNROLL
    for (int m = 0; m < NSLOT; ++m) {
        const int v = lane + 32 * m;
        if (v < n) w.vec[v] = bv[m];
    }
    wp::sync();
}

// ---------------------------------------------------------------------------------------------
// reduced gradient r = -Z' gam into vec (precision T); returns |r|_inf (fp64)
template <class T>
MPCQ_DEV double reduced_gradient(Work<T>& w) {
    const int lane = wp::lane();
    double rmax = 0;
    for (int v = lane; v < w.n; v += 32) {
        const int p = v / 3;
        double r = 0;
        if (p < w.ns) {
            const double* gp = w.gam + 3 * w.fk[p];
            const T* z = w.zt + 3 * v;
            r = -((double)z[0] * gp[0] + (double)z[1] * gp[1] + (double)z[2] * gp[2]);
        }
        w.vec[v] = (T)r;
        rmax = dmax(rmax, dabs(r));
    }
    rmax = wp::reduce_max(rmax);
    wp::sync();
    return rmax;
}

// u += Z w  (w = vec); Z is rebuilt from the face codes in fp64 so the face equalities hold
// to double precision whatever T is
template <class T>
MPCQ_DEV void apply_step(const Consts& cs, Work<T>& w) {
    const int lane = wp::lane();
    for (int p = lane; p < w.ns; p += 32) {
        const int sx = w.face[3 * p], sy = w.face[3 * p + 1], sz = w.face[3 * p + 2];
        if (sz < 0) continue;
        double* up = w.u + 3 * w.fk[p];
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
    wp::sync();
}

// refine u on the current factorisation until the reduced gradient is below tol; leaves gam = Hu+g
template <class T, int NSLOT>
MPCQ_DEV double refine(const Consts& cs, Work<T>& w, double tol_abs, bool u_is_zero) {
    const int lane = wp::lane();
    double rmax = 0, prev = 0;
    for (int it = 0;; ++it) {
        if (u_is_zero) {
            for (int idx = lane; idx < 12 * cs.horizon; idx += 32) w.gam[idx] = w.g[idx];
            wp::sync();
            u_is_zero = false;
        } else {
            hess_apply(cs, w);
        }
        rmax = reduced_gradient(w);
        if (rmax <= tol_abs || it >= cs.refine_max || (it > 1 && rmax > 0.5 * prev)) break;   // done / cap / stagnating
        prev = rmax;
        tri_solve<T, NSLOT>(w);
        apply_step(cs, w);
    }
    return rmax;
}

// ---------------------------------------------------------------------------------------------
// face tests.  mode 0 (primal-dual round): rewrite every offending face, return counts.
struct FaceCheck { int n_primal, n_dual; };

template <class T>
MPCQ_DEV FaceCheck pdas_update(const Consts& cs, Work<T>& w, bool write) {
    const int lane = wp::lane();
    const double mu = cs.mu;
    int npv = 0, ndv = 0;
    for (int p = lane; p < w.ns; p += 32) {
        const double* f = w.u + 3 * w.fk[p];
        const double* ga = w.gam + 3 * w.fk[p];
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
        if (write) { w.face[3 * p] = (int8_t)sx; w.face[3 * p + 1] = (int8_t)sy; w.face[3 * p + 2] = (int8_t)sz; }
    }
    FaceCheck fc;
    fc.n_primal = wp::reduce_sum(npv);
    fc.n_dual = wp::reduce_sum(ndv);
    wp::sync();
    return fc;
}

// one factor-and-solve on the current faces: u = argmin on the faces (to tol), gam = Hu+g
template <class T, int NSLOT>
MPCQ_DEV bool face_solve(const Consts& cs, Work<T>& w, double tol_abs, double& rmax) {
    const bool cnz = build_slots(cs, w);
    const bool ok = chol_factor<T, NSLOT>(cs, w);
    rmax = refine<T, NSLOT>(cs, w, tol_abs, !cnz);
    return ok;
}

// ---------------------------------------------------------------------------------------------
// fallback: feasible primal active-set method started from the clamped last iterate.
// The start point and its faces are derived from the point alone: every foot is moved into K and
// a row is taken active when the point sits on it (or beyond it) within the primal tolerance.
template <class T>
MPCQ_DEV void clamp_to_feasible(const Consts& cs, Work<T>& w) {
    const int lane = wp::lane();
    const double mu = cs.mu;
    for (int idx = lane; idx < 12 * cs.horizon; idx += 32) w.ucur[idx] = 0.0;
    wp::sync();
    for (int p = lane; p < w.ns; p += 32) {
        const double* f = w.u + 3 * w.fk[p];
        double* o = w.ucur + 3 * w.fk[p];
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
        w.face[3 * p] = (int8_t)sx; w.face[3 * p + 1] = (int8_t)sy; w.face[3 * p + 2] = (int8_t)sz;
    }
    wp::sync();
}

MPCQ_DEV void row_slacks(const double* f, double mu, double fm, double (&s)[6]) {
    s[0] = f[0] + mu * f[2]; s[1] = -f[0] + mu * f[2];
    s[2] = f[1] + mu * f[2]; s[3] = -f[1] + mu * f[2];
    s[4] = f[2]; s[5] = fm - f[2];
}

// returns: 0 = moved / face changed, keep going; 1 = optimal
template <class T>
MPCQ_DEV int active_set_step(const Consts& cs, Work<T>& w) {
    const int lane = wp::lane();
    const double mu = cs.mu;
    // --- ratio test from ucur towards the face minimiser u
    double alpha = 1.0;
    int tag = 0x7fffffff;
    for (int p = lane; p < w.ns; p += 32) {
        const int sx = w.face[3 * p], sy = w.face[3 * p + 1], sz = w.face[3 * p + 2];
        if (sz < 0) continue;
        const double* f0 = w.ucur + 3 * w.fk[p];
        const double* f1 = w.u + 3 * w.fk[p];
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
    wp::reduce_argmin(alpha, tag);
    const bool blocked = tag != 0x7fffffff;
    if (!blocked) alpha = 1.0;
#if defined(MPCQ_HOST_EMU) && defined(MPCQ_TRACE)
    if (lane == 0 && blocked) printf("  AS block p=%d (k=%d) row=%d alpha=%.3e\n", tag >> 3, w.fk[tag >> 3], tag & 7, alpha);
#endif
    for (int idx = lane; idx < 12 * cs.horizon; idx += 32) w.ucur[idx] += alpha * (w.u[idx] - w.ucur[idx]);
    wp::sync();
    if (blocked) {
        if (lane == 0) {
            const int p = tag >> 3, r = tag & 7;
            int sx = w.face[3 * p], sy = w.face[3 * p + 1], sz = w.face[3 * p + 2];
            bool apex = (r == 4) || (r == 0 && sx == 1) || (r == 1 && sx == -1) || (r == 2 && sy == 1) || (r == 3 && sy == -1);
            if (r == 0) sx = -1; else if (r == 1) sx = 1; else if (r == 2) sy = -1; else if (r == 3) sy = 1; else if (r == 5) sz = 1;
            double* f = w.ucur + 3 * w.fk[p];
            if (apex || !(f[2] > 0.0)) { sz = -1; f[0] = f[1] = f[2] = 0.0; }
            w.face[3 * p] = (int8_t)sx; w.face[3 * p + 1] = (int8_t)sy; w.face[3 * p + 2] = (int8_t)sz;
        }
        wp::sync();
        return 0;
    }
    // --- at the face minimiser (gam is current): release the most negative multiplier
    double worst = 0.0;
    int rel = 0x7fffffff;
    for (int p = lane; p < w.ns; p += 32) {
        const int sx = w.face[3 * p], sy = w.face[3 * p + 1], sz = w.face[3 * p + 2];
        const double* ga = w.gam + 3 * w.fk[p];
        const double gs = 1.0 + dmax(dabs(ga[0]), dmax(dabs(ga[1]), dabs(ga[2])));
        if (sz < 0) {
            const double v = (ga[2] - mu * (dabs(ga[0]) + dabs(ga[1]))) / gs;
            if (v < -cs.tol_d && v < worst) { worst = v; rel = p * 8 + 4; }
        } else {
            const double lx = sx != 0 ? -sx * ga[0] : 0.0, ly = sy != 0 ? -sy * ga[1] : 0.0;
            if (sx != 0 && lx / gs < -cs.tol_d && lx / gs < worst) { worst = lx / gs; rel = p * 8 + 0; }
            if (sy != 0 && ly / gs < -cs.tol_d && ly / gs < worst) { worst = ly / gs; rel = p * 8 + 2; }
            if (sz > 0) {
                const double lt = (-ga[2] + mu * (lx + ly)) / gs;
                if (lt < -cs.tol_d && lt < worst) { worst = lt; rel = p * 8 + 5; }
            }
        }
    }
    wp::reduce_argmin(worst, rel);
#if defined(MPCQ_HOST_EMU) && defined(MPCQ_TRACE)
    if (lane == 0 && rel != 0x7fffffff) printf("  AS release p=%d (k=%d) row=%d mult=%.3e\n", rel >> 3, w.fk[rel >> 3], rel & 7, worst);
#endif
    if (rel == 0x7fffffff) return 1;
    if (lane == 0) {
        const int p = rel >> 3, r = rel & 7;
        const double* ga = w.gam + 3 * w.fk[p];
        if (r == 4) {
            w.face[3 * p] = (int8_t)(ga[0] > 0 ? -1 : (ga[0] < 0 ? 1 : 0));
            w.face[3 * p + 1] = (int8_t)(ga[1] > 0 ? -1 : (ga[1] < 0 ? 1 : 0));
            w.face[3 * p + 2] = 0;
        } else if (r == 0) w.face[3 * p] = 0;
        else if (r == 2) w.face[3 * p + 1] = 0;
        else w.face[3 * p + 2] = 0;
    }
    wp::sync();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// the whole path for environment b.  NCAP = slot capacity of this size class.
template <class T, int NCAP>
MPCQ_DEV void solve_env(const Consts& cs, const IO<T>& io, int b, char* smem, T* l_global, int ns_lo, int ns_hi) {
    constexpr int NSLOT = NCAP / 32;
    const int lane = wp::lane();
    const int H = cs.horizon;
    Work<T> w;
    carve(w, smem, l_global, H, NCAP);
    // ---- K3a: stance list from the contact table (ub_fz = gait * fz_max > 0)
    const float* gait = io.gait + (size_t)b * 4 * H;
    int ns = 0;
    for (int k0 = 0; k0 < 4 * H; k0 += 32) {
        const int k = k0 + lane;
        const double fm = k < 4 * H ? (double)gait[k] * cs.fz_max : 0.0;
        const bool st = fm > 0.0;
        const unsigned bal = wp::ballot(st);
        const int pos = ns + wp::popc(bal & ((1u << lane) - 1u));
        if (st && pos < NCAP / 3) { w.fk[pos] = (uint8_t)k; w.fmax[pos] = fm; w.face[3 * pos] = 0; w.face[3 * pos + 1] = 0; w.face[3 * pos + 2] = 0; }
        ns += wp::popc(bal);
    }
    if (ns < ns_lo || ns > ns_hi) return;                       // another size class owns this env
    w.ns = ns;
    w.n = (3 * ns + 3) & ~3;
    wp::sync();
    int status = 0, nfac = 0, nas = 0;
    double rmax = 0.0, pviol = 0.0;
    if (ns == 0) {
        status = ST_NO_STANCE | ST_VERIFIED;
        for (int idx = lane; idx < 12 * H; idx += 32) w.u[idx] = 0.0;
        wp::sync();
    } else {
        const double yaw = io.yaw ? (double)io.yaw[b] : (double)io.x0[(size_t)b * 13 + 2];
        setup_model(cs, w, io.x0 + (size_t)b * 13, yaw, io.r_feet + (size_t)b * 12, io.x_ref + (size_t)b * 13 * H);
        double gsc = 0.0;
        for (int idx = lane; idx < 12 * H; idx += 32) gsc = dmax(gsc, dabs(w.g[idx]));
        gsc = 1.0 + wp::reduce_max(gsc);
        const double tol_loose = cs.tol_r_loose * gsc, tol_tight = cs.tol_r_tight * gsc;
        bool numeric_ok = (gsc == gsc) && (gsc < 1e300);
        bool done = false;
        // ---- primal-dual active-set rounds
        for (int round = 0; round <= cs.pdas_cap && numeric_ok && !done; ++round) {
            numeric_ok = face_solve<T, NSLOT>(cs, w, tol_loose, rmax) && numeric_ok;
            ++nfac;
            FaceCheck fc = pdas_update(cs, w, false);
            if (fc.n_primal == 0 && fc.n_dual == 0) {
                rmax = refine<T, NSLOT>(cs, w, tol_tight, false);   // tighten on the same factor, re-test
                fc = pdas_update(cs, w, false);
                if (fc.n_primal == 0 && fc.n_dual == 0) { done = true; break; }
            }
#if defined(MPCQ_HOST_EMU) && defined(MPCQ_TRACE)
            if (lane == 0) printf(" PDAS round %d: primal %d dual %d rmax %.2e\n", round, fc.n_primal, fc.n_dual, rmax);
#endif
            pdas_update(cs, w, true);
        }
        // ---- fallback: feasible primal active set (monotone, cannot cycle)
        if (!done && numeric_ok) {
            status |= ST_FALLBACK;
            clamp_to_feasible(cs, w);
            for (nas = 1; nas <= cs.as_cap; ++nas) {
                numeric_ok = face_solve<T, NSLOT>(cs, w, tol_tight, rmax) && numeric_ok;
                ++nfac;
                if (!numeric_ok) break;
                if (active_set_step(cs, w) == 1) { done = true; break; }
            }
            if (!done) {                                   // return the feasible iterate
                for (int idx = lane; idx < 12 * H; idx += 32) w.u[idx] = w.ucur[idx];
                wp::sync();
                hess_apply(cs, w);
            }
        }
        if (done) status |= ST_VERIFIED; else status |= numeric_ok ? ST_MAXITER : ST_NUMERIC;
    }
    // ---- outputs: forces, activity (on primal slack, like the oracle's kkt_report), residuals
    const double mu = cs.mu;
    for (int k = lane; k < 4 * H; k += 32) {
        const double* f = w.u + 3 * k;
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
    }
    pviol = wp::reduce_max(pviol);
    if (io.u_full)
        for (int idx = lane; idx < 12 * H; idx += 32) io.u_full[(size_t)b * 12 * H + idx] = (T)w.u[idx];
    if (lane < 12) io.f_out[(size_t)b * 12 + lane] = (T)w.u[lane];
    if (lane == 0) {
        if (io.iters) { io.iters[2 * b] = nfac; io.iters[2 * b + 1] = nas; }
        if (io.resid) { io.resid[2 * b] = rmax; io.resid[2 * b + 1] = dmax(pviol, 0.0); }
        if (io.status) io.status[b] = status;
    }
}

}  // namespace mpcq
