// Warp-level primitives of the MPC QP engine.
//
// The solver is written warp-synchronously: one 32-lane warp owns one environment and all
// cross-lane traffic goes through the handful of collectives below.  Under nvcc they are the
// sm_100a intrinsics.  When the same source is compiled by g++ with MPCQ_HOST_EMU (tests/emu
// only - a debugging harness for the device code, never linked into libmpcq.so) the 32 lanes
// run as cooperative coroutines and the collectives are a lock-step exchange, so the kernel
// logic can be unit-tested in the CPU-only build container.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef MPCQ_HOST_EMU
#define MPCQ_DEV inline
#define MPCQ_HD inline
#define MPCQ_UNROLL
#define MPCQ_UNROLL2
#define MPCQ_NOUNROLL
namespace mpcq_emu {
int lane_id();                            // lane within the warp
int thread_id();                          // thread within the team (CTA)
uint64_t exchange(uint64_t v, int src);   // every lane of the warp publishes v, returns the value of lane `src`
void team_barrier();                      // all threads of the team
}
#else
#include <cuda_runtime.h>
#define MPCQ_DEV __device__ __forceinline__
#define MPCQ_HD __host__ __device__ inline
#define MPCQ_UNROLL _Pragma("unroll")
#define MPCQ_UNROLL2 _Pragma("unroll 2")
#define MPCQ_NOUNROLL _Pragma("unroll 1")
#endif

namespace wp {

constexpr unsigned FULL = 0xffffffffu;

#ifdef MPCQ_HOST_EMU
MPCQ_DEV int lane() { return mpcq_emu::lane_id(); }
MPCQ_DEV void sync() { mpcq_emu::exchange(0, 0); }
template <class V> MPCQ_DEV V shfl(V v, int src) {
    static_assert(sizeof(V) <= 8, "shfl payload");
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(V));
    raw = mpcq_emu::exchange(raw, src & 31);
    V out;
    memcpy(&out, &raw, sizeof(V));
    return out;
}
template <class V> MPCQ_DEV V shfl_xor(V v, int m) { return shfl(v, lane() ^ m); }
MPCQ_DEV unsigned ballot(bool p) {
    unsigned out = 0;
    for (int l = 0; l < 32; ++l) out |= (shfl<int>(p ? 1 : 0, l) ? 1u : 0u) << l;
    return out;
}
MPCQ_DEV int popc(unsigned x) { return __builtin_popcount(x); }
MPCQ_DEV float rsqrt_(float x) { return 1.0f / sqrtf(x); }
MPCQ_DEV double rsqrt_(double x) { return 1.0 / sqrt(x); }
#else
MPCQ_DEV int lane() { return threadIdx.x & 31; }
MPCQ_DEV void sync() { __syncwarp(); }
template <class V> MPCQ_DEV V shfl(V v, int src) { return __shfl_sync(FULL, v, src); }
template <class V> MPCQ_DEV V shfl_xor(V v, int m) { return __shfl_xor_sync(FULL, v, m); }
MPCQ_DEV unsigned ballot(bool p) { return __ballot_sync(FULL, p); }
MPCQ_DEV int popc(unsigned x) { return __popc(x); }
// reciprocal square roots: hardware approximation (MUFU.RSQ, ~2 ulp) plus one Newton step, which brings it to
// within 1 ulp with a much shorter dependency chain than 1 / sqrt (the pivot chain is the serial part of the
// panel factorisation)
MPCQ_DEV float rsqrt_(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // bare MUFU.RSQ: pivots are normal numbers (checked > 0 before)
    return y * (1.5f - 0.5f * x * y * y);
}
MPCQ_DEV double rsqrt_(double x) {
    const double y = rsqrt(x);
    return y * (1.5 - 0.5 * x * y * y);
}
#endif

// sel(p, a, b) = p ? a : b as ONE select instruction.  nvcc turns nested ternaries on per-lane predicates into divergent
// branch regions (BSSY / BRA / BSYNC) - measured in the triangular sweeps - so the hot loops spell the select out.
#ifdef MPCQ_HOST_EMU
template <class V> MPCQ_DEV V sel(bool p, V a, V b) { return p ? a : b; }
#else
MPCQ_DEV float sel(bool p, float a, float b) {
    float r;
    asm("{ .reg .pred q; setp.ne.s32 q, %3, 0; selp.f32 %0, %1, %2, q; }" : "=f"(r) : "f"(a), "f"(b), "r"((int)p));
    return r;
}
MPCQ_DEV double sel(bool p, double a, double b) {
    double r;
    asm("{ .reg .pred q; setp.ne.s32 q, %3, 0; selp.f64 %0, %1, %2, q; }" : "=d"(r) : "d"(a), "d"(b), "r"((int)p));
    return r;
}
MPCQ_DEV int sel(bool p, int a, int b) {
    int r;
    asm("{ .reg .pred q; setp.ne.s32 q, %3, 0; selp.s32 %0, %1, %2, q; }" : "=r"(r) : "r"(a), "r"(b), "r"((int)p));
    return r;
}
#endif

// a[c] -= l * p[c], c = 0..3.  fp32 on sm_100a: two packed FFMA2 (fma.rn.f32x2, two IEEE fmas per instruction - same
// results as four scalar FFMAs) with the negated multiplier duplicated into a register pair.
template <class V> MPCQ_DEV void fma4_sub(V (&a)[4], V l, const V (&p)[4]) {
    a[0] -= l * p[0]; a[1] -= l * p[1]; a[2] -= l * p[2]; a[3] -= l * p[3];
}
#if !defined(MPCQ_HOST_EMU) && !defined(MPCQ_NO_FFMA2)
MPCQ_DEV void fma4_sub(float (&a)[4], float l, const float (&p)[4]) {
    const float2 nl = make_float2(-l, -l);
    const float2 r01 = __ffma2_rn(nl, make_float2(p[0], p[1]), make_float2(a[0], a[1]));
    const float2 r23 = __ffma2_rn(nl, make_float2(p[2], p[3]), make_float2(a[2], a[3]));
    a[0] = r01.x; a[1] = r01.y; a[2] = r23.x; a[3] = r23.y;
}
#endif

#ifdef MPCQ_HOST_EMU
MPCQ_DEV int team_tid(int) { return mpcq_emu::thread_id(); }
MPCQ_DEV void team_sync(int, int) { mpcq_emu::team_barrier(); }
#else
MPCQ_DEV int team_tid(int nt) { return threadIdx.x % nt; }
// a team is the whole CTA (barrier 0) or one of several teams sharing a CTA, each with its own named barrier (1..15)
MPCQ_DEV void team_sync(int bar, int nt) { asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nt) : "memory"); }
#endif

MPCQ_DEV bool any(bool p) { return ballot(p) != 0u; }
MPCQ_DEV bool all(bool p) { return ballot(p) == FULL; }

// Reductions must hand EVERY lane the same value - control flow is decided on them, and lanes that disagree would
// diverge around later full-mask collectives (a hang on real hardware).  `o > v ? o : v` is not commutative when a NaN
// is involved, so NaN is made to win explicitly, and the result is taken from lane 0 in the end.
template <class V> MPCQ_DEV V reduce_max(V v) {
    MPCQ_UNROLL
    for (int m = 16; m > 0; m >>= 1) {
        V o = shfl_xor(v, m);
        v = (o > v || o != o) ? o : v;                         // NaN propagates
    }
    return shfl(v, 0);
}
template <class V> MPCQ_DEV V reduce_sum(V v) {
    MPCQ_UNROLL
    for (int m = 16; m > 0; m >>= 1) v += shfl_xor(v, m);
    return shfl(v, 0);
}
// (value, payload) arg-min; ties resolved towards the smaller payload so the result is independent of lane order;
// a NaN value counts as smaller than everything (it must be noticed, not skipped)
MPCQ_DEV void reduce_argmin(double& v, int& tag) {
    MPCQ_UNROLL
    for (int m = 16; m > 0; m >>= 1) {
        double ov = shfl_xor(v, m);
        int ot = shfl_xor(tag, m);
        const bool o_nan = ov != ov, v_nan = v != v;
        const bool take = (o_nan && !v_nan) || (o_nan == v_nan && (ov < v || ((ov == v || o_nan) && ot < tag)));
        if (take) { v = ov; tag = ot; }
    }
    v = shfl(v, 0);
    tag = shfl(tag, 0);
}

}  // namespace wp

// A team = the warps (one CTA of nt threads) that own one environment.  With one warp the team collectives are the
// warp collectives; with more they go through a few words of shared memory and the CTA barrier.
namespace team {

struct Ctx {
    int tid, nt, wid;      // thread in team, team size, warp in team
    int bar;               // hardware barrier of the team (0 = the whole CTA)
    double* red;           // [16] scratch, one slot per warp
    int* redi;             // [16] scratch, one slot per warp (bcast uses the last)
};

MPCQ_DEV void sync(const Ctx& c) {
    if (c.nt == 32) wp::sync(); else wp::team_sync(c.bar, c.nt);
}

MPCQ_DEV double reduce_max(const Ctx& c, double v) {
    v = wp::reduce_max(v);
    if (c.nt == 32) return v;
    if (wp::lane() == 0) c.red[c.wid] = v;
    wp::team_sync(c.bar, c.nt);
    double r = c.red[0];
    for (int i = 1; i < (c.nt >> 5); ++i) r = (c.red[i] > r || c.red[i] != c.red[i]) ? c.red[i] : r;
    wp::team_sync(c.bar, c.nt);
    return r;
}

MPCQ_DEV double reduce_sum(const Ctx& c, double v) {
    v = wp::reduce_sum(v);
    if (c.nt == 32) return v;
    if (wp::lane() == 0) c.red[c.wid] = v;
    wp::team_sync(c.bar, c.nt);
    double r = c.red[0];
    for (int i = 1; i < (c.nt >> 5); ++i) r += c.red[i];
    wp::team_sync(c.bar, c.nt);
    return r;
}

MPCQ_DEV int reduce_sum(const Ctx& c, int v) {
    v = wp::reduce_sum(v);
    if (c.nt == 32) return v;
    if (wp::lane() == 0) c.redi[c.wid] = v;
    wp::team_sync(c.bar, c.nt);
    int r = c.redi[0];
    for (int i = 1; i < (c.nt >> 5); ++i) r += c.redi[i];
    wp::team_sync(c.bar, c.nt);
    return r;
}

MPCQ_DEV void reduce_argmin(const Ctx& c, double& v, int& tag) {
    wp::reduce_argmin(v, tag);
    if (c.nt == 32) return;
    if (wp::lane() == 0) { c.red[c.wid] = v; c.redi[c.wid] = tag; }
    wp::team_sync(c.bar, c.nt);
    double rv = c.red[0];
    int rt = c.redi[0];
    for (int i = 1; i < (c.nt >> 5); ++i) {                  // same order as wp::reduce_argmin: a NaN wins (it must be noticed)
        const double ov = c.red[i];
        const int ot = c.redi[i];
        const bool o_nan = ov != ov, v_nan = rv != rv;
        if ((o_nan && !v_nan) || (o_nan == v_nan && (ov < rv || ((ov == rv || o_nan) && ot < rt)))) { rv = ov; rt = ot; }
    }
    wp::team_sync(c.bar, c.nt);
    v = rv;
    tag = rt;
}

MPCQ_DEV bool any(const Ctx& c, bool p) { return reduce_sum(c, p ? 1 : 0) != 0; }

// value held by warp 0 -> every thread (also a barrier: what warp 0 wrote before is visible after)
MPCQ_DEV int bcast(const Ctx& c, int v) {
    if (c.nt == 32) { wp::sync(); return v; }
    if (c.tid == 0) c.redi[15] = v;
    wp::team_sync(c.bar, c.nt);
    const int r = c.redi[15];
    wp::team_sync(c.bar, c.nt);
    return r;
}

}  // namespace team
