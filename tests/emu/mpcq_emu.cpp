// TEST HARNESS ONLY - lock-step host emulation of one warp, for unit-testing the device code of
// pympc_quadruped_b200/csrc/mpcq_core.cuh in the CPU-only build container.  The product
// (libmpcq.so) never contains or loads this file: it has no CPU path.  Built by tests/emu/build.py
// into tests/emu/_build/libmpcq_emu.so and used only by `-m "not gpu"` tests.
//
// The threads of a team (1-6 warps) run as coroutines (ucontext) on one thread, switched round-robin at
// every collective; values are exchanged through a double-buffered slot array.
#define MPCQ_HOST_EMU 1
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <ucontext.h>

#include <vector>

#include "../../pympc_quadruped_b200/csrc/mpcq_host.h"
#include "../../pympc_quadruped_b200/csrc/mpcq_legs.cuh"

namespace mpcq_emu {
// A team of NT threads (NT / 32 warps) runs as coroutines on one host thread, switched round-robin at every
// collective.  Warp collectives exchange values among the 32 threads of one warp; the team barrier spans all.
constexpr int MAXT = 512;
static ucontext_t main_ctx, ctx[MAXT];
static int cur = 0, NT = 32;
static uint64_t slots[2][MAXT];
static long gen[MAXT], bgen[MAXT];
static bool finished[MAXT];
static void (*lane_fn)(void*) = nullptr;
static void* lane_arg = nullptr;

int lane_id() { return cur & 31; }
int thread_id() { return cur; }

static void yield_next() {
    int old = cur;
    for (int step = 1; step <= NT; ++step) {
        int nxt = (old + step) % NT;
        if (!finished[nxt]) {
            if (nxt == old) return;
            cur = nxt;
            swapcontext(&ctx[old], &ctx[nxt]);
            return;
        }
    }
    cur = old;
    swapcontext(&ctx[old], &main_ctx);
}

uint64_t exchange(uint64_t v, int src) {
    const long g = gen[cur];
    const int base = cur & ~31;
    slots[g & 1][cur] = v;
    gen[cur] = g + 1;
    for (;;) {
        bool all = true;
        for (int l = base; l < base + 32; ++l)
            if (!finished[l] && gen[l] <= g) { all = false; break; }
        if (all) break;
        yield_next();
    }
    return slots[g & 1][base + src];
}

void team_barrier() {
    const long g = bgen[cur];
    bgen[cur] = g + 1;
    for (;;) {
        bool all = true;
        for (int l = 0; l < NT; ++l)
            if (!finished[l] && bgen[l] <= g) { all = false; break; }
        if (all) break;
        yield_next();
    }
}

static void trampoline() {
    lane_fn(lane_arg);
    finished[cur] = true;
    for (;;) yield_next();
}

void run_team(int nthreads, void (*fn)(void*), void* arg) {
    static std::vector<char> stacks;
    const size_t STK = 512 * 1024;
    if (stacks.empty()) stacks.resize(MAXT * STK);
    NT = nthreads;
    lane_fn = fn;
    lane_arg = arg;
    for (int l = 0; l < NT; ++l) {
        finished[l] = false;
        gen[l] = 0;
        bgen[l] = 0;
        getcontext(&ctx[l]);
        ctx[l].uc_stack.ss_sp = stacks.data() + l * STK;
        ctx[l].uc_stack.ss_size = STK;
        ctx[l].uc_link = &main_ctx;
        makecontext(&ctx[l], trampoline, 0);
    }
    cur = 0;
    swapcontext(&main_ctx, &ctx[0]);
    // every lane of a warp must have gone through the same number of warp collectives, every thread of the team
    // through the same number of barriers: anything else is divergence around a full-mask collective, which
    // deadlocks on real hardware even if this emulation happens to get through it
    for (int l = 0; l < NT; ++l) {
        if (gen[l] != gen[l & ~31] || bgen[l] != bgen[0]) {
            fprintf(stderr, "mpcq_emu: DIVERGENT COLLECTIVES: thread %d did %ld warp collectives / %ld barriers, thread %d did %ld / %ld\n",
                    l, gen[l], bgen[l], l & ~31, gen[l & ~31], bgen[0]);
            abort();
        }
    }
}
}  // namespace mpcq_emu

namespace {
template <class T> struct Job {
    mpcq::Consts cs;
    mpcq::IO<T> io;
    int b;
    char* smem;
    T* lglobal;
    int ncap;
};

template <class T> void lane_entry(void* p) {
    Job<T>* j = static_cast<Job<T>*>(p);
    const mpcq::SizeClass& sc = mpcq::kClasses[j->ncap];
    switch (j->ncap) {
        case 0: mpcq::solve_env<T, 64, MPCQ_NW0>(j->cs, j->io, j->b, j->smem, j->lglobal, sc.ns_lo, sc.ns_hi); break;
        case 1: mpcq::solve_env<T, 128, MPCQ_NW1>(j->cs, j->io, j->b, j->smem, j->lglobal, sc.ns_lo, sc.ns_hi); break;
        case 2: mpcq::solve_env<T, 192, MPCQ_NW2>(j->cs, j->io, j->b, j->smem, j->lglobal, sc.ns_lo, sc.ns_hi); break;
        default: mpcq::solve_env<T, 384, MPCQ_NW3>(j->cs, j->io, j->b, j->smem, j->lglobal, sc.ns_lo, sc.ns_hi); break;
    }
}

template <class T>
int run(const mpcq_config* cfg, int B, const T* x0, const T* yaw, const T* feet, const float* gait, const T* xref,
        T* f_out, T* u_full, int32_t* iters, double* resid, int32_t* status, uint8_t* active,
        const uint8_t* faces_in = nullptr, uint8_t* faces_out = nullptr) {
    mpcq::Consts consts;
    std::string err;
    if (!mpcq::consts_from_config(*cfg, consts, err)) return MPCQ_ERR_INVALID;
    const mpcq::Consts* cs = &consts;
    for (int ci = 0; ci < mpcq::num_classes(cs->horizon); ++ci) {
        const int ncap = mpcq::kClasses[ci].ncap;
        std::vector<char> smem(mpcq::work_bytes<T>(cs->horizon, ncap, true, false, mpcq::class_nmax(mpcq::kClasses[ci]), mpcq::kClasses[ci].nw) + 64);
        for (int b = 0; b < B; ++b) {
            Job<T> j;
            j.cs = *cs;
            j.io = mpcq::IO<T>{x0, yaw, feet, gait, xref, f_out, u_full, iters, resid, status, active, nullptr, B, faces_in, faces_out};
            j.b = b;
            j.smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem.data()) + 31) & ~uintptr_t(31));
            j.lglobal = nullptr;
            j.ncap = ci;
            mpcq_emu::run_team(32 * mpcq::kClasses[ci].nw, lane_entry<T>, &j);
        }
    }
    return 0;
}
}  // namespace

extern "C" {
// read and clear the phase-call counters (thread 0 of every team): 0 setup 1 hess_apply 2 build_slots 3 chol_factor
// 4 tri_solve 5 reduced_gradient 6 apply_step 7 pdas_update 8 reorder_feet 9 solve_env
void mpcq_emu_phase_calls(long* out16) {
    for (int i = 0; i < 16; ++i) { out16[i] = mpcq::g_emu_phase_calls[i]; mpcq::g_emu_phase_calls[i] = 0; }
}
int mpcq_emu_solve_f32(const mpcq_config* cs, int B, const float* x0, const float* yaw, const float* feet, const float* gait,
                       const float* xref, float* f_out, float* u_full, int32_t* iters, double* resid, int32_t* status,
                       uint8_t* active) {
    return run<float>(cs, B, x0, yaw, feet, gait, xref, f_out, u_full, iters, resid, status, active);
}
int mpcq_emu_solve_warm_f32(const mpcq_config* cs, int B, const float* x0, const float* yaw, const float* feet, const float* gait,
                            const float* xref, float* f_out, float* u_full, int32_t* iters, double* resid, int32_t* status,
                            uint8_t* active, const uint8_t* faces_in, uint8_t* faces_out) {
    return run<float>(cs, B, x0, yaw, feet, gait, xref, f_out, u_full, iters, resid, status, active, faces_in, faces_out);
}
int mpcq_emu_solve_f64(const mpcq_config* cs, int B, const double* x0, const double* yaw, const double* feet, const float* gait,
                       const double* xref, double* f_out, double* u_full, int32_t* iters, double* resid, int32_t* status,
                       uint8_t* active) {
    return run<double>(cs, B, x0, yaw, feet, gait, xref, f_out, u_full, iters, resid, status, active);
}
}

// team collectives on their own: every thread of an nw-warp team contributes (vals[tid], tid) to team::reduce_argmin and
// writes what it got back (all threads must agree; a NaN anywhere must win)
namespace {
struct ArgminJob { const double* vals; double* out_v; int* out_tag; int nt; double red[16]; int redi[16]; };
void argmin_entry(void* p) {
    ArgminJob* j = static_cast<ArgminJob*>(p);
    team::Ctx c;
    c.tid = wp::team_tid(j->nt); c.nt = j->nt; c.wid = c.tid >> 5; c.bar = 0; c.red = j->red; c.redi = j->redi;
    double v = j->vals[c.tid];
    int tag = c.tid;
    team::reduce_argmin(c, v, tag);
    j->out_v[c.tid] = v; j->out_tag[c.tid] = tag;
}
}  // namespace
extern "C" void mpcq_emu_team_argmin(int nw, const double* vals, double* out_v, int* out_tag) {
    ArgminJob j{vals, out_v, out_tag, 32 * nw, {}, {}};
    mpcq_emu::run_team(32 * nw, argmin_entry, &j);
}

// the per-leg layer (mpcq_legs.cuh): the device bodies run over all (environment, leg) pairs in index order
extern "C" {
void mpcq_emu_swing_targets(int B, const mpcq_leg_params* lp, const double* pos_base, const double* lin_vel_base, const double* R_base,
                            const double* thighs, const double* pos_feet, const double* swing_state, const double* v_des,
                            const double* yaw_rate, const double* swing_time, const double* stance_time, uint8_t* active,
                            double* remaining, double* foot_init, double* foot_final, double* pos_t, double* vel_t) {
    mpcq::SwingArgs a{pos_base, lin_vel_base, R_base, thighs, pos_feet, swing_state, v_des, yaw_rate, swing_time, stance_time,
                      active, remaining, foot_init, foot_final, pos_t, vel_t,
                      lp->swing_height, lp->dt_control, lp->gravity, lp->foot_z_final, B};
    for (int idx = 0; idx < 4 * B; ++idx) mpcq::swing_leg(a, idx);
}
void mpcq_emu_leg_torques(int B, const mpcq_leg_params* lp, const double* Jv, int ncol, const double* R_base, const double* bpf,
                          const double* bvf, const void* forces, int forces_f64, const double* swing_state, const double* pos_t,
                          const double* vel_t, float* tau) {
    mpcq::TorqueArgs a{Jv, R_base, bpf, bvf, swing_state, pos_t, vel_t, forces, tau, {}, {}, B, ncol};
    for (int i = 0; i < 9; ++i) { a.kp[i] = lp->kp_swing[i]; a.kd[i] = lp->kd_swing[i]; }
    for (int idx = 0; idx < 4 * B; ++idx) {
        if (forces_f64) mpcq::torque_leg<double>(a, idx);
        else mpcq::torque_leg<float>(a, idx);
    }
}
}
