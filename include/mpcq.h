/*
 * mpcq.h - C ABI of the B200 batched convex-MPC engine (libmpcq.so).
 *
 * The reference (yinghansun/pympc-quadruped) has no FFI layer: its boundary for this
 * path is the Python class `ModelPredictiveController` (linear_mpc/mpc.py:22).  Each
 * entry point below names the reference method it replaces; the Python host side
 * (pympc_quadruped_b200/controller.py) binds them with ctypes and keeps the reference's
 * class API on top (INTEGRATION.md shows the binding a maintainer would add).
 *
 * Conventions
 *   - every array pointer is a DEVICE pointer owned by the caller (contiguous, row-major,
 *     environment-major) unless the function name ends in `_host`;
 *   - `real` arrays are float when the handle was created with MPCQ_F32, double with
 *     MPCQ_F64; the gait table is always float32 like the reference's (gait.py:87);
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*) and never
 *     synchronise or throw: 0 = success, negative = error (text via mpcq_last_error);
 *   - the factor workspace is allocated by mpcq_create; a launch-order buffer (5 bytes per environment) and the pinned /
 *     device staging of the `*_host` calls grow to the largest batch seen (a growing call synchronises the device);
 *   - mpcq_solve may run its size-class kernels side by side on private streams of the handle; they are forked from and joined
 *     back to `stream` with events, so to the caller the call is still one ordered operation on `stream` (graph-capturable);
 *   - one call in flight per handle: do not overlap two mpcq_solve calls of one handle on different streams;
 *   - a handle is bound to one device and is not thread-safe;
 *   - there is no CPU fallback: without a CUDA device mpcq_create fails.
 */
#ifndef MPCQ_H_
#define MPCQ_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPCQ_VERSION 100            /* 0.1.0 */

enum { MPCQ_F32 = 0, MPCQ_F64 = 1 };

/* error codes */
enum {
    MPCQ_OK = 0,
    MPCQ_ERR_INVALID = -1,          /* bad argument / configuration */
    MPCQ_ERR_CUDA = -2,             /* a CUDA runtime call failed */
    MPCQ_ERR_NO_DEVICE = -3,        /* no usable CUDA device */
    MPCQ_ERR_UNSUPPORTED = -4       /* size beyond what the kernels support */
};

/* per-environment status bits written to `status` */
enum {
    MPCQ_ST_VERIFIED = 1,           /* KKT conditions (primal, dual, stationarity) hold in fp64 on the returned point */
    MPCQ_ST_FALLBACK = 2,           /* the primal active-set fallback ran (primal-dual rounds cycled) */
    MPCQ_ST_MAXITER = 4,            /* iteration cap reached: the point is feasible but not verified optimal */
    MPCQ_ST_NUMERIC = 8,            /* non-finite input, failed factorisation or unreachable residual: forces are returned as 0 */
    MPCQ_ST_NO_STANCE = 32          /* every foot-step is swing: u = 0 */
};

/*
 * Constants of `ModelPredictiveController._load_parameters` (linear_mpc/mpc.py:35-52) and of
 * the config classes it reads (config/linear_mpc_configs.py:4-24, config/robot_configs.py:9-56).
 */
typedef struct mpcq_config {
    int32_t horizon;                /* LinearMpcConfig.horizon (1..32) */
    int32_t dtype;                  /* MPCQ_F32 | MPCQ_F64: factorisation precision and I/O type */
    int32_t device;                 /* CUDA device ordinal */
    int32_t reserved0;
    double dt;                      /* MPC step; the reference hard-codes 0.05 (mpc.py:38) */
    double mu;                      /* friction_coef */
    double fz_max;                  /* RobotConfig.fz_max */
    double mass;                    /* RobotConfig.mass_base */
    double gravity;                 /* LinearMpcConfig.gravity (positive) */
    double inertia[9];              /* RobotConfig.base_inertia_base, row-major (float32 values) */
    double q_diag[13];              /* diag(LinearMpcConfig.Q) */
    double r_diag[12];              /* diag(LinearMpcConfig.R), all > 0 */
    /* solver knobs (0 = default) */
    int32_t max_pdas_rounds;        /* primal-dual active-set rounds before the fallback; default 8 */
    int32_t max_as_iter;            /* fallback active-set iterations; default 12*horizon + 30 */
    int32_t max_refine;             /* preconditioned-CG refinement steps per factorisation; default 20 */
    int32_t schedule;               /* expected-work-first launch order (Q-weighted tracking error, hardest envs first); 0 = on (default), < 0 = natural order */
    double tol_primal;              /* relative feasibility tolerance of the face tests; default 1e-7 (f32) / 1e-9 (f64) */
    double tol_dual;                /* relative multiplier-sign tolerance; default 1e-7 (f32) / 1e-9 (f64) */
    double tol_residual;            /* reduced-gradient tolerance relative to 1+|g|_inf; default 1e-9 (f32) / 1e-12 (f64) */
    double tol_active;              /* slack tolerance of the reported activity; default 1e-6 */
    double tol_residual_loose;      /* reduced-gradient tolerance of intermediate active-set rounds; default 1e-6 (f32) / 1e-9 (f64) */
    double dt_control;              /* LinearMpcConfig.dt_control (used by mpcq_assemble only) */
    double com_height_des;          /* RobotConfig.base_height_des (used by mpcq_assemble only) */
} mpcq_config;

typedef struct mpcq_handle mpcq_handle;

int mpcq_version(void);

/* replaces ModelPredictiveController.__init__ / _load_parameters (mpc.py:24-52) */
int mpcq_create(const mpcq_config* cfg, mpcq_handle** out);
void mpcq_destroy(mpcq_handle* h);
/* last error text of this handle (or of the failed mpcq_create when h == NULL) */
const char* mpcq_last_error(const mpcq_handle* h);

/*
 * replaces ModelPredictiveController._solve_mpc (mpc.py:262-290, drake branch) for B robots:
 * state-space model (:173-192), discretisation (:194-208), condensed cost (:211-235),
 * friction/contact rows (:237-260) and the QP solve (:277-286).
 *
 *   x0      [B,13]   current_state of update_robot_state (mpc.py:55-79): rpy, pos, omega, vel, -g
 *   yaw     [B]      optional (NULL -> x0[:,2]); the reference keeps the unrounded float64 yaw (mpc.py:77)
 *   r_feet  [B,4,3]  pos_base_feet: world-frame base->foot, legs FL,FR,RL,RR (robot_data.py:144-149)
 *   gait    [B,4H]   float32 contact table, step-major/leg-minor, 1 = stance (gait.py:81-100)
 *   x_ref   [B,13H]  generate_reference_trajectory output (mpc.py:110-170)
 *   f_out   [B,12]   first-step ground reaction forces = _solve_mpc(...)[0:12] (mpc.py:99)
 * optional outputs (NULL to skip):
 *   u_full  [B,12H]  the whole optimum
 *   iters   [B,2]    factorisations, fallback active-set iterations
 *   resid   [B,2]    fp64 KKT residuals of the returned point: reduced gradient, primal violation
 *   status  [B]      MPCQ_ST_* bits
 *   active  [B,4H]   per foot-step bit mask of tight rows: bits 0-4 rows 5k..5k+4 at their lower
 *                    bound, bit 5 the fz row at its upper bound (constraint activity)
 */
int mpcq_solve(mpcq_handle* h, int32_t B,
               const void* x0, const void* yaw, const void* r_feet, const float* gait, const void* x_ref,
               void* f_out, void* u_full, int32_t* iters, double* resid, int32_t* status, uint8_t* active,
               void* stream);

/*
 * Warm start of the active-set rounds (no reference counterpart: its solver starts cold at every update, mpc.py:277-286).
 * `faces_in` / `faces_out` are device arrays [B,4H], one byte per foot-step: (sx & 3) | (sy & 3) << 2 | (sz & 3) << 4 with
 * sx, sy in {-1,0,+1} (fx = sx mu fz / free; -1 stored as 3), sz in {-1,0,+1} (f = 0 / free / fz = fz_max); a zero byte = all
 * free, so a zeroed buffer is a cold start.  Until changed, every following mpcq_solve of this handle starts from `faces_in`
 * (NULL = cold) and writes the faces of the point it returns to `faces_out` (NULL = not wanted; zeros for unverified
 * environments).  In a control loop hand the previous update's `faces_out`, shifted by one horizon step (4 bytes), back as
 * `faces_in`: consecutive updates share most of their active set, and a correct guess is verified with ONE factorisation.
 * The optimum is unique, so the result does not depend on the guess (only the number of rounds does).
 * mpcq_solve_host always starts cold.
 */
int mpcq_set_warm_start(mpcq_handle* h, const uint8_t* faces_in, uint8_t* faces_out);

/*
 * replaces, for B robots and in ONE elementwise kernel, everything the reference does between the simulator
 * state and _solve_mpc: ModelPredictiveController.update_robot_state (mpc.py:55-79, quat -> ZYX angles,
 * kinematics.py:40-49), the preamble of update_mpc_if_needed (:83-93, command rotation and the dt_control
 * integrators) and generate_reference_trajectory (:110-170, clamp, roll/pitch compensation, X_ref fill), with the
 * reference's float32 storage / float64 scalar arithmetic.  All inputs and the controller state are float64
 * device arrays (RobotData holds float64):
 *   quat [B,4] (w,x,y,z), pos [B,3], omega [B,3] (world), vel [B,3] (world), R_base [B,9] or NULL (from quat),
 *   v_des_body [B,3], yaw_rate_des [B]
 *   in/out state: xy_des [B,2], yaw_des [B], rp_init [B,2] (roll_init, pitch_init)
 *   first_run = 1: desired pose initialised as mpc.py:84-88 (x = y = 0, yaw = current);  first_run = 2 (no reference counterpart: a
 *   simulator env that was just reset): desired x / y / yaw = the CURRENT pose and the roll / pitch compensation integrators restart
 *   from 0;  do_mpc != 0: this tick recomputes X_ref (mpc.py:95-96)
 * outputs (`real`): x0 [B,13] and yaw [B] always, x_ref [B,13H] when do_mpc != 0 - exactly the arrays mpcq_solve takes.
 */
int mpcq_assemble(mpcq_handle* h, int32_t B,
                  const double* quat, const double* pos, const double* omega, const double* vel, const double* R_base,
                  const double* v_des_body, const double* yaw_rate_des,
                  double* xy_des, double* yaw_des, double* rp_init, int32_t first_run, int32_t do_mpc,
                  void* x0, void* yaw, void* x_ref, void* stream);

/*
 * replaces, for B robots, the contact schedule the reference computes per robot on the host before every MPC update:
 * Gait.set_iteration + Gait.get_gait_table (linear_mpc/gait.py:76-100) and, when asked for, Gait.get_swing_state /
 * get_stance_state (:102-135) - one elementwise kernel.  Device arrays:
 *   stance_offsets [B,4], stance_durations [B,4], num_segment [B] (int32; the Enum payload, gait.py:16-22),
 *   cur_iteration [B] (int32 control tick), iterations_between_mpc (LinearMpcConfig.iteration_between_mpc)
 *   table [B,4H] float32, step-major / leg-minor, 1 = stance, first entry = step t+1 - exactly the `gait` mpcq_solve takes
 *   swing_state, stance_state [B,4] float64 or NULL: phase progress in (0,1], 0 = not in that phase
 * Rows with num_segment < 1 are undefined behaviour like the reference's division by zero; validate on the host.
 */
int mpcq_gait_tables(mpcq_handle* h, int32_t B, const int32_t* stance_offsets, const int32_t* stance_durations,
                     const int32_t* num_segment, const int32_t* cur_iteration, int32_t iterations_between_mpc, float* table,
                     double* swing_state, double* stance_state, void* stream);

/*
 * Constants of the per-leg layer: RobotConfig.Kp_swing / Kd_swing / swing_height (config/robot_configs.py:16-18,35-37,54-56),
 * LinearMpcConfig.dt_control / gravity (config/linear_mpc_configs.py), and the landing height the reference hard-codes
 * (linear_mpc/swing_foot_trajectory_generator.py:117).  Note the reference reads AliengoConfig.swing_height for every robot (:33).
 */
typedef struct mpcq_leg_params {
    double kp_swing[9];             /* row-major 3x3 */
    double kd_swing[9];
    double swing_height;
    double dt_control;
    double gravity;                 /* positive */
    double foot_z_final;            /* -0.0255 in the reference */
} mpcq_leg_params;

/*
 * replaces, for B robots x 4 legs, SwingFootTrajectoryGenerator.set_foot_placement + compute_traj_swingfoot
 * (linear_mpc/swing_foot_trajectory_generator.py:82-129, :65-80, :38-63) as called from the simulator loop
 * (scripts/isaacgym_a1.py:143-158): one elementwise kernel, one thread per leg.  Device arrays, float64:
 *   pos_base, lin_vel_base [B,3], R_base [B,9] (RobotData, utils/robot_data.py:70-76), base_pos_base_thighs [B,4,3] (:178-184),
 *   pos_feet [B,4,3] world frame (:135-142), swing_state [B,4] (Gait.get_swing_state, e.g. from mpcq_gait_tables),
 *   v_des_body [B,3], yaw_rate_des [B], swing_time, stance_time [B] (Gait.swing_time / stance_time, gait.py:68-74)
 *   generator state, in/out, zero-initialised by the caller: swing_active [B,4] uint8 (0 = the next swing tick starts a new swing,
 *   the reference's is_first_swing), remaining_swing_time [B,4], footpos_init, footpos_final [B,4,3] (world frame)
 *   outputs: pos_targets, vel_targets [B,4,3]: swing-foot target relative to the base in the base frame; zero for legs with
 *   swing_state <= 0, whose state is left untouched (the reference only calls the generator for swinging legs).
 * The trajectory is Drake's PiecewisePolynomial.CubicHermite through (init, midpoint at swing_height, final) with zero slopes
 * and float32 break points, evaluated with the time clamped to the break range.
 */
int mpcq_swing_targets(mpcq_handle* h, int32_t B, const mpcq_leg_params* lp, const double* pos_base, const double* lin_vel_base,
                       const double* R_base, const double* base_pos_base_thighs, const double* pos_feet, const double* swing_state,
                       const double* v_des_body, const double* yaw_rate_des, const double* swing_time, const double* stance_time,
                       uint8_t* swing_active, double* remaining_swing_time, double* footpos_init, double* footpos_final,
                       double* pos_targets, double* vel_targets, void* stream);

/*
 * replaces, for B robots, LegController.update (linear_mpc/leg_controller.py:38-91): stance legs tau = Jv^T (-f), swing legs
 * tau = Jv^T (Kp (R p_des - R p) + Kd (R v_des - R v)), joint torques float32 [B,12] like the reference's torque_cmds.
 *   Jv_feet [B,4,3,ncol] float64: ncol = 18 is RobotData.Jv_feet (utils/robot_data.py:119-133; columns 6+3 leg .. 6+3 leg+2 are
 *   used, leg_controller.py:84,88), ncol = 3 is those joint blocks alone
 *   R_base [B,9], base_pos_base_feet, base_vel_base_feet [B,4,3] float64 (utils/robot_data.py:151-167)
 *   contact_forces [B,12] `real` (the f_out of mpcq_solve), swing_state [B,4] float64 (non-zero = swing),
 *   pos_targets, vel_targets [B,4,3] float64 (mpcq_swing_targets)
 */
int mpcq_leg_torques(mpcq_handle* h, int32_t B, const mpcq_leg_params* lp, const double* Jv_feet, int32_t ncol, const double* R_base,
                     const double* base_pos_base_feet, const double* base_vel_base_feet, const void* contact_forces,
                     const double* swing_state, const double* pos_targets, const double* vel_targets, float* torque_cmds,
                     void* stream);

/*
 * Same call with HOST buffers (what a CPU-side simulator loop such as scripts/isaacgym_a1.py:119-164
 * would hand over): inputs are staged through pinned memory, copied to the device, solved and the
 * requested outputs copied back (page-locked result buffers are written in place by the kernels, without a copy);
 * returns after the results are in the caller's buffers.
 */
int mpcq_solve_host(mpcq_handle* h, int32_t B,
                    const void* x0, const void* yaw, const void* r_feet, const float* gait, const void* x_ref,
                    void* f_out, void* u_full, int32_t* iters, double* resid, int32_t* status, uint8_t* active);

/*
 * One MPC update from HOST state: for B robots, the body of the simulator loop up to the forces -
 * Gait.set_iteration + get_gait_table (linear_mpc/gait.py:76-100), ModelPredictiveController.update_robot_state
 * (mpc.py:55-79) and update_mpc_if_needed on an MPC tick (:81-108), i.e. scripts/isaacgym_a1.py:119-144 without the
 * per-robot Python loop.  Only the RobotData fields cross the bus (272 B per robot instead of the 784 B of assembled
 * state + reference trajectory + contact table that mpcq_solve_host takes); the gait schedule, state assembly,
 * reference trajectory and the solve run on the device (mpcq_gait_tables + mpcq_assemble + mpcq_solve kernels).
 *   state_cmd   [B,29] float64 host: quat (w,x,y,z) 4 | pos_base 3 | ang_vel_base 3 | lin_vel_base 3 (world) |
 *               pos_base_feet 12 (world-frame base->foot, FL FR RL RR) | v_des_body 3 | yaw_rate_des
 *   gait_params [B,10] int32 host: stance_offsets 4 | stance_durations 4 | num_segment | cur_iteration (control tick)
 *   first_run: 0 = regular tick, 1 = mpc.py:84-88, 2 = respawn (desired pose = current pose, integrators zeroed); see mpcq_assemble
 *   f_out [B,12] `real` host, status [B] int32 host or NULL; page-locked buffers are read by DMA / written in place.
 * The controller state of mpc.py (x/y/yaw desired, roll/pitch compensation integrators) is kept per robot inside the
 * handle between calls (slot i = robot i; a larger B than before re-allocates and zeroes it; mpcq_tick_reset zeroes it).
 * Returns after the results are in the caller's buffers.
 */
int mpcq_tick_host(mpcq_handle* h, int32_t B, const double* state_cmd, const int32_t* gait_params,
                   int32_t iterations_between_mpc, int32_t first_run, void* f_out, int32_t* status);
int mpcq_tick_reset(mpcq_handle* h);
/*
 * The same tick, asynchronous, on one of TWO independent pipelines (slot 0 / 1: own device buffers, own controller state, own
 * stream): submit returns as soon as the copies and kernels are queued, wait returns when the results of that slot are in the
 * caller's buffers.  Two groups of robots can thus alternate - the transfers and the small kernels of one group hide behind the
 * solve of the other (e.g. two simulator instances stepping while the other one's forces are computed).  All four buffers must be
 * page-locked (MPCQ_ERR_INVALID otherwise); mpcq_tick_host is submit + wait on slot 0.  The two slots may carry different B.
 * The caller's buffers of a slot belong to the library from submit to wait.  Submitting to a slot that is still in flight waits
 * for it first; the synchronous host entry points (mpcq_tick_host, mpcq_solve_host) and mpcq_tick_reset first finish whatever is
 * in flight on either slot.
 */
int mpcq_tick_host_submit(mpcq_handle* h, int32_t slot, int32_t B, const double* state_cmd, const int32_t* gait_params,
                          int32_t iterations_between_mpc, int32_t first_run, void* f_out, int32_t* status);
int mpcq_tick_host_wait(mpcq_handle* h, int32_t slot);

/*
 * Stage entry point for parity tests: the QP data the reference would hand to the solver.
 *   H_out [B,12H,12H] (_generate_QP_cost :232), g_out [B,12H] (:233), ub_out [B,20H] (:248-258;
 *   lb is identically 0 and C = kron(I_4H, pyramid(mu)) is constant).  Always float64; +inf in
 *   ub_out marks the one-sided friction rows exactly like the reference.
 */
int mpcq_build_qp(mpcq_handle* h, int32_t B,
                  const void* x0, const void* yaw, const void* r_feet, const float* gait, const void* x_ref,
                  double* H_out, double* g_out, double* ub_out, void* stream);

/* number of kernels the last mpcq_solve / mpcq_build_qp call launched (for bench.py's gpu_launches) */
int mpcq_last_launch_count(const mpcq_handle* h);

/*
 * Measurement hooks (no reference counterpart; the reference only prints time.time() deltas, mpc.py:98-101).
 * With profiling on, mpcq_solve brackets every kernel launch with CUDA events on the launching stream.
 * mpcq_last_kernel_ms synchronises on those events and writes the duration of each launch of the last
 * mpcq_solve call (one per size class, in class order) into ms[0..cap); returns the number of launches.
 */
int mpcq_set_profiling(mpcq_handle* h, int32_t enable);
int mpcq_last_kernel_ms(mpcq_handle* h, float* ms, int32_t cap);
/*
 * Denominators for the roofline of the solve kernel that MEASURED_PEAKS.json does not hold: runs three
 * saturating micro-kernels on `device` (about 50 ms in total, synchronous) and writes
 *   out4[0] = FP32 FMA TFLOP/s, out4[1] = FP64 FMA TFLOP/s, out4[2] = shared-memory load GB/s (LDS.128), out4[3] = SM count.
 */
int mpcq_measure_peaks(int32_t device, double* out4);

#ifdef __cplusplus
}
#endif
#endif /* MPCQ_H_ */
