"""BASELINE configs[0]: A1 trot, single robot, horizon 10 - 1000 consecutive MPC updates through the drop-in class
(`ModelPredictiveController`, numpy in / numpy out, one synchronous call per update) on synthetic states; wall time per update.
The CPU reference path (oracle port) is timed on the same states for comparison."""
import sys, time, types
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from pympc_quadruped_b200 import A1Config, Gait, ModelPredictiveController, with_horizon
from pympc_quadruped_b200.synth import synth_states

N, H = 1000, 10
cfg = with_horizon(H)
st = synth_states(N, A1Config, "nominal", seed=20261018)
gait = Gait.TROTTING10.with_horizon(H)
mpc = ModelPredictiveController(cfg, A1Config)
ibm = mpc.iterations_between_mpc
times, forces = [], []
for t in range(N):
    rd = types.SimpleNamespace(quat_base=st["quat_base"][t], pos_base=st["pos_base"][t], ang_vel_base=st["ang_vel_base"][t],
                               lin_vel_base=st["lin_vel_base"][t], pos_base_feet=list(st["pos_base_feet"][t]), R_base=st["R_base"][t])
    gait.set_iteration(ibm, t * ibm)
    tab = gait.get_gait_table()
    t0 = time.perf_counter()
    mpc.update_robot_state(rd)
    f = mpc.update_mpc_if_needed(t * ibm, st["vel_cmd_body"][t], float(st["yaw_rate_cmd"][t]), tab)
    times.append(time.perf_counter() - t0)
    forces.append(f)
ts = np.sort(np.array(times[20:])) * 1e3
print(f"drop-in class, B=1: p50 {ts[len(ts)//2]:.3f} ms  p90 {ts[int(.9*len(ts))]:.3f} ms  mean {ts.mean():.3f} ms per MPC update ({1e3/ts.mean():.0f} updates/s)")

# the same updates through the CPU reference path (restated construction + exact solver), first 200
from oracle.mpc_oracle import OracleMPC, RobotState
from oracle.qp_exact import solve_qp_exact
m = OracleMPC(cfg, A1Config)
tc, worst = [], 0.0
for t in range(200):
    rd = RobotState(st["quat_base"][t], st["pos_base"][t], st["ang_vel_base"][t], st["lin_vel_base"][t], st["pos_base_feet"][t])
    gait.set_iteration(ibm, t * ibm)
    tab = gait.get_gait_table()
    t0 = time.perf_counter()
    m.update_robot_state(rd)
    f = m.update_mpc_if_needed(t * ibm, st["vel_cmd_body"][t], float(st["yaw_rate_cmd"][t]), tab)
    tc.append(time.perf_counter() - t0)
    worst = max(worst, float(np.abs(np.asarray(f) - forces[t]).max()))
tc = np.array(tc[5:]) * 1e3
print(f"CPU reference path (oracle port), one core: mean {tc.mean():.3f} ms per update; max |df| vs the engine over 200 updates {worst:.2e} N")
