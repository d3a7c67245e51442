"""The bench's input sets whose launch is slow: per set the launch time and the robots with the most rounds / fallback iterations;
saves the inputs of the worst robots to gpurun_out/stragglers.npz (to be traced on the warp emulator)."""
import sys, os, argparse
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from pympc_quadruped_b200.controller import BatchedModelPredictiveController, BatchedRobotData
from pympc_quadruped_b200.configs import with_horizon
from pympc_quadruped_b200.synth import SEED_BASE
a = argparse.Namespace(robot="A1Config", gait="trot", regime="mixed", envs=4096, horizon=10, dtype="f32", sets=64)
dev = torch.device("cuda", 0)
robot, st, tabs = bench.host_states(a, SEED_BASE + 2)
H, B, S = a.horizon, a.envs, a.sets
n = B * S
ctrl = BatchedModelPredictiveController(with_horizon(H), robot, n, device=dev, dtype=torch.float32)
eng = ctrl.engine
ctrl.update_robot_state(BatchedRobotData(st["quat_base"], st["pos_base"], st["ang_vel_base"], st["lin_vel_base"], st["pos_base_feet"], st["R_base"]))
vcmd = torch.as_tensor(st["vel_cmd_body"], device=dev)
vel = torch.einsum("bij,bj->bi", torch.as_tensor(st["R_base"], device=dev), vcmd)
pos32 = torch.as_tensor(st["pos_base"], device=dev).float().double()
ctrl._xy_des.copy_(pos32[:, 0:2] - ctrl.dt_control * vel[:, 0:2])
ctrl.is_first_run = False
eng.assemble(ctrl._quat, ctrl._pos, ctrl._omega, ctrl._vel, vcmd.contiguous(), torch.as_tensor(st["yaw_rate_cmd"], device=dev).contiguous(),
             ctrl._xy_des, ctrl.yaw_desired, ctrl._rp_init, False, True, ctrl.current_state, ctrl.yaw, ctrl.ref_traj, R_base=ctrl._R_given)
x0 = ctrl.current_state.reshape(S, B, 13); yaw = ctrl.yaw.reshape(S, B); feet = ctrl.pos_base_feet.float().reshape(S, B, 12).contiguous()
xref = ctrl.ref_traj.reshape(S, B, 13 * H); gait = torch.as_tensor(tabs, device=dev).reshape(S, B, 4 * H).contiguous()
rows = []
res = eng.solve(x0[0], feet[0], gait[0], xref[0], yaw=yaw[0], want=("iters", "status"))
for k in range(S):
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); eng.solve(x0[k], feet[k], gait[k], xref[k], yaw=yaw[k], out=res); e1.record(); torch.cuda.synchronize()
    it = res.iters.cpu().numpy(); stt = res.status.cpu().numpy()
    rows.append((e0.elapsed_time(e1), k, it.copy(), stt.copy()))
rows.sort(key=lambda r: -r[0])
save = {}
for ms, k, it, stt in rows[:8]:
    cost = it[:, 0] + 3 * it[:, 1]
    w = np.argsort(-cost)[:3]
    print(f"set {k}: {ms:.3f} ms | worst robots {[(int(b), int(it[b,0]), int(it[b,1]), int(stt[b])) for b in w]}")
    b = int(w[0])
    save[f"s{k}_x0"] = x0[k, b].cpu().numpy(); save[f"s{k}_yaw"] = yaw[k, b].cpu().numpy(); save[f"s{k}_feet"] = feet[k, b].cpu().numpy()
    save[f"s{k}_gait"] = gait[k, b].cpu().numpy(); save[f"s{k}_xref"] = xref[k, b].cpu().numpy()
print("median set:", f"{rows[len(rows)//2][0]:.3f} ms")
np.savez("gpurun_out/stragglers.npz", **save)
