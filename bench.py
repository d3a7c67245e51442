#!/usr/bin/env python
"""Benchmark of the batched convex-MPC hot path (BASELINE.json metric: MPC QP solves/sec at
4096 robots x H=10 on 1/2/4/8 B200; p50 step latency).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA engine through the C ABI)
  python bench.py --impl reference ...                           the reference's per-robot CPU loop on the host cores
  torchrun --nproc-per-node N bench.py --gpus N ...              one rank per GPU (weak scaling: --envs per GPU)

A step = one pass of the hot path (QP build + solve, 12 GRFs out) over one batch of --envs synthetic
robots per GPU.  Consecutive steps use different input sets (--sets of them, > L2 in total).
Prints ONE JSON line (rank 0).  Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mpc_qp_solves_per_sec"
UNIT = "solves/s"
# algorithmic HBM bytes of one solve: 4*(13 x0 + 1 yaw + 12 feet + 13H x_ref) + 4*4H gait + 4*12 out
ALGO_BYTES = lambda H: 4 * (13 + 1 + 12 + 13 * H) + 16 * H + 48


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--envs", type=int, default=4096, help="robots per GPU per step")
    ap.add_argument("--sets", type=int, default=64, help="distinct input sets rotated through (64 x 4096 x 708 B = 186 MB > L2)")
    ap.add_argument("--horizon", type=int, default=10)
    ap.add_argument("--robot", default="A1Config")
    ap.add_argument("--regime", default="mixed", choices=("mixed", "nominal", "aggressive"))
    ap.add_argument("--gait", default="trot", choices=("trot", "mix", "stand"))
    ap.add_argument("--dtype", default="f32", choices=("f32", "f64"))
    ap.add_argument("--gather", action="store_true", help="include the optional NCCL all-gather of GRFs in the step")
    ap.add_argument("--config", type=int, default=None, choices=(0, 1, 2, 3, 4),
                    help="BASELINE.json configs[k] preset (robot / gait / envs / horizon / dtype and the workload label); "
                         "default: configs[1], the configuration the metric is quoted on")
    ap.add_argument("--total-envs", type=int, default=0,
                    help="strong scaling: this many robots in total, sharded contiguously over the ranks (configs[4]: 262144)")
    ap.add_argument("--lean", action="store_true", help="only the headline measurements (device-resident, e2e, kernel time, roofline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="envs in the CPU-baseline sample (0 = ~20 s of CPU work)")
    a = ap.parse_args()
    # BASELINE.json configs[k].  Explicit flags given together with --config are overridden by the preset.
    presets = {
        0: dict(robot="A1Config", gait="trot", regime="nominal", envs=1, horizon=10, dtype="f32", sets=1000, steps=1000, lean=True),
        1: dict(robot="A1Config", gait="trot", regime="mixed", envs=4096, horizon=10, dtype="f32"),
        2: dict(robot="AliengoConfig", gait="mix", regime="mixed", envs=16384, horizon=10, dtype="f64", sets=16, lean=True),
        3: dict(robot="A1Config", gait="trot", regime="mixed", envs=4096, horizon=30, dtype="f32", sets=16, steps=50, lean=True),
        4: dict(robot="A1Config", gait="trot", regime="mixed", horizon=10, dtype="f32", total_envs=262144, sets=2, steps=20, lean=True),
    }
    a.config_index = 1 if a.config is None else a.config
    if a.config is not None:
        given = {t.split("=")[0] for t in sys.argv[1:] if t.startswith("--")}
        for k, v in presets[a.config].items():
            if k in ("steps", "sets") and ("--" + k) in given:
                continue                                            # an explicit --steps / --sets wins over the preset
            setattr(a, k, v)
    a.scaling = "weak"
    if a.total_envs:
        world = int(os.environ.get("WORLD_SIZE", "1")) if a.impl == "ours" else max(1, a.gpus)
        if a.total_envs % world:
            ap.error("--total-envs must be divisible by the number of ranks")
        a.envs = a.total_envs // world
        a.scaling = "strong"
    return a


def gaits_for(name):
    from pympc_quadruped_b200.gait import Gait
    from pympc_quadruped_b200.synth import GAIT_MIX
    return {"trot": (Gait.TROTTING10,), "mix": GAIT_MIX, "stand": (Gait.STANDING,)}[name]


def workload_name(a):
    default = dict(robot="A1Config", gait="trot", regime="mixed", envs=4096, horizon=10, dtype="f32")
    tag = f"BASELINE configs[{a.config_index}]" if (a.config is not None or all(getattr(a, k) == v for k, v in default.items())) \
        else "custom workload, not a BASELINE config"
    size = f"{a.total_envs} envs in total over the ranks ({a.envs} per GPU)" if a.total_envs else f"{a.envs} envs/GPU"
    if a.config_index == 0 and a.config is not None:
        size = "ONE robot, 1000 consecutive MPC updates over a synthetic recorded trajectory"
    return f"{a.robot[:-6]} {a.gait} {a.regime} states, {size}, horizon {a.horizon}, {a.dtype} ({tag})"


def config_dict(a, world):
    """The `config` object of the JSON line: identical keys and values in both arms (ours / reference)."""
    return {"workload": workload_name(a), "baseline_config": a.config_index if a.config is not None else None,
            "envs_per_gpu": a.envs, "total_envs": a.envs * world, "horizon": a.horizon, "robot": a.robot, "gait": a.gait,
            "regime": a.regime}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an in-process NVML polling thread (2 ms period, so
    that a 0.2 s timed region still yields ~100 samples); `nvidia-smi -lms` as the fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, uuid=None):
        self.idx = gpu_index
        self.uuid = uuid
        self.proc = None
        self.thread = None
        self.samples = []          # (sm_mhz, reasons bit mask)
        self.sm_max = None
        self._stop = False
        self.source = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        if self.uuid is not None:
            try:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(self.uuid)).encode())
            except Exception:
                pass
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = self.idx
        if vis:
            ent = vis.split(",")[self.idx].strip()
            if ent.isdigit():
                idx = int(ent)
            else:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(ent.encode())
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def start(self):
        import threading
        try:
            nv, h = self._nvml_handle()
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons

            def poll():
                while not self._stop:
                    try:
                        self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(reasons_fn(h))))
                    except Exception:
                        pass
                    time.sleep(0.002)

            self._nv = nv
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            self.source = "nvml"
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
        except OSError:
            self.proc = None

    @staticmethod
    def _summary(sm, mx, reasons, source):
        # the first samples may precede the load; take the median of the upper half
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm), "source": source}

    def stop(self):
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            nv = self._nv
            names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            reasons = {k for k, bit in names.items() if any(r & bit for _, r in self.samples)}
            return self._summary([s for s, _ in self.samples], self.sm_max, reasons, "nvml")
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            p = [x.strip() for x in line.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return self._summary(sm, max(mx) if mx else None, reasons, "nvidia-smi")


def host_states(a, seed):
    from pympc_quadruped_b200 import configs
    from pympc_quadruped_b200.synth import synth_gait_tables, synth_states
    robot = getattr(configs, a.robot)
    n = a.envs * a.sets
    st = synth_states(n, robot, a.regime, seed=seed)
    tabs = synth_gait_tables(n, a.horizon, gaits_for(a.gait), seed=seed)
    return robot, st, tabs


def run_reference(a):
    """The reference's Python loop (restated; Drake replaced by the exact oracle solver) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle.cpu_baseline import run_parallel
    from pympc_quadruped_b200.synth import SEED_BASE
    cores = os.cpu_count() or 1
    per_step = max(cores * 24, 96)                                  # bounded sample of the workload per step
    a2 = argparse.Namespace(**vars(a))
    a2.envs, a2.sets = per_step, a.steps + a.warmup
    robot, st, tabs = host_states(a2, SEED_BASE + 2)
    keys = ("quat_base", "pos_base", "ang_vel_base", "lin_vel_base", "pos_base_feet", "vel_cmd_body", "yaw_rate_cmd")
    times = []
    with mp.get_context("spawn").Pool(cores) as pool:
        for s in range(a.warmup + a.steps):
            sl = slice(s * per_step, (s + 1) * per_step)
            _, wall, tb, ts = run_parallel(pool, a.horizon, a.robot, {k: st[k][sl] for k in keys}, tabs[sl], cores)
            if s >= a.warmup:
                times.append(wall)
    total = sum(times)
    value = per_step * a.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": a.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(config_dict(a, max(1, a.gpus)), sample_per_step=per_step,
                       note="reference construction restated in numpy (pinned bit-for-bit to the reference) + exact fp64 "
                            "solver; Drake/OSQP is not installable offline; each step = a bounded sample of the workload"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} envs per step x {a.steps} steps of the same seeded workload"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(a):
    import torch
    import torch.distributed as dist
    from pympc_quadruped_b200.controller import BatchedModelPredictiveController, BatchedRobotData
    from pympc_quadruped_b200.configs import with_horizon
    from pympc_quadruped_b200.synth import SEED_BASE

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line: whatever NCCL prints (the version banner with NCCL_DEBUG=VERSION, INFO logs)
        # goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    tdt = torch.float64 if a.dtype == "f64" else torch.float32
    H, B, S = a.horizon, a.envs, a.sets

    # ---- synthetic inputs: seeded states -> the controller's own state / reference-trajectory code (device)
    robot, st, tabs = host_states(a, SEED_BASE + 2 + 1000 * rank)
    n = B * S
    ctrl = BatchedModelPredictiveController(with_horizon(H), robot, n, device=dev, dtype=tdt)
    eng = ctrl.engine
    ctrl.update_robot_state(BatchedRobotData(st["quat_base"], st["pos_base"], st["ang_vel_base"], st["lin_vel_base"],
                                             st["pos_base_feet"], st["R_base"]))
    # desired xy = current xy: preset the integrators so that this (non-first) tick lands there, then let the
    # fused kernel (mpcq_assemble) do state assembly + command integration + reference trajectory on the device
    vcmd = torch.as_tensor(st["vel_cmd_body"], device=dev)
    vel = torch.einsum("bij,bj->bi", torch.as_tensor(st["R_base"], device=dev), vcmd)
    pos32 = torch.as_tensor(st["pos_base"], device=dev).float().double()
    ctrl._xy_des.copy_(pos32[:, 0:2] - ctrl.dt_control * vel[:, 0:2])
    ctrl.is_first_run = False
    eng.assemble(ctrl._quat, ctrl._pos, ctrl._omega, ctrl._vel, vcmd.contiguous(),
                 torch.as_tensor(st["yaw_rate_cmd"], device=dev).contiguous(), ctrl._xy_des, ctrl.yaw_desired, ctrl._rp_init,
                 False, True, ctrl.current_state, ctrl.yaw, ctrl.ref_traj, R_base=ctrl._R_given)
    xref = ctrl.ref_traj
    x0 = ctrl.current_state.to(tdt).reshape(S, B, 13).contiguous()
    yaw = ctrl.yaw.to(tdt).reshape(S, B).contiguous()
    feet = ctrl.pos_base_feet.to(tdt).reshape(S, B, 12).contiguous()
    xref = xref.to(tdt).reshape(S, B, 13 * H).contiguous()
    gait = torch.as_tensor(tabs, device=dev).reshape(S, B, 4 * H).contiguous()
    from pympc_quadruped_b200.engine import SolveResult
    out = SolveResult(forces=torch.empty((B, 12), dtype=tdt, device=dev), u=None,
                      iters=torch.empty((B, 2), dtype=torch.int32, device=dev), resid=None,
                      status=torch.empty((B,), dtype=torch.int32, device=dev), active=None)
    gathered = torch.empty((world * B, 12), dtype=tdt, device=dev) if (a.gather and world > 1) else None

    def step(s):
        k = s % S
        eng.solve(x0[k], feet[k], gait[k], xref[k], yaw=yaw[k], out=out)
        if gathered is not None:
            dist.all_gather_into_tensor(gathered, out.forces)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing
    for s in range(a.warmup):
        step(s)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    sampler = ClockSampler(local, getattr(torch.cuda.get_device_properties(local), 'uuid', None))
    barrier()
    sampler.start()
    t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_all0.record()
    for s in range(a.steps):
        ev[s][0].record()
        step(a.warmup + s)
        ev[s][1].record()
    t_all1.record()
    barrier()
    clocks = sampler.stop()
    total_ms = t_all0.elapsed_time(t_all1)
    lat_seq = [e0.elapsed_time(e1) for e0, e1 in ev]
    if os.environ.get("MPCQ_BENCH_DUMP_LAT"):
        print("lat_ms_by_step", " ".join(f"{v:.3f}" for v in lat_seq), file=sys.stderr)
    lat = sorted(lat_seq)
    launches_per_step = eng.last_launch_count
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax.item())
    value = world * B * a.steps / (total_ms * 1e-3)

    # ---- solver statistics over every input set (outside the timed region)
    facts, fallback, unverified = [], 0, 0
    for k in range(S):
        eng.solve(x0[k], feet[k], gait[k], xref[k], yaw=yaw[k], out=out)
        it = out.iters[:, 0].double()
        facts.append(it.cpu().numpy())
        fallback += int(((out.status & 2) != 0).sum())
        unverified += int(((out.status & 1) == 0).sum())
    facts = np.concatenate(facts)

    # ---- dominant-kernel duration: CUDA events on the launching stream around each class launch
    eng.set_profiling(True)
    kms = []
    for s in range(a.steps):
        step(a.warmup + s)
        kms.append(eng.last_kernel_ms())
    eng.set_profiling(False)
    kms = np.array(kms)                                            # [steps, classes]
    kmean = kms.mean(axis=0)
    dom = int(np.argmax(kmean))

    graph_replay = None
    if not a.lean:
        # ---- the same steps replayed from CUDA graphs (one graph per input set: mpcq_solve is one capturable operation on the
        # caller's stream, forked class streams included); information beside the headline, which times plain launches
        graph_replay = None
        try:
            graphs = []
            for k in range(min(S, 16)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    eng.solve(x0[k], feet[k], gait[k], xref[k], yaw=yaw[k], out=out)
                graphs.append(g)
            for s in range(a.warmup):
                graphs[s % len(graphs)].replay()
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for s in range(a.steps):
                graphs[(a.warmup + s) % len(graphs)].replay()
            g1.record()
            barrier()
            gms = g0.elapsed_time(g1) / a.steps
            graph_replay = {"ms_per_step": gms, "solves_per_s": world * B / (gms * 1e-3), "graphs": len(graphs),
                            "note": "per-rank time, not reduced over ranks; first 16 input sets"}
            del graphs
        except Exception as ex:                                       # never let the extra measurement break the bench line
            graph_replay = {"error": repr(ex)[:200]}
            torch.cuda.synchronize(dev)

    # ---- end to end through the C ABI with HOST buffers (mpcq_solve_host: pinned staging, H2D, solve, D2H)
    # inputs live in pinned host memory (one array per input, all sets), results land in pinned host arrays
    pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t).numpy()
    hx0, hyaw, hfeet, hxref, hgait = (pin(t) for t in (x0, yaw, feet, xref, gait))
    hout = {"forces": torch.empty((B, 12), dtype=tdt, pin_memory=True).numpy(),
            "status": torch.empty((B,), dtype=torch.int32, pin_memory=True).numpy()}
    rs = 8 if a.dtype == "f64" else 4
    h2d = B * (rs * (13 + 1 + 12 + 13 * H) + 16 * H)
    d2h = B * (12 * rs + 4)
    for s in range(a.warmup):
        eng.solve_host(hx0[s % S], hfeet[s % S], hgait[s % S], hxref[s % S], yaw=hyaw[s % S], out=hout)
    barrier()
    t0 = time.perf_counter()
    for s in range(a.steps):
        k = (a.warmup + s) % S
        r = eng.solve_host(hx0[k], hfeet[k], hgait[k], hxref[k], yaw=hyaw[k], out=hout)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_launches = eng.last_launch_count
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * a.steps / float(te.item())
    assert np.array_equal(r["forces"], hout["forces"])
    sh_e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * float(te.item()) / a.steps,
              "launches_per_step": e2e_launches,
              "api": "mpcq_solve_host (C ABI, pinned host buffers: assembled state + reference trajectory + contact table in, forces out)"}

    # ---- the headline end-to-end number: mpcq_tick_host, the reference's per-robot loop body (gait schedule, update_robot_state,
    # update_mpc_if_needed on an MPC tick) for the batch, from HOST RobotData fields: 272 B per robot over the bus, gait
    # table + state assembly + reference trajectory + solve on the device, forces + status written into page-locked host arrays
    from pympc_quadruped_b200.synth import synth_gait_params
    goff, gdur, gseg, git = synth_gait_params(n, gaits_for(a.gait), seed=SEED_BASE + 2 + 1000 * rank)
    ibm = int(ctrl.iterations_between_mpc)
    sc = np.zeros((n, 29))
    sc[:, 0:4], sc[:, 4:7], sc[:, 7:10], sc[:, 10:13] = st["quat_base"], st["pos_base"], st["ang_vel_base"], st["lin_vel_base"]
    sc[:, 13:25] = st["pos_base_feet"].reshape(n, 12)
    sc[:, 25:28], sc[:, 28] = st["vel_cmd_body"], st["yaw_rate_cmd"]
    gp = np.concatenate([goff, gdur, gseg[:, None], (git * ibm)[:, None]], axis=1).astype(np.int32)
    hsc = torch.empty((S, B, 29), dtype=torch.float64, pin_memory=True); hsc.copy_(torch.as_tensor(sc).reshape(S, B, 29)); hsc = hsc.numpy()
    hgp = torch.empty((S, B, 10), dtype=torch.int32, pin_memory=True); hgp.copy_(torch.as_tensor(gp).reshape(S, B, 10)); hgp = hgp.numpy()
    # every step is a batch of freshly (re)spawned robots (first_run = 2: desired pose = current pose, integrators at zero) -
    # the synthetic states of consecutive steps are unrelated, and this is exactly the SURVEY 8d workload the device-resident
    # number is measured on
    for s in range(a.warmup):
        eng.tick_host(hsc[s % S], hgp[s % S], ibm, first_run=2, out=hout, validate=False)
    barrier()
    t0 = time.perf_counter()
    for s in range(a.steps):
        k = (a.warmup + s) % S
        r = eng.tick_host(hsc[k], hgp[k], ibm, first_run=2, out=hout, validate=False)
    barrier()
    tick_s = time.perf_counter() - t0
    tick_launches = eng.last_launch_count
    assert np.all(r["status"] & 1), "mpcq_tick_host returned unverified robots"
    te = torch.tensor([tick_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * a.steps / float(te.item())
    h2d, d2h = B * (29 * 8 + 10 * 4), B * (12 * rs + 4)
    e2e_launches = tick_launches
    # the same ticks through the asynchronous two-slot form: batch k+1 is submitted (H2D + kernels queued) before batch k is
    # waited for, so the transfers and small kernels of one batch hide behind the solve of the other - two groups of robots
    # alternating; every batch still crosses the bus both ways inside the timed region
    houts = [hout, {"forces": torch.empty((B, 12), dtype=tdt, pin_memory=True).numpy(),
                    "status": torch.empty((B,), dtype=torch.int32, pin_memory=True).numpy()}]
    def pipelined(nsteps):
        eng.tick_submit(0, hsc[a.warmup % S], hgp[a.warmup % S], ibm, 2, houts[0])
        for s in range(nsteps):
            if s + 1 < nsteps:
                k = (a.warmup + s + 1) % S
                eng.tick_submit((s + 1) & 1, hsc[k], hgp[k], ibm, 2, houts[(s + 1) & 1])
            eng.tick_wait(s & 1)
    pipelined(max(a.warmup, 2))
    barrier()
    t0 = time.perf_counter()
    pipelined(a.steps)
    barrier()
    tp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    e2e_pipe = {"value": world * B * a.steps / float(tp.item()), "unit": UNIT, "ms_per_step": 1e3 * float(tp.item()) / a.steps,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "mpcq_tick_host_submit / mpcq_tick_host_wait on two alternating slots (two groups of robots): batch k+1 is queued "
                       "before batch k is waited for; the headline e2e above is the synchronous call"}

    ctrl_ms, robot_tick, closed_loop = None, None, None
    if not a.lean:
        # ---- the batched controller API on device tensors: gait schedule + update_robot_state + update_mpc_if_needed (gait kernel,
        # fused assembly kernel, solve), i.e. the call sequence of scripts/isaacgym_a1.py:136-143 for the whole batch
        c2 = BatchedModelPredictiveController(with_horizon(H), robot, B, device=dev, dtype=tdt)
        dv = lambda key, k: torch.as_tensor(st[key][k * B:(k + 1) * B], device=dev)
        rds = [BatchedRobotData(dv("quat_base", k), dv("pos_base", k), dv("ang_vel_base", k), dv("lin_vel_base", k),
                                dv("pos_base_feet", k), dv("R_base", k)) for k in range(min(S, 8))]
        cmds = [(dv("vel_cmd_body", k), dv("yaw_rate_cmd", k)) for k in range(min(S, 8))]
        # contact schedule on the device too (mpcq_gait_tables): same per-env patterns and phases as the precomputed tables
        from pympc_quadruped_b200.synth import synth_gait_params
        goff, gdur, gseg, git = synth_gait_params(n, gaits_for(a.gait), seed=SEED_BASE + 2 + 1000 * rank)
        ibm = int(c2.iterations_between_mpc)
        i32 = lambda v: torch.as_tensor(np.ascontiguousarray(v).astype(np.int32), device=dev)
        gparams = [(i32(goff[k * B:(k + 1) * B]), i32(gdur[k * B:(k + 1) * B]), i32(gseg[k * B:(k + 1) * B]),
                    i32(git[k * B:(k + 1) * B] * ibm)) for k in range(min(S, 8))]
        gtab = torch.empty((B, 4 * H), dtype=torch.float32, device=dev)

        def ctrl_step(s):
            k = s % len(rds)
            eng2 = c2.engine
            eng2.gait_tables(gparams[k][0], gparams[k][1], gparams[k][2], gparams[k][3], ibm, table=gtab)
            c2.update_robot_state(rds[k])
            return c2.update_mpc_if_needed(0, cmds[k][0], cmds[k][1], gtab)
        ctrl_step(0)
        assert torch.equal(gtab, gait[0]), "device gait tables differ from the host tables of the same schedule"
        for s in range(a.warmup):
            ctrl_step(s)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for s in range(a.steps):
            ctrl_step(a.warmup + s)
        c1.record()
        barrier()
        ctrl_ms = c0.elapsed_time(c1) / a.steps

        # ---- the whole loop body of scripts/isaacgym_a1.py:136-162 for the batch on the device: controller tick above + swing-foot
        # targets + joint torque map (mpcq_swing_targets, mpcq_leg_torques); synthetic kinematics (random foot
        # Jacobians in the reference's 3x18 layout, thigh / foot positions) since pinocchio is the caller's side of the boundary
        from pympc_quadruped_b200 import BatchedGaitSchedule, BatchedLegController, BatchedLegKinematics, BatchedSwingFootTrajectoryGenerator
        from pympc_quadruped_b200.gait import GaitSchedule
        f64 = torch.float64
        lrng = torch.Generator(device=dev); lrng.manual_seed(SEED_BASE + 77 + rank)
        ru = lambda *sh: torch.rand(*sh, generator=lrng, device=dev, dtype=f64) - 0.5
        gs = BatchedGaitSchedule(c2.engine, [GaitSchedule("bench", int(gseg[i]), goff[i], gdur[i], horizon=H) for i in range(B)])
        kins = [BatchedLegKinematics(rd.pos_base, rd.lin_vel_base, rd.R_base, 0.4 * ru(B, 4, 3), rd.pos_base[:, None, :] + rd.pos_base_feet,
                                     ru(B, 4, 3), 2 * ru(B, 4, 3), ru(B, 4, 3, 18)) for rd in rds]
        swing_gen = BatchedSwingFootTrajectoryGenerator(c2.engine, B, robot_config=robot)
        leg_ctrl = BatchedLegController(c2.engine, B, robot.Kp_swing, robot.Kd_swing)
        tick0 = gparams[0][3].clone()

        def leg_step(s, forces):
            k = s % len(rds)
            gs.set_iteration(ibm, tick0 + s)                                           # swing states advance one control tick per step
            pt, vt = swing_gen.update(kins[k], gs, cmds[k][0], cmds[k][1])
            return leg_ctrl.update(kins[k], forces, gs.get_swing_state(), pt, vt)

        def tick_step(s):
            return leg_step(s, ctrl_step(s))
        for s in range(a.warmup):
            tick_step(s)
        barrier()
        c0.record()
        for s in range(a.steps):
            tick_step(a.warmup + s)
        c1.record()
        barrier()
        tick_ms = c0.elapsed_time(c1) / a.steps
        f_last = ctrl_step(0)
        barrier()
        c0.record()
        for s in range(a.steps):
            leg_step(a.warmup + a.steps + s, f_last)
        c1.record()
        barrier()
        leg_ms = c0.elapsed_time(c1) / a.steps
        # algorithmic HBM bytes per robot of the two leg kernels (all four legs, every array the call reads or writes once; of the
        # 3x18 Jacobian only the leg's own 3x3 block is algorithmic): swing 392 in + 360 state + 192 out, torque 824 in + 48 out
        LEG_BYTES = 944 + 872
        robot_tick = {"value": world * B / (tick_ms * 1e-3), "unit": "robot control ticks/s (each with an MPC update)", "ms_per_step": tick_ms,
                      "api": "BatchedGaitSchedule.set_iteration + update_robot_state + update_mpc_if_needed + "
                             "BatchedSwingFootTrajectoryGenerator.update + BatchedLegController.update on device tensors (9 kernel launches)",
                      "leg_layer_ms": leg_ms, "leg_layer_launches": 3,
                      "leg_layer_hbm_gbs": B * LEG_BYTES / (leg_ms * 1e-3) / 1e9,
                      "note": "leg layer = gait kernel + mpcq_swing_targets + mpcq_leg_torques, launch-latency bound at this batch "
                              "(1.8 KB per robot); see tools/leg_layer_stream.py for the HBM-stream figure at 1M robots"}

        # ---- closed loop in the MPC's own model (warm start, mpcq_set_warm_start): from input set 0 the state advances one
        # horizon step per update under the first-step forces (x+ = x + dt w + dt^2/2 Ac w, w = Ac x + Bc u: Ac is nilpotent),
        # the contact table and the reference trajectory shift by one step; every update is solved cold and warm-started
        # from the previous update's faces.  Synthetic (no simulator), but consistent with the model the MPC optimises over.
        closed_loop = None
        if a.gait == "trot" and a.robot == "A1Config":
            T = 12
            f64 = torch.float64
            cx0, cfeet, cyaw = x0[0].to(f64).clone(), feet[0].to(f64).reshape(B, 4, 3).clone(), yaw[0].to(f64).clone()
            cxr = xref[0].to(f64).reshape(B, H, 13).clone()
            Ib = torch.as_tensor(np.asarray(robot.base_inertia_base, dtype=np.float64), device=dev)
            mass, dtm = float(robot.mass_base), float(ctrl.dt)
            seq = []
            for t in range(T):
                eng.gait_tables(gparams[0][0], gparams[0][1], gparams[0][2], gparams[0][3] + t * ibm, ibm, table=gtab)
                inp = (cx0.to(tdt), cfeet.reshape(B, 12).to(tdt).contiguous(), gtab.clone(), cxr.reshape(B, 13 * H).to(tdt).contiguous(), cyaw.to(tdt))
                seq.append(inp)
                r = eng.solve(inp[0], inp[1], inp[2], inp[3], yaw=inp[4], want=())
                u0 = r.forces.to(f64).reshape(B, 4, 3)
                c, s_ = torch.cos(cyaw), torch.sin(cyaw)
                Rz = torch.zeros((B, 3, 3), dtype=f64, device=dev)
                Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1], Rz[:, 2, 2] = c, -s_, s_, c, 1.0
                Iw = Rz @ Ib @ Rz.transpose(1, 2)
                tau = torch.cross(cfeet, u0, dim=2).sum(dim=1)
                wacc = torch.linalg.solve(Iw, tau)
                vacc = u0.sum(dim=1) / mass
                vacc[:, 2] += cx0[:, 12]
                w_rpy = torch.einsum("bji,bj->bi", Rz, cx0[:, 6:9])
                nx = cx0.clone()
                nx[:, 0:3] += dtm * w_rpy + 0.5 * dtm * dtm * torch.einsum("bji,bj->bi", Rz, wacc)
                nx[:, 3:6] += dtm * cx0[:, 9:12] + 0.5 * dtm * dtm * vacc
                nx[:, 6:9] += dtm * wacc
                nx[:, 9:12] += dtm * vacc
                cfeet -= (nx[:, 3:6] - cx0[:, 3:6])[:, None, :]
                cx0, cyaw = nx, nx[:, 2].clone()
                cxr[:, :-1] = cxr[:, 1:].clone()
                cxr[:, -1, 2:5] += cxr[:, -1, 2:5] - cxr[:, -3, 2:5]
            fbuf = [torch.zeros((B, 4 * H), dtype=torch.uint8, device=dev) for _ in range(2)]
            it_out = torch.empty((B, 2), dtype=torch.int32, device=dev)
            res_cl = SolveResult(forces=out.forces, u=None, iters=it_out, resid=None, status=out.status, active=None)

            def run_loop(warm_mode):
                nf, bad = [], 0
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                fbuf[0].zero_(); fbuf[1].zero_()
                barrier(); e0.record()
                for t, inp in enumerate(seq):
                    kw = {}
                    if warm_mode:
                        prev, cur = fbuf[(t + 1) % 2], fbuf[t % 2]
                        fin = torch.cat([prev[:, 4:], prev[:, 4 * (H - 1):]], dim=1)
                        kw = dict(faces_in=fin, faces_out=cur)
                    eng.solve(inp[0], inp[1], inp[2], inp[3], yaw=inp[4], out=res_cl, **kw)
                    nf.append(it_out[:, 0].float().mean())
                e1.record(); barrier()
                return e0.elapsed_time(e1) / len(seq), float(torch.stack(nf[1:]).mean())

            run_loop(False); run_loop(True)
            cold_ms, cold_nf = run_loop(False)
            warm_ms, warm_nf = run_loop(True)
            closed_loop = {"updates": T, "cold": {"ms_per_update": cold_ms, "rounds_mean": cold_nf, "solves_per_s": world * B / (cold_ms * 1e-3)},
                           "warm_start": {"ms_per_update": warm_ms, "rounds_mean": warm_nf, "solves_per_s": world * B / (warm_ms * 1e-3)},
                           "note": "model-consistent synthetic rollout from input set 0 (state advanced one horizon step per update under the "
                                   "first-step forces, contact table and reference shifted by one step); warm start = previous update's "
                                   "faces shifted by one step (mpcq_set_warm_start); NOT the headline metric, which is cold"}

    # ---- roofline of the dominant kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    kern_s = float(kmean[dom]) * 1e-3
    achieved = ALGO_BYTES(H) * (rs / 4) * B / kern_s / 1e9
    # executed arithmetic of the dominant kernel (model): per robot ONE Cholesky of the stance-slot Hessian
    # (n^3/6 FMAs), its inverse (n^3/3) and the panel assembly (2 n^2); the Schur-complement rounds (q^3/6 + 2 n q each, q =
    # active rows, not reported per round) are NOT counted - a lower bound
    nred = float(3.0 * tabs.reshape(S, B, -1).sum(axis=2).mean())
    fma_fixed = nred ** 3 / 6 + nred ** 3 / 3 + 2 * nred ** 2
    exec_flops = 2.0 * fma_fixed * B / kern_s

    # ---- measured denominators MEASURED_PEAKS.json does not hold (fp32 / fp64 FMA, shared-memory loads): micro-kernels
    # through the C ABI, run after every timed region
    import ctypes
    pk = (ctypes.c_double * 4)()
    peaks_ok = eng.lib.mpcq_measure_peaks(local, pk) == 0
    fma_peak = (pk[1] if a.dtype == "f64" else pk[0]) if peaks_ok else None
    # algorithmic flops of one solve as SURVEY.md 8d defines them for the reference's dense formulation:
    # build 2 n^2 k + 2 n k, one Cholesky n^3/3, 2 n^2 per triangular-solve pair (n = 12H, k = 13H)
    nfull, kfull = 12 * H, 13 * H
    algo_flops = 2.0 * nfull * nfull * kfull + 2.0 * nfull * kfull + nfull ** 3 / 3.0 + 2.0 * nfull * nfull * float(facts.mean())
    algo_tflops = algo_flops * B / kern_s / 1e12

    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture of this same command
    traffic, traffic_src, smem = None, None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        tj = next((e for e in (tj if isinstance(tj, list) else [tj])
                   if e.get("envs") == B and e.get("horizon") == H and e.get("dtype") == a.dtype and e.get("gait", "trot") == a.gait), None)
        if tj is not None:
            traffic, traffic_src = tj["dram_bytes_read"] + tj["dram_bytes_write"], tj.get("source")
            if peaks_ok and tj.get("smem_wavefronts"):
                # shared-memory data pipe: wavefronts per launch (ncu capture of this command; one wavefront = one 128-byte pass
                # of the LSU data pipe, 1 per clock per SM) over the kernel duration measured live, against the LDS.128 probe
                ach = tj["smem_wavefronts"] * 128.0 / kern_s / 1e9
                smem = {"bound": "shared-memory data pipe (LSU wavefronts x 128 B)", "achieved": ach, "peak": pk[2], "unit": "GB/s",
                        "frac": ach / pk[2], "wavefronts_per_launch": tj["smem_wavefronts"],
                        "ncu_pct_of_peak_elapsed": tj.get("smem_pipe_pct_of_peak_elapsed_ncu"),
                        "ncu_l1tex_pct_of_peak_active": tj.get("l1tex_throughput_pct_active_ncu"),
                        "note": "shared-memory data pipe of the solve kernel (l1tex__data_pipe_lsu_wavefronts_mem_shared of the ncu capture of "
                                "this command, over the kernel duration measured live): the factor, its inverse and the Schur block live "
                                "in shared memory"}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_baseline(a, st, tabs, out, eng, (x0, yaw, feet, xref, gait))

    set_mb = S * B * ALGO_BYTES(H) * rs // 4 / 1e6
    l2_note = f"{S} distinct input sets rotated ({set_mb:.0f} MB total, " + (
        "> 126 MB L2)" if set_mb > 126 else "BELOW the 126 MB L2: raise --sets for an L2-cold number)")
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
            "dtype": a.dtype, "data": "synthetic",
            "config": dict(config_dict(a, world),
                           parallelism=f"env-sharded x{world}, no collective in the solve loop" + (" + all-gather of GRFs" if gathered is not None else ""),
                           l2=l2_note,
                           precision="Cholesky, inverse and Schur-complement rounds in " + a.dtype + ", residuals + KKT tests in f64"),
            "latency_ms": {"p50": lat[len(lat) // 2], "p90": lat[int(len(lat) * 0.9)], "max": lat[-1]},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "mpcq_tick_host (C ABI, page-locked host buffers: RobotData fields + command + gait parameters in - 272 B per robot -, "
                           "forces + status out, written in place by the kernels; gait table, state assembly, reference trajectory "
                           "and solve on the device = the reference loop body scripts/isaacgym_a1.py:119-144 for the batch)",
                    "ms_per_step": 1e3 * float(te.item()) / a.steps, "launches_per_step": e2e_launches},
            "e2e_solve_host": sh_e2e,
            "e2e_pipelined": e2e_pipe,
            "controller_api": None if ctrl_ms is None else {"value": world * B / (ctrl_ms * 1e-3), "unit": UNIT, "ms_per_step": ctrl_ms,
                               "api": "BatchedModelPredictiveController.update_robot_state + update_mpc_if_needed on device tensors "
                                      "(mpcq_gait_tables + mpcq_assemble + mpcq_solve: 6 kernel launches, the two size classes side by side)"},
            "robot_tick": robot_tick,
            "graph_replay": graph_replay,
            "gpu_launches": launches_per_step * a.steps,
            "kernel_ms": {"per_class_mean": [float(v) for v in kmean], "dominant_class": dom,
                          "share_of_step": float(kmean[dom] / (total_ms / a.steps)),
                          "note": "solve kernels by size class; the schedule pre-pass (two small launches: score + scatter) is not in this list; the classes run side by side"},
            # the binding roofline of the dominant kernel: fp32 (fp64 in f64 mode) multiply-add throughput of the CUDA cores.
            # achieved = SURVEY.md 8d algorithmic flops per solve (the reference's dense formulation) x robots per launch / the
            # kernel's duration measured live with CUDA events; peak = the FMA micro-kernel of this run (mpcq_measure_peaks;
            # MEASURED_PEAKS.json holds no fp32 / fp64 figure).  frac is recomputable from kernel_ms.  HBM beside it.
            "roofline": {"bound": "fp64 fma (cuda cores)" if a.dtype == "f64" else "fp32 fma (cuda cores)",
                         "achieved": algo_tflops, "peak": fma_peak, "unit": "TFLOP/s",
                         "frac": (algo_tflops / fma_peak) if fma_peak else None, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_mflop_per_solve": algo_flops / 1e6, "kernel_ms": float(kmean[dom]), "robots_per_launch": B,
                         "peak_source": "mpcq_measure_peaks micro-kernel, this run (fp32 / fp64 FMA are not in MEASURED_PEAKS.json)",
                         "executed_tflops_model": exec_flops / 1e12, "frac_executed": (exec_flops / 1e12 / fma_peak) if fma_peak else None,
                         "note": "algorithmic = 2 n^2 k build + n^3/3 Cholesky + 2 n^2 per solve pair x rounds (n = 12H, k = 13H); executed = "
                                 "what the kernel does per robot (closed-form Hessian on the stance slots only, n_red^3/6 Cholesky + n_red^3/3 "
                                 "inverse once, Schur rounds not counted)"
                                 + ("; frac > 1: the kernel does not execute the dense formulation's flops (at long horizons the stance-only "
                                    "system is half the size, an eighth of the cubic work) - frac_executed is the honest occupancy of the pipe"
                                    if fma_peak and algo_tflops > fma_peak else "")},
            "roofline_hbm": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                             "traffic": traffic, "algorithmic_bytes_per_launch": ALGO_BYTES(H) * (rs // 4) * B, "peak_source": peak_src,
                             "note": "not the binding roofline: the path moves 732 B per solve (H = 10, f32) and re-reads nothing"},
            "roofline_compute": {"fp32_fma_tflops": pk[0] if peaks_ok else None, "fp64_fma_tflops": pk[1] if peaks_ok else None,
                                 "smem_load_gbs": pk[2] if peaks_ok else None,
                                 "rounds_per_solve_mean": float(facts.mean()), "rounds_p50": float(np.median(facts)),
                                 "rounds_max": float(facts.max()), "reduced_dim_mean": float(nred),
                                 "note": "rounds = active-set solves per robot (the first on the free faces, then Schur-complement rounds); "
                                         "ONE Hessian factorisation + inverse per robot whatever the rounds"},
            "solver": {"fallback_envs": fallback, "unverified_envs": unverified, "envs_checked": int(S * B)},
            "clocks": clocks,
        }
        if smem is not None:
            line["roofline_smem"] = smem
        if closed_loop is not None:
            line["closed_loop"] = closed_loop
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line, default=float), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(a, st, tabs, out, eng, dev_inputs):
    """Oracle port on the host cores over a bounded sample of the same workload; also cross-checks the GPU forces."""
    import multiprocessing as mp
    import torch
    from oracle.cpu_baseline import run_parallel
    cores = os.cpu_count() or 1
    sample = min(a.cpu_sample or (3072 if a.horizon <= 16 else 768), a.envs * a.sets)   # ~20 s of CPU work
    keys = ("quat_base", "pos_base", "ang_vel_base", "lin_vel_base", "pos_base_feet", "vel_cmd_body", "yaw_rate_cmd")
    sub = {k: st[k][:sample] for k in keys}
    with mp.get_context("spawn").Pool(cores) as pool:
        run_parallel(pool, a.horizon, a.robot, {k: v[:cores * 2] for k, v in sub.items()}, tabs[:cores * 2], cores)   # warm the workers
        f_cpu, wall, tb, ts = run_parallel(pool, a.horizon, a.robot, sub, tabs[:sample], cores)
    x0, yaw, feet, xref, gait = dev_inputs
    eng.solve(x0[0], feet[0], gait[0], xref[0], yaw=yaw[0], out=out)
    nchk = min(sample, a.envs)
    f_gpu = out.forces[:nchk].double().cpu().numpy()
    f_cpu_chk = f_cpu[:nchk]
    err = np.abs(f_gpu - f_cpu_chk).max(axis=1)
    tol = np.maximum(1e-3, 1e-4 * np.abs(f_cpu_chk).max(axis=1))
    return {"value": sample / wall, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {sample} envs of the seeded workload (input set 0 onwards), one pass, {cores} worker processes",
            "cpu_ms_per_solve": {"build": 1e3 * tb / sample, "solve": 1e3 * ts / sample},
            "gpu_vs_cpu_max_abs_df_N": float(err.max()), "gpu_vs_cpu_within_tolerance": bool(np.all(err <= tol))}


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
