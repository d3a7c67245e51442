"""HBM-stream rate of the two leg-layer kernels at a batch large enough to leave launch latency behind:
`mpcq_swing_targets` and `mpcq_leg_torques` over B robots (default 2^20), CUDA events, 20 launches after 3 warm-ups, inputs far
larger than L2.  Algorithmic bytes per robot (every array read or written once, steady state = mid-swing; checked against the
ncu capture profiles/r01_ncu_legs_v14.txt: 799 B per robot measured when the kernel still re-wrote footpos_init, 96 B more than this):
  swing  : 168 B per robot (pos, vel, R, commands, times; once if any leg swings) + per swinging leg 146 B (thigh 24, phase 8,
           state in/out 42, final 24, targets 48) + per stance leg 56 B (phase 8, zero targets 48) = 752 B with all four legs swinging
  torque : 4 x 72 B Jacobian blocks + 72 R + 96 + 96 + 48 forces + 32 + 96 + 96 in + 48 out = 872 B (ncol = 3); in the reference's
           18-column layout the DRAM reads nearly the whole 1 728 B Jacobian (ncu: 2.19 GB per 2^20 robots = 6.9 TB/s, the HBM peak)
Usage: python tools/leg_layer_stream.py [B]"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from pympc_quadruped_b200 import A1Config, _capi, with_horizon
from pympc_quadruped_b200.engine import MpcqEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev, f64 = "cuda:0", torch.float64
eng = MpcqEngine(with_horizon(10), A1Config, dtype=torch.float32, device=dev)
lp = _capi.make_leg_params(A1Config.Kp_swing, A1Config.Kd_swing, 0.1, 0.001, 9.81)
g = torch.Generator(device=dev); g.manual_seed(1)
ru = lambda *s: torch.rand(*s, generator=g, device=dev, dtype=f64) - 0.5
pos, vel, R, th, feet, vdes, yr = ru(B, 3), ru(B, 3), ru(B, 9), ru(B, 4, 3), ru(B, 4, 3), ru(B, 3), ru(B)
tsw, tst = torch.full((B,), 0.1, device=dev, dtype=f64), torch.full((B,), 0.1, device=dev, dtype=f64)
state = (torch.zeros((B, 4), dtype=torch.uint8, device=dev), torch.zeros((B, 4), dtype=f64, device=dev),
         torch.zeros((B, 4, 3), dtype=f64, device=dev), torch.zeros((B, 4, 3), dtype=f64, device=dev))
bpf, bvf, frc = ru(B, 4, 3), ru(B, 4, 3), ru(B, 12).float()
peak = 6650.0
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
out = {"robots": B, "hbm_peak_gbs": peak}
for label, frac_swing in (("all legs swinging", 1.0), ("trot (half the legs swinging)", 0.5)):
    ss = (torch.rand(B, 4, generator=g, device=dev, dtype=f64) < frac_swing).to(f64) * 0.5
    pt, vt = torch.empty((B, 4, 3), dtype=f64, device=dev), torch.empty((B, 4, 3), dtype=f64, device=dev)
    for ncol in (18, 3):
        Jv = ru(B, 4, 3, ncol)
        tau = torch.empty((B, 12), dtype=torch.float32, device=dev)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ms_s = ms_t = 0.0
        for it in range(23):
            e[0].record()
            eng.swing_targets(lp, pos, vel, R, th, feet, ss, vdes, yr, tsw, tst, state, pos_targets=pt, vel_targets=vt)
            e[1].record()
            eng.leg_torques(lp, Jv, R, bpf, bvf, frc, ss, pt, vt, torque_cmds=tau)
            e[2].record()
            torch.cuda.synchronize()
            if it >= 3:
                ms_s += e[0].elapsed_time(e[1]) / 20; ms_t += e[1].elapsed_time(e[2]) / 20
        sw_bytes = B * (168 * (1 - (1 - frac_swing) ** 4) + 4 * (frac_swing * 146 + (1 - frac_swing) * 56))
        tq_alg = B * (4 * 8 + 48 + 4 * (72 + frac_swing * (72 / 4 + 24 * 4) + (1 - frac_swing) * 12))
        out[f"{label}, ncol={ncol}"] = {
            "swing_ms": ms_s, "swing_gbs": sw_bytes / ms_s / 1e6, "swing_frac_hbm": sw_bytes / ms_s / 1e6 / peak,
            "torque_ms": ms_t, "torque_gbs_algorithmic": tq_alg / ms_t / 1e6, "torque_frac_hbm_algorithmic": tq_alg / ms_t / 1e6 / peak,
            "robots_per_s_both": B / ((ms_s + ms_t) * 1e-3)}
        del Jv
print(json.dumps(out, indent=1))
