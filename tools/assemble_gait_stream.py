"""HBM-stream rate of the two input kernels at a batch large enough to leave launch latency behind:
`mpcq_gait_tables` (40 B in + 16 H B out per robot; with the phase states + 64 B) and `mpcq_assemble` (17 + 5 float64 in,
5 float64 state in/out, (14 + 13 H) reals out per robot) over B robots (default 2^20), CUDA events, 20 launches after 3 warm-ups.
Algorithmic bytes per robot, H = 10, f32: gait 40 + 160 = 200 B;  assemble 8 * (4+3+3+3+3+1) + 2 * 40 + 4 * (13 + 1 + 130) = 792 B.
Usage: python tools/assemble_gait_stream.py [B]"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from pympc_quadruped_b200 import A1Config, with_horizon
from pympc_quadruped_b200.engine import MpcqEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
H = 10
dev, f64 = "cuda:0", torch.float64
peak = 6650.0
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
out = {"robots": B, "horizon": H, "hbm_peak_gbs": peak}
g = torch.Generator(device=dev); g.manual_seed(1)
ru = lambda *s: torch.rand(*s, generator=g, device=dev, dtype=f64) - 0.5
for dtype, rs in ((torch.float32, 4), (torch.float64, 8)):
    eng = MpcqEngine(with_horizon(H), A1Config, dtype=dtype, device=dev)
    quat = torch.nn.functional.normalize(ru(B, 4) + torch.tensor([2.0, 0, 0, 0], device=dev, dtype=f64), dim=1).contiguous()
    pos, omega, vel, vdes, yr = ru(B, 3), ru(B, 3), ru(B, 3), ru(B, 3), ru(B)
    xy, yd, rp = torch.zeros((B, 2), dtype=f64, device=dev), torch.zeros(B, dtype=f64, device=dev), torch.zeros((B, 2), dtype=f64, device=dev)
    x0, yaw, xref = torch.empty((B, 13), dtype=dtype, device=dev), torch.empty(B, dtype=dtype, device=dev), torch.empty((B, 13 * H), dtype=dtype, device=dev)
    offs = torch.tensor([0, 5, 5, 0], dtype=torch.int32, device=dev).repeat(B, 1).contiguous()
    durs = torch.full((B, 4), 5, dtype=torch.int32, device=dev)
    seg = torch.full((B,), 10, dtype=torch.int32, device=dev)
    cur = torch.randint(0, 5000, (B,), generator=g, device=dev, dtype=torch.int32)
    tab = torch.empty((B, 4 * H), dtype=torch.float32, device=dev)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ms_g = ms_a = 0.0
    for it in range(23):
        e[0].record()
        eng.gait_tables(offs, durs, seg, cur, 20, table=tab)
        e[1].record()
        eng.assemble(quat, pos, omega, vel, vdes, yr, xy, yd, rp, False, True, x0, yaw, xref)
        e[2].record()
        torch.cuda.synchronize()
        if it >= 3:
            ms_g += e[0].elapsed_time(e[1]) / 20; ms_a += e[1].elapsed_time(e[2]) / 20
    gait_bytes = B * (40 + 16 * H)
    asm_bytes = B * (8 * 17 + 2 * 40 + rs * (14 + 13 * H))
    out["f32" if rs == 4 else "f64"] = {
        "gait_ms": ms_g, "gait_gbs": gait_bytes / ms_g / 1e6, "gait_frac_hbm": gait_bytes / ms_g / 1e6 / peak, "gait_bytes_per_robot": gait_bytes // B,
        "assemble_ms": ms_a, "assemble_gbs": asm_bytes / ms_a / 1e6, "assemble_frac_hbm": asm_bytes / ms_a / 1e6 / peak, "assemble_bytes_per_robot": asm_bytes // B}
print(json.dumps(out, indent=1))
