#!/usr/bin/env python
"""Developer tool: GPU timeline of one step (torch.profiler / CUPTI) — every kernel with its
duration and the idle gap in front of it, to find host-side bubbles.

    python tools/step_trace.py [fused|raster]      (config B)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import splat_one_b200 as S  # noqa: E402
from splat_one_b200 import synthetic  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "fused"
dev = torch.device("cuda:0")
N, W, H = 1_000_000, 1920, 1080
scene = synthetic.pinhole_scene(N, W, H, seed=42)
raw = {
    "means": scene["means"], "quats": scene["quats"], "scales": torch.log(scene["scales"]),
    "opacities": torch.logit(scene["opacities"].clamp(1e-4, 1 - 1e-4)),
    "sh0": scene["sh"][:, :1].contiguous(), "shN": scene["sh"][:, 1:].contiguous(),
}
P = {k: v.to(dev).requires_grad_() for k, v in raw.items()}
A = {k: scene[k].to(dev).requires_grad_() for k in ("means", "quats", "scales", "opacities", "sh")}
viewmats, Ks = scene["viewmats"].to(dev), scene["Ks"].to(dev)
c2w = torch.inverse(viewmats)
g = torch.Generator().manual_seed(1)
pixels = torch.rand(1, H, W, 3, generator=g).to(dev)
vc, va = torch.randn(1, H, W, 3, generator=g).to(dev), torch.randn(1, H, W, 1, generator=g).to(dev)


def step():
    if mode == "fused":
        for p in P.values():
            p.grad = None
        rc, ra, _ = S.rasterize_splats(P, c2w, Ks, W, H, sh_degree=3, packed=False)
        S.l1_ssim_loss(rc, pixels, 0.2).backward()
    else:
        for p in A.values():
            p.grad = None
        rc, ra, _ = S.rasterization(A["means"], A["quats"], A["scales"], A["opacities"], A["sh"], viewmats, Ks, W, H,
                                    sh_degree=3, packed=False)
        torch.autograd.backward([rc, ra], [vc, va])


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
n = len(evs) // 3
last = evs[2 * n:]   # third step
t0 = last[0].time_range.start
prev_end = t0
busy = gap = 0.0
print(f"{'start_us':>9} {'gap_us':>7} {'dur_us':>8}  kernel")
for e in last:
    s_, e_ = e.time_range.start, e.time_range.end
    gp = max(0.0, s_ - prev_end)
    print(f"{s_ - t0:9.1f} {gp:7.1f} {e_ - s_:8.1f}  {e.name[:100]}")
    busy += e_ - s_
    gap += gp
    prev_end = max(prev_end, e_)
print(f"kernels {len(last)}  busy {busy:.1f} us  idle {gap:.1f} us  span {prev_end - t0:.1f} us")
