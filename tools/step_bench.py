#!/usr/bin/env python
"""Developer benchmark of the f4 row (not the contract bench): one TRAINING step's forward +
backward on config B (1 M Gaussians, SH3, 1080p) from RAW parameters to parameter gradients,

  composed : exp / sigmoid / cat (torch, as R/utils/gsplat_utils/gsplat_trainer.py:456-474 writes
             them) -> rasterization() -> l1_loss + SSIM (torch conv2d composite; fused_ssim is
             not installed) -> backward
  fused    : splat_one_b200.rasterize_splats (one activation kernel, split SH table) ->
             l1_ssim_loss (one kernel per direction) -> backward

and the colour stage alone, staged vs thread-per-row kernels.   python tools/step_bench.py [N W H]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import splat_one_b200 as S  # noqa: E402
from splat_one_b200 import synthetic, wrapper  # noqa: E402

dev = torch.device("cuda:0")
N, W, H = (int(a) for a in (sys.argv[1:4] + ["1000000", "1920", "1080"][len(sys.argv) - 1:]))
scene = synthetic.pinhole_scene(N, W, H, seed=42)
raw = {
    "means": scene["means"], "quats": scene["quats"], "scales": torch.log(scene["scales"]),
    "opacities": torch.logit(scene["opacities"].clamp(1e-4, 1 - 1e-4)),
    "sh0": scene["sh"][:, :1].contiguous(), "shN": scene["sh"][:, 1:].contiguous(),
}
P = {k: v.to(dev).requires_grad_() for k, v in raw.items()}
viewmats, Ks = scene["viewmats"].to(dev), scene["Ks"].to(dev)
c2w = torch.inverse(viewmats)
pixels = torch.rand(1, H, W, 3, generator=torch.Generator().manual_seed(1)).to(dev)


def _win():
    x = torch.arange(11, dtype=torch.float64) - 5
    g = torch.exp(-x * x / (2 * 1.5 ** 2))
    g = (g / g.sum()).float().to(dev)
    return (g[:, None] * g[None, :])[None, None].expand(3, 1, -1, -1).contiguous()


WIN = _win()


def torch_ssim_valid(a, b):
    blur = lambda x: F.conv2d(x, WIN, groups=3)  # noqa: E731  (valid correlation)
    mu1, mu2 = blur(a), blur(b)
    s1, s2, s12 = blur(a * a) - mu1 * mu1, blur(b * b) - mu2 * mu2, blur(a * b) - mu1 * mu2
    return (((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 * mu1 + mu2 * mu2 + 1e-4) * (s1 + s2 + 9e-4))).mean()


def step_composed():
    scales, opac = torch.exp(P["scales"]), torch.sigmoid(P["opacities"])
    colors = torch.cat([P["sh0"], P["shN"]], 1)
    rc, ra, _ = S.rasterization(P["means"], P["quats"], scales, opac, colors, viewmats, Ks, W, H, sh_degree=3,
                                packed=False)
    loss = F.l1_loss(rc, pixels) * 0.8 + (1 - torch_ssim_valid(rc.permute(0, 3, 1, 2), pixels.permute(0, 3, 1, 2))) * 0.2
    loss.backward()
    return loss


def step_fused():
    rc, ra, _ = S.rasterize_splats(P, c2w, Ks, W, H, sh_degree=3, packed=False)
    loss = S.l1_ssim_loss(rc, pixels, 0.2)
    loss.backward()
    return loss


def timeit(fn, reps=20, warm=4):
    for _ in range(warm):
        for p in P.values():
            p.grad = None
        out = fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        for p in P.values():
            p.grad = None
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


res = {"N": N, "W": W, "H": H}
t, l = timeit(step_composed)
res["composed_ms"], res["composed_loss"] = round(t, 4), l.item()
g_ref = {k: p.grad.clone() for k, p in P.items()}
t, l = timeit(step_fused)
res["fused_ms"], res["fused_loss"] = round(t, 4), l.item()
res["grad_rel_diff"] = {k: ((P[k].grad - g_ref[k]).abs().max() / (g_ref[k].abs().max() + 1e-20)).item() for k in P}

# stage timings of the fused step
wrapper.profiler.enabled = True
wrapper.profiler.reset()
for _ in range(10):
    for p in P.values():
        p.grad = None
    step_fused()
torch.cuda.synchronize()
res["fused_stages_ms"] = {k: round(v["avg_ms"], 4) for k, v in wrapper.profiler.summary_ms().items()}
wrapper.profiler.reset()
for _ in range(10):
    for p in P.values():
        p.grad = None
    step_composed()
torch.cuda.synchronize()
res["composed_stages_ms"] = {k: round(v["avg_ms"], 4) for k, v in wrapper.profiler.summary_ms().items()}
wrapper.profiler.enabled = False

# colour stage alone: thread-per-row vs staged, whole table vs split
with torch.no_grad():
    radii = S.fully_fused_projection(P["means"], None, P["quats"], torch.exp(P["scales"]), viewmats, Ks, W, H)[0]
table = torch.cat([P["sh0"], P["shN"]], 1).detach().requires_grad_()
v = torch.randn(1, N, 3, device=dev)


def colour(fn):
    def run():
        out = fn()
        out.backward(v)
        return out
    return run


wrapper.profiler.enabled = True
for name, fn in (("rows", lambda: wrapper.sh_view_colors(3, P["means"], viewmats, table, radii)),
                 ("staged_whole", lambda: wrapper.sh_view_colors_split(3, P["means"], viewmats, None, table, radii)),
                 ("staged_split", lambda: wrapper.sh_view_colors_split(3, P["means"], viewmats, P["sh0"], P["shN"], radii))):
    wrapper.profiler.reset()
    for i in range(13):
        if i == 3:
            torch.cuda.synchronize()
            wrapper.profiler.reset()
        colour(fn)()
    torch.cuda.synchronize()
    res["colour_" + name] = {k: round(v_["avg_ms"], 4) for k, v_ in wrapper.profiler.summary_ms().items() if "sh_" in k}
print(json.dumps(res))
