#!/usr/bin/env python
"""Developer diagnostic: where do the spherical (config D) projections of this library and of the
reference's CUDA differ, and does the rasterizer agree when fed identical projection outputs?"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import splat_one_b200 as S  # noqa: E402
from oracle import ref_cuda  # noqa: E402
from splat_one_b200 import synthetic  # noqa: E402

R = ref_cuda.load()
dev = "cuda:0"
N, W, H = 2_000_000, 2048, 1024
sc = synthetic.spherical_scene(N, W, H, seed=44)
P = {k: sc[k].to(dev) for k in ("means", "quats", "scales", "opacities", "sh", "viewmats", "Ks")}
cm = ref_cuda.camera_model(R, "spherical")
r_radii, r_m2, r_dep, r_con, _ = R.fully_fused_projection_fwd(P["means"], None, P["quats"], P["scales"], P["viewmats"],
                                                              P["Ks"], W, H, 0.3, 0.01, 1e10, 0.0, False, cm)
radii, m2, dep, con, _ = S.fully_fused_projection(P["means"], None, P["quats"], P["scales"], P["viewmats"], P["Ks"], W, H,
                                                  camera_model="spherical")
vis = (radii > 0) & (r_radii > 0)
print("radii mismatches", int((radii != r_radii).sum()))
dm = (m2 - r_m2).abs().amax(-1)[vis]
dc = ((con - r_con).abs() / (r_con.abs().amax(-1, keepdim=True) + 1e-30)).amax(-1)[vis]
dd = (dep - r_dep).abs()[vis]
print("means2d: max", dm.max().item(), "n>1e-4", int((dm > 1e-4).sum()), "n>1e-3", int((dm > 1e-3).sum()), "bit-equal frac",
      (dm == 0).float().mean().item())
print("conics rel: max", dc.max().item(), "n>1e-5", int((dc > 1e-5).sum()), "bit-equal frac", (dc == 0).float().mean().item())
print("depths: max", dd.max().item(), "bit-equal frac", (dd == 0).float().mean().item())
mn = P["means"][None].expand(1, -1, -1)[vis]
t = (mn[:, 1] / mn.norm(dim=-1)).abs()
worst = dm.topk(10).indices
print("worst means2d diffs:", dm[worst].tolist())
print("their |y/r|:", t[worst].tolist())
for lo, hi in ((0, 0.9), (0.9, 0.99), (0.99, 0.999), (0.999, 0.9999), (0.9999, 1.1)):
    sel = (t >= lo) & (t < hi)
    if sel.any():
        print(f"|y/r| in [{lo},{hi}): n={int(sel.sum())} max dmean={dm[sel].max().item():.3e} max dconic_rel={dc[sel].max().item():.3e}")
# x/y components separately
dxy = (m2 - r_m2).abs()[vis]
print("dx max", dxy[:, 0].max().item(), "dy max", dxy[:, 1].max().item())
