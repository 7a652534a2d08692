#!/usr/bin/env python
"""Developer benchmark: packed-mode training step with the SH table read in place (split) against the same
step with torch.cat([sh0, shN], 1) in front of rasterization() (R/utils/gsplat_utils/gsplat_trainer.py:474)."""
import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splat_one_b200 as S
from splat_one_b200 import synthetic
dev = "cuda:0"; N, W, H = 2_000_000, 1920, 1080
sc = synthetic.pinhole_scene(N, W, H, seed=42)
raw = {"means": sc["means"], "quats": sc["quats"], "scales": torch.log(sc["scales"]),
       "opacities": torch.logit(sc["opacities"].clamp(1e-4, 1 - 1e-4)),
       "sh0": sc["sh"][:, :1].contiguous(), "shN": sc["sh"][:, 1:].contiguous()}
P = {k: v.to(dev).requires_grad_() for k, v in raw.items()}
vm, Ks = sc["viewmats"].to(dev), sc["Ks"].to(dev); c2w = torch.inverse(vm)
pix = torch.rand(1, H, W, 3).to(dev)
def fused():
    rc, _, _ = S.rasterize_splats(P, c2w, Ks, W, H, sh_degree=3, packed=True, sparse_grad=False)
    S.l1_ssim_loss(rc, pix, 0.2).backward()
def cat():
    sc_, op = S.splat_activations(P["scales"], P["opacities"])
    rc, _, _ = S.rasterization(P["means"], P["quats"], sc_, op, torch.cat([P["sh0"], P["shN"]], 1), vm, Ks, W, H, sh_degree=3, packed=True)
    S.l1_ssim_loss(rc, pix, 0.2).backward()
for name, fn in (("split", fused), ("cat", cat), ("split", fused), ("cat", cat)):
    for _ in range(3):
        for p in P.values(): p.grad = None
        fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(15):
        for p in P.values(): p.grad = None
        fn()
    b.record(); torch.cuda.synchronize(); print(name, round(a.elapsed_time(b) / 15, 3), "ms/step (2M Gaussians, packed, 1080p)")
