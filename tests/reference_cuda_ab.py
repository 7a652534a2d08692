#!/usr/bin/env python
"""A/B checker (test infrastructure; extra evidence, not the contract's reference arm): the reference's OWN CUDA
kernels (oracle/_ref/gsplat_ref_csrc.so, built for sm_100a by oracle/build_ref.py) chained the way
G/rendering.py + the autograd Functions of G/cuda/_wrapper.py chain them, on config B, on the same
B200 and the same synthetic scene as bench.py — next to this repo's rasterization()+backward.

The reference chain is driven with raw pybind calls (no reference Python, which does not exist on the
GPU box): forward = projection, dirs / masks / SH / +0.5 / clamp, isect_tiles, offset encode, raster;
backward = raster bwd, clamp / SH bwd / v_dirs, projection bwd — i.e. the reference WITHOUT its Python
and autograd overhead, which favours the reference.      python tests/reference_cuda_ab.py [N W H [model]]
"""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import splat_one_b200 as S  # noqa: E402
from oracle import ref_cuda  # noqa: E402
from splat_one_b200 import synthetic  # noqa: E402

R = ref_cuda.load()
if R is None:
    print(json.dumps({"unavailable": "oracle/_ref/gsplat_ref_csrc.so not built (python oracle/build_ref.py)"}))
    sys.exit(0)
dev = torch.device("cuda:0")
N, W, H = (int(a) for a in (sys.argv[1:4] + ["1000000", "1920", "1080"][len(sys.argv[1:4]):]))
MODEL = sys.argv[4] if len(sys.argv) > 4 else "pinhole"
sc = synthetic.spherical_scene(N, W, H, seed=42) if MODEL == "spherical" else synthetic.pinhole_scene(N, W, H, seed=42)
P = {k: sc[k].to(dev) for k in ("means", "quats", "scales", "opacities", "sh", "viewmats", "Ks")}
g = torch.Generator().manual_seed(1)
vc = torch.randn(1, H, W, 3, generator=g).to(dev)
va = torch.randn(1, H, W, 1, generator=g).to(dev)
tw, th = math.ceil(W / 16), math.ceil(H / 16)
PIN = ref_cuda.camera_model(R, MODEL)
stages = {}


class T:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.a, self.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.a.record()

    def __exit__(self, *e):
        self.b.record()
        stages.setdefault(self.name, []).append((self.a, self.b))


def ref_step():
    with T("projection_fwd"):
        radii, m2d, dep, con, _ = R.fully_fused_projection_fwd(P["means"], None, P["quats"], P["scales"], P["viewmats"],
                                                               P["Ks"], W, H, 0.3, 0.01, 1e10, 0.0, False, PIN)
    with T("sh_fwd(+glue)"):
        campos = torch.inverse(P["viewmats"])[:, :3, 3]
        dirs = P["means"][None] - campos[:, None]
        masks = radii > 0
        shs = P["sh"][None]
        sh_col = R.compute_sh_fwd(3, dirs, shs, masks)
        col = torch.clamp_min(sh_col + 0.5, 0.0)
    opac = P["opacities"][None]
    with T("isect_tiles"):
        _, ids, flat = R.isect_tiles(m2d, radii, dep, None, None, 1, 16, tw, th, True, True)
    with T("isect_offset_encode"):
        offs = R.isect_offset_encode(ids, 1, tw, th)
    with T("rasterize_fwd"):
        rc, ra, last = R.rasterize_to_pixels_fwd(m2d, con, col, opac, None, None, W, H, 16, offs, flat)
    with T("rasterize_bwd"):
        _, v_m2d, v_con, v_col, v_op = R.rasterize_to_pixels_bwd(m2d, con, col, opac, None, None, W, H, 16, offs, flat,
                                                                  ra, last, vc, va, False)
    with T("sh_bwd(+glue)"):
        v_sh_col = torch.where(col > 0, v_col, torch.zeros_like(v_col))
        v_coeffs, v_dirs = R.compute_sh_bwd(16, 3, dirs, shs, masks, v_sh_col, True)
        v_means_dirs = v_dirs.sum(0)
    with T("projection_bwd"):
        v_means, _, v_quats, v_scales, _ = R.fully_fused_projection_bwd(
            P["means"], None, P["quats"], P["scales"], P["viewmats"], P["Ks"], W, H, 0.3, PIN, radii, con, None, v_m2d,
            torch.zeros_like(dep), v_con, None, False)
        v_means = v_means + v_means_dirs
    return rc, (v_means, v_quats, v_scales, v_op.sum(0), v_coeffs[0]), flat.numel()


A = {k: P[k].clone().requires_grad_() for k in ("means", "quats", "scales", "opacities", "sh")}


def our_step():
    for p in A.values():
        p.grad = None
    rc, ra, _ = S.rasterization(A["means"], A["quats"], A["scales"], A["opacities"], A["sh"], P["viewmats"], P["Ks"], W,
                                H, sh_degree=3, packed=False, camera_model=MODEL)
    torch.autograd.backward([rc, ra], [vc, va])
    return rc


def timeit(fn, reps=20, warm=5):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    stages.clear()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


t_ref, (rc_ref, g_ref, n_isects) = timeit(ref_step)
ref_stages = {k: round(sum(x.elapsed_time(y) for x, y in v) / len(v), 4) for k, v in stages.items()}
t_our, rc_our = timeit(our_step)
err = (rc_our - rc_ref).abs()
names = ("means", "quats", "scales", "opacities", "sh")
gdiff = {n: ((A[n].grad - gr).abs().max() / (gr.abs().max() + 1e-20)).item() for n, gr in zip(names, g_ref)}
print(json.dumps({
    "workload": f"{N} Gaussians, SH3, {W}x{H} {MODEL}, fwd+bwd, n_isects={n_isects}",
    "reference_cuda_ms": round(t_ref, 4), "reference_cuda_Mpix_s": round(W * H / t_ref / 1e3, 1),
    "reference_cuda_stages_ms": ref_stages,
    "b200splat_ms": round(t_our, 4), "b200splat_Mpix_s": round(W * H / t_our / 1e3, 1),
    "speedup": round(t_ref / t_our, 3),
    "image_frac_outside_1e-4": (err > 1e-4 + 1e-4 * rc_ref.abs()).float().mean().item(), "image_max_err": err.max().item(),
    "grad_max_rel_diff": gdiff,
    "note": "reference = the fork's CUDA kernels built for sm_100a, raw pybind calls without its Python/autograd overhead",
}))
