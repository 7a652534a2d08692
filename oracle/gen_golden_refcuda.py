"""ORACLE (test infrastructure): golden vectors from the reference's OWN CUDA kernels.

Run on a GPU box that has oracle/_ref/gsplat_ref_csrc.so (oracle/build_ref.py):

    python oracle/gen_golden_refcuda.py gpurun_out/golden      # then copy *.npz to tests/golden/

Produces small fixtures for the parts of the path that the reference's CPU code cannot pin
(SURVEY.md §8c): `camera_model="spherical"` projection forward + closed-form VJP, the packed
projection's rules, and the CUDA rasterizer itself.  The committed vectors let the CPU oracle and
the kernels be checked against the real reference even where oracle/_ref is absent.
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_cuda  # noqa: E402
from splat_one_b200 import synthetic  # noqa: E402

out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden"
os.makedirs(out_dir, exist_ok=True)
R = ref_cuda.load()
assert R is not None, "oracle/_ref not built"
dev = "cuda:0"


def save(name, **kw):
    arrs = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in kw.items()}
    np.savez_compressed(os.path.join(out_dir, name), **arrs)
    print(name, {k: v.shape for k, v in arrs.items()})


def projection(model, name, N=2500, W=256, H=128, C=2, comp=True):
    if model == "spherical":
        sc = synthetic.spherical_scene(N, W, H, seed=31)
        sc["viewmats"] = sc["viewmats"].repeat(C, 1, 1)
        sc["viewmats"][1, :3, 3] = torch.tensor([0.3, -0.2, 0.1])
        sc["Ks"] = sc["Ks"].repeat(C, 1, 1)
    else:
        sc = synthetic.pinhole_scene(N, W, H, seed=31, n_cameras=C)
    g = {k: v.to(dev) for k, v in sc.items() if isinstance(v, torch.Tensor)}
    cm = ref_cuda.camera_model(R, model)
    radii, m2d, dep, con, cmp_ = R.fully_fused_projection_fwd(g["means"], None, g["quats"], g["scales"], g["viewmats"],
                                                              g["Ks"], W, H, 0.3, 0.01, 1e10, 0.0, comp, cm)
    gen = torch.Generator().manual_seed(5)
    v_m2d = torch.randn(C, N, 2, generator=gen).to(dev)
    v_dep = torch.randn(C, N, generator=gen).to(dev)
    v_con = (torch.randn(C, N, 3, generator=gen) * 0.1).to(dev)
    v_cmp = torch.randn(C, N, generator=gen).to(dev)
    v_means, _, v_quats, v_scales, v_vm = R.fully_fused_projection_bwd(
        g["means"], None, g["quats"], g["scales"], g["viewmats"], g["Ks"], W, H, 0.3, cm, radii, con,
        cmp_ if comp else None, v_m2d, v_dep, v_con, v_cmp if comp else None, True)
    vis = radii > 0
    z = lambda t: torch.where(vis.reshape(vis.shape + (1,) * (t.dim() - 2)), t, torch.zeros_like(t))  # noqa: E731
    save(name, means=g["means"], quats=g["quats"], scales=g["scales"], viewmats=g["viewmats"], Ks=g["Ks"],
         width=W, height=H, camera_model=model, radii=radii, means2d=z(m2d), depths=z(dep), conics=z(con),
         compensations=z(cmp_), v_means2d=v_m2d, v_depths=v_dep, v_conics=v_con, v_compensations=v_cmp,
         v_means=v_means, v_quats=v_quats, v_scales=v_scales, v_viewmats=v_vm)
    # packed rules on the same inputs
    outp = R.fully_fused_projection_packed_fwd(g["means"], None, g["quats"], g["scales"], g["viewmats"], g["Ks"], W, H,
                                               0.3, 0.01, 1e10, 0.0, comp, cm)
    indptr, cam, gid, pr, pm, pd, pc, pcmp = outp
    save(name.replace(".npz", "_packed.npz"), indptr=indptr, camera_ids=cam, gaussian_ids=gid, radii=pr, means2d=pm,
         depths=pd, conics=pc, compensations=pcmp)


def raster(name, N=3000, W=96, H=64, D=3):
    sc = synthetic.pinhole_scene(N, W, H, seed=33, footprint_px=4.0)
    g = {k: v.to(dev) for k, v in sc.items() if isinstance(v, torch.Tensor)}
    radii, m2d, dep, con, _ = R.fully_fused_projection_fwd(g["means"], None, g["quats"], g["scales"], g["viewmats"],
                                                           g["Ks"], W, H, 0.3, 0.01, 1e10, 0.0, False,
                                                           R.CameraModelType.PINHOLE)
    vis = radii > 0
    m2d = torch.where(vis[..., None], m2d, torch.zeros_like(m2d))
    con = torch.where(vis[..., None], con, torch.zeros_like(con))
    dep = torch.where(vis, dep, torch.zeros_like(dep))
    gen = torch.Generator().manual_seed(6)
    colors = torch.rand(1, N, D, generator=gen).to(dev)
    opac = g["opacities"][None].contiguous()
    bg = torch.rand(1, D, generator=gen).to(dev)
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    tpg, ids, flat = R.isect_tiles(m2d, radii, dep, None, None, 1, 16, tw, th, True, True)
    offs = R.isect_offset_encode(ids, 1, tw, th)
    rc, ra, last = R.rasterize_to_pixels_fwd(m2d, con, colors, opac, bg, None, W, H, 16, offs, flat)
    v_rc = torch.randn(rc.shape, generator=gen).to(dev)
    v_ra = torch.randn(ra.shape, generator=gen).to(dev)
    v_abs, v_m2d, v_con, v_col, v_op = R.rasterize_to_pixels_bwd(m2d, con, colors, opac, bg, None, W, H, 16, offs, flat,
                                                                 ra, last, v_rc, v_ra, True)
    save(name, means2d=m2d, conics=con, colors=colors, opacities=opac, backgrounds=bg, radii=radii, depths=dep,
         width=W, height=H, tile_size=16, tiles_per_gauss=tpg, isect_ids=ids, flatten_ids=flat, isect_offsets=offs,
         render_colors=rc, render_alphas=ra, last_ids=last, v_render_colors=v_rc, v_render_alphas=v_ra,
         v_means2d_abs=v_abs, v_means2d=v_m2d, v_conics=v_con, v_colors=v_col, v_opacities=v_op)


projection("spherical", "refcuda_projection_spherical.npz")
projection("pinhole", "refcuda_projection_pinhole.npz")
raster("refcuda_raster_d3.npz")
