"""Golden vectors produced on a B200 by the reference's OWN CUDA kernels
(oracle/gen_golden_refcuda.py over oracle/_ref/gsplat_ref_csrc.so; tests/golden/refcuda_*.npz).
They pin what the reference's CPU code cannot: the spherical camera model incl. its
closed-form VJP, the packed projection's rules and the CUDA rasterizer.

  * CPU (`-m "not gpu"`): the oracle restatement against the vectors;
  * GPU: the kernels against the same vectors (independent of oracle/_ref being present).
Tolerances: radii within +-1 (ceil of a rounded value) and equal on >= 99.5 %; 1e-4 abs/rel on
means2d (5e-3 px next to the poles of the spherical model), depths, conics; 1e-3 rel on gradients;
intersection ids bit-exact."""
import math

import pytest
import torch

from oracle import raster_ref as RC
from oracle import torch_ref as O
from parity import assert_grad_close

DEV = "cuda:0"


def _check_projection(d, radii, m2d, dep, con, comp, model):
    r_radii = d["radii"]
    vis = (r_radii > 0) & (radii.cpu() > 0)
    assert ((r_radii > 0) != (radii.cpu() > 0)).float().mean().item() < 2e-3
    diff = (r_radii - radii.cpu()).abs()[vis]
    assert (diff > 1).sum().item() == 0 and (diff > 0).float().mean().item() < 5e-3
    atol = {"means2d": 5e-3 if model == "spherical" else 1e-4, "depths": 1e-4, "conics": 1e-4, "compensations": 1e-4}
    for name, a in (("means2d", m2d), ("depths", dep), ("conics", con), ("compensations", comp)):
        a, b = a.detach().cpu()[vis], d[name][vis]
        assert torch.allclose(a, b, rtol=2e-4, atol=atol[name]), (model, name, (a - b).abs().max().item())
    return vis


def _projection_loss(d, vis, m2d, dep, con, comp):
    dev = m2d.device
    v = lambda k: d[k].to(dev)  # noqa: E731
    w = vis.to(dev)
    return ((m2d * v("v_means2d") * w[..., None]).sum() + (dep * v("v_depths") * w).sum()
            + (con * v("v_conics") * w[..., None]).sum() + (comp * v("v_compensations") * w).sum())


@pytest.mark.parametrize("model", ["spherical", "pinhole"])
def test_oracle_projection_matches_reference_cuda_vectors(golden, model):
    d = golden(f"refcuda_projection_{model}.npz")
    P = [d[k].clone().requires_grad_() for k in ("means", "quats", "scales")]
    radii, m2d, dep, con, comp = O.fully_fused_projection(P[0], None, P[1], P[2], d["viewmats"], d["Ks"], d["width"],
                                                          d["height"], calc_compensations=True, camera_model=model)
    vis = _check_projection(d, radii, m2d, dep, con, comp, model)
    same = ((d["radii"] > 0) == (radii > 0)).all(dim=0)
    g = torch.autograd.grad(_projection_loss(d, vis, m2d, dep, con, comp), P)
    for name, a in zip(("v_means", "v_quats", "v_scales"), g):
        assert_grad_close(a[same], d[name][same], what=f"oracle {model} {name} vs reference CUDA", frac_ok=0.998)


@pytest.mark.parametrize("model", ["spherical", "pinhole"])
def test_oracle_packed_rules_match_reference_cuda_vectors(golden, model):
    d = golden(f"refcuda_projection_{model}.npz")
    p = golden(f"refcuda_projection_{model}_packed.npz")
    radii, m2d, dep, con, comp = O.fully_fused_projection(d["means"], None, d["quats"], d["scales"], d["viewmats"],
                                                          d["Ks"], d["width"], d["height"], calc_compensations=True,
                                                          camera_model=model, packed_rules=True)
    cam, gid, pr, pm, pd, pc, pcmp = O.pack_projection(radii, m2d, dep, con, comp)
    ours = set(zip(cam.tolist(), gid.tolist()))
    theirs = set(zip(p["camera_ids"].tolist(), p["gaussian_ids"].tolist()))
    assert len(ours ^ theirs) <= 2e-3 * len(theirs) + 2
    # compare common rows (COO order is (camera, gaussian) row-major in both)
    key_o = cam * d["means"].shape[0] + gid
    key_r = p["camera_ids"] * d["means"].shape[0] + p["gaussian_ids"]
    common = sorted(set(key_o.tolist()) & set(key_r.tolist()))
    io = torch.searchsorted(key_o, torch.tensor(common))
    ir = torch.searchsorted(key_r, torch.tensor(common))
    assert ((pr[io] - p["radii"][ir]).abs() > 1).sum().item() == 0
    assert torch.allclose(pd[io], p["depths"][ir], rtol=2e-4, atol=1e-4)   # packed depth = z (negative behind a 360 camera)
    assert torch.allclose(pc[io], p["conics"][ir], rtol=2e-4, atol=1e-4)
    assert torch.allclose(pm[io], p["means2d"][ir], rtol=2e-4, atol=5e-3 if model == "spherical" else 1e-4)


def test_oracle_isect_and_raster_match_reference_cuda_vectors(golden):
    d = golden("refcuda_raster_d3.npz")
    W, H, ts = d["width"], d["height"], d["tile_size"]
    tw, th = math.ceil(W / ts), math.ceil(H / ts)
    tpg, ids, flat = O.isect_tiles(d["means2d"], d["radii"], d["depths"], ts, tw, th)
    assert torch.equal(tpg, d["tiles_per_gauss"]) and torch.equal(ids, d["isect_ids"])
    assert torch.equal(flat, d["flatten_ids"])
    assert torch.equal(O.isect_offset_encode(ids, 1, tw, th), d["isect_offsets"])
    P = [d[k].clone().requires_grad_() for k in ("means2d", "conics", "colors", "opacities")]
    rc, ra = RC.rasterize_to_pixels(*P, W, H, ts, d["isect_offsets"], d["flatten_ids"], backgrounds=d["backgrounds"])
    for a, b in ((rc, d["render_colors"]), (ra, d["render_alphas"])):
        err = (a.detach() - b).abs()
        assert (err > 1e-4 + 1e-4 * b.abs()).float().mean().item() < 2e-3 and err.max().item() < 2e-2, err.max()
    g = torch.autograd.grad((rc * d["v_render_colors"]).sum() + (ra * d["v_render_alphas"]).sum(), P)
    for name, a in zip(("v_means2d", "v_conics", "v_colors", "v_opacities"), g):
        assert_grad_close(a, d[name], what=f"oracle raster {name} vs reference CUDA", frac_ok=0.995)


# --------------------------------------------------------------------------------------------
# the kernels against the same vectors
# --------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("model", ["spherical", "pinhole"])
def test_kernels_projection_match_reference_cuda_vectors(golden, model):
    import splat_one_b200 as S

    d = golden(f"refcuda_projection_{model}.npz")
    P = [d[k].to(DEV).requires_grad_() for k in ("means", "quats", "scales")]
    radii, m2d, dep, con, comp = S.fully_fused_projection(P[0], None, P[1], P[2], d["viewmats"].to(DEV), d["Ks"].to(DEV),
                                                          d["width"], d["height"], calc_compensations=True,
                                                          camera_model=model)
    vis = _check_projection(d, radii, m2d, dep, con, comp, model)
    same = ((d["radii"] > 0) == (radii.cpu() > 0)).all(dim=0)
    g = torch.autograd.grad(_projection_loss(d, vis, m2d, dep, con, comp), P)
    for name, a in zip(("v_means", "v_quats", "v_scales"), g):
        assert_grad_close(a.cpu()[same], d[name][same], what=f"kernel {model} {name} vs reference CUDA", frac_ok=0.998)
    p = golden(f"refcuda_projection_{model}_packed.npz")
    cam, gid, pr, pm, pd, pc, pcmp = S.fully_fused_projection(
        d["means"].to(DEV), None, d["quats"].to(DEV), d["scales"].to(DEV), d["viewmats"].to(DEV), d["Ks"].to(DEV),
        d["width"], d["height"], calc_compensations=True, camera_model=model, packed=True)
    ours = set(zip(cam.tolist(), gid.tolist()))
    theirs = set(zip(p["camera_ids"].tolist(), p["gaussian_ids"].tolist()))
    assert len(ours ^ theirs) <= 2e-3 * len(theirs) + 2
    if len(ours ^ theirs) == 0:
        assert torch.allclose(pd.cpu(), p["depths"], rtol=2e-4, atol=1e-4)
        assert ((pr.cpu() - p["radii"]).abs() > 1).sum().item() == 0


@pytest.mark.gpu
def test_kernels_isect_and_raster_match_reference_cuda_vectors(golden):
    import splat_one_b200 as S

    d = golden("refcuda_raster_d3.npz")
    W, H, ts = d["width"], d["height"], d["tile_size"]
    tw, th = math.ceil(W / ts), math.ceil(H / ts)
    g_ = lambda k: d[k].to(DEV)  # noqa: E731
    tpg, ids, flat = S.isect_tiles(g_("means2d"), g_("radii"), g_("depths"), ts, tw, th)
    assert torch.equal(tpg.cpu(), d["tiles_per_gauss"]) and torch.equal(ids.cpu(), d["isect_ids"])
    assert torch.equal(flat.cpu(), d["flatten_ids"])
    assert torch.equal(S.isect_offset_encode(ids, 1, tw, th).cpu(), d["isect_offsets"])
    P = [g_(k).clone().requires_grad_() for k in ("means2d", "conics", "colors", "opacities")]
    rc, ra = S.rasterize_to_pixels(*P, W, H, ts, g_("isect_offsets"), g_("flatten_ids"), backgrounds=g_("backgrounds"),
                                   absgrad=True)
    for a, b in ((rc, d["render_colors"]), (ra, d["render_alphas"])):
        err = (a.detach().cpu() - b).abs()
        assert (err > 1e-4 + 1e-4 * b.abs()).float().mean().item() < 2e-3 and err.max().item() < 2e-2, err.max()
    g = torch.autograd.grad((rc * g_("v_render_colors")).sum() + (ra * g_("v_render_alphas")).sum(), P)
    for name, a in zip(("v_means2d", "v_conics", "v_colors", "v_opacities"), g):
        assert_grad_close(a, d[name], what=f"kernel raster {name} vs reference CUDA", frac_ok=0.995)
    assert_grad_close(P[0].absgrad, d["v_means2d_abs"], what="absgrad vs reference CUDA", frac_ok=0.995)
