"""End-to-end parity of `rasterization()` (the call splat_one makes) against the CPU
oracle pipeline, plus size-independent properties at BASELINE config B's full size."""
import math

import pytest
import torch

import splat_one_b200 as S
from oracle import raster_ref as RC
from oracle import torch_ref as O
from parity import assert_grad_close, assert_image_close
from splat_one_b200 import synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run_both(scene, packed, sh=True, render_mode="RGB", bg=False, rasterize_mode="classic", absgrad=False,
              per_view_color=False, C=None):
    C = scene["viewmats"].shape[0]
    names = ["means", "quats", "scales", "opacities"]
    if sh:
        colors = scene["sh"]
        if per_view_color:
            colors = colors[None].repeat(C, 1, 1, 1) * torch.linspace(0.5, 1.0, C)[:, None, None, None]
    else:
        g = torch.Generator().manual_seed(1)
        colors = torch.rand(scene["means"].shape[0], 3, generator=g)
        if per_view_color:
            colors = colors[None].repeat(C, 1, 1) * torch.linspace(0.5, 1.0, C)[:, None, None]
    P_c = [scene[k].clone().requires_grad_() for k in names] + [colors.clone().requires_grad_()]
    P_g = [scene[k].to(DEV).requires_grad_() for k in names] + [colors.to(DEV).requires_grad_()]
    g = torch.Generator().manual_seed(2)
    bgc = torch.rand(C, 3, generator=g) if bg else None
    kw = dict(width=scene["width"], height=scene["height"], sh_degree=scene["sh_degree"] if sh else None,
              packed=packed, render_mode=render_mode, rasterize_mode=rasterize_mode,
              camera_model=scene["camera_model"])
    ref = O.rasterization(*P_c, scene["viewmats"], scene["Ks"], backgrounds=bgc, raster_fn=RC.rasterize_to_pixels, **kw)
    got = S.rasterization(*P_g, scene["viewmats"].to(DEV), scene["Ks"].to(DEV),
                          backgrounds=None if bgc is None else bgc.to(DEV), absgrad=absgrad, **kw)
    return ref, got, P_c, P_g


def _margin_of(ref_meta, scene, bgc=None):
    m = ref_meta
    out = RC.raster_fwd(m["means2d"], m["conics"], m["colors"], m["opacities"], scene["width"], scene["height"],
                        m["tile_size"], m["isect_offsets"], m["flatten_ids"], None, None, want_margin=True)
    return out[3]


@pytest.mark.parametrize("packed", [False, True])
@pytest.mark.parametrize("sh", [True, False])
@pytest.mark.parametrize("render_mode", ["RGB", "RGB+D", "D", "RGB+ED", "ED"])
def test_rasterization_pinhole_matches_oracle(packed, sh, render_mode):
    scene = synthetic.pinhole_scene(6000, 208, 152, seed=7, n_cameras=2)
    ref, got, P_c, P_g = _run_both(scene, packed, sh, render_mode, bg=True)
    (rc_r, ra_r, m_r), (rc, ra, m) = ref, got
    X = {"RGB": 3, "RGB+D": 4, "RGB+ED": 4, "D": 1, "ED": 1}[render_mode]
    assert rc.shape == (2, 152, 208, X) and ra.shape == (2, 152, 208, 1)
    for k in ["camera_ids", "gaussian_ids", "radii", "means2d", "depths", "conics", "opacities", "tile_width",
              "tile_height", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets", "width", "height",
              "tile_size", "n_cameras"]:
        assert k in m, k
    margin = _margin_of(m_r, scene)
    tol = 1e-4 if render_mode in ("RGB",) else 1e-3  # depth channels carry values ~10
    assert_image_close(rc, rc_r, margin, atol=tol, rtol=1e-4, what=f"colors[{render_mode}]")
    assert_image_close(ra, ra_r, margin[..., None], what="alphas")
    # the integer stage must agree exactly whenever the projection agrees exactly
    if torch.equal(m["radii"].cpu(), m_r["radii"]) and torch.equal(m["means2d"].cpu(), m_r["means2d"].detach()) \
            and torch.equal(m["depths"].cpu(), m_r["depths"].detach()):
        assert torch.equal(m["isect_ids"].cpu(), m_r["isect_ids"])
        assert torch.equal(m["flatten_ids"].cpu(), m_r["flatten_ids"])
    g = torch.Generator().manual_seed(4)
    keep = (margin >= 1e-3).float()[..., None]
    vc = torch.randn(rc_r.shape, generator=g) * keep
    va = torch.randn(ra_r.shape, generator=g) * keep
    g_ref = torch.autograd.grad((rc_r * vc).sum() + (ra_r * va).sum(), P_c, allow_unused=True)
    g_got = torch.autograd.grad((rc * vc.to(DEV)).sum() + (ra * va.to(DEV)).sum(), P_g, allow_unused=True)
    for n, a, b in zip(["means", "quats", "scales", "opacities", "colors"], g_got, g_ref):
        assert (a is None) == (b is None), n  # colours are unused in the depth-only modes
        if a is not None:
            assert_grad_close(a, b, rtol=2e-3, what=f"{render_mode}/{n}", frac_ok=0.995)


@pytest.mark.parametrize("packed", [False, True])
def test_rasterization_spherical_matches_oracle(packed):
    scene = synthetic.spherical_scene(6000, 256, 128, seed=3, footprint_px=2.0)
    ref, got, P_c, P_g = _run_both(scene, packed, True, "RGB+ED")
    (rc_r, ra_r, m_r), (rc, ra, m) = ref, got
    margin = _margin_of(m_r, scene)
    assert_image_close(rc, rc_r, margin, atol=1e-3, rtol=1e-4, what="colors")
    assert_image_close(ra, ra_r, margin[..., None], what="alphas")
    g = torch.Generator().manual_seed(4)
    keep = (margin >= 1e-3).float()[..., None]
    vc = torch.randn(rc_r.shape, generator=g) * keep
    va = torch.randn(ra_r.shape, generator=g) * keep
    g_ref = torch.autograd.grad((rc_r * vc).sum() + (ra_r * va).sum(), P_c)
    g_got = torch.autograd.grad((rc * vc.to(DEV)).sum() + (ra * va.to(DEV)).sum(), P_g)
    for n, a, b in zip(["means", "quats", "scales", "opacities", "colors"], g_got, g_ref):
        assert_grad_close(a, b, rtol=2e-3, what=f"spherical/{n}", frac_ok=0.995)


def test_rasterization_antialiased_per_view_colors_absgrad():
    scene = synthetic.pinhole_scene(4000, 160, 128, seed=9, n_cameras=3)
    ref, got, P_c, P_g = _run_both(scene, False, True, "RGB", bg=False, rasterize_mode="antialiased", absgrad=True,
                                   per_view_color=True)
    (rc_r, ra_r, m_r), (rc, ra, m) = ref, got
    margin = _margin_of(m_r, scene)
    assert_image_close(rc, rc_r, margin, what="colors")
    m["means2d"].retain_grad()
    (rc.sum() + ra.sum()).backward()
    assert m["means2d"].grad is not None and m["means2d"].grad.shape == (3, 4000, 2)
    assert m["means2d"].absgrad.shape == (3, 4000, 2)
    assert (m["means2d"].absgrad >= m["means2d"].grad.abs() - 1e-4).all()


def test_golden_pipeline_vectors(golden):
    g = golden("pipeline_oracle.npz")
    P = [g[k].to(DEV).requires_grad_() for k in ("means", "quats", "scales", "opacities", "sh")]
    rc, ra, meta = S.rasterization(*P, g["viewmats"].to(DEV), g["Ks"].to(DEV), g["width"], g["height"], sh_degree=3,
                                   packed=False)
    assert_image_close(rc, g["render_colors"], None, atol=2e-4, what="golden colours")
    assert_image_close(ra, g["render_alphas"], None, atol=2e-4, what="golden alphas")
    grads = torch.autograd.grad((rc * g["v_render_colors"].to(DEV)).sum() + (ra * g["v_render_alphas"].to(DEV)).sum(), P)
    for n, a in zip(["v_means", "v_quats", "v_scales", "v_opacities", "v_sh"], grads):
        assert_grad_close(a, g[n], rtol=2e-3, what=n, frac_ok=0.99)


def test_garden_fixture_renders(golden):
    """The reference's own test scene (subset of assets/test_garden.npz, 3 cameras)."""
    g = golden("garden_subset.npz")
    N = g["means"].shape[0]
    gen = torch.Generator().manual_seed(42)
    scales = torch.rand(N, 3, generator=gen) * 0.02
    quats = torch.nn.functional.normalize(torch.randn(N, 4, generator=gen), dim=-1)
    opac = torch.rand(N, generator=gen)
    args_c = (g["means"], quats, scales, opac, g["colors"], g["viewmats"], g["Ks"], g["width"], g["height"])
    ref = O.rasterization(*args_c, packed=False, raster_fn=RC.rasterize_to_pixels)
    got = S.rasterization(*[a.to(DEV) if isinstance(a, torch.Tensor) else a for a in args_c], packed=False)
    margin = _margin_of(ref[2], dict(width=g["width"], height=g["height"]))
    assert_image_close(got[0], ref[0], margin, what="garden colours")
    assert_image_close(got[1], ref[1], margin[..., None], what="garden alphas")


def test_config_b_full_size_properties():
    """1 M Gaussians, SH3, 1920x1080 (BASELINE config B): size-independent properties."""
    scene = synthetic.to_device(synthetic.pinhole_scene(1_000_000, 1920, 1080, seed=42), DEV)
    P = [scene[k].clone().requires_grad_() for k in ("means", "quats", "scales", "opacities", "sh")]
    rc, ra, m = S.rasterization(*P, scene["viewmats"], scene["Ks"], 1920, 1080, sh_degree=3, packed=False)
    assert torch.isfinite(rc).all() and torch.isfinite(ra).all()
    assert (ra >= 0).all() and (ra <= 1.0).all() and (rc >= -1e-6).all()
    ids, fl, offs = m["isect_ids"], m["flatten_ids"], m["isect_offsets"].flatten().long()
    n_isects = ids.numel()
    assert n_isects == int(m["tiles_per_gauss"].sum()) and n_isects > 1_000_000
    assert (ids[1:] >= ids[:-1]).all(), "isect_ids must be sorted"
    # stable: equal keys keep ascending flat index
    eq = ids[1:] == ids[:-1]
    assert (fl[1:][eq] > fl[:-1][eq]).all()
    assert (offs[1:] >= offs[:-1]).all() and offs[0] == 0 and offs[-1] <= n_isects
    # offsets really are the first index of each tile
    tile_of = (ids >> 32)
    k = torch.randint(0, offs.numel(), (2000,), device=DEV)
    start = offs[k]
    end = torch.where(k + 1 < offs.numel(), offs[(k + 1).clamp(max=offs.numel() - 1)], torch.full_like(start, n_isects))
    nz = end > start
    assert (tile_of[start[nz]] == k[nz]).all() and (tile_of[end[nz] - 1] == k[nz]).all()
    # depth bits of every key equal the depth of the Gaussian it points to
    assert torch.equal((ids & 0xFFFFFFFF).int(), m["depths"].flatten()[fl.long()].view(torch.int32))
    # linearity in the colours for fixed geometry: render(a*c) = a*render(c)
    rc2, _ = S.rasterize_to_pixels(m["means2d"].detach(), m["conics"].detach(),
                                   torch.ones_like(m["means2d"][..., :1]).expand(-1, -1, 3).contiguous() * 0.5,
                                   m["opacities"].detach(), 1920, 1080, 16, m["isect_offsets"], fl)
    torch.testing.assert_close(rc2, (ra * 0.5).expand(-1, -1, -1, 3), rtol=1e-4, atol=1e-5)
    # gradients: finite, SH0 gradient equals C0 * d(colour) summed, opacity grads only on visible Gaussians
    vc = torch.randn_like(rc)
    g = torch.autograd.grad((rc * vc).sum() + ra.sum(), P)
    for t in g:
        assert torch.isfinite(t).all()
    invisible = (m["radii"][0] == 0)
    assert g[3][invisible].abs().max() == 0 and g[0][invisible].abs().max() == 0
    assert (g[3].abs() > 0).float().mean() > 0.05
