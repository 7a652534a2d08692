"""GPU parity of the f4 row: splat activations, the staged / split colour
stage, rasterize_splats and the fused L1 + SSIM loss, each against the CPU oracle
(oracle/step_ref.py, oracle/torch_ref.py).  Floating point: 1e-5 abs on activations and
colours, 1e-5 abs on the loss, 1e-3 rel on gradients (written below)."""
import pytest
import torch

import splat_one_b200 as S
from oracle import raster_ref as RC
from oracle import step_ref as SR
from oracle import torch_ref as O
from parity import assert_grad_close, assert_image_close
from splat_one_b200 import synthetic, wrapper

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_activations_match_torch():
    g = torch.Generator().manual_seed(0)
    sr = (torch.randn(1001, 3, generator=g) * 2 - 3).requires_grad_()
    orr = (torch.randn(1001, generator=g) * 3).requires_grad_()
    sg, og = sr.detach().to(DEV).requires_grad_(), orr.detach().to(DEV).requires_grad_()
    s, o = S.splat_activations(sg, og)
    s_ref, o_ref = SR.splat_activations(sr, orr)
    assert torch.allclose(s.cpu(), s_ref, rtol=2e-6, atol=1e-7) and torch.allclose(o.cpu(), o_ref, rtol=2e-6, atol=1e-7)
    vs, vo = torch.randn(1001, 3, generator=g), torch.randn(1001, generator=g)
    gs, go = torch.autograd.grad((s * vs.to(DEV)).sum() + (o * vo.to(DEV)).sum(), (sg, og))
    gs_ref, go_ref = torch.autograd.grad((s_ref * vs).sum() + (o_ref * vo).sum(), (sr, orr))
    assert_grad_close(gs, gs_ref, what="v_scales_raw", frac_ok=1.0)
    assert_grad_close(go, go_ref, what="v_opacities_raw", frac_ok=1.0)


@pytest.mark.parametrize("K,deg,split,C,N", [(16, 3, True, 1, 1000), (16, 3, False, 2, 333), (16, 1, True, 2, 97),
                                             (25, 4, True, 1, 130), (9, 2, True, 3, 64), (4, 1, False, 1, 33),
                                             (16, 0, True, 1, 40), (7, 1, True, 1, 50)])
def test_staged_colour_stage_matches_unstaged_and_oracle(K, deg, split, C, N):
    """Staged kernels (every staging mode: rows of 45/48/72/24/12/18 floats, ragged last warp)
    against the thread-per-row kernels and the CPU oracle, forward and gradients."""
    g = torch.Generator().manual_seed(K * 100 + N)
    means = torch.randn(N, 3, generator=g) + torch.tensor([0.0, 0.0, 4.0])
    table = torch.randn(N, K, 3, generator=g) * 0.3
    viewmats = torch.eye(4).repeat(C, 1, 1)
    viewmats[:, :3, 3] = torch.randn(C, 3, generator=g) * 0.2
    radii = (torch.rand(C, N, generator=g) > 0.2).int() * 3
    v = torch.randn(C, N, 3, generator=g)

    mg = means.to(DEV).requires_grad_()
    tg = table.to(DEV).requires_grad_()
    ref = wrapper.sh_view_colors(deg, mg, viewmats.to(DEV), tg, radii.to(DEV))
    gm_ref, gt_ref = torch.autograd.grad((ref * v.to(DEV)).sum(), (mg, tg))

    m2 = means.to(DEV).requires_grad_()
    if split:
        sh0 = table[:, :1].contiguous().to(DEV).requires_grad_()
        shN = table[:, 1:].contiguous().to(DEV).requires_grad_()
        got = wrapper.sh_view_colors_split(deg, m2, viewmats.to(DEV), sh0, shN, radii.to(DEV))
        gm, g0, gN = torch.autograd.grad((got * v.to(DEV)).sum(), (m2, sh0, shN))
        gt = torch.cat([g0, gN], 1)
    else:
        t2 = table.to(DEV).requires_grad_()
        got = wrapper.sh_view_colors_split(deg, m2, viewmats.to(DEV), None, t2, radii.to(DEV))
        gm, gt = torch.autograd.grad((got * v.to(DEV)).sum(), (m2, t2))
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6), (got - ref).abs().max()
    assert_grad_close(gt, gt_ref, what="v_table", frac_ok=1.0)
    assert_grad_close(gm, gm_ref, what="v_means", frac_ok=1.0)

    # CPU oracle: dirs -> SH -> +0.5 -> clamp (G/rendering.py:368-392)
    mc, tc = means.clone().requires_grad_(), table.clone().requires_grad_()
    campos = torch.inverse(viewmats)[:, :3, 3]
    dirs = mc[None] - campos[:, None]
    col = O.spherical_harmonics(deg, dirs, tc[None].expand(C, -1, -1, -1), masks=radii > 0)
    col = torch.clamp_min(col + 0.5, 0.0) * (radii > 0)[..., None]
    assert torch.allclose(got.cpu(), col, rtol=1e-4, atol=1e-5), (got.cpu() - col).abs().max()
    gm_o, gt_o = torch.autograd.grad((col * v).sum(), (mc, tc), allow_unused=True)
    gm_o = torch.zeros_like(mc) if gm_o is None else gm_o  # degree 0 does not depend on the direction
    assert_grad_close(gt, gt_o, what="v_table vs oracle", frac_ok=1.0)
    assert_grad_close(gm, gm_o, what="v_means vs oracle", frac_ok=1.0)


@pytest.mark.parametrize("C,H,W", [(1, 64, 80), (2, 45, 33), (1, 11, 11), (1, 270, 480)])
def test_l1_ssim_loss_matches_oracle(C, H, W):
    g = torch.Generator().manual_seed(H * 7 + W)
    a = torch.rand(C, H, W, 3, generator=g)
    b = (a + 0.15 * torch.randn(C, H, W, 3, generator=g)).clamp(0, 1)
    a[0, :3, :3] = b[0, :3, :3]  # exact ties: sign(0) = 0 in the L1 gradient
    ac = a.clone().requires_grad_()
    loss_ref, l1_ref, ssim_ref = SR.l1_ssim_loss(ac, b, 0.2)
    (g_ref,) = torch.autograd.grad(loss_ref * 1.7, ac)
    ag = a.to(DEV).requires_grad_()
    loss, terms = S.l1_ssim_loss(ag, b.to(DEV), 0.2, return_terms=True)
    (g_got,) = torch.autograd.grad(loss * 1.7, ag)
    assert abs(loss.item() - loss_ref.item()) < 1e-5, (loss.item(), loss_ref.item())
    assert abs(terms[1].item() - l1_ref.item()) < 1e-5 and abs(terms[2].item() - ssim_ref.item()) < 1e-5
    assert_grad_close(g_got, g_ref, what="v_colors of the loss", frac_ok=0.9999)
    # no-grad call: no derivative maps, same value
    with torch.no_grad():
        loss2 = S.l1_ssim_loss(a.to(DEV), b.to(DEV), 0.2)
    assert abs(loss2.item() - loss.item()) < 1e-7


@pytest.mark.parametrize("a,b", [(0.3, 0.7), (0.05, 0.9)])
def test_l1_ssim_kernel_known_answer_on_constant_images(a, b):
    """The fused L1 + SSIM kernel against the hand-derived closed form for constant images
    (tests/parity.py::ssim_constant_images_closed_form), not against any restated convolution."""
    from parity import ssim_constant_images_closed_form as kat

    H, W = 45, 70
    x = torch.full((1, H, W, 3), a, device=DEV)
    y = torch.full((1, H, W, 3), b, device=DEV)
    loss, terms = S.l1_ssim_loss(x, y, 0.2, return_terms=True)
    ssim = kat(a, b, H, W, "valid")
    # 3e-4: with zero true variance the fp32 rounding noise of E[x^2] - mu^2 (~1e-7) is measured against
    # C2 = 9e-4; the fp32 oracle shows the same spread (tests/test_oracle_step.py)
    assert abs(terms[2].item() - ssim) < 3e-4, (terms[2].item(), ssim)
    assert abs(terms[1].item() - abs(a - b)) < 1e-6
    assert abs(loss.item() - (0.8 * abs(a - b) + 0.2 * (1 - ssim))) < 1e-4


def test_rasterize_splats_and_loss_match_oracle_step():
    """One whole step (activations -> split SH -> rasterization -> L1+SSIM -> backward) against
    the oracle step on the same raw parameters."""
    W, H, N = 160, 112, 4000
    scene = synthetic.pinhole_scene(N, W, H, seed=5)
    raw = {
        "means": scene["means"], "quats": scene["quats"], "scales": torch.log(scene["scales"]),
        "opacities": torch.logit(scene["opacities"].clamp(1e-4, 1 - 1e-4)),
        "sh0": scene["sh"][:, :1].contiguous(), "shN": scene["sh"][:, 1:].contiguous(),
    }
    names = list(raw)
    g = torch.Generator().manual_seed(9)
    pixels = torch.rand(1, H, W, 3, generator=g)
    c2w = torch.inverse(scene["viewmats"])

    Pc = {k: v.clone().requires_grad_() for k, v in raw.items()}
    rc_r, ra_r, _ = SR.rasterize_splats(Pc, c2w, scene["Ks"], W, H, sh_degree=3, packed=False,
                                        raster_fn=RC.rasterize_to_pixels)
    loss_r, _, _ = SR.l1_ssim_loss(rc_r, pixels, 0.2)
    g_r = torch.autograd.grad(loss_r, [Pc[k] for k in names])

    Pg = {k: v.to(DEV).requires_grad_() for k, v in raw.items()}
    rc, ra, info = S.rasterize_splats(Pg, c2w.to(DEV), scene["Ks"].to(DEV), W, H, sh_degree=3, packed=False)
    loss = S.l1_ssim_loss(rc, pixels.to(DEV), 0.2)
    g_g = torch.autograd.grad(loss, [Pg[k] for k in names])
    err = (rc.cpu() - rc_r).abs()
    assert (err > 1e-4 + 1e-4 * rc_r.abs()).float().mean().item() < 2e-3, err.max()
    assert abs(loss.item() - loss_r.item()) < 2e-5, (loss.item(), loss_r.item())
    for n, a, b in zip(names, g_g, g_r):
        assert_grad_close(a, b, rtol=2e-3, what=f"step grad {n}", frac_ok=0.995)
    # packed mode takes the concatenating route: identical to rasterization() on the cat'd table
    det = {k: v.detach() for k, v in Pg.items()}
    rc_p, _, _ = S.rasterize_splats(det, c2w.to(DEV), scene["Ks"].to(DEV), W, H, sh_degree=3, packed=True)
    sc, op = S.splat_activations(det["scales"], det["opacities"])
    rc_q, _, _ = S.rasterization(det["means"], det["quats"], sc, op, torch.cat([det["sh0"], det["shN"]], 1),
                                 scene["viewmats"].to(DEV), scene["Ks"].to(DEV), W, H, sh_degree=3, packed=True)
    assert (rc_p - rc_q).abs().max().item() < 1e-6


@pytest.mark.parametrize("opts", [dict(antialiased=True), dict(render_mode="RGB+ED"), dict(absgrad=True),
                                  dict(camera_model="fisheye"), dict(packed=True, sparse_grad=True)])
def test_rasterize_splats_options_equal_composed_call(opts):
    """rasterize_splats(...) == activations + cat + rasterization(...) written out as
    gsplat_trainer.py:446-497 does, for the options the trainer forwards (incl. masks)."""
    W, H, N = 128, 96, 3000
    scene = synthetic.pinhole_scene(N, W, H, seed=8, n_cameras=2)
    raw = {
        "means": scene["means"], "quats": scene["quats"], "scales": torch.log(scene["scales"]),
        "opacities": torch.logit(scene["opacities"].clamp(1e-4, 1 - 1e-4)),
        "sh0": scene["sh"][:, :1].contiguous(), "shN": scene["sh"][:, 1:].contiguous(),
    }
    Pg = {k: v.to(DEV).requires_grad_() for k, v in raw.items()}
    Pc = {k: v.to(DEV).requires_grad_() for k, v in raw.items()}
    vm, Ks = scene["viewmats"].to(DEV), scene["Ks"].to(DEV)
    c2w = torch.inverse(vm)
    g = torch.Generator().manual_seed(4)
    masks = (torch.rand(2, H, W, generator=g) > 0.1).to(DEV)
    o = dict(opts)
    rc, ra, info = S.rasterize_splats(Pg, c2w, Ks, W, H, masks=masks, sh_degree=3, **o)
    aa = o.pop("antialiased", False)
    rc2, ra2, info2 = S.rasterization(
        Pc["means"], Pc["quats"], torch.exp(Pc["scales"]), torch.sigmoid(Pc["opacities"]),
        torch.cat([Pc["sh0"], Pc["shN"]], 1), torch.linalg.inv(c2w), Ks, W, H, sh_degree=3,
        rasterize_mode="antialiased" if aa else "classic", **{"packed": False, **o})
    rc2 = rc2.clone()
    rc2[~masks] = 0
    # the two pose inverses differ in the last bit, which can flip isolated alpha >= 1/255 decisions
    for a, b in ((rc, rc2), (ra, ra2)):
        err = (a - b).abs()
        assert (err > 2e-6 + 1e-5 * b.abs()).float().mean().item() < 1e-3, err.max()
        assert err.max().item() < 5e-3, err.max()
    assert (rc[~masks] == 0).all()
    v = torch.randn(rc.shape, generator=g).to(DEV)
    if o.get("absgrad"):
        info["means2d"].retain_grad()
        info2["means2d"].retain_grad()
    (rc * v).sum().backward()
    (rc2 * v).sum().backward()
    for k in raw:
        a, b = Pg[k].grad, Pc[k].grad
        a = a.to_dense() if a.is_sparse else a
        b = b.to_dense() if b.is_sparse else b
        assert_grad_close(a, b, what=f"rasterize_splats grad {k} ({opts})", frac_ok=0.999)
    if o.get("absgrad"):
        assert_grad_close(info["means2d"].absgrad, info2["means2d"].absgrad, what="absgrad", frac_ok=0.999)


@pytest.mark.parametrize("C", [1, 2])
@pytest.mark.parametrize("deg", [0, 3])
def test_packed_split_colour_stage_equals_concatenated_table(C, deg):
    """COO colour kernels reading sh0 / shN in place == the same kernels on torch.cat([sh0, shN], 1)
    (C = 1: rows written once; C = 2: rows accumulated with atomics), forward and gradients."""
    N, K = 3000, 16
    g = torch.Generator().manual_seed(31 + C)
    means = (torch.randn(N, 3, generator=g) + torch.tensor([0.0, 0.0, 4.0])).to(DEV)
    table = (torch.randn(N, K, 3, generator=g) * 0.3).to(DEV)
    viewmats = torch.eye(4).repeat(C, 1, 1)
    viewmats[:, :3, 3] = torch.randn(C, 3, generator=g) * 0.2
    viewmats = viewmats.to(DEV)
    vis = torch.rand(C, N, generator=g) > 0.3
    cam, gid = torch.nonzero(vis, as_tuple=True)
    cam, gid = cam.to(DEV), gid.to(DEV)
    v = torch.randn(cam.numel(), 3, generator=g).to(DEV)

    m1, t1 = means.clone().requires_grad_(), table.clone().requires_grad_()
    ref = wrapper.sh_view_colors_packed(deg, m1, viewmats, t1, cam, gid)
    gm_ref, gt_ref = torch.autograd.grad((ref * v).sum(), (m1, t1))
    m2 = means.clone().requires_grad_()
    sh0, shN = table[:, :1].contiguous().requires_grad_(), table[:, 1:].contiguous().requires_grad_()
    got = wrapper.sh_view_colors_packed_split(deg, m2, viewmats, sh0, shN, cam, gid)
    gm, g0, gN = torch.autograd.grad((got * v).sum(), (m2, sh0, shN))
    assert torch.allclose(got, ref, rtol=1e-6, atol=1e-6)
    assert_grad_close(torch.cat([g0, gN], 1), gt_ref, what="v_table", frac_ok=1.0)
    assert_grad_close(gm, gm_ref, what="v_means", frac_ok=1.0)
