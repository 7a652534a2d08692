"""GPU parity of (1) rasterize_to_pixels against vectors produced by the reference's own
pure-PyTorch compositing, and (2) the §8(f) operators: rasterize_to_indices_in_range +
accumulate and the un-fused projection chain, against reference-generated vectors and the
CPU oracle.  Everything goes through the public Python API -> C ABI of libb200splat.so."""
import math
import os

import numpy as np

import pytest
import torch

import splat_one_b200 as S
from oracle import torch_ref as O
from parity import assert_grad_close, assert_image_close
from test_gpu_ops import _raster_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _g(t):
    return t.to(DEV) if isinstance(t, torch.Tensor) else t


# ------------------------------------------------------------------------------------
# a8 / a9 against the reference's _rasterize_to_pixels + accumulate (tests/golden/raster_ref_*)
# ------------------------------------------------------------------------------------
@pytest.mark.parametrize("D", [3, 1])
def test_rasterize_matches_reference_compositing(golden, D):
    d = golden(f"raster_ref_d{D}.npz")
    P = [_g(d[k]).clone().requires_grad_() for k in ("means2d", "conics", "colors", "opacities", "backgrounds")]
    rc, ra = S.rasterize_to_pixels(P[0], P[1], P[2], P[3], d["width"], d["height"], d["tile_size"],
                                   _g(d["isect_offsets"]), _g(d["flatten_ids"]), backgrounds=P[4])
    assert_image_close(rc, d["ref_render_colors"], d["margin"], what="colors vs reference")
    assert_image_close(ra, d["ref_render_alphas"], d["margin"], what="alphas vs reference")
    grads = torch.autograd.grad((rc * _g(d["v_render_colors"])).sum() + (ra * _g(d["v_render_alphas"])).sum(), P)
    for name, g in zip(("means2d", "conics", "colors", "opacities", "backgrounds"), grads):
        assert_grad_close(g, d[f"ref_v_{name}"], what=f"v_{name} vs reference autograd")


# ------------------------------------------------------------------------------------
# f1: rasterize_to_indices_in_range / accumulate
# ------------------------------------------------------------------------------------
def _robust_equal(got, ref, margin, H, W, what):
    """Lists agree exactly on every pixel whose decisions are not within 1e-3 of a threshold;
    few pixels are ambiguous."""
    robust = (margin.flatten() >= 1e-3)
    assert (~robust).float().mean() < 5e-2  # a property of the data: ~40 decisions per pixel
    gg, gp, gc = (t.cpu() for t in got)
    rg, rp, rc_ = ref
    kg = robust[gc * H * W + gp]
    kr = robust[rc_ * H * W + rp]
    assert torch.equal(gg[kg], rg[kr]), what
    assert torch.equal(gp[kg], rp[kr]) and torch.equal(gc[kg], rc_[kr]), what
    assert abs(gg.numel() - rg.numel()) <= max(8, int(2e-3 * rg.numel())), (gg.numel(), rg.numel())


def test_indices_match_golden(golden):
    d = golden("raster_ref_d3.npz")
    C, H, W = d["means2d"].shape[0], d["height"], d["width"]
    got = S.rasterize_to_indices_in_range(0, 10**10, torch.ones(C, H, W, device=DEV), _g(d["means2d"]), _g(d["conics"]),
                                          _g(d["opacities"]), W, H, d["tile_size"], _g(d["isect_offsets"]),
                                          _g(d["flatten_ids"]))
    assert all(t.dtype == torch.int64 and t.is_cuda for t in got)
    _robust_equal(got, (d["idx_gaussian_ids"], d["idx_pixel_ids"], d["idx_camera_ids"]), d["margin"], H, W, "golden")


@pytest.mark.parametrize("ts,W,H", [(16, 200, 136), (8, 100, 70), (4, 37, 29)])
@pytest.mark.parametrize("rng", [(0, 1), (1, 3), (0, 10**10), (50, 60)])
def test_indices_ranges_match_oracle(ts, W, H, rng):
    x = _raster_inputs(C=2, N=3000, W=W, H=H, ts=ts, D=3, seed=ts)
    C = x["C"]
    g = torch.Generator().manual_seed(3)
    T0 = torch.rand(C, H, W, generator=g) * 0.9 + 0.1
    args = (x["m2"], x["con"], x["op"], W, H, ts, x["offs"], x["fl"])
    *ref, margin = O.rasterize_to_indices_in_range(rng[0], rng[1], T0, *args, return_margin=True)
    got = S.rasterize_to_indices_in_range(rng[0], rng[1], _g(T0), *[_g(a) for a in args])
    _robust_equal(got, ref, margin, H, W, f"range {rng}")


def test_reference_style_compositing_loop_equals_fused_kernel():
    """The reference's own check (tests/test_basic.py::test_rasterize_to_pixels): iterate
    rasterize_to_indices_in_range + accumulate like `_rasterize_to_pixels`
    (_torch_impl.py:620-668) and compare with the fused kernel, forward and gradients."""
    x = _raster_inputs(C=2, N=3000, W=120, H=88, ts=16, D=3, seed=21)
    C, W, H, ts = x["C"], x["W"], x["H"], x["ts"]
    offs, fl = _g(x["offs"]), _g(x["fl"])
    bg0 = torch.rand(C, 3)

    def leaves():
        return [_g(x[k]).clone().requires_grad_() for k in ("m2", "con", "col", "op")] + [_g(bg0).clone().requires_grad_()]

    Pa, Pb = leaves(), leaves()
    rc, ra = S.rasterize_to_pixels(Pa[0], Pa[1], Pa[2], Pa[3], W, H, ts, offs, fl, backgrounds=Pa[4])
    # torch loop
    n_isects = fl.numel()
    render_colors = torch.zeros((C, H, W, 3), device=DEV)
    render_alphas = torch.zeros((C, H, W, 1), device=DEV)
    block = ts * ts
    ofl = torch.cat([offs.flatten(), torch.tensor([n_isects], device=DEV, dtype=torch.int32)])
    num_batches = (int((ofl[1:] - ofl[:-1]).max()) + block - 1) // block
    per_iter = 100  # the reference's default batch_per_iter (_torch_impl.py:586): one walk per pixel here
    for step in range(0, num_batches, per_iter):
        trans = 1.0 - render_alphas[..., 0]
        gi, pi, ci = S.rasterize_to_indices_in_range(step, step + per_iter, trans, Pb[0], Pb[1], Pb[3], W, H, ts,
                                                     offs, fl)
        if len(gi) == 0:
            break
        r_s, a_s = S.accumulate(Pb[0], Pb[1], Pb[3], Pb[2], gi, pi, ci, W, H)
        render_colors = render_colors + r_s * trans[..., None]
        render_alphas = render_alphas + a_s * trans[..., None]
    render_colors = render_colors + Pb[4][:, None, None, :] * (1.0 - render_alphas)
    assert (rc - render_colors).abs().max() < 5e-4
    assert (ra - render_alphas).abs().max() < 5e-4
    g = torch.Generator().manual_seed(0)
    vc, va = _g(torch.randn(rc.shape, generator=g)), _g(torch.randn(ra.shape, generator=g))
    ga = torch.autograd.grad((rc * vc).sum() + (ra * va).sum(), Pa)
    gb = torch.autograd.grad((render_colors * vc).sum() + (render_alphas * va).sum(), Pb)
    for name, a, b in zip(("means2d", "conics", "colors", "opacities", "backgrounds"), ga, gb):
        assert_grad_close(a, b, rtol=3e-3, what=f"fused vs compositing loop v_{name}", frac_ok=0.995)


def test_indices_empty_cases():
    C, N, W, H = 1, 5, 40, 24
    offs = torch.zeros(C, 2, 3, dtype=torch.int32, device=DEV)
    fl = torch.zeros(0, dtype=torch.int32, device=DEV)
    out = S.rasterize_to_indices_in_range(0, 10, torch.ones(C, H, W, device=DEV), torch.zeros(C, N, 2, device=DEV),
                                          torch.ones(C, N, 3, device=DEV), torch.ones(C, N, device=DEV), W, H, 16,
                                          offs, fl)
    assert all(t.numel() == 0 and t.dtype == torch.int64 for t in out)
    rc, ra = S.accumulate(torch.zeros(C, N, 2, device=DEV), torch.ones(C, N, 3, device=DEV),
                          torch.ones(C, N, device=DEV), torch.ones(C, N, 3, device=DEV), *out, W, H)
    assert rc.shape == (C, H, W, 3) and rc.abs().sum() == 0 and ra.abs().sum() == 0
    with pytest.raises(AssertionError):
        S.rasterize_to_indices_in_range(0, 10, torch.ones(C, H, W, device=DEV), torch.zeros(C, N, 2, device=DEV),
                                        torch.ones(C, N, 3, device=DEV), torch.ones(C, N, device=DEV), W, H, 8, offs, fl)
    with pytest.raises(RuntimeError):
        S.rasterize_to_indices_in_range(0, 10, torch.ones(C, H, W), torch.zeros(C, N, 2), torch.ones(C, N, 3),
                                        torch.ones(C, N), W, H, 16, offs.cpu(), fl.cpu())


# ------------------------------------------------------------------------------------
# f2: quat_scale_to_covar_preci / world_to_cam / proj
# ------------------------------------------------------------------------------------
@pytest.mark.parametrize("triu", [False, True])
@pytest.mark.parametrize("which", ["both", "covar", "preci"])
def test_quat_scale_to_covar_preci(golden, triu, which):
    d = golden("unfused_ref.npz")
    t = "triu" if triu else "full"
    cc, cp = which in ("both", "covar"), which in ("both", "preci")
    q, s = _g(d["quats"]).clone().requires_grad_(), _g(d["scales"]).clone().requires_grad_()
    cov, pre = S.quat_scale_to_covar_preci(q, s, cc, cp, triu)
    assert (cov is None) == (not cc) and (pre is None) == (not cp)
    loss = 0
    if cc:
        torch.testing.assert_close(cov.cpu(), d[f"ref_covars_{t}"], rtol=1e-4, atol=1e-6)
        loss = loss + (cov * _g(d[f"v_covars_{t}"])).sum()
    if cp:
        torch.testing.assert_close(pre.cpu(), d[f"ref_precis_{t}"], rtol=1e-4, atol=1e-2)
        loss = loss + (pre * _g(d[f"v_precis_{t}"])).sum()
    gq, gs = torch.autograd.grad(loss, (q, s))
    if which == "both":
        assert_grad_close(gq, d[f"ref_v_quats_{t}"], what="v_quats vs reference")
        assert_grad_close(gs, d[f"ref_v_scales_{t}"], what="v_scales vs reference")
    else:  # partial outputs: against autograd of the oracle
        qc, sc = d["quats"].clone().requires_grad_(), d["scales"].clone().requires_grad_()
        oc, op_ = O.quat_scale_to_covar_preci(qc, sc, cc, cp, triu)
        lo = (oc * d[f"v_covars_{t}"]).sum() if cc else (op_ * d[f"v_precis_{t}"]).sum()
        rq, rs = torch.autograd.grad(lo, (qc, sc))
        assert_grad_close(gq, rq, what="v_quats")
        assert_grad_close(gs, rs, what="v_scales")


def test_world_to_cam(golden):
    d = golden("unfused_ref.npz")
    m, cv, vm = (_g(d[k]).clone().requires_grad_() for k in ("means", "covars", "viewmats"))
    mc, cc = S.world_to_cam(m, cv, vm)
    torch.testing.assert_close(mc.cpu(), d["ref_means_c"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(cc.cpu(), d["ref_covars_c"], rtol=1e-4, atol=1e-6)
    g = torch.autograd.grad((mc * _g(d["v_means_c"])).sum() + (cc * _g(d["v_covars_c"])).sum(), (m, cv, vm))
    for got, k in zip(g, ("ref_w2c_v_means", "ref_w2c_v_covars", "ref_w2c_v_viewmats")):
        assert_grad_close(got, d[k], what=k)
    # only some inputs need gradients
    m2 = _g(d["means"]).clone().requires_grad_()
    mc2, cc2 = S.world_to_cam(m2, _g(d["covars"]), _g(d["viewmats"]))
    (g2,) = torch.autograd.grad((mc2 * _g(d["v_means_c"])).sum() + cc2.sum(), (m2,))
    assert_grad_close(g2, d["ref_w2c_v_means"], what="v_means only")


@pytest.mark.parametrize("cm", ["pinhole", "ortho", "fisheye", "spherical"])
def test_proj(golden, cm):
    d = golden("unfused_ref.npz")
    W, H = d["width"], d["height"]
    a, b = _g(d["ref_means_c"]).clone().requires_grad_(), _g(d["ref_covars_c"]).clone().requires_grad_()
    m2, c2 = S.proj(a, b, _g(d["Ks"]), W, H, cm)
    ga, gb = torch.autograd.grad((m2 * _g(d["v_means2d"])).sum() + (c2 * _g(d["v_covars2d"])).sum(), (a, b))
    if cm != "spherical":  # the reference's _torch_impl has no faithful spherical model
        torch.testing.assert_close(m2.cpu(), d[f"ref_means2d_{cm}"], rtol=1e-4, atol=2e-3)
        torch.testing.assert_close(c2.cpu(), d[f"ref_covars2d_{cm}"], rtol=1e-3, atol=1e-3)
        assert_grad_close(ga, d[f"ref_proj_v_means_{cm}"], what=f"proj v_means {cm} vs reference", rtol=2e-3)
        assert_grad_close(gb, d[f"ref_proj_v_covars_{cm}"], what=f"proj v_covars {cm} vs reference", rtol=2e-3)
    ac, bc = d["ref_means_c"].clone().requires_grad_(), d["ref_covars_c"].clone().requires_grad_()
    om, oc = O.proj(ac, bc, d["Ks"], W, H, cm)
    torch.testing.assert_close(m2.cpu(), om, rtol=1e-4, atol=2e-3)
    torch.testing.assert_close(c2.cpu(), oc, rtol=1e-3, atol=1e-3)
    ra, rb = torch.autograd.grad((om * d["v_means2d"]).sum() + (oc * d["v_covars2d"]).sum(), (ac, bc))
    assert_grad_close(ga, ra, what=f"proj v_means {cm}", rtol=2e-3)
    assert_grad_close(gb, rb, what=f"proj v_covars {cm}", rtol=2e-3)
    with pytest.raises(AttributeError):
        S.proj(a, b, _g(d["Ks"]), W, H, "lens")


def test_unfused_chain_equals_fused_projection():
    """quat_scale_to_covar_preci -> world_to_cam -> proj reproduces the means2d of
    fully_fused_projection and (after blur + inverse) its conics, as tests/test_basic.py
    composes them."""
    g = torch.Generator().manual_seed(4)
    N, C, W, H = 2000, 2, 320, 240
    means = torch.rand(N, 3, generator=g) * 2 - 1
    means[:, 2] += 3
    quats, scales = torch.randn(N, 4, generator=g), torch.rand(N, 3, generator=g) * 0.1 + 0.01
    vm = torch.eye(4).expand(C, -1, -1).contiguous().clone()
    vm[1, :3, 3] = torch.tensor([0.1, 0.0, 0.3])
    Ks = torch.tensor([[300.0, 0, 160], [0, 300.0, 120], [0, 0, 1]]).expand(C, -1, -1).contiguous()
    covars, _ = S.quat_scale_to_covar_preci(_g(quats), _g(scales), True, False)
    mc, cc = S.world_to_cam(_g(means), covars, _g(vm))
    m2, c2 = S.proj(mc, cc, _g(Ks), W, H)
    radii, fm2, dep, con, _ = S.fully_fused_projection(_g(means), None, _g(quats), _g(scales), _g(vm), _g(Ks), W, H)
    sel = radii > 0
    assert sel.float().mean() > 0.3
    torch.testing.assert_close(m2[sel], fm2[sel], rtol=1e-4, atol=1e-3)
    c2 = c2.clone()
    c2[..., 0, 0] += 0.3
    c2[..., 1, 1] += 0.3
    inv = torch.inverse(c2[sel])
    torch.testing.assert_close(torch.stack([inv[:, 0, 0], inv[:, 0, 1], inv[:, 1, 1]], -1), con[sel], rtol=2e-3,
                               atol=1e-5)
    # empty inputs
    e = S.quat_scale_to_covar_preci(torch.zeros(0, 4, device=DEV), torch.zeros(0, 3, device=DEV))
    assert e[0].shape == (0, 3, 3) and e[1].shape == (0, 3, 3)
    e = S.world_to_cam(torch.zeros(0, 3, device=DEV), torch.zeros(0, 3, 3, device=DEV), _g(vm))
    assert e[0].shape == (C, 0, 3) and e[1].shape == (C, 0, 3, 3)


# ------------------------------------------------------------------------------------
# f3: selective_adam_update / SelectiveAdam / compute_relocation
# ------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(1000, 3), (777, 15, 3), (512,), (64, 4)])
def test_selective_adam_matches_oracle(shape):
    g = torch.Generator().manual_seed(len(shape))
    N = shape[0]
    param, grad = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
    m, v = torch.randn(shape, generator=g) * 0.1, torch.rand(shape, generator=g) * 0.1
    vis = torch.rand(N, generator=g) > 0.4
    lr, b1, b2, eps = 1e-2, 0.9, 0.999, 1e-8
    rp, rm, rv = O.selective_adam_update(param, grad, m, v, vis, lr, b1, b2, eps)
    gp, gm, gv = _g(param).clone(), _g(m).clone(), _g(v).clone()
    S.selective_adam_update(gp, _g(grad), gm, gv, _g(vis), lr, b1, b2, eps, N, param.numel() // N)
    torch.testing.assert_close(gp.cpu(), rp, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(gm.cpu(), rm, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(gv.cpu(), rv, rtol=2e-6, atol=2e-7)
    # invisible rows are bit-identical to the input
    assert torch.equal(gp.cpu()[~vis], param[~vis]) and torch.equal(gm.cpu()[~vis], m[~vis])


def test_selective_adam_optimizer_class():
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(100, 3, generator=g)
    p = torch.nn.Parameter(_g(p0).clone())
    opt = S.SelectiveAdam([{"params": [p], "lr": 0.05}], eps=1e-8, betas=(0.9, 0.999))
    vis = _g(torch.cat([torch.ones(50), torch.zeros(50)]).bool())
    ref_p, ref_m, ref_v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for _ in range(3):
        opt.zero_grad()
        (p ** 2).sum().backward()
        opt.step(visibility=vis)
        ref_p, ref_m, ref_v = O.selective_adam_update(ref_p, 2 * ref_p, ref_m, ref_v, vis.cpu(), 0.05, 0.9, 0.999, 1e-8)
    torch.testing.assert_close(p.detach().cpu(), ref_p, rtol=1e-5, atol=1e-6)
    assert torch.equal(p.detach().cpu()[50:], p0[50:])
    with pytest.raises(RuntimeError):
        S.selective_adam_update(p0, p0, p0, p0, vis.cpu(), 0.1, 0.9, 0.99, 1e-8, 100, 3)


def test_compute_relocation_matches_oracle():
    g = torch.Generator().manual_seed(2)
    N, n_max = 4000, 51
    binoms = torch.zeros(n_max, n_max)
    for n in range(n_max):
        for k in range(n + 1):
            binoms[n, k] = math.comb(n, k)
    op = torch.rand(N, generator=g) * 0.98 + 0.01
    sc = torch.rand(N, 3, generator=g)
    ratios = torch.randint(0, 12, (N,), generator=g)  # 0 is clamped to 1; modest n keeps fp32 well conditioned
    r_op, r_sc = O.compute_relocation(op, sc, ratios.clone(), binoms)
    rt = _g(ratios).clone()
    g_op, g_sc = S.compute_relocation(_g(op), _g(sc), rt, _g(binoms))
    assert int(rt.min()) == 1  # clamped in place like the reference (G/relocation.py:48)
    torch.testing.assert_close(g_op.cpu(), r_op, rtol=2e-4, atol=1e-6)
    torch.testing.assert_close(g_sc.cpu(), r_sc, rtol=2e-3, atol=1e-6)
    # n = 1 is the identity
    one = torch.ones(N, dtype=torch.int64)
    i_op, i_sc = S.compute_relocation(_g(op), _g(sc), _g(one), _g(binoms))
    torch.testing.assert_close(i_op.cpu(), op, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(i_sc.cpu(), sc, rtol=1e-4, atol=1e-6)
    e = S.compute_relocation(torch.zeros(0, device=DEV), torch.zeros(0, 3, device=DEV),
                             torch.zeros(0, device=DEV, dtype=torch.int64), _g(binoms))
    assert e[0].shape == (0,) and e[1].shape == (0, 3)


# ------------------------------------------------------------------------------------------------
# f3: DefaultStrategy._update_state — stand-alone kernel and the form folded into the projection backward
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["unpacked", "packed"])
def test_strategy_update_state_matches_reference_outputs(mode):
    """tests/golden/strategy_state.npz = outputs of the reference's own DefaultStrategy._update_state."""
    from types import SimpleNamespace

    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "strategy_state.npz"))
    N, C, W, H = int(G["N"]), int(G["C"]), int(G["width"]), int(G["height"])
    state = {"grad2d": None, "count": None, "radii": None}
    for call in range(2):
        grads = torch.from_numpy(G[f"{mode}_grads{call}"]).to(DEV)
        info = dict(width=W, height=H, n_cameras=C if mode == "unpacked" else 1,
                    radii=torch.from_numpy(G[f"{mode}_radii{call}"]).to(DEV),
                    gaussian_ids=torch.from_numpy(G[f"{mode}_ids{call}"]).to(DEV) if mode == "packed" else None,
                    means2d=SimpleNamespace(grad=grads, absgrad=grads.abs()))
        S.update_strategy_state(state, info, packed=mode == "packed", n_gaussians=N)
        assert torch.allclose(state["grad2d"].cpu(), torch.from_numpy(G[f"{mode}_grad2d_after{call}"]), rtol=2e-6, atol=1e-9)
        assert torch.equal(state["count"].cpu(), torch.from_numpy(G[f"{mode}_count_after{call}"]))
        assert torch.equal(state["radii"].cpu(), torch.from_numpy(G[f"{mode}_radii_after{call}"]))


@pytest.mark.parametrize("n_cams", [1, 3])
def test_strategy_state_folded_into_projection_backward(n_cams):
    """`strategy_state_sink` around backward() == the reference sequence: retain means2d.grad, then
    _update_state (restated in oracle/strategy_ref.py, itself pinned to the reference's outputs), and
    == the stand-alone kernel; the parameter gradients are untouched by the sink."""
    from oracle import strategy_ref as SRf

    from splat_one_b200 import synthetic

    W_, H_, N_ = 320, 240, 30000
    sc = synthetic.pinhole_scene(N_, W_, H_, seed=5, n_cameras=n_cams)
    names = ("means", "quats", "scales", "opacities", "sh")
    g = torch.Generator().manual_seed(8)
    vc = torch.randn(n_cams, H_, W_, 3, generator=g).to(DEV)
    va = torch.randn(n_cams, H_, W_, 1, generator=g).to(DEV)

    def run(sink_state):
        P = [sc[k].to(DEV).requires_grad_() for k in names]
        rc, ra, meta = S.rasterization(*P, sc["viewmats"].to(DEV), sc["Ks"].to(DEV), W_, H_, sh_degree=3, packed=False)
        meta["means2d"].retain_grad()
        if sink_state is not None:
            with S.strategy_state_sink(sink_state, N_, W_, H_, DEV) as sink:
                torch.autograd.backward([rc, ra], [vc, va])
            assert sink.updates == 1
        else:
            torch.autograd.backward([rc, ra], [vc, va])
        return P, meta

    fused = {"grad2d": None, "count": None, "radii": None}
    for _ in range(2):  # two steps: accumulation and the running maximum
        P1, meta1 = run(fused)
    ref = [torch.zeros(N_), torch.zeros(N_), torch.zeros(N_)]
    alone = {"grad2d": None, "count": None, "radii": None}
    for _ in range(2):
        P0, meta0 = run(None)
        SRf.update_state(*ref, meta0["means2d"].grad.cpu(), meta0["radii"].cpu(), W_, H_, n_cams)
        S.update_strategy_state(alone, meta0)
    for a, b, n in zip(P0, P1, names):  # same kernels; the raster backward's atomics reorder fp32 sums run to run
        assert_grad_close(b.grad, a.grad, what=f"{n} with the state sink", frac_ok=1.0)
    for k, r in zip(("grad2d", "count", "radii"), ref):
        assert torch.allclose(fused[k].cpu(), r, rtol=1e-5, atol=1e-7), k
        assert torch.allclose(alone[k].cpu(), r, rtol=1e-5, atol=1e-7), k
    assert int((fused["count"] > 0).sum()) > 1000
