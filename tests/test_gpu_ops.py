"""GPU parity of every operator of the hot path against the CPU oracle, called through
the public Python API (which goes through the C ABI of libb200splat.so)."""
import math

import pytest
import torch

import splat_one_b200 as S
from oracle import raster_ref as RC
from oracle import torch_ref as O
from parity import assert_grad_close, assert_image_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cams(C):
    vm = torch.eye(4).expand(C, -1, -1).contiguous().clone()
    for c in range(1, C):
        a = 0.1 * c
        vm[c, :3, :3] = torch.tensor([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]])
        vm[c, :3, 3] = torch.tensor([0.1 * c, -0.05 * c, 0.2])
    return vm


def _scene(N=3000, seed=0, spherical=False):
    g = torch.Generator().manual_seed(seed)
    if spherical:
        d = torch.randn(N, 3, generator=g)
        means = d / d.norm(dim=-1, keepdim=True) * (torch.rand(N, 1, generator=g) * 4 + 1)
    else:
        means = torch.rand(N, 3, generator=g) * 2 - 1
        means[:, 2] = means[:, 2] * 2 + 3.5
    quats = torch.randn(N, 4, generator=g)
    scales = torch.rand(N, 3, generator=g) * 0.1 + 0.005
    return means, quats, scales


# ------------------------------------------------------------------------------------
# a2 / a3 / a4
# ------------------------------------------------------------------------------------
@pytest.mark.parametrize("cm", ["pinhole", "ortho", "fisheye", "spherical"])
@pytest.mark.parametrize("use_covars", [False, True])
@pytest.mark.parametrize("comp", [False, True])
def test_projection_unpacked(cm, use_covars, comp):
    C, W, H = 3, 320, 200
    means, quats, scales = _scene(spherical=(cm == "spherical"))
    vm = _cams(C)
    Ks = torch.tensor([[260.0, 0, 160], [0, 250.0, 100], [0, 0, 1]]).expand(C, -1, -1).contiguous()
    if cm == "ortho":
        Ks = Ks.clone()
        Ks[:, 0, 0] = Ks[:, 1, 1] = 80.0
    covars = None
    if use_covars:
        c3 = O.quat_scale_to_covar(quats, scales)
        covars = c3[:, [0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2]].contiguous()
    leaves_c = [means.clone().requires_grad_(), vm.clone().requires_grad_()]
    leaves_g = [means.to(DEV).requires_grad_(), vm.to(DEV).requires_grad_()]
    if use_covars:
        leaves_c += [covars.clone().requires_grad_()]
        leaves_g += [covars.to(DEV).requires_grad_()]
        args_c = (leaves_c[0], leaves_c[2], None, None, leaves_c[1])
        args_g = (leaves_g[0], leaves_g[2], None, None, leaves_g[1])
    else:
        leaves_c += [quats.clone().requires_grad_(), scales.clone().requires_grad_()]
        leaves_g += [quats.to(DEV).requires_grad_(), scales.to(DEV).requires_grad_()]
        args_c = (leaves_c[0], None, leaves_c[2], leaves_c[3], leaves_c[1])
        args_g = (leaves_g[0], None, leaves_g[2], leaves_g[3], leaves_g[1])
    ref = O.fully_fused_projection(*args_c, Ks, W, H, calc_compensations=comp, camera_model=cm, radius_clip=1.0)
    got = S.fully_fused_projection(*args_g, Ks.to(DEV), W, H, calc_compensations=comp, camera_model=cm,
                                   radius_clip=1.0)
    r_ref, r_got = ref[0], got[0].cpu()
    # radii may differ by 1 where ceil() sits on a rounding boundary (tests/test_basic.py:242)
    assert (r_ref - r_got).abs().max() <= 1
    # compare where neither side culled (a ±1 radius can flip the radius_clip / bounds cull)
    sel = (r_ref > 0) & (r_got > 0)
    assert ((r_ref > 0) != (r_got > 0)).float().mean() < 2e-3
    assert sel.sum() > 500
    torch.testing.assert_close(got[1].cpu()[sel], ref[1][sel], rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(got[2].cpu()[sel], ref[2][sel], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(got[3].cpu()[sel], ref[3][sel], rtol=2e-3, atol=1e-5)
    if comp:
        torch.testing.assert_close(got[4].cpu()[sel], ref[4][sel], rtol=1e-3, atol=1e-5)
    else:
        assert got[4] is None
    # backward with fixed cotangents on the commonly visible set
    g = torch.Generator().manual_seed(5)
    v_m2, v_d, v_c = torch.randn(C, means.shape[0], 2, generator=g), torch.randn(C, means.shape[0], generator=g), \
        torch.randn(C, means.shape[0], 3, generator=g) * 1e-2
    v_cp = torch.randn(C, means.shape[0], generator=g)
    selg = sel.to(DEV)

    def loss(o, s, dev):
        l = (o[1] * v_m2.to(dev))[s].sum() + (o[2] * v_d.to(dev))[s].sum() + (o[3] * v_c.to(dev))[s].sum()
        if comp:
            l = l + (o[4] * v_cp.to(dev))[s].sum()
        return l

    g_ref = torch.autograd.grad(loss(ref, sel, "cpu"), leaves_c)
    g_got = torch.autograd.grad(loss(got, selg, DEV), leaves_g)
    names = ["means", "viewmats"] + (["covars"] if use_covars else ["quats", "scales"])
    for n, a, b in zip(names, g_got, g_ref):
        assert_grad_close(a, b, rtol=2e-3, what=f"{cm}/{n}", frac_ok=0.995)


@pytest.mark.parametrize("cm", ["pinhole", "spherical"])
@pytest.mark.parametrize("sparse_grad", [False, True])
def test_projection_packed(cm, sparse_grad):
    C, W, H = 2, 320, 200
    means, quats, scales = _scene(spherical=(cm == "spherical"))
    vm = _cams(C)
    Ks = torch.tensor([[260.0, 0, 160], [0, 250.0, 100], [0, 0, 1]]).expand(C, -1, -1).contiguous()
    P_c = [t.clone().requires_grad_() for t in (means, quats, scales)]
    P_g = [t.to(DEV).requires_grad_() for t in (means, quats, scales)]
    dense = O.fully_fused_projection(P_c[0], None, P_c[1], P_c[2], vm, Ks, W, H, camera_model=cm, packed_rules=True,
                                     calc_compensations=True)
    cam_r, gau_r, rad_r, m2_r, dep_r, con_r, comp_r = O.pack_projection(*dense)
    cam, gau, rad, m2, dep, con, comp = S.fully_fused_projection(
        P_g[0], None, P_g[1], P_g[2], vm.to(DEV), Ks.to(DEV), W, H, camera_model=cm, packed=True,
        sparse_grad=sparse_grad, calc_compensations=True)
    assert cam.dtype == torch.int64 and gau.dtype == torch.int64 and rad.dtype == torch.int32
    # identical visible sets up to ±1-radius boundary cases: compare on the intersection of (cam, gauss) pairs
    key_r = (cam_r * means.shape[0] + gau_r)
    key_g = (cam.cpu() * means.shape[0] + gau.cpu())
    assert torch.equal(key_g, torch.sort(key_g).values), "COO rows must be (camera, gaussian) row-major"
    common = torch.isin(key_g, key_r)
    common_r = torch.isin(key_r, key_g)
    assert common.float().mean() > 0.998 and common_r.float().mean() > 0.998
    assert (rad.cpu()[common] - rad_r[common_r]).abs().max() <= 1
    torch.testing.assert_close(m2.cpu()[common], m2_r[common_r], rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(dep.cpu()[common], dep_r[common_r], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(con.cpu()[common], con_r[common_r], rtol=2e-3, atol=1e-5)
    torch.testing.assert_close(comp.cpu()[common], comp_r[common_r], rtol=1e-3, atol=1e-5)
    g = torch.Generator().manual_seed(5)
    n_r = key_r.numel()
    v_m2, v_d, v_c = torch.randn(n_r, 2, generator=g), torch.randn(n_r, generator=g), torch.randn(n_r, 3, generator=g) * 1e-2
    # scatter the cotangents onto the GPU row order through the common keys
    pos_in_r = torch.searchsorted(key_r, key_g[common])
    vg_m2 = torch.zeros(key_g.numel(), 2); vg_d = torch.zeros(key_g.numel()); vg_c = torch.zeros(key_g.numel(), 3)
    vg_m2[common], vg_d[common], vg_c[common] = v_m2[pos_in_r], v_d[pos_in_r], v_c[pos_in_r]
    mask_r = common_r.float()
    l_ref = (m2_r * v_m2 * mask_r[:, None]).sum() + (dep_r * v_d * mask_r).sum() + (con_r * v_c * mask_r[:, None]).sum()
    l_got = (m2 * vg_m2.to(DEV)).sum() + (dep * vg_d.to(DEV)).sum() + (con * vg_c.to(DEV)).sum()
    g_ref = torch.autograd.grad(l_ref, P_c)
    g_got = torch.autograd.grad(l_got, P_g)
    for n, a, b in zip(["means", "quats", "scales"], g_got, g_ref):
        if sparse_grad:
            assert a.is_sparse and a.layout == torch.sparse_coo
            a = a.to_dense()
        assert_grad_close(a, b, rtol=2e-3, what=f"packed/{cm}/{n}", frac_ok=0.995)


def test_projection_empty_and_all_culled():
    vm, Ks = torch.eye(4, device=DEV)[None], torch.eye(3, device=DEV)[None]
    out = S.fully_fused_projection(torch.zeros(0, 3, device=DEV), None, torch.zeros(0, 4, device=DEV),
                                   torch.zeros(0, 3, device=DEV), vm, Ks, 16, 16)
    assert out[0].shape == (1, 0)
    behind = torch.tensor([[0.0, 0.0, -5.0]] * 7, device=DEV)
    q, s = torch.ones(7, 4, device=DEV), torch.ones(7, 3, device=DEV) * 0.1
    out = S.fully_fused_projection(behind, None, q, s, vm, Ks, 16, 16)
    assert out[0].abs().sum() == 0
    packed = S.fully_fused_projection(behind, None, q, s, vm, Ks, 16, 16, packed=True)
    assert all(t.shape[0] == 0 for t in packed[:6])


# ------------------------------------------------------------------------------------
# a5
# ------------------------------------------------------------------------------------
@pytest.mark.parametrize("deg", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("K", [25, 16])
def test_sh_fwd_bwd(deg, K, golden):
    if (deg + 1) ** 2 > K:
        pytest.skip("K too small")
    g = golden("sh_ref.npz")
    coeffs, dirs, v_colors = g["coeffs"][:, :K].contiguous(), g["dirs"], g["v_colors"]
    c = coeffs.to(DEV).requires_grad_()
    d = dirs.to(DEV).requires_grad_()
    col = S.spherical_harmonics(deg, d, c)
    torch.testing.assert_close(col.cpu(), g[f"ref_colors_{deg}"], rtol=1e-4, atol=1e-4)  # tests/test_basic.py:590
    gc, gd = torch.autograd.grad((col * v_colors.to(DEV)).sum(), (c, d), allow_unused=True)
    torch.testing.assert_close(gc.cpu(), g[f"ref_v_coeffs_{deg}"][:, :K], rtol=1e-4, atol=1e-4)
    if deg > 0:
        torch.testing.assert_close(gd.cpu(), g[f"ref_v_dirs_{deg}"], rtol=1e-4, atol=1e-4)


def test_sh_masks_batch_dims_and_table():
    torch.manual_seed(0)
    C, N, K = 3, 500, 16
    dirs = torch.randn(C, N, 3)
    table = torch.randn(N, K, 3)
    masks = torch.rand(C, N) > 0.3
    v = torch.randn(C, N, 3)
    tr = table.clone().requires_grad_()
    dr = dirs.clone().requires_grad_()
    ref = O.spherical_harmonics(3, dr, tr.expand(C, -1, -1, -1), masks)
    g_ref = torch.autograd.grad((ref * v).sum(), (tr, dr))
    for mode in ("table", "expand", "materialised"):
        tg = table.to(DEV).requires_grad_()
        dg = dirs.to(DEV).requires_grad_()
        if mode == "table":
            got = S.spherical_harmonics_table(3, dg, tg, masks.to(DEV))
        elif mode == "expand":
            got = S.spherical_harmonics(3, dg, tg.expand(C, -1, -1, -1), masks.to(DEV))
        else:
            got = S.spherical_harmonics(3, dg, tg[None].repeat(C, 1, 1, 1), masks.to(DEV))
        torch.testing.assert_close(got.cpu(), ref, rtol=1e-4, atol=1e-4)
        assert got.cpu()[~masks].abs().max() == 0
        g_got = torch.autograd.grad((got * v.to(DEV)).sum(), (tg, dg))
        torch.testing.assert_close(g_got[0].cpu(), g_ref[0], rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(g_got[1].cpu(), g_ref[1], rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------------------------
# a6 / a7: bit-exact
# ------------------------------------------------------------------------------------
def test_isect_reference_test_shapes_exact(golden):
    g = golden("isect_ref.npz")
    tpg, ids, fl = S.isect_tiles(g["means2d"].to(DEV), g["radii"].to(DEV), g["depths"].to(DEV), g["tile_size"],
                                 g["tile_width"], g["tile_height"])
    assert tpg.dtype == torch.int32 and ids.dtype == torch.int64 and fl.dtype == torch.int32
    assert torch.equal(tpg.cpu(), g["ref_tiles_per_gauss"])
    assert torch.equal(ids.cpu(), g["ref_isect_ids"])
    assert torch.equal(fl.cpu(), g["ref_flatten_ids"])
    offs = S.isect_offset_encode(ids, 3, g["tile_width"], g["tile_height"])
    assert torch.equal(offs.cpu(), g["ref_offsets"])
    # unsorted variant
    tpg2, ids2, fl2 = S.isect_tiles(g["means2d"].to(DEV), g["radii"].to(DEV), g["depths"].to(DEV), g["tile_size"],
                                    g["tile_width"], g["tile_height"], sort=False)
    o = O.isect_tiles(g["means2d"], g["radii"], g["depths"], g["tile_size"], g["tile_width"], g["tile_height"],
                      sort=False)
    assert torch.equal(ids2.cpu(), o[1]) and torch.equal(fl2.cpu(), o[2])


@pytest.mark.parametrize("C,N,W,H,ts", [(1, 20000, 640, 360, 16), (4, 5000, 200, 120, 16), (2, 3000, 97, 61, 8),
                                        (1, 50000, 1920, 1080, 16), (9, 700, 64, 64, 4)])
def test_isect_random_exact(C, N, W, H, ts):
    g = torch.Generator().manual_seed(C * 1000 + N)
    m2 = torch.rand(C, N, 2, generator=g) * torch.tensor([W * 1.2, H * 1.2]) - torch.tensor([W * 0.1, H * 0.1])
    radii = torch.randint(0, 40, (C, N), generator=g, dtype=torch.int32)
    depths = torch.rand(C, N, generator=g) * 10
    depths[:, : N // 10] = 1.5  # many exact depth ties: stability of the sort
    tw, th = math.ceil(W / ts), math.ceil(H / ts)
    ref = O.isect_tiles(m2, radii, depths, ts, tw, th)
    got = S.isect_tiles(m2.to(DEV), radii.to(DEV), depths.to(DEV), ts, tw, th)
    for a, b in zip(got, ref):
        assert torch.equal(a.cpu(), b)
    ref_offs = O.isect_offset_encode(ref[1], C, tw, th)
    assert torch.equal(S.isect_offset_encode(got[1], C, tw, th).cpu(), ref_offs)
    # fused variant used by rasterization(): same four tensors from one call
    from splat_one_b200.wrapper import isect_tiles_and_offsets
    fused = isect_tiles_and_offsets(m2.to(DEV), radii.to(DEV), depths.to(DEV), ts, tw, th)
    for a, b in zip(fused, list(ref) + [ref_offs]):
        assert torch.equal(a.cpu(), b)
    # wide rectangles: radii up to the whole image (long per-Gaussian tile runs)
    radii2 = torch.randint(0, max(W, H), (C, N), generator=g, dtype=torch.int32)
    radii2[:, N // 4:] = 0
    ref2 = O.isect_tiles(m2, radii2, depths, ts, tw, th)
    fused2 = isect_tiles_and_offsets(m2.to(DEV), radii2.to(DEV), depths.to(DEV), ts, tw, th)
    for a, b in zip(fused2, list(ref2) + [O.isect_offset_encode(ref2[1], C, tw, th)]):
        assert torch.equal(a.cpu(), b)


def test_isect_packed_and_edges():
    g = torch.Generator().manual_seed(3)
    C, nnz, W, H, ts = 3, 4000, 300, 200, 16
    cam = torch.sort(torch.randint(0, C, (nnz,), generator=g)).values
    gau = torch.randint(0, 10000, (nnz,), generator=g)
    m2 = torch.rand(nnz, 2, generator=g) * torch.tensor([float(W), float(H)])
    radii = torch.randint(1, 30, (nnz,), generator=g, dtype=torch.int32)
    depths = torch.rand(nnz, generator=g) * 5
    tw, th = math.ceil(W / ts), math.ceil(H / ts)
    ref = O.isect_tiles(m2, radii, depths, ts, tw, th, packed=True, n_cameras=C, camera_ids=cam, gaussian_ids=gau)
    got = S.isect_tiles(m2.to(DEV), radii.to(DEV), depths.to(DEV), ts, tw, th, packed=True, n_cameras=C,
                        camera_ids=cam.to(DEV), gaussian_ids=gau.to(DEV))
    for a, b in zip(got, ref):
        assert torch.equal(a.cpu(), b)
    assert torch.equal(S.isect_offset_encode(got[1], C, tw, th).cpu(), O.isect_offset_encode(ref[1], C, tw, th))
    from splat_one_b200.wrapper import isect_tiles_and_offsets
    fused = isect_tiles_and_offsets(m2.to(DEV), radii.to(DEV), depths.to(DEV), ts, tw, th, packed=True, n_cameras=C,
                                    camera_ids=cam.to(DEV), gaussian_ids=gau.to(DEV))
    assert torch.equal(fused[3].cpu(), O.isect_offset_encode(ref[1], C, tw, th))
    zf = isect_tiles_and_offsets(torch.zeros(2, 9, 2, device=DEV), torch.zeros(2, 9, dtype=torch.int32, device=DEV),
                                 torch.ones(2, 9, device=DEV), 16, 2, 2)
    assert zf[3].shape == (2, 2, 2) and zf[3].abs().sum() == 0
    # empty inputs and all-invisible inputs
    e = S.isect_tiles(torch.zeros(2, 0, 2, device=DEV), torch.zeros(2, 0, dtype=torch.int32, device=DEV),
                      torch.zeros(2, 0, device=DEV), 16, 2, 2)
    assert e[0].shape == (2, 0) and e[1].numel() == 0 and e[2].numel() == 0
    z = S.isect_tiles(torch.zeros(2, 9, 2, device=DEV), torch.zeros(2, 9, dtype=torch.int32, device=DEV),
                      torch.ones(2, 9, device=DEV), 16, 2, 2)
    assert z[0].sum() == 0 and z[1].numel() == 0
    offs = S.isect_offset_encode(z[1], 2, 2, 2)
    assert offs.shape == (2, 2, 2) and offs.abs().sum() == 0


# ------------------------------------------------------------------------------------
# a8 / a9
# ------------------------------------------------------------------------------------
def _raster_inputs(C=2, N=4000, W=200, H=136, ts=16, D=3, seed=0, packed=False):
    g = torch.Generator().manual_seed(seed)
    means, quats, scales = _scene(N, seed)
    scales = scales * 1.5
    vm = _cams(C)
    Ks = torch.tensor([[180.0, 0, W / 2], [0, 180.0, H / 2], [0, 0, 1]]).expand(C, -1, -1).contiguous()
    radii, m2, dep, con, _ = O.fully_fused_projection(means, None, quats, scales, vm, Ks, W, H, packed_rules=packed)
    op = torch.rand(C, N, generator=g)
    col = torch.randn(C, N, D, generator=g)
    cam = gau = None
    if packed:
        cam, gau, radii, m2, dep, con, _ = O.pack_projection(radii, m2, dep, con, None)
        op, col = op[cam, gau], col[cam, gau]
    tw, th = math.ceil(W / ts), math.ceil(H / ts)
    tpg, ids, fl = O.isect_tiles(m2, radii, dep, ts, tw, th, packed=packed, n_cameras=C, camera_ids=cam,
                                 gaussian_ids=gau)
    offs = O.isect_offset_encode(ids, C, tw, th)
    return dict(m2=m2.contiguous(), con=con.contiguous(), col=col.contiguous(), op=op.contiguous(), W=W, H=H, ts=ts,
                offs=offs, fl=fl, C=C)


@pytest.mark.parametrize("D", [1, 2, 3, 4, 5, 7, 8, 17, 32, 40])
@pytest.mark.parametrize("bg", [False, True])
def test_rasterize_fwd_bwd_channels(D, bg):
    x = _raster_inputs(D=D, seed=D)
    _check_raster(x, bg=bg, absgrad=(D == 3))


@pytest.mark.parametrize("W,H,ts", [(200, 136, 16), (100, 70, 8), (37, 29, 4), (65, 33, 16), (48, 48, 32)])
def test_rasterize_tile_sizes_and_ragged_images(W, H, ts):
    x = _raster_inputs(C=1, N=2500, W=W, H=H, ts=ts, D=3, seed=W)
    _check_raster(x, bg=True, absgrad=False)


def test_rasterize_packed_and_masks():
    x = _raster_inputs(packed=True, seed=11)
    _check_raster(x, bg=True, absgrad=True, packed=True)
    x = _raster_inputs(seed=12)
    g = torch.Generator().manual_seed(1)
    masks = torch.rand(x["offs"].shape, generator=g) > 0.3
    _check_raster(x, bg=True, absgrad=False, masks=masks)
    _check_raster(x, bg=False, absgrad=False, masks=masks)


def test_rasterize_no_intersections():
    C, N, W, H = 1, 8, 40, 24
    offs = torch.zeros(C, 2, 3, dtype=torch.int32, device=DEV)
    bgc = torch.rand(C, 3, device=DEV)
    rc, ra = S.rasterize_to_pixels(torch.zeros(C, N, 2, device=DEV), torch.ones(C, N, 3, device=DEV),
                                   torch.ones(C, N, 3, device=DEV), torch.ones(C, N, device=DEV), W, H, 16, offs,
                                   torch.zeros(0, dtype=torch.int32, device=DEV), backgrounds=bgc)
    assert ra.abs().max() == 0
    torch.testing.assert_close(rc, bgc[:, None, None, :].expand(C, H, W, 3))


def _check_raster(x, bg, absgrad, packed=False, masks=None):
    D = x["col"].shape[-1]
    g = torch.Generator().manual_seed(99)
    bgc = torch.rand(x["C"], D, generator=g) if bg else None
    leaves_c = [x[k].clone().requires_grad_() for k in ("m2", "con", "col", "op")]
    leaves_g = [x[k].to(DEV).requires_grad_() for k in ("m2", "con", "col", "op")]
    bg_c = bgc.clone().requires_grad_() if bg else None
    bg_g = bgc.to(DEV).requires_grad_() if bg else None
    rc_ref, ra_ref, last_ref, margin = RC.raster_fwd(*leaves_c, x["W"], x["H"], x["ts"], x["offs"], x["fl"], bg_c,
                                                     masks, want_margin=True)
    rc, ra = S.rasterize_to_pixels(*leaves_g, x["W"], x["H"], x["ts"], x["offs"].to(DEV), x["fl"].to(DEV),
                                   backgrounds=bg_g, masks=None if masks is None else masks.to(DEV), packed=packed,
                                   absgrad=absgrad)
    assert rc.shape == (x["C"], x["H"], x["W"], D) and ra.shape == (x["C"], x["H"], x["W"], 1)
    assert_image_close(rc, rc_ref, margin, what="render_colors", rtol=1e-4)
    assert_image_close(ra, ra_ref, margin[..., None], what="render_alphas")
    vc = torch.randn(rc_ref.shape, generator=g)
    va = torch.randn(ra_ref.shape, generator=g)
    # pixels whose skip/stop decisions are ambiguous get no cotangent, so a flipped
    # decision cannot leak into the gradient comparison
    keep = (margin >= 1e-3).float()[..., None]
    vc, va = vc * keep, va * keep
    inputs_g = leaves_g + ([bg_g] if bg else [])
    grads = torch.autograd.grad((rc * vc.to(DEV)).sum() + (ra * va.to(DEV)).sum(), inputs_g)
    ref = RC.raster_bwd(*[t.detach() for t in leaves_c], x["W"], x["H"], x["ts"], x["offs"], x["fl"], ra_ref,
                        last_ref, vc, va, bgc, masks, absgrad=absgrad)
    names = ["v_means2d", "v_conics", "v_colors", "v_opacities"]
    for n, a, b in zip(names, grads[:4], ref[1:]):
        assert_grad_close(a, b, what=n)
    if absgrad:
        assert_grad_close(leaves_g[0].absgrad, ref[0], what="absgrad")
    if bg:
        v_bg_ref = (vc * (1.0 - ra_ref)).sum(dim=(1, 2))
        assert_grad_close(grads[4], v_bg_ref, what="v_backgrounds")


def test_isect_both_sort_paths_and_negative_depths(monkeypatch):
    """The depth-first ordering and the generic full-key sort give identical bits; depths
    with the sign bit set (outside the contract, CS/isect_tiles.cu:92 sign-extends them)
    are routed to the generic sort and still match the reference semantics."""
    from splat_one_b200 import wrapper

    g = torch.Generator().manual_seed(17)
    C, N, W, H, ts = 2, 20000, 640, 360, 16
    m2 = torch.rand(C, N, 2, generator=g) * torch.tensor([float(W), float(H)])
    radii = torch.randint(0, 30, (C, N), generator=g, dtype=torch.int32)
    depths = torch.rand(C, N, generator=g) * 10
    tw, th = math.ceil(W / ts), math.ceil(H / ts)
    fast = S.isect_tiles(m2.to(DEV), radii.to(DEV), depths.to(DEV), ts, tw, th)
    monkeypatch.setattr(wrapper, "_FORCE_GENERIC_SORT", True)
    slow = S.isect_tiles(m2.to(DEV), radii.to(DEV), depths.to(DEV), ts, tw, th)
    monkeypatch.setattr(wrapper, "_FORCE_GENERIC_SORT", False)
    for a, b in zip(fast, slow):
        assert torch.equal(a, b)
    depths[0, :50] = -depths[0, :50]
    ref = O.isect_tiles(m2, radii, depths, ts, tw, th)
    got = S.isect_tiles(m2.to(DEV), radii.to(DEV), depths.to(DEV), ts, tw, th)
    for a, b in zip(got, ref):
        assert torch.equal(a.cpu(), b)


def test_raster_fast_and_generic_kernels_agree(monkeypatch):
    """tile_size 16 / <= 4 channels runs the warp-per-tile kernels; they must agree with the
    generic kernels (and both with the oracle, tested above)."""
    from splat_one_b200 import wrapper

    x = _raster_inputs(D=3, seed=21)
    args = [x[k].to(DEV) for k in ("m2", "con", "col", "op")]
    outs = []
    for generic in (False, True):
        monkeypatch.setattr(wrapper, "_FORCE_GENERIC_RASTER", generic)
        leaves = [a.clone().requires_grad_() for a in args]
        rc, ra = S.rasterize_to_pixels(*leaves, x["W"], x["H"], x["ts"], x["offs"].to(DEV), x["fl"].to(DEV))
        gs = torch.autograd.grad(rc.sum() + 2 * ra.sum(), leaves)
        outs.append((rc, ra, gs))
    monkeypatch.setattr(wrapper, "_FORCE_GENERIC_RASTER", False)
    (rc0, ra0, g0), (rc1, ra1, g1) = outs
    assert (rc0 - rc1).abs().max() < 5e-3 and ((rc0 - rc1).abs() > 1e-4).float().mean() < 1e-3
    assert ((ra0 - ra1).abs() > 1e-4).float().mean() < 1e-3
    for a, b in zip(g0, g1):
        assert_grad_close(a, b, rtol=2e-3, frac_ok=0.995)


@pytest.mark.parametrize("per_view", [False, True])
@pytest.mark.parametrize("deg", [0, 2, 3])
def test_sh_view_colors_fused(per_view, deg):
    """Fused dirs + SH + 0.5 + clamp (+ its backward) == the unfused reference chain."""
    from splat_one_b200.wrapper import camera_centers, sh_view_colors

    torch.manual_seed(deg)
    C, N, K = 3, 2000, 16
    means = torch.randn(N, 3)
    vm = _cams(C)
    vm[2, :3, :3] = vm[2, :3, :3] * 1.3  # non-rigid view matrix: the centre needs a general inverse
    table = torch.randn(C, N, K, 3) * 0.4 if per_view else torch.randn(N, K, 3) * 0.4
    radii = (torch.rand(C, N) > 0.25).int() * 5
    v = torch.randn(C, N, 3)
    torch.testing.assert_close(camera_centers(vm.to(DEV)).cpu(), torch.inverse(vm)[:, :3, 3], rtol=1e-5, atol=1e-5)
    m_c, t_c = means.clone().requires_grad_(), table.clone().requires_grad_()
    dirs = m_c[None] - torch.inverse(vm)[:, None, :3, 3]
    shs = t_c if per_view else t_c.expand(C, -1, -1, -1)
    ref = torch.clamp_min(O.spherical_harmonics(deg, dirs, shs, radii > 0) + 0.5, 0.0)
    ref = torch.where((radii > 0)[..., None], ref, torch.zeros_like(ref))
    g_ref = list(torch.autograd.grad((ref * v).sum(), (m_c, t_c), allow_unused=True))
    if g_ref[0] is None:  # degree 0 does not depend on the view direction
        g_ref[0] = torch.zeros_like(means)
    m_g, t_g = means.to(DEV).requires_grad_(), table.to(DEV).requires_grad_()
    got = sh_view_colors(deg, m_g, vm.to(DEV), t_g, radii.to(DEV))
    torch.testing.assert_close(got.cpu(), ref, rtol=1e-4, atol=1e-4)
    g_got = torch.autograd.grad((got * v.to(DEV)).sum(), (m_g, t_g))
    torch.testing.assert_close(g_got[0].cpu(), g_ref[0], rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(g_got[1].cpu(), g_ref[1], rtol=1e-4, atol=1e-4)


def test_sh_colors_bwd_premasked_camera_exchange():
    """The camera-parallel exchange form of the fused colour backward: pre-masked colour
    cotangents of ALL cameras in, coefficient gradient summed over all of them out, means
    gradient only from the local camera range (splat_one_b200/distributed.py)."""
    from splat_one_b200 import wrapper
    from splat_one_b200._lib import get_lib
    from splat_one_b200.wrapper import _ptr, camera_centers, native, sh_view_colors

    torch.manual_seed(0)
    C, N, K, deg = 3, 3000, 16, 3
    means = torch.randn(N, 3, device=DEV)
    vm = _cams(C).to(DEV)
    table = (torch.randn(N, K, 3) * 0.4).to(DEV)
    radii = ((torch.rand(C, N) > 0.25).int() * 5).to(DEV)
    v = torch.randn(C, N, 3, device=DEV)
    # reference: the ordinary fused path over all C cameras (coefficients) and over camera 1 only (means)
    m_all, t_all = means.clone().requires_grad_(), table.clone().requires_grad_()
    col = sh_view_colors(deg, m_all, vm, t_all, radii)
    g_all = torch.autograd.grad((col * v).sum(), (m_all, t_all))
    m_1 = means.clone().requires_grad_()
    col1 = sh_view_colors(deg, m_1, vm[1:2], table, radii[1:2])
    g_m1 = torch.autograd.grad((col1 * v[1:2]).sum(), m_1)[0]
    # exchange form: what rank 1 of 3 would run after the all-gather
    g_masked = torch.where(col.detach() > 0, v, torch.zeros_like(v)).contiguous()
    v_coeffs = torch.empty_like(table)
    v_means = torch.empty_like(means)
    campos = camera_centers(vm)
    native("sh_colors_bwd", get_lib(), means.device, C, N, K, deg, 0, _ptr(means), _ptr(campos), _ptr(table), None, None,
           _ptr(g_masked), _ptr(v_coeffs), _ptr(v_means), 1, 2)
    torch.testing.assert_close(v_coeffs, g_all[1], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(v_means, g_m1, rtol=1e-4, atol=1e-6)
    assert wrapper._CAMERA_PARALLEL == {}


def test_odd_sized_batches_and_misaligned_means2d_views():
    """C*N odd: the projection outputs are views of one allocation and means2d is accessed as float2 by every
    consumer — it must stay 8-byte aligned; a caller-provided means2d view at an odd float offset is re-packed
    by the wrappers (the C ABI rejects it with an error instead of faulting)."""
    from splat_one_b200 import synthetic
    from splat_one_b200._lib import get_lib

    sc = synthetic.pinhole_scene(1001, 160, 120, seed=3)
    P = [sc[k].to(DEV) for k in ("means", "quats", "scales", "opacities", "sh")]
    rc, ra, meta = S.rasterization(*P, sc["viewmats"].to(DEV), sc["Ks"].to(DEV), 160, 120, sh_degree=3, packed=False)
    torch.cuda.synchronize()
    assert meta["means2d"].data_ptr() % 8 == 0 and torch.isfinite(rc).all()
    # the same intersections from a means2d view that starts at an odd float offset
    m2 = meta["means2d"].detach()
    buf = torch.empty(m2.numel() + 1, device=DEV)
    odd = buf[1:].view_as(m2)
    odd.copy_(m2)
    assert odd.data_ptr() % 8 == 4
    a = S.isect_tiles(m2, meta["radii"], meta["depths"].detach(), 16, 10, 8)
    b = S.isect_tiles(odd, meta["radii"], meta["depths"].detach(), 16, 10, 8)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    lib = get_lib()
    import ctypes

    tpg = torch.empty(1001, dtype=torch.int32, device=DEV)
    tot = torch.zeros(2, dtype=torch.int64, device=DEV)
    rc_ = lib.b200splat_isect_count(0, 1, 1001, 0, ctypes.c_void_p(odd.data_ptr()), ctypes.c_void_p(meta["radii"].data_ptr()),
                                    ctypes.c_void_p(meta["depths"].data_ptr()), 16, 10, 8, ctypes.c_void_p(tpg.data_ptr()), None,
                                    ctypes.c_void_p(tot.data_ptr()), None, 0, None)
    assert rc_ != 0 and b"8-byte aligned" in lib.b200splat_last_error()
