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


@pytest.mark.parametrize("packed", [False, True])
def test_rasterization_antialiased_per_view_colors_absgrad(packed):
    scene = synthetic.pinhole_scene(4000, 160, 128, seed=9, n_cameras=3)
    ref, got, P_c, P_g = _run_both(scene, packed, True, "RGB", bg=False, rasterize_mode="antialiased", absgrad=True,
                                   per_view_color=True)
    (rc_r, ra_r, m_r), (rc, ra, m) = ref, got
    margin = _margin_of(m_r, scene)
    assert_image_close(rc, rc_r, margin, what="colors")
    g = torch.Generator().manual_seed(4)
    keep = (margin >= 1e-3).float()[..., None]
    vc = torch.randn(rc_r.shape, generator=g) * keep
    g_ref = torch.autograd.grad((rc_r * vc).sum(), P_c)
    m["means2d"].retain_grad()
    (rc * vc.to(DEV)).sum().backward()
    for n, a, b in zip(["means", "quats", "scales", "opacities", "colors"], [p.grad for p in P_g], g_ref):
        assert_grad_close(a, b, rtol=2e-3, what=f"per-view/{n}", frac_ok=0.995)
    shape = (m["gaussian_ids"].numel(), 2) if packed else (3, 4000, 2)
    assert m["means2d"].grad is not None and m["means2d"].grad.shape == shape
    assert m["means2d"].absgrad.shape == shape
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


def _full_size_properties(scene, W, H, camera_model="pinhole", packed=False, sparse_grad=False, min_isects=1_000_000):
    """Size-independent properties of one full-size rasterization() + backward()."""
    scene = synthetic.to_device(scene, DEV)
    P = [scene[k].clone().requires_grad_() for k in ("means", "quats", "scales", "opacities", "sh")]
    N = P[0].shape[0]
    rc, ra, m = S.rasterization(*P, scene["viewmats"], scene["Ks"], W, H, sh_degree=3, packed=packed,
                                sparse_grad=sparse_grad, camera_model=camera_model)
    assert rc.shape == (1, H, W, 3) and ra.shape == (1, H, W, 1)
    assert torch.isfinite(rc).all() and torch.isfinite(ra).all()
    assert (ra >= 0).all() and (ra <= 1.0).all() and (rc >= -1e-6).all()
    ids, fl, offs = m["isect_ids"], m["flatten_ids"], m["isect_offsets"].flatten().long()
    n_isects = ids.numel()
    assert n_isects == int(m["tiles_per_gauss"].sum()) and n_isects > min_isects
    assert (ids[1:] >= ids[:-1]).all(), "isect_ids must be sorted"
    # stable: equal keys keep ascending flat index
    eq = ids[1:] == ids[:-1]
    assert (fl[1:][eq] > fl[:-1][eq]).all()
    assert (offs[1:] >= offs[:-1]).all() and offs[0] == 0 and offs[-1] <= n_isects
    # the fused offsets equal the stand-alone operator's, and really are the first index of each tile
    assert torch.equal(m["isect_offsets"], S.isect_offset_encode(ids, 1, m["tile_width"], m["tile_height"]))
    tile_of = (ids >> 32)
    k = torch.randint(0, offs.numel(), (2000,), device=DEV)
    start = offs[k]
    end = torch.where(k + 1 < offs.numel(), offs[(k + 1).clamp(max=offs.numel() - 1)], torch.full_like(start, n_isects))
    nz = end > start
    assert (tile_of[start[nz]] == k[nz]).all() and (tile_of[end[nz] - 1] == k[nz]).all()
    # depth bits of every key equal the depth of the Gaussian row it points to
    assert torch.equal((ids & 0xFFFFFFFF).int(), m["depths"].flatten()[fl.long()].view(torch.int32))
    if packed:
        gi, ci = m["gaussian_ids"], m["camera_ids"]
        assert gi.dtype == torch.int64 and (gi[1:] > gi[:-1]).all() and (ci == 0).all()  # (camera, gaussian) order
        assert (m["radii"] > 0).all()
    # linearity in the colours for fixed geometry: render(0.5) = 0.5 * alpha
    half = torch.full(m["means2d"].shape[:-1] + (3,), 0.5, device=DEV)
    rc2, ra2 = S.rasterize_to_pixels(m["means2d"].detach(), m["conics"].detach(), half, m["opacities"].detach(),
                                     W, H, 16, m["isect_offsets"], fl, packed=packed)
    torch.testing.assert_close(ra2, ra.detach(), rtol=0, atol=0)  # same kernel, same inputs: deterministic forward
    torch.testing.assert_close(rc2, (ra * 0.5).expand(-1, -1, -1, 3), rtol=1e-4, atol=1e-5)
    # gradients: finite, zero on invisible Gaussians, linear in the cotangent
    vc = torch.randn_like(rc)
    g1 = torch.autograd.grad((rc * vc).sum() + ra.sum(), P, retain_graph=True)
    g2 = torch.autograd.grad((rc * (2 * vc)).sum() + (2 * ra).sum(), P)
    visible = torch.zeros(N, dtype=torch.bool, device=DEV)
    if packed:
        visible[m["gaussian_ids"]] = True
    else:
        visible = m["radii"][0] > 0
    for a, b in zip(g1, g2):
        if a.is_sparse:
            assert sparse_grad and a.values().shape[0] == m["gaussian_ids"].numel()
            a, b = a.to_dense(), b.to_dense()
        assert torch.isfinite(a).all()
        if (~visible).any():
            assert a[~visible].abs().max() == 0
        scale = a.abs().max()
        assert ((2 * a - b).abs().max() <= 2e-3 * scale), "gradient is not linear in the cotangent"
    assert (g1[3].to_dense() if g1[3].is_sparse else g1[3]).abs().gt(0).float().mean() > 0.05
    return m


def test_config_b_full_size_properties():
    """BASELINE config B: 1 M Gaussians, SH3, 1920x1080 pinhole."""
    _full_size_properties(synthetic.pinhole_scene(1_000_000, 1920, 1080, seed=42), 1920, 1080)


def test_config_c_per_rank_full_size_properties():
    """BASELINE config C, one rank's share: 3 M Gaussians, one 1080p camera."""
    _full_size_properties(synthetic.pinhole_scene(3_000_000, 1920, 1080, seed=43), 1920, 1080, min_isects=5_000_000)


def test_config_d_spherical_full_size_properties():
    """BASELINE config D: 2 M Gaussians, equirectangular 2048x1024 (the fork's camera model)."""
    _full_size_properties(synthetic.spherical_scene(2_000_000, 2048, 1024, seed=44), 2048, 1024,
                          camera_model="spherical")


def test_config_e_packed_sparse_full_size_properties():
    """BASELINE config E: 6 M Gaussians, 3840x2160, packed mode with sparse gradients."""
    m = _full_size_properties(synthetic.pinhole_scene(6_000_000, 3840, 2160, seed=45), 3840, 2160, packed=True,
                              sparse_grad=True, min_isects=10_000_000)
    assert m["tile_width"] == 240 and m["tile_height"] == 135


def test_gradient_sink_writes_parameter_grads_in_place():
    """With a GradArena sink the backward kernels produce quats/scales/SH gradients inside the
    flat all-reduce arena (no gather copy); values are identical to the plain path."""
    from splat_one_b200.distributed import GradArena

    scene = synthetic.to_device(synthetic.pinhole_scene(30000, 320, 240, seed=3, n_cameras=2), DEV)
    names = ("means", "quats", "scales", "opacities", "sh")

    def run(use_sink):
        P = [scene[k].clone().requires_grad_() for k in names]
        rc, ra, _ = S.rasterization(*P, scene["viewmats"], scene["Ks"], 320, 240, sh_degree=3, packed=False)
        g = torch.Generator(device="cpu").manual_seed(1)
        vc = torch.randn(rc.shape, generator=g).to(DEV)
        arena = GradArena(P)
        if use_sink:
            with arena.sink():
                torch.autograd.backward([rc, ra], [vc, torch.ones_like(ra)])
        else:
            torch.autograd.backward([rc, ra], [vc, torch.ones_like(ra)])
        in_place = [p.grad.data_ptr() == v.data_ptr() for p, v in zip(P, arena.views)]
        arena.gather_from_params()
        return [p.grad.clone() for p in P], in_place, arena.flat.clone()

    g0, in0, flat0 = run(False)
    g1, in1, flat1 = run(True)
    assert in0 == [False] * 5 and in1 == [False, True, True, False, True]
    for n, a, b in zip(names, g0, g1):
        assert_grad_close(b, a, rtol=1e-4, what=n)
    assert_grad_close(flat1, flat0, rtol=1e-4, what="arena")


@pytest.mark.parametrize("n_cameras", [1, 2])
def test_rasterization_packed_sparse_grad_matches_oracle(n_cameras):
    """packed=True, sparse_grad=True through rasterization(): quats / scales gradients are COO
    tensors over gaussian_ids (G/cuda/_wrapper.py:1163-1203), means is dense (it also feeds the
    SH view directions); values equal the oracle's dense gradients."""
    scene = synthetic.pinhole_scene(5000, 176, 144, seed=11, n_cameras=n_cameras)
    names = ["means", "quats", "scales", "opacities", "sh"]
    P_c = [scene[k].clone().requires_grad_() for k in names]
    P_g = [scene[k].to(DEV).requires_grad_() for k in names]
    kw = dict(width=176, height=144, sh_degree=3, packed=True)
    rc_r, ra_r, m_r = O.rasterization(*P_c, scene["viewmats"], scene["Ks"], raster_fn=RC.rasterize_to_pixels, **kw)
    rc, ra, m = S.rasterization(*P_g, scene["viewmats"].to(DEV), scene["Ks"].to(DEV), sparse_grad=True, **kw)
    margin = _margin_of(m_r, scene)
    assert_image_close(rc, rc_r, margin, what="colors")
    g = torch.Generator().manual_seed(4)
    keep = (margin >= 1e-3).float()[..., None]
    vc = torch.randn(rc_r.shape, generator=g) * keep
    g_ref = torch.autograd.grad((rc_r * vc).sum() + (ra_r * keep).sum(), P_c)
    g_got = torch.autograd.grad((rc * vc.to(DEV)).sum() + (ra * keep.to(DEV)).sum(), P_g)
    assert not g_got[0].is_sparse and g_got[1].is_sparse and g_got[2].is_sparse
    assert g_got[1].is_coalesced() == (n_cameras == 1)
    assert torch.equal(g_got[1]._indices()[0], m["gaussian_ids"])
    for n, a, b in zip(names, g_got, g_ref):
        a = a.to_dense() if a.is_sparse else a
        assert_grad_close(a, b, rtol=2e-3, what=f"sparse/{n}", frac_ok=0.995)


def test_two_python_threads_render_concurrently():
    """splat_one calls the rasterizer from its training thread and from the Qt viewer thread at the
    same time (R/app/gsplat_manager.py:185 vs :204); ctypes releases the GIL, so the library, the
    per-thread pinned read-back ring and the per-thread CUDA streams must be re-entrant.  Two
    threads render different scenes in a loop on their own streams; every result must equal the
    single-threaded one bit for bit (integer outputs) / exactly (same kernels, same inputs)."""
    import threading

    dev = "cuda:0"
    scenes = [synthetic.pinhole_scene(6000 + 500 * i, 160 + 16 * i, 120, seed=20 + i) for i in range(2)]
    args = []
    for sc in scenes:
        P = [sc[k].to(dev) for k in ("means", "quats", "scales", "opacities", "sh")]
        args.append((P, sc["viewmats"].to(dev), sc["Ks"].to(dev), sc["width"], sc["height"]))

    def render(i, packed):
        P, vm, Ks, W, H = args[i]
        with torch.no_grad():
            rc, ra, meta = S.rasterization(*P, vm, Ks, W, H, sh_degree=3, packed=packed)
        return rc, ra, meta["flatten_ids"], meta["isect_offsets"]

    expected = [[render(i, packed) for packed in (False, True)] for i in range(2)]
    torch.cuda.synchronize()
    errors = []

    def worker(i):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for it in range(12):
                    packed = bool(it & 1)
                    rc, ra, flat, offs = render(i, packed)
                    stream.synchronize()
                    e = expected[i][int(packed)]
                    assert torch.equal(flat, e[2]) and torch.equal(offs, e[3]), f"thread {i} iter {it}: ids differ"
                    assert torch.equal(rc, e[0]) and torch.equal(ra, e[1]), f"thread {i} iter {it}: image differs"
        except Exception as ex:  # noqa: BLE001
            errors.append(repr(ex))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
