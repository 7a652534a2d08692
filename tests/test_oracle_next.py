"""Oracle vs the reference-generated golden vectors for the rasterizer itself and for the
§8(f) operators (CPU, no GPU needed).

`raster_ref_d*.npz` are outputs of the reference's OWN `_rasterize_to_pixels` + `accumulate`
(gsplat/cuda/_torch_impl.py) run on the CPU by oracle/gen_golden.py — this is what pins the
fused-raster restatements (torch + C) that the GPU parity tests use as their oracle."""
import pytest
import torch

from oracle import raster_ref as RC
from oracle import torch_ref as O
from parity import assert_grad_close


def _raster_inputs(d):
    return (d["means2d"], d["conics"], d["colors"], d["opacities"], d["width"], d["height"], d["tile_size"],
            d["isect_offsets"], d["flatten_ids"])


@pytest.mark.parametrize("D", [3, 1])
@pytest.mark.parametrize("impl", ["torch", "c"])
def test_fused_raster_oracles_match_reference_compositing(golden, D, impl):
    d = golden(f"raster_ref_d{D}.npz")
    assert "_rasterize_to_pixels" in d["source"]
    m2, con, col, op, W, H, ts, offs, fl = _raster_inputs(d)
    P = [t.clone().requires_grad_() for t in (m2, con, col, op, d["backgrounds"])]
    fn = O.rasterize_to_pixels if impl == "torch" else RC.rasterize_to_pixels
    rc, ra = fn(P[0], P[1], P[2], P[3], W, H, ts, offs, fl, backgrounds=P[4])
    # same decisions (the restatements use the same fp32 exp on the CPU), fp32 summation order differs
    assert (rc - d["ref_render_colors"]).abs().max() < 5e-6
    assert (ra - d["ref_render_alphas"]).abs().max() < 5e-6
    grads = torch.autograd.grad((rc * d["v_render_colors"]).sum() + (ra * d["v_render_alphas"]).sum(), P)
    for name, g in zip(("means2d", "conics", "colors", "opacities", "backgrounds"), grads):
        assert_grad_close(g, d[f"ref_v_{name}"], rtol=1e-4, what=f"{impl} v_{name}", frac_ok=0.9999)


def test_indices_and_accumulate_reproduce_reference_render(golden):
    d = golden("raster_ref_d3.npz")
    m2, con, col, op, W, H, ts, offs, fl = _raster_inputs(d)
    C = m2.shape[0]
    gi, pi, ci = O.rasterize_to_indices_in_range(0, 10**9, torch.ones(C, H, W), m2, con, op, W, H, ts, offs, fl)
    assert torch.equal(gi, d["idx_gaussian_ids"]) and torch.equal(pi, d["idx_pixel_ids"])
    assert torch.equal(ci, d["idx_camera_ids"])
    rc, ra = O.accumulate(m2, con, op, col, gi, pi, ci, W, H)
    rc = rc + d["backgrounds"][:, None, None, :] * (1.0 - ra)
    assert (rc - d["ref_render_colors"]).abs().max() < 5e-6
    assert (ra - d["ref_render_alphas"]).abs().max() < 5e-6


def test_indices_range_batches_partition_the_full_list(golden):
    """[0,1) + [1,2) + ... with the transmittance carried between calls == one full call
    (how the reference's `_rasterize_to_pixels` iterates, _torch_impl.py:628-662)."""
    d = golden("raster_ref_d3.npz")
    m2, con, col, op, W, H, ts, offs, fl = _raster_inputs(d)
    C = m2.shape[0]
    full = O.rasterize_to_indices_in_range(0, 10**9, torch.ones(C, H, W), m2, con, op, W, H, ts, offs, fl)
    T = torch.ones(C, H, W)
    parts = []
    for b in range(0, 8):
        gi, pi, ci = O.rasterize_to_indices_in_range(b, b + 1, T, m2, con, op, W, H, ts, offs, fl)
        if gi.numel() == 0:
            continue
        parts.append((gi, pi, ci))
        _, acc = O.accumulate(m2, con, op, col, gi, pi, ci, W, H)
        T = T * (1.0 - acc[..., 0])
    # The "done" flag is not carried between calls (CS/rasterize_to_indices_in_range.cu:62): a
    # pixel whose walk stopped (exclusively) in batch b is walked again in batch b+1 from the same
    # transmittance, so the batched lists are a SUPERSET of the one-call list, and every extra
    # pair belongs to a pixel that had stopped.
    key = lambda g, p, c: set(zip((c * H * W + p).tolist(), g.tolist()))
    a, b = key(*full), set().union(*[key(*p) for p in parts])
    assert len(a - b) <= 4
    last_full = {}
    for pix, g in zip((full[2] * H * W + full[1]).tolist(), full[0].tolist()):
        last_full[pix] = last_full.get(pix, 0) + 1
    _, acc = O.accumulate(m2, con, op, col, *full, W, H)
    T_full = (1.0 - acc[..., 0]).flatten()
    extra_pixels = {pix for pix, _ in (b - a)}
    assert all(T_full[p] < 0.5 for p in extra_pixels)


def test_unfused_oracle_matches_reference_torch_impl(golden):
    d = golden("unfused_ref.npz")
    for triu in (False, True):
        t = "triu" if triu else "full"
        q, s = d["quats"].clone().requires_grad_(), d["scales"].clone().requires_grad_()
        cov, pre = O.quat_scale_to_covar_preci(q, s, True, True, triu)
        torch.testing.assert_close(cov, d[f"ref_covars_{t}"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(pre, d[f"ref_precis_{t}"], rtol=1e-4, atol=1e-3)
        gq, gs = torch.autograd.grad((cov * d[f"v_covars_{t}"]).sum() + (pre * d[f"v_precis_{t}"]).sum(), (q, s))
        assert_grad_close(gq, d[f"ref_v_quats_{t}"], rtol=1e-4)
        assert_grad_close(gs, d[f"ref_v_scales_{t}"], rtol=1e-4)
    m, cv, vm = (d[k].clone().requires_grad_() for k in ("means", "covars", "viewmats"))
    mc, cc = O.world_to_cam(m, cv, vm)
    torch.testing.assert_close(mc, d["ref_means_c"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(cc, d["ref_covars_c"], rtol=1e-5, atol=1e-6)
    g = torch.autograd.grad((mc * d["v_means_c"]).sum() + (cc * d["v_covars_c"]).sum(), (m, cv, vm))
    for got, k in zip(g, ("ref_w2c_v_means", "ref_w2c_v_covars", "ref_w2c_v_viewmats")):
        assert_grad_close(got, d[k], rtol=1e-4, what=k)
    for cm in ("pinhole", "ortho", "fisheye"):
        a, b = d["ref_means_c"].clone().requires_grad_(), d["ref_covars_c"].clone().requires_grad_()
        m2, c2 = O.proj(a, b, d["Ks"], d["width"], d["height"], cm)
        torch.testing.assert_close(m2, d[f"ref_means2d_{cm}"], rtol=1e-5, atol=1e-3)
        torch.testing.assert_close(c2, d[f"ref_covars2d_{cm}"], rtol=1e-4, atol=1e-3)
        ga, gb = torch.autograd.grad((m2 * d["v_means2d"]).sum() + (c2 * d["v_covars2d"]).sum(), (a, b))
        assert_grad_close(ga, d[f"ref_proj_v_means_{cm}"], rtol=1e-3, what=f"proj v_means {cm}")
        assert_grad_close(gb, d[f"ref_proj_v_covars_{cm}"], rtol=1e-3, what=f"proj v_covars {cm}")


def test_relocation_and_adam_oracle_known_answers():
    """n = 1 relocation is the identity; n = 2 has the closed form of Eq. (9); selective Adam
    equals torch.optim.Adam's first step without bias correction on the visible rows."""
    import math

    n_max = 51
    binoms = torch.zeros(n_max, n_max)
    for n in range(n_max):
        for k in range(n + 1):
            binoms[n, k] = math.comb(n, k)
    op = torch.tensor([0.3, 0.7, 0.95])
    sc = torch.tensor([[1.0, 2.0, 3.0]] * 3)
    no, ns = O.compute_relocation(op, sc, torch.tensor([1, 1, 1]), binoms)
    torch.testing.assert_close(no, op)
    torch.testing.assert_close(ns, sc)
    no, ns = O.compute_relocation(op, sc, torch.tensor([2, 2, 2]), binoms)
    exp_no = 1 - torch.sqrt(1 - op)
    denom = exp_no + (exp_no - exp_no ** 2 / math.sqrt(2))
    torch.testing.assert_close(no, exp_no)
    torch.testing.assert_close(ns, (op / denom)[:, None] * sc)
    p, g = torch.tensor([[1.0, 2.0], [3.0, 4.0]]), torch.tensor([[0.5, -0.5], [1.0, 1.0]])
    z = torch.zeros_like(p)
    np_, nm, nv = O.selective_adam_update(p, g, z, z, torch.tensor([True, False]), 0.1, 0.9, 0.999, 1e-8)
    assert torch.equal(np_[1], p[1]) and torch.equal(nm[1], z[1])
    m, v = 0.1 * g[0], 0.001 * g[0] ** 2
    torch.testing.assert_close(np_[0], p[0] - 0.1 * m / (v.sqrt() + 1e-8), rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------------------------------------
# f3: DefaultStrategy._update_state restatement against outputs of the reference's own method
# ------------------------------------------------------------------------------------------------
def _strategy_golden():
    import os

    import numpy as np

    return np.load(os.path.join(os.path.dirname(__file__), "golden", "strategy_state.npz"))


@pytest.mark.parametrize("mode", ["unpacked", "packed"])
def test_strategy_state_oracle_matches_reference_outputs(mode):
    from oracle import strategy_ref as SRf

    G = _strategy_golden()
    N, C, W, H = int(G["N"]), int(G["C"]), int(G["width"]), int(G["height"])
    grad2d, count, radii = torch.zeros(N), torch.zeros(N), torch.zeros(N)
    for call in range(2):
        grads = torch.from_numpy(G[f"{mode}_grads{call}"])
        rad = torch.from_numpy(G[f"{mode}_radii{call}"])
        ids = torch.from_numpy(G[f"{mode}_ids{call}"]) if mode == "packed" else None
        SRf.update_state(grad2d, count, radii, grads, rad, W, H, C if mode == "unpacked" else 1, ids)
        assert torch.allclose(grad2d, torch.from_numpy(G[f"{mode}_grad2d_after{call}"]), rtol=1e-6, atol=1e-9)
        assert torch.equal(count, torch.from_numpy(G[f"{mode}_count_after{call}"]))
        assert torch.equal(radii, torch.from_numpy(G[f"{mode}_radii_after{call}"]))
