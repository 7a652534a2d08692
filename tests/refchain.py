"""Test infrastructure (not product code): the reference's OWN CUDA kernels (oracle/_ref, built by
oracle/build_ref.py from /root/reference/submodules/gsplat/gsplat/cuda/csrc) chained exactly the way
G/rendering.py:297-558 chains its operators and G/cuda/_wrapper.py's autograd Functions run their
backwards — driven through the raw pybind entry points (CS/ext.cpp:11-56), because the reference's
Python does not exist on the GPU box — plus the classification helpers of the full-size parity tests:

  * `pixel_margins`: for a list of pixels, the smallest relative distance of any evaluated
    (pixel, Gaussian) pair to one of the two decision thresholds of the rasterizer
    (alpha = 1/255, T(1-alpha) = 1e-4; CS/rasterize_to_pixels_fwd.cu:143-155), in float64, and the
    list entries each pixel evaluates.  Same definition as the `margin` output of oracle/raster_ref.c
    (tests/test_oracle.py pins one against the other).
  * `compare_full_size`: image / gradient comparison with every outlier classified.
"""
import json
import os

import torch

ALPHA_MIN, ALPHA_MAX, T_EPS = 1.0 / 255.0, 0.999, 1e-4


from oracle.ref_cuda import reference_chain  # noqa: E402,F401  (the chain itself lives next to the loader)


@torch.no_grad()
def pixel_margins(pix, means2d, conics, opacities, offsets, flatten_ids, W, H, tile=16, pad=16, chunk_elems=1 << 25):
    """pix: [F] int64 linear pixel indices (camera-major).  Returns (margin [F] float64, elems): `elems`
    = the flat Gaussian rows (indices into means2d.view(-1, 2)) the flagged pixels evaluate up to their
    stop, plus `pad` entries beyond it (a flipped stop decision moves the end of the list)."""
    dev = means2d.device
    C, th, tw = offsets.shape
    n_isects = flatten_ids.numel()
    m2 = means2d.reshape(-1, 2).double()
    cn = conics.reshape(-1, 3).double()
    op = opacities.reshape(-1).double()
    offs = torch.cat([offsets.reshape(-1).long(), torch.tensor([n_isects], device=dev)])
    c = pix // (H * W)
    y = (pix % (H * W)) // W
    x = pix % W
    t = c * (th * tw) + (y // tile) * tw + (x // tile)
    start, end = offs[t], offs[t + 1]
    margins = torch.ones(pix.numel(), dtype=torch.float64, device=dev)
    touched = []
    if pix.numel() == 0:
        return margins, torch.zeros(0, dtype=torch.long, device=dev)
    L = int((end - start).max())
    rows = max(1, chunk_elems // max(L, 1))
    ar = torch.arange(L, device=dev)
    for a in range(0, pix.numel(), rows):
        s, e = start[a:a + rows], end[a:a + rows]
        idx = s[:, None] + ar[None]
        valid = idx < e[:, None]
        g = flatten_ids[idx.clamp(max=max(n_isects - 1, 0))].long()
        dx = m2[g, 0] - (x[a:a + rows, None].double() + 0.5)
        dy = m2[g, 1] - (y[a:a + rows, None].double() + 0.5)
        sigma = 0.5 * (cn[g, 0] * dx * dx + cn[g, 2] * dy * dy) + cn[g, 1] * dx * dy
        alpha = torch.clamp_max(op[g] * torch.exp(-sigma), ALPHA_MAX)
        ok = valid & (sigma >= 0) & (alpha >= ALPHA_MIN)
        next_T = torch.cumprod(torch.where(ok, 1.0 - alpha, torch.ones_like(alpha)), dim=1)
        stop = ok & (next_T <= T_EPS)
        before_stop = (stop.cumsum(1) - stop.long()) == 0          # up to and including the first stop
        evaluated = valid & before_stop
        big = torch.full_like(alpha, 1.0)
        m_a = torch.where(evaluated & (sigma >= 0), (alpha * 255.0 - 1.0).abs(), big)
        m_t = torch.where(evaluated & ok, (next_T / T_EPS - 1.0).abs(), big)
        margins[a:a + rows] = torch.minimum(m_a, m_t).min(dim=1).values
        n_eval = evaluated.sum(1, keepdim=True)
        near = valid & (ar[None] < n_eval + pad)
        touched.append(g[near])
    return margins, torch.unique(torch.cat(touched))


def _dense_rows(ref, t, n_gauss):
    """[C=1, N, ...] or packed [nnz, ...] rows -> [N, ...] (zeros for rows that were culled)."""
    g_ids = ref["gaussian_ids"]
    if g_ids is None:
        return t.reshape((n_gauss,) + t.shape[2:])
    out = torch.zeros((n_gauss,) + t.shape[1:], dtype=t.dtype, device=t.device)
    out[g_ids] = t
    return out


def _dense_radii(ref, n_gauss):
    return _dense_rows(ref, ref["radii"], n_gauss)


@torch.no_grad()
def compare_full_size(name, ours, ref, n_gauss, W, H, image_atol=1e-4, grad_rtol=1e-3, margin_tol=1e-3):
    """ours / ref: dicts with image, alpha, grads{name: dense tensor}; ref also carries the
    intermediates of `reference_chain`.  Returns a report dict; raises AssertionError on a parity
    violation.  Every pixel outside `image_atol` must sit within `margin_tol` (relative) of a decision
    threshold (a different-but-valid rounding of sigma / ex2 flips the skip or stop
    decision there); every gradient row outside `grad_rtol` (of the tensor's scale) must belong to a
    Gaussian evaluated by such a pixel."""
    d_img = (ours["image"] - ref["image"]).abs().amax(dim=-1)
    d_alpha = (ours["alpha"] - ref["alpha"]).abs().amax(dim=-1)
    flagged = (d_img > image_atol) | (d_alpha > image_atol)
    pix = flagged.reshape(-1).nonzero().squeeze(1)
    n_pix = flagged.numel()
    margins, elems = pixel_margins(pix, ref["means2d"], ref["conics"], ref["opacities"], ref["offsets"],
                                   ref["flatten_ids"], W, H)
    g_ids = ref["gaussian_ids"]
    affected = torch.zeros(n_gauss, dtype=torch.bool, device=d_img.device)
    if elems.numel():
        affected[(g_ids[elems] if g_ids is not None else elems % n_gauss)] = True
    # second class: Gaussians whose integer radius (or visibility) differs between the two projections
    # (3 sigma within an ulp of an integer, or a cull test on the fence): they are listed in a
    # different set of tiles, so pixels inside their bounding square may differ without any threshold
    # being involved (one camera per call here)
    r_ours, r_ref = ours["radii_dense"], _dense_radii(ref, n_gauss)
    rmis = (r_ours != r_ref).nonzero().squeeze(1)
    affected[rmis] = True
    unexplained = margins >= margin_tol
    if rmis.numel() and unexplained.any():
        m2 = _dense_rows(ref, ref["means2d"], n_gauss)[rmis]
        m2o = ours["means2d_dense"][rmis]
        rad = torch.maximum(r_ours[rmis], r_ref[rmis]).float() + 1.0
        px = (pix[unexplained] % W).float() + 0.5
        py = ((pix[unexplained] % (H * W)) // W).float() + 0.5
        inside = torch.zeros_like(px, dtype=torch.bool)
        for c2, r2 in ((m2, rad), (m2o, rad)):
            inside |= (((px[:, None] - c2[None, :, 0]).abs() <= r2[None]) &
                       ((py[:, None] - c2[None, :, 1]).abs() <= r2[None])).any(dim=1)
        unexplained[unexplained.clone()] = ~inside
    rep = {
        "config": name, "pixels": n_pix, "pixels_outside_1e-4": int(pix.numel()),
        "frac_pixels_outside_1e-4": pix.numel() / n_pix,
        "max_image_err": float(torch.maximum(d_img.max(), d_alpha.max())),
        "max_margin_of_outside_pixels": float(margins.max()) if pix.numel() else 0.0,
        "outside_pixels_not_at_a_threshold": int(unexplained.sum()),
        "radius_mismatches": int(rmis.numel()),
        "gaussians_in_those_pixels": int(affected.sum()), "n_isects_ref": int(ref["flatten_ids"].numel()),
        "grads": {},
    }
    for k, g_ref in ref["grads"].items():
        g = ours["grads"][k]
        scale = float(g_ref.abs().max()) + 1e-30
        err = (g - g_ref).abs().reshape(n_gauss, -1).amax(dim=1) / scale
        strict_bad = (err > grad_rtol) & ~affected
        rep["grads"][k] = {
            "scale": scale, "max_rel_err_all": float(err.max()),
            "max_rel_err_outside_flagged": float(err[~affected].max()) if (~affected).any() else 0.0,
            "rows_outside_1e-3": int((err > grad_rtol).sum()),
            "rows_outside_1e-3_not_explained": int(strict_bad.sum()),
        }
    line = json.dumps(rep)
    print("PARITY " + line)
    out = os.environ.get("B200SPLAT_PARITY_REPORT")
    if out:
        with open(out, "a") as f:
            f.write(line + "\n")
    assert rep["frac_pixels_outside_1e-4"] <= 1e-3, rep
    assert rep["outside_pixels_not_at_a_threshold"] == 0, rep
    assert rep["max_image_err"] <= 2e-2, rep  # one flipped contribution: alpha*T*colour <= ~0.999/255 * colour
    for k, r in rep["grads"].items():
        assert r["rows_outside_1e-3_not_explained"] == 0, (k, rep)
        assert r["max_rel_err_all"] <= 5e-2, (k, rep)
        assert r["rows_outside_1e-3"] <= max(64, rep["gaussians_in_those_pixels"]), (k, rep)
    return rep
