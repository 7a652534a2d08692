"""Shared helpers for the GPU parity tests (test infrastructure)."""
import torch

# north-star tolerances (BASELINE.json): 1e-4 abs on images, 1e-3 rel on gradients
IMG_ATOL = 1e-4
GRAD_RTOL = 1e-3


def assert_image_close(got, ref, margin=None, atol=IMG_ATOL, rtol=1e-4, max_ambiguous_frac=2e-3, what="image"):
    """|got-ref| <= atol + rtol|ref| on every pixel, except pixels the oracle flags as
    sitting within 1e-3 (relative) of a skip/stop threshold — there a different-but-valid
    rounding of exp() may flip a decision.  Those are counted, bounded,
    and must still agree to the size of one flipped contribution."""
    got, ref = got.detach().cpu().float(), ref.detach().cpu().float()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    if margin is not None:
        amb = (margin.detach().cpu() < 1e-3)
        while amb.dim() < bad.dim():
            amb = amb[..., None]
        amb = amb.expand_as(bad)
        strict_bad = bad & ~amb
        n_amb_bad = int((bad & amb).sum())
        assert n_amb_bad <= max_ambiguous_frac * bad.numel() + 8, f"{what}: {n_amb_bad} ambiguous pixels differ"
        bad = strict_bad
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.numel()} elements differ, max err "
                           f"{err[bad].max().item():.3e}, max ref {ref.abs().max().item():.3e}")


def assert_grad_close(got, ref, rtol=GRAD_RTOL, what="grad", frac_ok=0.999):
    """Gradient parity: relative to the tensor's scale (atomics reorder fp32 sums, and a
    flipped threshold decision moves single entries), 1e-3·max|ref| absolute + 1e-3 rel on
    at least `frac_ok` of the entries, and 1e-2 of the scale everywhere."""
    got, ref = got.detach().cpu().double(), ref.detach().cpu().double()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    scale = max(ref.abs().max().item(), 1e-12)
    err = (got - ref).abs()
    ok = err <= rtol * scale + rtol * ref.abs()
    frac = ok.double().mean().item() if ok.numel() else 1.0
    assert frac >= frac_ok, f"{what}: only {frac:.5f} of entries within {rtol} (scale {scale:.3e}, max err {err.max():.3e})"
    assert err.max().item() <= 20 * rtol * scale, f"{what}: max err {err.max().item():.3e} vs scale {scale:.3e}"
