"""CPU tests of the f4 oracle (oracle/step_ref.py): the SSIM restatement against an
independent float64 scipy implementation and SSIM identities, and the activation / loss
plumbing.  No GPU, no product code on the compute path."""
import numpy as np
import pytest
import torch
from scipy.ndimage import correlate1d

from oracle import step_ref as SR


def _ssim_valid_scipy(a, b):
    """Independent restatement: float64, separable correlate1d with zero ('constant') padding,
    [H,W,3] arrays, mean over the interior."""
    x = np.arange(11, dtype=np.float64) - 5
    g = np.exp(-x * x / (2 * 1.5 ** 2))
    g /= g.sum()

    def blur(z):
        z = correlate1d(z, g, axis=0, mode="constant", cval=0.0)
        return correlate1d(z, g, axis=1, mode="constant", cval=0.0)

    a, b = a.astype(np.float64), b.astype(np.float64)
    mu1, mu2 = blur(a), blur(b)
    s1, s2, s12 = blur(a * a) - mu1 ** 2, blur(b * b) - mu2 ** 2, blur(a * b) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 ** 2 + mu2 ** 2 + C1) * (s1 + s2 + C2))
    return m[5:-5, 5:-5].mean()


@pytest.mark.parametrize("H,W", [(24, 31), (64, 48)])
def test_ssim_restatement_matches_independent_float64(H, W):
    g = torch.Generator().manual_seed(3)
    a = torch.rand(2, H, W, 3, generator=g)
    b = (a + 0.1 * torch.randn(2, H, W, 3, generator=g)).clamp(0, 1)
    got = SR.fused_ssim(a.permute(0, 3, 1, 2), b.permute(0, 3, 1, 2)).item()
    ref = np.mean([_ssim_valid_scipy(a[i].numpy(), b[i].numpy()) for i in range(2)])
    assert abs(got - ref) < 2e-6, (got, ref)


def test_ssim_identities():
    g = torch.Generator().manual_seed(4)
    a = torch.rand(1, 3, 40, 40, generator=g)
    b = torch.rand(1, 3, 40, 40, generator=g)
    assert abs(SR.fused_ssim(a, a).item() - 1.0) < 1e-6
    assert abs(SR.fused_ssim(a, b).item() - SR.fused_ssim(b, a).item()) < 1e-6
    assert -1.0 <= SR.fused_ssim(a, b).item() <= 1.0
    w = SR.gaussian_window()
    assert abs(w.sum().item() - 1.0) < 1e-6 and torch.equal(w, w.flip(0))
    # the normalised window fused-ssim hard-codes (its published constants, float32)
    assert abs(w[5].item() - 0.26601171493530273) < 1e-7 and abs(w[0].item() - 0.001028380123898387) < 1e-9


@pytest.mark.parametrize("a,b", [(0.3, 0.7), (0.05, 0.9), (0.5, 0.5)])
def test_ssim_known_answer_on_constant_images(a, b):
    """Closed-form known answer (tests/parity.py::ssim_constant_images_closed_form): pins the window, the
    constants, the zero padding of the local statistics ("same": border pixels see a partial window) and the
    5-pixel "valid" crop, independently of any convolution code."""
    from parity import ssim_constant_images_closed_form as kat

    H, W = 23, 37
    # float64: on constant images the variances are differences of equal numbers, and fp32 rounding noise
    # (~1e-7) is measured against C2 = 9e-4
    x = torch.full((1, 3, H, W), a, dtype=torch.float64)
    y = torch.full((1, 3, H, W), b, dtype=torch.float64)
    assert abs(SR.fused_ssim(x, y, "valid").item() - kat(a, b, H, W, "valid")) < 1e-10
    assert abs(SR.fused_ssim(x, y, "same").item() - kat(a, b, H, W, "same")) < 1e-10
    assert abs(SR.fused_ssim(x.float(), y.float(), "valid").item() - kat(a, b, H, W, "valid")) < 3e-4
    assert abs(kat(a, b, H, W, "valid") - (2 * a * b + 1e-4) / (a * a + b * b + 1e-4)) < 1e-12


def test_loss_composition_and_gradient_flows_to_first_argument_only():
    g = torch.Generator().manual_seed(5)
    a = torch.rand(1, 20, 22, 3, generator=g, requires_grad=True)
    b = torch.rand(1, 20, 22, 3, generator=g)
    loss, l1, ssim = SR.l1_ssim_loss(a, b, 0.2)
    assert abs(loss.item() - (0.8 * l1.item() + 0.2 * (1 - ssim.item()))) < 1e-6
    loss.backward()
    assert a.grad is not None and a.grad.abs().sum() > 0


def test_activations():
    s, o = SR.splat_activations(torch.tensor([[0.0, 1.0, -1.0]]), torch.tensor([0.0]))
    assert torch.allclose(s, torch.tensor([[1.0, 2.718281828, 0.36787944]])) and abs(o.item() - 0.5) < 1e-7
