"""Full-size, ASSERTING parity of rasterization() forward + backward against the reference's own CUDA
kernels (oracle/_ref) on BASELINE.json's configs B, C (one rank's camera), D and E — the sizes the
benchmark numbers are quoted on.  Same synthetic scenes as bench.py / tools/config_bench.py.

Bounds (north star): image and alpha 1e-4 absolute; parameter gradients 1e-3 of each tensor's scale.
Every element outside those bounds must be EXPLAINED, and the explanation is checked, not assumed
(tests/refchain.py): a pixel outside 1e-4 must evaluate a (pixel, Gaussian) pair within 1e-3
(relative, float64) of one of the rasterizer's two decision thresholds, or lie inside the bounding
square of a Gaussian whose integer radius differs between the two projections; a gradient row
outside 1e-3 must belong to a Gaussian evaluated by such a pixel.  The counts of both classes are
bounded and printed (`PARITY {...}` lines; B200SPLAT_PARITY_REPORT=<file> collects them)."""
import pytest
import torch

import splat_one_b200 as S
from oracle import ref_cuda
from refchain import compare_full_size, reference_chain
from splat_one_b200 import synthetic

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_cuda.available(),
                                                  reason="oracle/_ref not built (python oracle/build_ref.py)")]
DEV = "cuda:0"
NAMES = ("means", "quats", "scales", "opacities", "sh")


def _run(name, scene, W, H, model="pinhole", packed=False, sparse_grad=False):
    R = ref_cuda.load()
    N = scene["means"].shape[0]
    P = {k: scene[k].to(DEV) for k in NAMES + ("viewmats", "Ks")}
    g = torch.Generator().manual_seed(1)
    vc = torch.randn(1, H, W, 3, generator=g).to(DEV)
    va = torch.randn(1, H, W, 1, generator=g).to(DEV)
    A = {k: P[k].clone().requires_grad_() for k in NAMES}
    rc, ra, meta = S.rasterization(A["means"], A["quats"], A["scales"], A["opacities"], A["sh"], P["viewmats"], P["Ks"],
                                   W, H, sh_degree=3, packed=packed, sparse_grad=sparse_grad, camera_model=model)
    torch.autograd.backward([rc, ra], [vc, va])
    grads = {k: (A[k].grad.to_dense() if A[k].grad.is_sparse else A[k].grad) for k in NAMES}
    if packed:
        radii_dense = torch.zeros(N, dtype=torch.int32, device=DEV)
        radii_dense[meta["gaussian_ids"]] = meta["radii"]
        m2_dense = torch.zeros(N, 2, device=DEV)
        m2_dense[meta["gaussian_ids"]] = meta["means2d"].detach()
    else:
        radii_dense, m2_dense = meta["radii"][0], meta["means2d"].detach()[0]
    ours = dict(image=rc.detach(), alpha=ra.detach(), grads=grads, radii_dense=radii_dense, means2d_dense=m2_dense)
    del A
    ref = reference_chain(R, P, W, H, model, vc, va, packed=packed, sparse_grad=sparse_grad)
    # the bit-exact part of the contract, end to end: same projection outputs => same intersections
    if torch.equal(radii_dense, (ref["radii"][0] if not packed else
                                 torch.zeros_like(radii_dense).index_put_((ref["gaussian_ids"],), ref["radii"]))) \
            and torch.equal(meta["means2d"].detach().reshape(-1, 2), ref["means2d"].reshape(-1, 2)) \
            and torch.equal(meta["depths"].detach().reshape(-1), ref["depths"].reshape(-1)):
        assert torch.equal(meta["isect_ids"], ref["isect_ids"]) and torch.equal(meta["flatten_ids"], ref["flatten_ids"])
        assert torch.equal(meta["isect_offsets"], ref["offsets"])
    rep = compare_full_size(name, ours, ref, N, W, H)
    assert abs(meta["flatten_ids"].numel() - rep["n_isects_ref"]) <= 1e-4 * rep["n_isects_ref"] + 64
    return rep


def test_config_b_vs_reference_cuda():
    """BASELINE config B: 1 M Gaussians, SH3, one 1920x1080 pinhole camera."""
    _run("B", synthetic.pinhole_scene(1_000_000, 1920, 1080, seed=42), 1920, 1080)


def test_config_c_one_rank_vs_reference_cuda():
    """BASELINE config C, the camera of one rank: 3 M Gaussians at 1080p."""
    _run("C_rank", synthetic.pinhole_scene(3_000_000, 1920, 1080, seed=43), 1920, 1080)


def test_config_d_spherical_vs_reference_cuda():
    """BASELINE config D: 2 M Gaussians, equirectangular 2048x1024 (the fork's camera model)."""
    _run("D", synthetic.spherical_scene(2_000_000, 2048, 1024, seed=44), 2048, 1024, model="spherical")


def test_config_e_packed_sparse_vs_reference_cuda():
    """BASELINE config E: 6 M Gaussians, 3840x2160, packed mode with sparse gradients."""
    _run("E", synthetic.pinhole_scene(6_000_000, 3840, 2160, seed=45), 3840, 2160, packed=True, sparse_grad=True)
