"""Host-side behaviour of the drop-in API that needs no GPU: shape asserts, error types
and messages follow the reference wrappers (gsplat/cuda/_wrapper.py)."""
import pytest
import torch

import splat_one_b200 as S


def test_rasterization_shape_asserts():
    N, C = 10, 1
    ok = dict(means=torch.zeros(N, 3), quats=torch.zeros(N, 4), scales=torch.zeros(N, 3),
              opacities=torch.zeros(N), colors=torch.zeros(N, 3), viewmats=torch.eye(4)[None],
              Ks=torch.eye(3)[None], width=32, height=32)
    for key, bad in [("means", torch.zeros(N, 2)), ("quats", torch.zeros(N, 3)), ("scales", torch.zeros(N, 4)),
                     ("opacities", torch.zeros(N, 1)), ("viewmats", torch.eye(3)[None]), ("Ks", torch.eye(4)[None]),
                     ("colors", torch.zeros(N + 1, 3))]:
        kw = dict(ok)
        kw[key] = bad
        with pytest.raises(AssertionError):
            S.rasterization(**kw)
    with pytest.raises(AssertionError):
        S.rasterization(**ok, render_mode="RGBD")
    with pytest.raises(AssertionError):  # (sh_degree+1)^2 <= K
        S.rasterization(**{**ok, "colors": torch.zeros(N, 4, 3)}, sh_degree=3)
    with pytest.raises(NotImplementedError):
        S.rasterization(**ok, distributed=True)


def test_no_cpu_fallback():
    N = 4
    with pytest.raises(RuntimeError, match="CUDA"):
        S.fully_fused_projection(torch.zeros(N, 3), None, torch.ones(N, 4), torch.ones(N, 3), torch.eye(4)[None],
                                 torch.eye(3)[None], 8, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        S.isect_tiles(torch.zeros(1, N, 2), torch.ones(1, N, dtype=torch.int32), torch.ones(1, N), 16, 1, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        S.isect_offset_encode(torch.zeros(3, dtype=torch.int64), 1, 1, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        S.spherical_harmonics(0, torch.zeros(N, 3), torch.zeros(N, 1, 3))


def test_projection_argument_checks():
    N = 4
    a = (torch.zeros(N, 3), None, torch.ones(N, 4), torch.ones(N, 3), torch.eye(4)[None], torch.eye(3)[None], 8, 8)
    with pytest.raises(AssertionError, match="sparse_grad"):
        S.fully_fused_projection(*a, sparse_grad=True, packed=False)
    with pytest.raises(AssertionError):
        S.fully_fused_projection(torch.zeros(N, 3), None, None, None, torch.eye(4)[None], torch.eye(3)[None], 8, 8)
    with pytest.raises(AttributeError):  # CameraModelType has no such member (_wrapper.py:796-798)
        S.fully_fused_projection(*a, camera_model="cylindrical")


def test_rasterize_to_pixels_argument_checks():
    C, N = 1, 5
    offs = torch.zeros(C, 2, 2, dtype=torch.int32)
    base = (torch.zeros(C, N, 2), torch.zeros(C, N, 3))
    with pytest.raises(ValueError, match="Unsupported number of color channels"):
        S.rasterize_to_pixels(*base, torch.zeros(C, N, 0), torch.zeros(C, N), 32, 32, 16, offs,
                              torch.zeros(0, dtype=torch.int32))
    with pytest.raises(ValueError, match="Unsupported number of color channels"):
        S.rasterize_to_pixels(*base, torch.zeros(C, N, 514), torch.zeros(C, N), 32, 32, 16, offs,
                              torch.zeros(0, dtype=torch.int32))
    with pytest.raises(AssertionError, match="Assert Failed"):
        S.rasterize_to_pixels(*base, torch.zeros(C, N, 3), torch.zeros(C, N), 64, 32, 16, offs,
                              torch.zeros(0, dtype=torch.int32))
    with pytest.raises(AssertionError):
        S.rasterize_to_pixels(*base, torch.zeros(C, N, 3), torch.zeros(C, N), 32, 32, 16, offs,
                              torch.zeros(0, dtype=torch.int32), backgrounds=torch.zeros(C, 4))


def test_isect_tiles_packed_requires_ids():
    with pytest.raises(AssertionError, match="camera_ids"):
        S.isect_tiles(torch.zeros(3, 2), torch.ones(3, dtype=torch.int32), torch.ones(3), 16, 1, 1, packed=True,
                      n_cameras=1)


def test_sh_argument_checks():
    with pytest.raises(AssertionError):
        S.spherical_harmonics(3, torch.zeros(4, 3), torch.zeros(4, 9, 3))
    with pytest.raises(AssertionError):
        S.spherical_harmonics(1, torch.zeros(4, 3), torch.zeros(5, 4, 3))


def test_step_api_argument_checks_and_no_cpu_fallback():
    """f4 row: the fused training-step entry points validate like the operators and have no CPU path."""
    N = 6
    with pytest.raises(AssertionError):
        S.splat_activations(torch.zeros(N, 2), torch.zeros(N))
    with pytest.raises(RuntimeError, match="CUDA"):
        S.splat_activations(torch.zeros(N, 3), torch.zeros(N))
    with pytest.raises(AssertionError, match="11x11"):
        S.l1_ssim_loss(torch.zeros(1, 8, 32, 3), torch.zeros(1, 8, 32, 3))
    with pytest.raises(AssertionError):
        S.l1_ssim_loss(torch.zeros(1, 16, 16, 3), torch.zeros(1, 16, 17, 3))
    with pytest.raises(AssertionError):
        S.l1_ssim_loss(torch.zeros(1, 16, 16, 4), torch.zeros(1, 16, 16, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        S.l1_ssim_loss(torch.zeros(1, 16, 16, 3), torch.zeros(1, 16, 16, 3))
    from splat_one_b200.step import invert_poses

    with pytest.raises(RuntimeError, match="CUDA"):
        invert_poses(torch.eye(4)[None])
    # poses that require a gradient go through torch (pose optimisation), CPU included
    p = torch.eye(4)[None].clone().requires_grad_()
    assert torch.allclose(invert_poses(p), torch.eye(4)[None])


def test_rasterization_split_sh_table_asserts():
    N = 10
    ok = dict(means=torch.zeros(N, 3), quats=torch.zeros(N, 4), scales=torch.zeros(N, 3), opacities=torch.zeros(N),
              viewmats=torch.eye(4)[None], Ks=torch.eye(3)[None], width=32, height=32)
    with pytest.raises(AssertionError, match="sh_degree"):
        S.rasterization(**ok, colors=(torch.zeros(N, 1, 3), torch.zeros(N, 15, 3)))
    with pytest.raises(AssertionError):
        S.rasterization(**ok, colors=(torch.zeros(N, 2, 3), torch.zeros(N, 15, 3)), sh_degree=3)
    with pytest.raises(AssertionError):
        S.rasterization(**ok, colors=(torch.zeros(N, 1, 3), torch.zeros(N + 1, 15, 3)), sh_degree=3)
    with pytest.raises(AssertionError):  # (sh_degree+1)^2 <= 1 + shN.shape[1]
        S.rasterization(**ok, colors=(torch.zeros(N, 1, 3), torch.zeros(N, 3, 3)), sh_degree=3)
    # CPU tensors take the concatenating route and then fail loudly at the first kernel
    with pytest.raises(RuntimeError, match="CUDA"):
        S.rasterization(**ok, colors=(torch.zeros(N, 1, 3), torch.zeros(N, 15, 3)), sh_degree=3)


def test_peer_exchange_layout_is_16_byte_aligned_and_disjoint():
    """Host-side layout of the symmetric buffer of the peer gradient exchange (flags | slot 0 | slot 1 |
    arena): every region starts on a 16-byte boundary and holds what the kernels index."""
    from splat_one_b200._lib import get_lib
    from splat_one_b200.distributed import PeerExchange

    lib = get_lib()
    for world in (2, 3, 8):
        fb = lib.b200splat_peer_flag_bytes(world)
        assert fb >= (1 + 2 * 148) * world * 4
        for n, cm in ((40000, 1), (1000003, 2), (7, 3)):
            hdr, slot, flags, total = PeerExchange.layout(n, cm, world, fb)
            assert hdr % 4 == 0 and hdr >= 3 * cm
            assert slot % 4 == 0 and slot >= hdr + 3 * cm * n
            assert flags % 4 == 0 and flags * 4 >= fb
            assert total == flags + 2 * slot
