"""Two-GPU parity of the camera-sharded data-parallel path, NCCL over NVLink:
rank r renders camera r of a 2-camera batch through `rasterization()`, backward runs inside
`GradArena.sink()` + `camera_parallel()` (all-gather of colour cotangents for the SH gradient,
one all-reduce of the arena for the rest), and every rank must end up with the gradients of a
single-process C = 2 batch (1e-3 rel, the north star's multi-GPU parity check).  Needs >= 2 GPUs
(skipped on the one-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path, sh_mode, defer, n_cams, packed_sparse, peer_mode=None):
    import torch.distributed as dist

    import splat_one_b200 as S
    from splat_one_b200 import synthetic
    from splat_one_b200.distributed import (GradArena, PeerExchange, allreduce_mixed_gradients, arena_layout, camera_parallel,
                                            shard_cameras)

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    W, H, N = 320, 240, 40000
    scene = synthetic.pinhole_scene(N, W, H, seed=7, n_cameras=n_cams)
    g = torch.Generator().manual_seed(3)
    vc_all = torch.randn(n_cams, H, W, 3, generator=g)
    va_all = torch.randn(n_cams, H, W, 1, generator=g)
    uneven = n_cams % world != 0
    peer_box = {}

    def grads(cam_ids, dp):
        vm, Ks = scene["viewmats"][cam_ids].to(dev), scene["Ks"][cam_ids].to(dev)
        vc, va = vc_all[cam_ids].to(dev), va_all[cam_ids].to(dev)
        kw = dict(sh_degree=3, packed=packed_sparse is not None, sparse_grad=packed_sparse is not None)
        if sh_mode in ("split", "cat"):
            names = ("means", "quats", "scales", "opacities", "sh0", "shN")
            raw = dict(means=scene["means"], quats=scene["quats"], scales=scene["scales"], opacities=scene["opacities"],
                       sh0=scene["sh"][:, :1].contiguous(), shN=scene["sh"][:, 1:].contiguous())
            P = [raw[k].to(dev).requires_grad_() for k in names]
            # "cat": the table is a temporary, as gsplat_trainer.py:474 builds it
            colors = (P[4], P[5]) if sh_mode == "split" else torch.cat([P[4], P[5]], 1)
            rc, ra, _ = S.rasterization(P[0], P[1], P[2], P[3], colors, vm, Ks, W, H, **kw)
        else:
            names = ("means", "quats", "scales", "opacities", "sh")
            P = [scene[k].to(dev).requires_grad_() for k in names]
            rc, ra, _ = S.rasterization(*P, vm, Ks, W, H, **kw)
        if dp and packed_sparse is not None:
            # packed mode, sparse gradients: (gaussian_ids, rows) exchange or its dense fallback
            if peer_mode is not None:
                # ... and the SH gradient through the colour-cotangent exchange of the packed colour stage
                if "x" not in peer_box:
                    peer_box["x"] = PeerExchange(N, (n_cams + world - 1) // world) if peer_mode != "nccl" else None
                with camera_parallel(peer=peer_box["x"], n_cameras_global=n_cams if uneven else None) as cp:
                    torch.autograd.backward([rc, ra], [vc, va])
                assert P[4].data_ptr() in cp.reduced_ptrs
                skip = cp.reduced_ptrs
            else:
                torch.autograd.backward([rc, ra], [vc, va])
                skip = ()
            assert P[1].grad.is_sparse and P[2].grad.is_sparse
            allreduce_mixed_gradients(P, dense_threshold=packed_sparse, skip_ptrs=skip)
            assert P[1].grad.is_sparse == (packed_sparse > 1.0)
        elif dp:
            peer = None
            if peer_mode is not None:
                # own kernels over NVLink peer memory (csrc/peer.cu): built once, reused by every step
                if "x" not in peer_box:
                    peer_box["x"] = PeerExchange(N, (n_cams + world - 1) // world, arena_floats=arena_layout(P)[1],
                                                 use_multicast=peer_mode == "multicast")
                peer = peer_box["x"]
            arena = GradArena(P, peer=peer)
            with arena.sink(), camera_parallel(defer=defer, n_cameras_global=n_cams if uneven else None,
                                               peer=peer) as cp:
                torch.autograd.backward([rc, ra], [vc, va])
            if defer:  # all-reduce overlapped with the colour backward (camera_parallel.finish)
                cp.finish(arena)
            else:
                arena.gather_from_params()
                arena.all_reduce(skip_ptrs=cp.reduced_ptrs)
            arena.scatter_to_params()
        else:
            torch.autograd.backward([rc, ra], [vc, va])
        return {n: (p.grad.to_dense() if p.grad.is_sparse else p.grad).detach().clone() for n, p in zip(names, P)}

    mine = shard_cameras(scene["viewmats"], scene["Ks"], rank, world, allow_uneven=uneven)[2].tolist()
    g_dp = grads(mine, dp=True)
    if peer_mode is not None:  # both cotangent slots and the self-resetting flags get reused
        for _ in range(3):
            g_dp = grads(mine, dp=True)
    g_ref = grads(list(range(n_cams)), dp=False)   # the whole batch on this GPU
    report = {}
    for n in g_ref:
        scale = g_ref[n].abs().max().item() + 1e-20
        err = (g_dp[n] - g_ref[n]).abs()
        report[n] = (err.max().item() / scale, int((err > 1e-3 * scale).sum()))
    gathered = [None] * world
    dist.all_gather_object(gathered, report)
    if rank == 0:
        torch.save(gathered, out_path)
    dist.destroy_process_group()


CASES = [
    ("table", False, 2, None), ("split", False, 2, None), ("table", True, 2, None), ("split", True, 2, None),
    # the trainer's own pattern: colors = torch.cat([sh0, shN], 1) — the exchange must mark the LEAVES as reduced
    ("cat", False, 2, None), ("cat", True, 2, None),
    # 3 cameras on 2 ranks: shards of 2 and 1 cameras
    ("table", False, 3, None), ("split", False, 3, None),
    # packed + sparse_grad (config E's mode): sparse (ids, rows) exchange and its dense fallback
    ("table", False, 2, 10.0), ("table", False, 2, 0.0),
]


PACKED_CP_CASES = [("table", 2, 0.0, "p2p"), ("table", 2, 10.0, "nccl"), ("cat", 3, 0.0, "p2p")]


@pytest.mark.parametrize("sh_mode,n_cams,packed_sparse,peer_mode", PACKED_CP_CASES)
def test_two_gpu_packed_camera_parallel_matches_single_process_batch(tmp_path, sh_mode, n_cams, packed_sparse, peer_mode):
    """Packed + sparse_grad (config E's mode) inside camera_parallel: the packed colour stage scatters its masked
    cotangents into the dense [C,N,3] layout and takes the same exchange as the un-packed path (peer kernels or
    NCCL all-gather), so the SH table is never all-reduced; projection gradients go through the sparse exchange."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    out = str(tmp_path / "report.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, sh_mode, False, n_cams, packed_sparse, peer_mode), nprocs=2, join=True)
    for r, rep in enumerate(torch.load(out)):
        for n, (max_rel, n_bad) in rep.items():
            assert max_rel < 1e-3 and n_bad == 0, f"rank {r} grad {n}: max rel {max_rel:.2e}, rows outside 1e-3: {n_bad}"


PEER_CASES = [("table", 2, "p2p", False), ("split", 2, "p2p", False), ("cat", 2, "p2p", False), ("table", 3, "p2p", False),
              ("split", 3, "multicast", False), ("table", 2, "multicast", False),
              # overlapped: colour backward on a side stream next to projection backward + all-reduce
              ("table", 2, "p2p", True), ("split", 3, "multicast", True), ("cat", 2, "p2p", True)]


@pytest.mark.parametrize("sh_mode,n_cams,peer_mode,defer", PEER_CASES)
def test_two_gpu_peer_exchange_matches_single_process_batch(tmp_path, sh_mode, n_cams, peer_mode, defer):
    """The same parity bar with both exchanges done by this library's own kernels over NVLink peer memory
    (`PeerExchange`: published cotangents read in place by the colour backward, two-shot arena all-reduce;
    "multicast" uses the switch-side reduction when the fabric maps a multicast address, else it is the
    peer load/store kernel again)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    out = str(tmp_path / "report.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, sh_mode, defer, n_cams, None, peer_mode), nprocs=2, join=True)
    for r, rep in enumerate(torch.load(out)):
        for n, (max_rel, n_bad) in rep.items():
            assert max_rel < 1e-3 and n_bad == 0, f"rank {r} grad {n}: max rel {max_rel:.2e}, rows outside 1e-3: {n_bad}"


@pytest.mark.parametrize("sh_mode,defer,n_cams,packed_sparse", CASES)
def test_two_gpu_camera_parallel_matches_single_process_batch(tmp_path, sh_mode, defer, n_cams, packed_sparse):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    out = str(tmp_path / "report.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, sh_mode, defer, n_cams, packed_sparse), nprocs=2, join=True)
    reports = torch.load(out)
    for r, rep in enumerate(reports):
        for n, (max_rel, n_bad) in rep.items():
            # same kernels, same decisions: only the order of the fp32 sums differs between the two ways
            assert max_rel < 1e-3 and n_bad == 0, f"rank {r} grad {n}: max rel {max_rel:.2e}, rows outside 1e-3: {n_bad}"
