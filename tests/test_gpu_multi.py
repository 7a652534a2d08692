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


def _worker(rank, world, port, out_path, split_sh, defer):
    import torch.distributed as dist

    import splat_one_b200 as S
    from splat_one_b200 import synthetic
    from splat_one_b200.distributed import GradArena, camera_parallel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    W, H, N = 320, 240, 40000
    scene = synthetic.pinhole_scene(N, W, H, seed=7, n_cameras=world)
    g = torch.Generator().manual_seed(3)
    vc_all = torch.randn(world, H, W, 3, generator=g)
    va_all = torch.randn(world, H, W, 1, generator=g)

    def grads(cam_ids, dp):
        vm, Ks = scene["viewmats"][cam_ids].to(dev), scene["Ks"][cam_ids].to(dev)
        vc, va = vc_all[cam_ids].to(dev), va_all[cam_ids].to(dev)
        if split_sh:
            names = ("means", "quats", "scales", "opacities", "sh0", "shN")
            raw = dict(means=scene["means"], quats=scene["quats"], scales=scene["scales"], opacities=scene["opacities"],
                       sh0=scene["sh"][:, :1].contiguous(), shN=scene["sh"][:, 1:].contiguous())
            P = [raw[k].to(dev).requires_grad_() for k in names]
            rc, ra, _ = S.rasterization(P[0], P[1], P[2], P[3], (P[4], P[5]), vm, Ks, W, H, sh_degree=3, packed=False)
        else:
            names = ("means", "quats", "scales", "opacities", "sh")
            P = [scene[k].to(dev).requires_grad_() for k in names]
            rc, ra, _ = S.rasterization(*P, vm, Ks, W, H, sh_degree=3, packed=False)
        if dp:
            arena = GradArena(P)
            with arena.sink(), camera_parallel(defer=defer) as cp:
                torch.autograd.backward([rc, ra], [vc, va])
            if defer:  # all-reduce overlapped with the colour backward (camera_parallel.finish)
                cp.finish(arena)
            else:
                arena.gather_from_params()
                arena.all_reduce(skip_ptrs=cp.reduced_ptrs)
            arena.scatter_to_params()
        else:
            torch.autograd.backward([rc, ra], [vc, va])
        return {n: p.grad.detach().clone() for n, p in zip(names, P)}

    g_dp = grads([rank], dp=True)
    g_ref = grads(list(range(world)), dp=False)   # the whole batch on this GPU
    report = {}
    for n in g_ref:
        scale = g_ref[n].abs().max().item() + 1e-20
        err = (g_dp[n] - g_ref[n]).abs()
        report[n] = (err.max().item() / scale, (err > 1e-3 * scale + 1e-3 * g_ref[n].abs()).float().mean().item())
    gathered = [None] * world
    dist.all_gather_object(gathered, report)
    if rank == 0:
        torch.save(gathered, out_path)
    dist.destroy_process_group()


@pytest.mark.parametrize("split_sh,defer", [(False, False), (True, False), (False, True), (True, True)])
def test_two_gpu_camera_parallel_matches_single_process_batch(tmp_path, split_sh, defer):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    out = str(tmp_path / "report.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, split_sh, defer), nprocs=2, join=True)
    reports = torch.load(out)
    for r, rep in enumerate(reports):
        for n, (max_rel, frac_bad) in rep.items():
            assert max_rel < 2e-2 and frac_bad < 2e-3, f"rank {r} grad {n}: max rel {max_rel:.2e}, outside 1e-3: {frac_bad:.2e}"
