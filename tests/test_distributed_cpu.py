"""N > 1 path on CPU (gloo, world_size 2): camera sharding + ONE all-reduce of the flat
gradient arena must reproduce the single-process gradients of the whole camera batch
(SURVEY.md §8e).  The render function here is the oracle (test infrastructure); on the
GPU box bench.py runs the same helpers over NCCL with the CUDA path."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NAMES = ("means", "quats", "scales", "opacities", "sh")
W, H, N, C = 48, 32, 400, 4


def _scene():
    from splat_one_b200 import synthetic

    return synthetic.pinhole_scene(N, W, H, seed=7, sh_degree=1, n_cameras=C)


def _cotangents():
    g = torch.Generator().manual_seed(5)
    return torch.randn(C, H, W, 3, generator=g), torch.randn(C, H, W, 1, generator=g)


def _render(params, vm, Ks):
    from oracle import torch_ref as O

    return O.rasterization(*params, vm, Ks, W, H, sh_degree=1, packed=False)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from splat_one_b200.distributed import GradArena, allreduce_gradients, rasterization_dp, shard_cameras

        torch.set_num_threads(1)
        scene = _scene()
        params = [scene[k].clone().requires_grad_() for k in NAMES]
        vm, Ks, ids = shard_cameras(scene["viewmats"], scene["Ks"])
        assert ids.tolist() == list(range(rank, C, world))
        rc, ra, meta, ids2 = rasterization_dp(
            lambda m, q, s, o, c, v, k, w, h, **kw: _render([m, q, s, o, c], v, k),
            dict(means=params[0], quats=params[1], scales=params[2], opacities=params[3], colors=params[4]),
            scene["viewmats"], scene["Ks"], W, H)
        assert torch.equal(ids, ids2) and rc.shape == (len(ids), H, W, 3)
        vc, va = _cotangents()
        ((rc * vc[ids]).sum() + (ra * va[ids]).sum()).backward()
        arena = allreduce_gradients(params)
        assert isinstance(arena, GradArena) and arena.nbytes >= sum(p.numel() for p in params) * 4
        for p, v in zip(params, arena.views):
            assert p.grad.data_ptr() == v.data_ptr()  # grads now alias the arena
        if rank == 0:
            torch.save([p.grad.clone() for p in params], os.path.join(out_dir, "dp.pt"))
            torch.save((rc.detach(), ids), os.path.join(out_dir, "img0.pt"))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(600)
def test_camera_sharded_dp_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = torch.load(os.path.join(tmp_path, "dp.pt"))
    img0, ids0 = torch.load(os.path.join(tmp_path, "img0.pt"))
    scene = _scene()
    params = [scene[k].clone().requires_grad_() for k in NAMES]
    rc, ra, _ = _render(params, scene["viewmats"], scene["Ks"])
    vc, va = _cotangents()
    ref = torch.autograd.grad((rc * vc).sum() + (ra * va).sum(), params)
    # the rank-0 shard renders exactly the same images as the same cameras in the full batch
    torch.testing.assert_close(img0, rc.detach()[ids0], rtol=0, atol=0)
    for n, a, b in zip(NAMES, got, ref):
        scale = b.abs().max().item() + 1e-12
        assert (a - b).abs().max().item() <= 1e-4 * scale, f"{n}: dp-summed gradient differs from the batch gradient"


def test_grad_arena_layout_and_sparse_grads():
    from splat_one_b200.distributed import GradArena

    p = [torch.zeros(5, 3, requires_grad=True), torch.zeros(7, requires_grad=True), torch.zeros(5, 4, requires_grad=True)]
    arena = GradArena(p)
    assert arena.offsets == [0, 16, 24] and arena.flat.numel() == 44  # segments padded to 16 bytes
    p[0].grad = torch.sparse_coo_tensor(torch.tensor([[1, 3, 1]]), torch.ones(3, 3), size=(5, 3))
    p[1].grad = torch.arange(7.0)
    arena.gather_from_params()
    assert arena.views[0][1].tolist() == [2.0, 2.0, 2.0] and arena.views[0][3].tolist() == [1.0, 1.0, 1.0]
    assert arena.views[1].tolist() == list(range(7)) and arena.views[2].abs().sum() == 0
    assert arena.all_reduce() is None  # no process group: no-op
    arena.scatter_to_params()
    assert p[2].grad.data_ptr() == arena.views[2].data_ptr()
