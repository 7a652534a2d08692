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


def _worker_exchange(rank, world, port, out_dir):
    """Identity behind `camera_parallel`: sum_r outer(B(dir_r), g_r) computed from all-gathered
    colour cotangents == all-reduce of the per-rank SH coefficient gradients."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import torch_ref as O
        from splat_one_b200.distributed import GradArena

        torch.manual_seed(0)
        Nn, K = 200, 16
        means, table = torch.randn(Nn, 3), torch.randn(Nn, K, 3) * 0.3
        campos = torch.randn(world, 3) * 3
        vis = torch.rand(world, Nn) > 0.3
        v = torch.randn(world, Nn, 3)

        def colours(t, c):
            col = O.spherical_harmonics(3, means - campos[c], t, vis[c])
            return torch.where(vis[c][:, None], torch.clamp_min(col + 0.5, 0.0), torch.zeros_like(col))

        # (a) what plain data parallelism does: local dense gradient, then all-reduce
        t = table.clone().requires_grad_()
        col = colours(t, rank)
        (col * v[rank]).sum().backward()
        other = torch.zeros(7, requires_grad=True)
        other.grad = torch.full((7,), float(rank + 1))
        arena = GradArena([other, t])
        arena.gather_from_params()
        dense = arena.views[1].clone()
        dist.all_reduce(dense)
        # (b) the exchange: all-gather the masked cotangents, evaluate every camera locally
        g_local = torch.where(col.detach() > 0, v[rank], torch.zeros_like(v[rank]))
        g_all = [torch.empty_like(g_local) for _ in range(world)]
        dist.all_gather(g_all, g_local)
        t2 = table.clone().requires_grad_()
        tot = sum((O.spherical_harmonics(3, means - campos[c], t2, None) * g_all[c]).sum() for c in range(world))
        tot.backward()
        torch.testing.assert_close(t2.grad, dense, rtol=1e-4, atol=1e-5)
        # the arena leaves an already-global segment alone and reduces the rest
        arena.views[1].copy_(t2.grad)
        arena.all_reduce(skip_ptrs={t.data_ptr()})
        assert torch.equal(arena.views[1], t2.grad)
        assert arena.views[0].tolist() == [float(sum(range(1, world + 1)))] * 7
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_colour_cotangent_exchange_equals_gradient_allreduce(tmp_path):
    mp.spawn(_worker_exchange, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)


def _async_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from splat_one_b200.distributed import GradArena

        P = [torch.zeros(5, 3, requires_grad=True), torch.zeros(5, requires_grad=True), torch.zeros(5, 2, 3, requires_grad=True)]
        for i, p in enumerate(P):
            p.grad = torch.full_like(p, float(rank + 1 + i))
        arena = GradArena(P)
        arena.gather_from_params()
        # async: one work handle per contiguous run of non-skipped segments (the middle parameter is skipped)
        works = arena.all_reduce(async_op=True, skip_ptrs={P[1].data_ptr()})
        assert isinstance(works, list) and len(works) == 2
        for w in works:
            w.wait()
        assert torch.all(arena.views[0] == 1 + 2) and torch.all(arena.views[2] == 3 + 4)
        assert torch.all(arena.views[1] == float(rank + 2))  # skipped: still the local value
        # sync call returns the last handle (or None) as before
        assert arena.all_reduce(skip_ptrs={p.data_ptr() for p in P}) is None
        # only_ptrs: just that parameter's segment (the split all-reduce of the overlapped peer exchange)
        before = [v.clone() for v in arena.views]
        arena.all_reduce(only_ptrs={P[1].data_ptr()})
        assert torch.equal(arena.views[0], before[0]) and torch.equal(arena.views[2], before[2])
        assert torch.all(arena.views[1] == 2 + 3)
        if rank == 0:
            open(os.path.join(out_dir, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_arena_async_all_reduce_returns_one_work_per_run(tmp_path):
    mp.spawn(_async_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok")


def _worker_sparse(rank, world, port, out_dir):
    """sparse_allreduce / allreduce_mixed_gradients: the (gaussian_ids, rows) exchange of packed mode
    with sparse gradients == the dense all-reduce, below and above the dense-fallback threshold, with
    ragged and empty shards."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from splat_one_b200.distributed import allreduce_mixed_gradients, sparse_allreduce

        N = 1000
        res = {}
        for case, nnz_of in (("sparse", lambda r: 37 + 11 * r), ("ragged_empty", lambda r: 0 if r == 1 else 55),
                             ("dense_fallback", lambda r: 400 + 50 * r)):
            g = torch.Generator().manual_seed(100 + rank)
            nnz = nnz_of(rank)
            ids = torch.randperm(N, generator=g)[:nnz].sort().values
            vq, vs = torch.randn(nnz, 4, generator=g), torch.randn(nnz, 3, generator=g)
            gq = torch.sparse_coo_tensor(ids[None], vq, size=(N, 4), is_coalesced=True)
            gs = torch.sparse_coo_tensor(ids[None], vs, size=(N, 3), is_coalesced=True)
            counts = []
            oq, os_ = sparse_allreduce([gq, gs], dense_threshold=0.4, counts_out=counts)
            assert counts == [nnz_of(r) for r in range(world)]
            assert oq.is_sparse == (case != "dense_fallback")
            dq, ds = gq.to_dense(), gs.to_dense()
            dist.all_reduce(dq)
            dist.all_reduce(ds)
            res[case] = ((oq.to_dense() if oq.is_sparse else oq) - dq).abs().max().item(), \
                        ((os_.to_dense() if os_.is_sparse else os_) - ds).abs().max().item()
        # mixed dense + sparse parameter gradients through one call
        g = torch.Generator().manual_seed(7 + rank)
        P = [torch.zeros(N, 3, requires_grad=True), torch.zeros(N, 4, requires_grad=True), torch.zeros(N, requires_grad=True)]
        ids = torch.randperm(N, generator=g)[:20].sort().values
        P[0].grad = torch.randn(N, 3, generator=g)
        P[1].grad = torch.sparse_coo_tensor(ids[None], torch.randn(20, 4, generator=g), size=(N, 4), is_coalesced=True)
        P[2].grad = torch.randn(N, generator=g)
        want = []
        for p in P:
            d = p.grad.to_dense().clone() if p.grad.is_sparse else p.grad.clone()
            dist.all_reduce(d)
            want.append(d)
        allreduce_mixed_gradients(P)
        res["mixed"] = tuple(((p.grad.to_dense() if p.grad.is_sparse else p.grad) - w).abs().max().item()
                             for p, w in zip(P, want))
        assert P[1].grad.is_sparse
        if rank == 0:
            torch.save(res, os.path.join(out_dir, "sparse.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sparse_gradient_exchange_equals_dense_allreduce(tmp_path):
    world = 2
    mp.spawn(_worker_sparse, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = torch.load(os.path.join(tmp_path, "sparse.pt"))
    for case, errs in res.items():
        assert max(errs) <= 1e-6, (case, errs)


def test_uneven_camera_shards_must_be_announced():
    from splat_one_b200.distributed import shard_cameras

    vm, Ks = torch.eye(4).repeat(5, 1, 1), torch.eye(3).repeat(5, 1, 1)
    with pytest.raises(ValueError):
        shard_cameras(vm, Ks, rank=0, world_size=2)
    a = shard_cameras(vm, Ks, rank=0, world_size=2, allow_uneven=True)[2].tolist()
    b = shard_cameras(vm, Ks, rank=1, world_size=2, allow_uneven=True)[2].tolist()
    assert a == [0, 2, 4] and b == [1, 3]
