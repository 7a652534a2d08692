"""Micro-benchmark of the gradient-exchange primitives at config-B sizes (1 M Gaussians, K = 16, one camera per
rank): the library collectives against this library's own peer kernels (csrc/peer.cu).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/peer_bench.py

Prints one JSON line on rank 0 (ms per call, max over ranks, back-to-back calls on one stream)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from splat_one_b200 import wrapper
    from splat_one_b200._lib import get_lib
    from splat_one_b200.distributed import PeerExchange

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    N, K = 1_000_000, 16
    n_arena = 11 * N  # means 3 + quats 4 + scales 3 + opacities 1 floats per Gaussian
    peer = PeerExchange(N, 1, arena_floats=n_arena, use_multicast=True)
    lib = get_lib()
    reps = int(os.environ.get("REPS", "50"))

    def timed(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return round(t.item(), 5)

    out = {"world": world, "multicast_available": bool(peer.multicast_base), "arena_MB": n_arena * 4 / 1e6}
    flat = torch.randn(n_arena, device=dev)
    out["nccl_allreduce_ms"] = timed(lambda: dist.all_reduce(flat))
    mc = peer.multicast_base
    if mc:
        out["peer_allreduce_multimem_ms"] = timed(lambda: peer.all_reduce_(peer.arena_off, n_arena))
    peer.multicast_base = 0
    out["peer_allreduce_p2p_ms"] = timed(lambda: peer.all_reduce_(peer.arena_off, n_arena))
    peer.multicast_base = mc
    # correctness of both variants on fresh data
    for label, base in (("multimem", mc), ("p2p", 0)):
        if (label == "multimem" and not mc) or int(os.environ.get("B200SPLAT_TUNING_VARIANT", "0")) >= 4000:
            continue
        peer.multicast_base = base
        src = torch.randn(n_arena, device=dev, generator=torch.Generator(device=dev).manual_seed(rank))
        ref = src.clone()
        dist.all_reduce(ref)
        peer.arena_view(n_arena).copy_(src)
        peer.all_reduce_(peer.arena_off, n_arena)
        out[f"allreduce_{label}_max_err"] = float((peer.arena_view(n_arena) - ref).abs().max())
    peer.multicast_base = mc
    if os.environ.get("AR_ONLY") == "1":
        if rank == 0:
            print(json.dumps(out))
        dist.destroy_process_group()
        return

    # colour cotangent exchange + colour backward
    g_local = torch.randn(1, N, 3, device=dev)
    g_all = torch.empty(world, N, 3, device=dev)
    out["nccl_allgather_ms"] = timed(lambda: dist.all_gather_into_tensor(g_all, g_local))
    means = torch.randn(N, 3, device=dev)
    coeffs = torch.randn(N, K, 3, device=dev)
    campos_all = torch.randn(world, 3, device=dev)
    colors = torch.rand(1, N, 3, device=dev)
    v_coeffs, v_means = torch.empty_like(coeffs), torch.empty_like(means)
    P = wrapper._ptr

    def local_bwd():
        wrapper.native("sh_colors_bwd", lib, dev, world, N, K, 3, 0, P(means), P(campos_all), P(coeffs), None, None,
                       P(g_all), P(v_coeffs), P(v_means), rank, rank + 1)

    out["sh_colors_bwd_local_W_cameras_ms"] = timed(local_bwd)
    out["nccl_allgather_plus_bwd_ms"] = timed(lambda: (dist.all_gather_into_tensor(g_all, g_local), local_bwd()))

    def peer_bwd():
        slot = peer.next_slot()
        wrapper.native("peer_publish_cotangents", lib, dev, 1, N, 1, peer.hdr, P(campos_all[rank:rank + 1].contiguous()),
                       P(colors), P(g_local), P(peer.slot_view(slot)))
        peer.barrier()
        wrapper.native("sh_colors_bwd_peer", lib, dev, world, N, K, 3, P(means), P(coeffs), peer.bases_dev,
                       4 * peer.slot_off[slot], 1, peer.hdr, P(v_coeffs), P(v_means), rank, rank + 1)

    out["peer_publish_barrier_bwd_ms"] = timed(peer_bwd)
    out["peer_barrier_ms"] = timed(peer.barrier)
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
