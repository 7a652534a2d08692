#!/usr/bin/env python
"""Developer tool: device timeline of one camera-sharded data-parallel step on rank 0
(torchrun --nproc-per-node N tools/dp_trace.py): every device operation with duration and the
idle gap in front of it, to see what the exchange and the host cost at N ranks.
DP_EXCHANGE=peer (default: own kernels over NVLink peer memory) | nccl; DP_DEFER=1: overlapped order."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import splat_one_b200 as S  # noqa: E402
from splat_one_b200 import synthetic  # noqa: E402
from splat_one_b200.distributed import GradArena, PeerExchange, arena_layout, camera_parallel  # noqa: E402

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, W, H = 1_000_000, 1920, 1080
scene = synthetic.pinhole_scene(N, W, H, seed=42, n_cameras=world)
params = [scene[k].to(dev).requires_grad_() for k in ("means", "quats", "scales", "opacities", "sh")]
vm, Ks = scene["viewmats"][rank::world].to(dev), scene["Ks"][rank::world].to(dev)
g = torch.Generator().manual_seed(1000 + rank)
vc, va = torch.randn(1, H, W, 3, generator=g).to(dev), torch.randn(1, H, W, 1, generator=g).to(dev)
peer = PeerExchange(N, 1, arena_floats=arena_layout(params)[1]) if os.environ.get("DP_EXCHANGE", "peer") == "peer" else None
defer = os.environ.get("DP_DEFER", "0") == "1"
arena = GradArena(params, peer=peer)


def step():
    for p in params:
        p.grad = None
    rc, ra, _ = S.rasterization(*params, vm, Ks, W, H, sh_degree=3, packed=False)
    with arena.sink(), camera_parallel(peer=peer, defer=defer) as cp:
        torch.autograd.backward([rc, ra], [vc, va])
    if defer:
        cp.finish(arena)
    else:
        arena.gather_from_params()
        arena.all_reduce(skip_ptrs=cp.reduced_ptrs)


for _ in range(8):
    step()
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    step()
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print(f"world {world}, exchange {'peer' if peer is not None else 'nccl'}{' overlapped' if defer else ''}: "
          f"{e0.elapsed_time(e1) / 20:.3f} ms/step (no profiler)")
dist.barrier()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    n = len(evs) // 3
    last = evs[2 * n:]
    t0 = last[0].time_range.start
    prev_end, busy, gap = t0, 0.0, 0.0
    print(f"{'start_us':>9} {'gap_us':>7} {'dur_us':>8}  op")
    for e in last:
        s_, e_ = e.time_range.start, e.time_range.end
        gp = max(0.0, s_ - prev_end)
        if gp > 3.0 or (e_ - s_) > 15.0 or "nccl" in e.name.lower() or "peer" in e.name.lower():
            print(f"{s_ - t0:9.1f} {gp:7.1f} {e_ - s_:8.1f}  {e.name[:90]}")
        busy += e_ - s_
        gap += gp
        prev_end = max(prev_end, e_)
    print(f"ops {len(last)}  busy {busy:.1f} us  idle {gap:.1f} us  span {prev_end - t0:.1f} us")
dist.destroy_process_group()
