#!/usr/bin/env python
"""Developer benchmark: fwd+bwd step time of rasterization() on the BASELINE.json configs
other than the contract bench's (one GPU; config C = one rank's share)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import splat_one_b200 as S  # noqa: E402
from splat_one_b200 import synthetic, wrapper  # noqa: E402

dev = torch.device("cuda:0")
CONFIGS = {
    "B": dict(scene=lambda: synthetic.pinhole_scene(1_000_000, 1920, 1080, seed=42), W=1920, H=1080, kw={}),
    "B_packed": dict(scene=lambda: synthetic.pinhole_scene(1_000_000, 1920, 1080, seed=42), W=1920, H=1080,
                     kw=dict(packed=True)),
    "C_rank": dict(scene=lambda: synthetic.pinhole_scene(3_000_000, 1920, 1080, seed=43), W=1920, H=1080, kw={}),
    "D": dict(scene=lambda: synthetic.spherical_scene(2_000_000, 2048, 1024, seed=44), W=2048, H=1024,
              kw=dict(camera_model="spherical")),
    "E": dict(scene=lambda: synthetic.pinhole_scene(6_000_000, 3840, 2160, seed=45), W=3840, H=2160,
              kw=dict(packed=True, sparse_grad=True)),
}
for name in (sys.argv[1:] or list(CONFIGS)):
    cfg = CONFIGS[name]
    scene = synthetic.to_device(cfg["scene"](), dev)
    P = [scene[k].clone().requires_grad_() for k in ("means", "quats", "scales", "opacities", "sh")]
    W, H = cfg["W"], cfg["H"]
    kw = dict(packed=False)
    kw.update(cfg["kw"])
    g = torch.Generator().manual_seed(0)
    vc = torch.randn(1, H, W, 3, generator=g).to(dev)
    va = torch.randn(1, H, W, 1, generator=g).to(dev)

    def step():
        for p in P:
            p.grad = None
        rc, ra, meta = S.rasterization(*P, scene["viewmats"], scene["Ks"], W, H, sh_degree=3, **kw)
        torch.autograd.backward([rc, ra], [vc, va])
        return meta

    for _ in range(3):
        meta = step()
    torch.cuda.synchronize()
    wrapper.profiler.reset()
    wrapper.profiler.enabled = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        meta = step()
    e1.record()
    torch.cuda.synchronize()
    wrapper.profiler.enabled = False
    ms = e0.elapsed_time(e1) / n
    V = int((meta["radii"] > 0).sum())
    st = {k: round(v["avg_ms"] * v["calls"] / n, 3) for k, v in wrapper.profiler.summary_ms().items()}
    print(f"config {name}: {ms:.3f} ms/step  {W * H / ms / 1e3:.0f} Mpix/s  V={V} I={meta['flatten_ids'].numel()} "
          f"peak_mem={torch.cuda.max_memory_allocated() / 2**30:.2f} GiB  stages(ms/step)={st}")
    del P, scene, vc, va, meta
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
