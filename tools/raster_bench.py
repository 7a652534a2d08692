#!/usr/bin/env python
"""Developer micro-benchmark (not the contract bench): times the raster forward and
backward kernels alone on the config-B scene, for each B200SPLAT_TUNING_VARIANT given.

    python tools/raster_bench.py [variants...]      e.g.  python tools/raster_bench.py 0 1
(one variant: B200SPLAT_TUNING_VARIANT=<v> python tools/raster_bench.py <v>; several: one subprocess each)
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import splat_one_b200 as S  # noqa: E402
from splat_one_b200 import synthetic, wrapper  # noqa: E402

dev = torch.device("cuda:0")
N, W, H = int(os.environ.get("RB_N", 1_000_000)), int(os.environ.get("RB_W", 1920)), int(os.environ.get("RB_H", 1080))
scene = synthetic.pinhole_scene(N, W, H, seed=42)
P = {k: scene[k].to(dev) for k in ("means", "quats", "scales", "opacities", "sh", "viewmats", "Ks")}
with torch.no_grad():
    radii, m2, dep, con, _ = S.fully_fused_projection(P["means"], None, P["quats"], P["scales"], P["viewmats"], P["Ks"], W, H)
    col = wrapper.sh_view_colors(3, P["means"], P["viewmats"], P["sh"], radii)
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    tpg, ids, fl = S.isect_tiles(m2, radii, dep, 16, tw, th)
    offs = S.isect_offset_encode(ids, 1, tw, th)
op = P["opacities"][None].contiguous()
g = torch.Generator().manual_seed(1)
vc = torch.randn(1, H, W, 3, generator=g).to(dev)
va = torch.randn(1, H, W, 1, generator=g).to(dev)
print(f"V={int((radii > 0).sum())} I={fl.numel()}")


def run(reps=10):
    leaves = [t.clone().requires_grad_() for t in (m2, con, col, op)]
    tf, tb = [], []
    out = None
    for i in range(reps + 3):
        for t in leaves:
            t.grad = None
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        rc, ra = S.rasterize_to_pixels(*leaves, W, H, 16, offs, fl)
        e[1].record()
        torch.autograd.backward([rc, ra], [vc, va])
        e[2].record()
        torch.cuda.synchronize()
        if i >= 3:
            tf.append(e[0].elapsed_time(e[1]))
            tb.append(e[1].elapsed_time(e[2]))
        out = (rc.detach(), ra.detach(), [t.grad.clone() for t in leaves])
    return sorted(tf)[len(tf) // 2], sorted(tb)[len(tb) // 2], out


# The tuning variant is read ONCE, when libb200splat.so is loaded: several variants = one process each
# (this script re-invokes itself), built with B200SPLAT_TUNING=1 python -m splat_one_b200.build --force.
if len(sys.argv) > 2:
    import subprocess

    for v in sys.argv[1:]:
        env = dict(os.environ, B200SPLAT_TUNING_VARIANT="0" if v == "generic" else v)
        subprocess.run([sys.executable, os.path.abspath(__file__), v], env=env, check=False)
    sys.exit(0)

base = None
for v in (sys.argv[1:] or ["0"]):
    if v == "generic":
        wrapper._FORCE_GENERIC_RASTER = True
    else:
        wrapper._FORCE_GENERIC_RASTER = False
        assert os.environ.get("B200SPLAT_TUNING_VARIANT", "0") == v or v == "0", \
            "set B200SPLAT_TUNING_VARIANT before the library loads (or pass several variants)"
    f, b, out = run()
    msg = f"variant {v:8s} fwd(+pack) {f:.3f} ms   bwd(+zero-fill) {b:.3f} ms"
    if base is None:
        base = out
    else:
        d_img = (out[0] - base[0]).abs().max().item()
        d_g = max(((a - c).abs().max() / (c.abs().max() + 1e-12)).item() for a, c in zip(out[2], base[2]))
        msg += f"   vs first: max|dC|={d_img:.2e} max rel dgrad={d_g:.2e}"
    print(msg)
