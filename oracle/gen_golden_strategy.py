"""Generates tests/golden/strategy_state.npz by running the REFERENCE's own `DefaultStrategy._update_state`
(/root/reference/submodules/gsplat/gsplat/strategy/default.py:203-262) on the CPU in this container:

    python oracle/gen_golden_strategy.py

Two calls in a row on the same state (so accumulation and the running maximum are exercised), unpacked with
C = 2 cameras (radii equal across the cameras of a Gaussian wherever both see it, so that the reference's
duplicate-index assignment is well defined) and packed with one camera.  Test infrastructure only."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference/submodules/gsplat")
from gsplat.strategy import DefaultStrategy  # noqa: E402  (the reference)


def main():
    g = torch.Generator().manual_seed(11)
    N, C, W, H = 257, 2, 96, 64
    out = dict(N=N, C=C, width=W, height=H)
    for mode in ("unpacked", "packed"):
        strat = DefaultStrategy(verbose=False, refine_scale2d_stop_iter=4000)
        state = strat.initialize_state()
        params = {"means": torch.zeros(N, 3)}
        for call in range(2):
            if mode == "unpacked":
                base = torch.randint(0, 40, (N,), generator=g)
                vis = torch.rand(C, N, generator=g) > 0.3
                radii = (base[None] * vis).to(torch.int32)           # equal across cameras where visible
                grads = torch.randn(C, N, 2, generator=g) * 1e-3
                info = dict(width=W, height=H, n_cameras=C, radii=radii, gaussian_ids=None,
                            means2d=SimpleNamespace(grad=grads, absgrad=grads.abs()))
                strat._update_state(params, state, info, packed=False)
            else:
                ids = torch.randperm(N, generator=g)[:150].sort().values
                radii = torch.randint(1, 40, (150,), generator=g).to(torch.int32)
                grads = torch.randn(150, 2, generator=g) * 1e-3
                info = dict(width=W, height=H, n_cameras=1, radii=radii, gaussian_ids=ids,
                            means2d=SimpleNamespace(grad=grads, absgrad=grads.abs()))
                strat._update_state(params, state, info, packed=True)
                out[f"{mode}_ids{call}"] = ids.numpy()
            out[f"{mode}_radii{call}"] = radii.numpy()
            out[f"{mode}_grads{call}"] = grads.numpy()
            for k in ("grad2d", "count", "radii"):
                out[f"{mode}_{k}_after{call}"] = state[k].numpy().copy()
    path = os.path.join(ROOT, "tests", "golden", "strategy_state.npz")
    np.savez_compressed(path, source="gsplat/strategy/default.py::DefaultStrategy._update_state (reference, CPU)", **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
