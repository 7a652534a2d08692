"""ORACLE (test infrastructure only — never imported by splat_one_b200/) for SURVEY.md §8 f3, third item:
`DefaultStrategy._update_state`, /root/reference/submodules/gsplat/gsplat/strategy/default.py:203-262, restated
in plain PyTorch on explicit arrays (no `info` dict, no autograd objects).

Parity status: PINNED — tests/golden/strategy_state.npz holds outputs of the reference's own method, run in this
container by oracle/gen_golden_strategy.py; tests/test_oracle_next.py checks this restatement against it.

One deliberate difference, documented: with several cameras seeing one Gaussian the reference's
`state["radii"][gs_ids] = maximum(state["radii"][gs_ids], r)` (:257-261) is an indexed assignment with duplicate
indices, i.e. it keeps an unspecified one of the candidates ("Should be ideally using scatter max").  `exact_max=
True` (default) takes the true maximum, which is what the kernels compute; the golden vectors are generated so
that both agree (one camera, or equal radii across cameras)."""
import torch


def update_state(grad2d, count, radii_state, grads, radii, width, height, n_cameras, gaussian_ids=None):
    """In place.  Unpacked: grads [C,N,2], radii [C,N].  Packed: grads [nnz,2], radii [nnz], gaussian_ids [nnz].
    `radii_state` may be None (refine_scale2d_stop_iter == 0)."""
    g = grads.clone().float()
    g[..., 0] *= width / 2.0 * n_cameras                      # :225
    g[..., 1] *= height / 2.0 * n_cameras                     # :226
    if gaussian_ids is None:
        sel = radii > 0.0                                     # :249
        ids = torch.where(sel)[1]                             # :250
        g = g[sel]                                            # :251
        r = radii[sel]                                        # :252
    else:
        ids, r = gaussian_ids, radii                          # :244-246
    grad2d.index_add_(0, ids, g.norm(dim=-1))                 # :254
    count.index_add_(0, ids, torch.ones_like(ids, dtype=torch.float32))   # :255-257
    if radii_state is not None:
        radii_state.scatter_reduce_(0, ids, r.float() / float(max(width, height)), reduce="amax", include_self=True)
    return grad2d, count, radii_state
