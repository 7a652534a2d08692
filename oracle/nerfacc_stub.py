"""ORACLE (test infrastructure) — restatement of the two `nerfacc` functions the reference's
pure-PyTorch compositing imports (`gsplat/cuda/_torch_impl.py:536`):

    from nerfacc import accumulate_along_rays, render_weight_from_alpha

nerfacc (github.com/nerfstudio-project/nerfacc) is a third-party dependency of the
reference that is NOT vendored under /root/reference and not pinned to a version (the
reference's examples/requirements.txt lists the git main branch, commented out); it is not
installed in this image.  Its documented algorithm for the flattened (`ray_indices`) layout:

  render_weight_from_alpha(alphas, ray_indices, n_rays) -> (weights, trans)
      trans_i  = prod over earlier samples j of the SAME ray of (1 - alpha_j)   (exclusive)
      weights_i = alpha_i * trans_i
  accumulate_along_rays(weights, values, ray_indices, n_rays) -> [n_rays, D]
      out[r] = sum over samples i of ray r of weights_i * values_i   (values=None -> D = 1, values = 1)

Samples of a ray are contiguous and ordered front to back.  `oracle/gen_golden.py` installs
this module as `sys.modules["nerfacc"]` so that the reference's own `accumulate` and
`_rasterize_to_pixels` run on the CPU and produce the committed golden vectors.
"""
import torch


def _exclusive_prod_segments(x: torch.Tensor, ray_indices: torch.Tensor) -> torch.Tensor:
    """Sequential (left to right) exclusive product inside each run of equal ray index."""
    out = torch.ones_like(x)
    if x.numel() == 0:
        return out
    first = torch.ones_like(ray_indices, dtype=torch.bool)
    first[1:] = ray_indices[1:] != ray_indices[:-1]
    starts = torch.nonzero(first).flatten()
    lens = torch.diff(torch.cat([starts, torch.tensor([x.numel()])]))
    # position of every sample inside its segment; walk the positions so each multiplication
    # happens in the same order a per-ray loop would do it (differentiable, no in-place ops)
    seg = torch.cumsum(first.long(), 0) - 1
    pos = torch.arange(x.numel()) - starts[seg]
    max_len = int(lens.max())
    cur = torch.ones(starts.numel(), dtype=x.dtype)
    pieces = []
    for k in range(max_len):
        sel = torch.nonzero(pos == k).flatten()
        s = seg[sel]
        pieces.append((sel, cur[s]))
        cur = cur.index_put((s,), cur[s] * x[sel])
    idx = torch.cat([p[0] for p in pieces])
    val = torch.cat([p[1] for p in pieces])
    return torch.zeros_like(x).index_put((idx,), val)


def render_weight_from_alpha(alphas, packed_info=None, ray_indices=None, n_rays=None, prefix_trans=None):
    assert packed_info is None and prefix_trans is None and ray_indices is not None
    trans = _exclusive_prod_segments(1.0 - alphas, ray_indices)
    return alphas * trans, trans


def accumulate_along_rays(weights, values=None, ray_indices=None, n_rays=None):
    assert ray_indices is not None and n_rays is not None
    src = weights[:, None] if values is None else weights[:, None] * values
    out = torch.zeros((n_rays, src.shape[-1]), dtype=src.dtype)
    return out.index_add(0, ray_indices, src)
