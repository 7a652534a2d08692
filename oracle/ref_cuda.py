"""ORACLE (test infrastructure): loader for the reference's own CUDA extension built by
oracle/build_ref.py into oracle/_ref/gsplat_ref_csrc.so (pybind11 entry points of
/root/reference/submodules/gsplat/gsplat/cuda/csrc/ext.cpp:11-56), and `reference_chain`: those
kernels chained the way G/rendering.py chains its operators.  Only tests/, tools/ and the labelled
`reference_cuda` leg of bench.py (outside its timed region, never on the product path) import this;
`load()` returns None when the library was not built (it needs /root/reference at build time, never
at run time)."""
import importlib.machinery
import importlib.util
import math
import os

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "gsplat_ref_csrc.so")
_mod = None


def available() -> bool:
    return os.path.exists(_SO)


def load():
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        return None
    import torch  # noqa: F401  (libtorch must be loaded before the extension)

    loader = importlib.machinery.ExtensionFileLoader("gsplat_ref_csrc", _SO)
    spec = importlib.util.spec_from_loader("gsplat_ref_csrc", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    _mod = mod
    return mod


def camera_model(mod, name: str):
    return getattr(mod.CameraModelType, name.upper())


def reference_chain(R, P, W, H, model, vc, va, packed=False, sparse_grad=False, sh_degree=3, tile=16):
    """Forward + backward of the reference.  P: dict of means/quats/scales/opacities/sh/viewmats/Ks
    (CUDA tensors).  Returns a dict with the rendered image / alpha, dense parameter gradients and the
    intermediates the classification needs."""
    import torch

    cm = camera_model(R, model)
    means, quats, scales, opac0, sh = P["means"], P["quats"], P["scales"], P["opacities"], P["sh"]
    vm, Ks = P["viewmats"], P["Ks"]
    C, N, K = vm.shape[0], means.shape[0], sh.shape[1]
    tw, th = math.ceil(W / tile), math.ceil(H / tile)
    campos = torch.inverse(vm)[:, :3, 3]
    if packed:  # rendering.py:297-331, 366-392 (packed branch)
        (_, cam_ids, g_ids, radii, m2d, dep, con, _) = R.fully_fused_projection_packed_fwd(
            means, None, quats, scales, vm, Ks, W, H, 0.3, 0.01, 1e10, 0.0, False, cm)
        dirs = means[g_ids] - campos[cam_ids]
        shs = sh[g_ids]
        opac = opac0[g_ids]
    else:
        radii, m2d, dep, con, _ = R.fully_fused_projection_fwd(means, None, quats, scales, vm, Ks, W, H, 0.3, 0.01, 1e10,
                                                               0.0, False, cm)
        cam_ids = g_ids = None
        dirs = means[None] - campos[:, None]
        shs = sh[None].expand(C, -1, -1, -1).contiguous()
        opac = opac0[None].repeat(C, 1)
    masks = radii > 0
    col = torch.clamp_min(R.compute_sh_fwd(sh_degree, dirs, shs, masks) + 0.5, 0.0)  # rendering.py:380-392
    tpg, ids, flat = R.isect_tiles(m2d, radii, dep, cam_ids, g_ids, C, tile, tw, th, True, True)
    offs = R.isect_offset_encode(ids, C, tw, th)
    rc, ra, last = R.rasterize_to_pixels_fwd(m2d, con, col, opac, None, None, W, H, tile, offs, flat)
    # backward (_RasterizeToPixels.backward :957-1028, clamp/add, _SphericalHarmonics.backward :1240,
    # _FullyFusedProjection(.Packed).backward :831-898 / :1100-1223)
    _, v_m2d, v_con, v_col, v_op = R.rasterize_to_pixels_bwd(m2d, con, col, opac, None, None, W, H, tile, offs, flat, ra,
                                                              last, vc, va, False)
    v_sh_col = torch.where(col > 0, v_col, torch.zeros_like(v_col))
    v_coeffs, v_dirs = R.compute_sh_bwd(K, sh_degree, dirs, shs, masks, v_sh_col, True)
    if packed:
        v_means, _, v_quats, v_scales, _ = R.fully_fused_projection_packed_bwd(
            means, None, quats, scales, vm, Ks, W, H, 0.3, cm, cam_ids, g_ids, con, None, v_m2d, torch.zeros_like(dep),
            v_con, None, False, sparse_grad)
        if sparse_grad:  # [nnz, .] value rows of the COO gradients (_wrapper.py:1163-1203)
            v_means = torch.zeros_like(means).index_add_(0, g_ids, v_means)
            v_quats = torch.zeros_like(quats).index_add_(0, g_ids, v_quats)
            v_scales = torch.zeros_like(scales).index_add_(0, g_ids, v_scales)
        v_means = v_means.index_add(0, g_ids, v_dirs)
        g_sh = torch.zeros_like(sh).index_add_(0, g_ids, v_coeffs)
        g_op = torch.zeros_like(opac0).index_add_(0, g_ids, v_op)
    else:
        v_means, _, v_quats, v_scales, _ = R.fully_fused_projection_bwd(
            means, None, quats, scales, vm, Ks, W, H, 0.3, cm, radii, con, None, v_m2d, torch.zeros_like(dep), v_con,
            None, False)
        v_means = v_means + v_dirs.sum(0)
        g_sh = v_coeffs.sum(0)
        g_op = v_op.sum(0)
    return dict(image=rc, alpha=ra, last_ids=last, radii=radii, means2d=m2d, depths=dep, conics=con, colors=col,
                opacities=opac, isect_ids=ids, flatten_ids=flat, offsets=offs, gaussian_ids=g_ids, tiles_per_gauss=tpg,
                grads=dict(means=v_means, quats=v_quats, scales=v_scales, opacities=g_op, sh=g_sh))
