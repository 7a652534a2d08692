"""Python operator API of the hot path — the drop-in for `gsplat/cuda/_wrapper.py`.

Same names, positional order, defaults, shape asserts, return arity and exception
types as the reference (G = /root/reference/submodules/gsplat/gsplat):

    spherical_harmonics      G/cuda/_wrapper.py:47-73     (+ _SphericalHarmonics :1226-1256)
    fully_fused_projection   G/cuda/_wrapper.py:203-339   (+ _FullyFusedProjection :775-898,
                                                            _FullyFusedProjectionPacked :1031-1223)
    isect_tiles              G/cuda/_wrapper.py:342-413
    isect_offset_encode      G/cuda/_wrapper.py:416-433
    rasterize_to_pixels      G/cuda/_wrapper.py:436-568   (+ _RasterizeToPixels :901-1028)
    rasterize_to_indices_in_range  G/cuda/_wrapper.py:571-643;  accumulate  G/cuda/_torch_impl.py:485-572
    quat_scale_to_covar_preci, proj, persp_proj, world_to_cam   G/cuda/_wrapper.py:76-200, 646-772

Every tensor (outputs, gradients, scan/sort workspaces) is allocated here with torch so
memory stays under the caching allocator and the current stream; the native library
owns nothing.  There is no CPU path: non-CUDA inputs raise RuntimeError like the
reference's TORCH_CHECK (CS/bindings.h:10-16).
"""
from __future__ import annotations

import ctypes
import threading
import os
from typing import Optional, Tuple

import torch
from torch import Tensor
from typing_extensions import Literal

from ._lib import check, get_lib

CAMERA_MODELS = {"pinhole": 0, "ortho": 1, "fisheye": 2, "spherical": 3}  # CS/bindings.h:34-40

# compiled channel instances of the raster kernels (csrc/raster_common.cuh pick_cdim)
_MAX_NATIVE_CHANNELS = 33
# debugging / A-B switch: B200SPLAT_GENERIC_RASTER=1 disables the warp-per-tile raster kernels
_FORCE_GENERIC_RASTER = os.environ.get("B200SPLAT_GENERIC_RASTER", "0") == "1"
# same for the depth-first intersection ordering (falls back to fill + full 64-bit key sort)
_FORCE_GENERIC_SORT = os.environ.get("B200SPLAT_GENERIC_SORT", "0") == "1"


class _Profiler:
    """Optional per-native-call CUDA-event timing + kernel-launch counting (bench.py).
    Disabled by default: `native()` is then a plain call."""

    # kernels launched by this library per native call (csrc/*.cu), memset nodes not counted;
    # the sorts are own onesweep passes (csrc/sort.cu): depth order = keys + 4 passes + scan,
    # tile order = expand + 2 passes (<= 16 key bits) + offsets, generic sort = histogram + 6 passes
    KERNELS = {
        "projection_fwd": 1, "projection_bwd": 1, "projection_packed_count": 3, "projection_packed_fill": 1,
        "projection_packed_bwd": 1, "sh_fwd": 1, "sh_bwd": 1, "camera_centers": 1, "sh_colors_fwd": 1,
        "sh_colors_bwd": 1, "sh_colors_packed_fwd": 1, "sh_colors_packed_bwd": 1, "sh_colors_packed_split_fwd": 1,
        "sh_colors_packed_split_bwd": 1, "isect_count": 1, "isect_fill": 1,
        "isect_sort": 7, "isect_sorted": 10, "isect_depth_order": 6, "isect_tile_order": 4, "invert_4x4": 1, "copy_small": 1, "isect_offset_encode": 1, "rasterize_pack": 1, "rasterize_fwd": 1, "rasterize_bwd": 1,
        "raster_indices_count": 2, "raster_indices_fill": 1, "quat_scale_to_covar_preci_fwd": 1,
        "quat_scale_to_covar_preci_bwd": 1, "world_to_cam_fwd": 1, "world_to_cam_bwd": 1, "proj_fwd": 1, "proj_bwd": 1,
        "selective_adam_update": 1, "compute_relocation": 1, "sh_colors_staged_fwd": 1, "sh_colors_staged_bwd": 1,
        "splat_activations_fwd": 1, "splat_activations_bwd": 1, "l1_ssim_fwd": 2, "l1_ssim_bwd": 1,
        "projection_bwd_state": 1, "strategy_update_state": 1, "peer_publish_cotangents": 1, "peer_barrier": 1, "peer_allreduce_f32": 1, "sh_colors_bwd_peer": 1,
        "sh_colors_staged_bwd_peer": 1,
    }

    def __init__(self):
        self.enabled = False
        self.only = None  # None: time every native call; else the set of names to bracket with events
        self.reset()

    def reset(self):
        self.events = {}
        self.calls = {}

    def launches(self) -> int:
        return sum(self.KERNELS.get(k, 0) * n for k, n in self.calls.items())

    def summary_ms(self):
        out = {}
        for name, evs in self.events.items():
            ts = [a.elapsed_time(b) for a, b in evs]
            out[name] = {"calls": len(ts), "avg_ms": sum(ts) / max(len(ts), 1), "total_ms": sum(ts)}
        return out


profiler = _Profiler()


def native(name: str, lib, device, *args):
    """Call `lib.b200splat_<name>(*args, stream)` on the current stream of `device`."""
    fn = getattr(lib, "b200splat_" + name)
    with torch.cuda.device(device):
        st = _stream(device)
        if profiler.enabled:
            profiler.calls[name] = profiler.calls.get(name, 0) + 1
            if profiler.only is None or name in profiler.only:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                rc = fn(*args, st)
                b.record()
                profiler.events.setdefault(name, []).append((a, b))
            else:
                rc = fn(*args, st)
        else:
            rc = fn(*args, st)
    check(rc, lib)


def _ptr(t: Optional[Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device: torch.device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


_PINNED = threading.local()


_SIDE_STREAMS: dict = {}  # process-wide (autograd engine threads and the caller must see the same streams)


def _side_stream(device, priority: int = 0) -> "torch.cuda.Stream":
    key = (torch.device(device), priority)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=device, priority=priority)
    return st


_RING = 256  # pinned read-back slots per thread (16 bytes each), used round-robin
_LEGACY_READBACK = os.environ.get("B200SPLAT_LEGACY_READBACK", "0") == "1"  # A/B switch: blocking .tolist()


def _start_read_back(t: Tensor, ready: Optional["torch.cuda.Event"], device):
    """Queue the read-back of a tiny device tensor (<= 16 bytes) and return a callable that waits
    for it and returns the values.  A one-warp kernel stores the words straight into a slot of a
    per-thread pinned (device-mapped) host buffer and the host waits on an event recorded behind
    it: no copy engine is involved, so the read-back cannot queue behind an unrelated large
    cudaMemcpyAsync of the application, and kernels queued behind the event keep running while
    the host waits.  (`ready` is kept for callers that recorded their own event; unused.)"""
    if _LEGACY_READBACK:
        return lambda: t.tolist()
    st = getattr(_PINNED, "state", None)
    if st is None:
        ring = torch.zeros((_RING, 2), dtype=torch.int64, pin_memory=True)
        st = _PINNED.state = {"ring": ring, "i64": ring.numpy(), "i32": ring.view(torch.int32).numpy(), "next": 0,
                              "events": [None] * _RING}
    n = t.numel()
    n_bytes = n * t.element_size()
    assert t.dtype in (torch.int64, torch.int32) and n_bytes <= 16 and t.is_contiguous(), (t.dtype, t.shape)
    k = st["next"] % _RING
    st["next"] += 1
    lib = get_lib()
    stream = torch.cuda.current_stream(device)
    with torch.cuda.device(device):
        check(lib.b200splat_copy_small(_ptr(t), ctypes.c_void_p(st["ring"].data_ptr() + 16 * k), n_bytes // 4,
                                       ctypes.c_void_p(stream.cuda_stream)), lib)
        if profiler.enabled:
            profiler.calls["copy_small"] = profiler.calls.get("copy_small", 0) + 1
        done = st["events"][k]
        if done is None:
            done = st["events"][k] = torch.cuda.Event()
        done.record(stream)
    view = st["i64"] if t.dtype == torch.int64 else st["i32"]

    def wait():
        done.synchronize()
        return view[k, :n].tolist()

    return wait


def _read_back(t: Tensor, ready: Optional["torch.cuda.Event"] = None):
    return _start_read_back(t, ready, t.device)()


def _check_cuda(*tensors: Optional[Tensor]) -> None:
    """GSPLAT_CHECK_INPUT: CUDA + contiguous + fp32/int dtype (CS/bindings.h:10-16)."""
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("b200splat: tensor must be a CUDA tensor (no CPU fallback exists)")
        if not t.is_contiguous():
            raise RuntimeError("b200splat: tensor must be contiguous")


def _f32(t: Optional[Tensor]) -> Optional[Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise RuntimeError(f"b200splat: expected float32, got {t.dtype}")
    return t


# ----------------------------------------------------------------------------------------
# gradient sink: let the backward kernels write parameter gradients in place
# ----------------------------------------------------------------------------------------
_GRAD_SINK: dict = {}  # param.data_ptr() -> destination buffer (global: autograd runs backward on its own threads)


class gradient_sink:
    """Context manager.  While active, parameter gradients that a backward kernel fully
    OVERWRITES (means / quats / scales / covars of the unpacked projection, the SH table of the
    fused colour stage) are written straight into the given buffers — e.g. the views of the
    flat all-reduce arena of `splat_one_b200.distributed.GradArena` — instead of into fresh
    allocations, so no gather copy is needed before the collective.  `pairs`: iterable of
    (parameter, destination) with equal shapes; destinations must be contiguous fp32."""

    def __init__(self, pairs):
        self.pairs = [(p, d) for p, d in pairs]

    def __enter__(self):
        for p, d in self.pairs:
            assert p.shape == d.shape and d.is_contiguous() and d.dtype == torch.float32, (p.shape, d.shape)
            _GRAD_SINK[p.data_ptr()] = d
        return self

    def __exit__(self, *exc):
        for p, _ in self.pairs:
            _GRAD_SINK.pop(p.data_ptr(), None)
        return False


# densification-state sink (set by splat_one_b200.strategy.strategy_state_sink): while active, the unpacked
# projection backward also updates the running grad2d / count / radii statistics of the reference's
# DefaultStrategy.  Process-wide for the same reason as _GRAD_SINK.
_STRATEGY_SINK: dict = {}


# camera-parallel mode (set by splat_one_b200.distributed.camera_parallel): process group + the
# data_ptr()s of parameters whose gradient came out of the backward already summed over ranks.
# Like _GRAD_SINK this is process-wide, not thread-local, on purpose: autograd runs the backward
# nodes on its own engine threads, which do not see the caller's thread-local state.  One training
# loop per process (one process per GPU) is the supported use; the viewer thread of splat_one only
# runs no-grad forwards and never reads either registry.
_CAMERA_PARALLEL: dict = {}


def _leaf_sources(t: Tensor):
    """The leaf tensors a gradient handed to `t` ends up in, when that is known to be a plain
    pass-through: `t` itself if it is a leaf, or the operands of a `torch.cat` of leaves (the
    `torch.cat([sh0, shN], 1)` of gsplat_trainer.py:474).  None otherwise."""
    if t.is_leaf:
        return [t]
    fn = t.grad_fn
    if fn is not None and type(fn).__name__.startswith("CatBackward"):
        leaves = []
        for nxt, _ in fn.next_functions:
            v = getattr(nxt, "variable", None)
            if v is None:
                return None
            leaves.append(v)
        return leaves
    return None


def _grad_out(like: Tensor) -> Tensor:
    """Destination for a gradient the kernel fully overwrites: the registered sink buffer of
    this parameter if there is one, else a fresh allocation."""
    d = _GRAD_SINK.get(like.data_ptr()) if _GRAD_SINK else None
    if d is not None and d.shape == like.shape and d.device == like.device:
        # a fresh alias of the buffer: autograd adopts a gradient without cloning it only when
        # nothing else references the tensor object it is handed
        return d.detach()
    return torch.empty_like(like)


def _aligned8(t: Optional[Tensor]) -> Optional[Tensor]:
    """means2d arrays are accessed as float2 by the kernels (8-byte aligned base, checked at the C ABI): a
    contiguous view that starts at an odd float offset is re-packed."""
    if t is None or t.data_ptr() % 8 == 0:
        return t
    return t.clone(memory_format=torch.contiguous_format)


def _aligned16(t: Optional[Tensor]) -> Optional[Tensor]:
    """128-bit loads need 16-byte aligned bases; views into odd offsets are re-packed."""
    if t is None or t.data_ptr() % 16 == 0:
        return t
    return t.clone(memory_format=torch.contiguous_format)


# ----------------------------------------------------------------------------------------
# spherical harmonics (a5)
# ----------------------------------------------------------------------------------------
def spherical_harmonics(
    degrees_to_use: int,
    dirs: Tensor,  # [..., 3]
    coeffs: Tensor,  # [..., K, 3]
    masks: Optional[Tensor] = None,
) -> Tensor:
    """Computes spherical harmonics (G/cuda/_wrapper.py:47-73).

    Returns colours [..., 3].  Masked-out elements are zero (the reference leaves them
    uninitialised, CS/compute_sh_fwd.cu:55).
    """
    assert (degrees_to_use + 1) ** 2 <= coeffs.shape[-2], coeffs.shape
    assert dirs.shape[:-1] == coeffs.shape[:-2], (dirs.shape, coeffs.shape)
    assert dirs.shape[-1] == 3, dirs.shape
    assert coeffs.shape[-1] == 3, coeffs.shape
    if masks is not None:
        assert masks.shape == dirs.shape[:-1], masks.shape
        masks = masks.contiguous()
    # An `expand`ed [C,N,K,3] view of a [N,K,3] table (rendering.py:386) is passed through
    # as the table itself: the kernel indexes it modulo N instead of materialising C copies.
    table = _unexpanded_table(coeffs)
    if table is not None:
        return _SphericalHarmonics.apply(degrees_to_use, dirs.contiguous(), table.contiguous(), masks, True)
    return _SphericalHarmonics.apply(degrees_to_use, dirs.contiguous(), coeffs.contiguous(), masks, False)


def spherical_harmonics_table(degrees_to_use: int, dirs: Tensor, table: Tensor,
                              masks: Optional[Tensor] = None) -> Tensor:
    """SH colours of C view-direction sets [C,N,3] against ONE coefficient table [N,K,3].

    Same result as `spherical_harmonics(deg, dirs, table.expand(C, -1, -1, -1), masks)` —
    what rendering.py:386-390 computes — without the C-fold copy the reference's
    `.contiguous()` makes (G/cuda/_wrapper.py:71-73) and with the coefficient gradient
    reduced over cameras in the backward."""
    assert dirs.dim() == 3 and table.dim() == 3, (dirs.shape, table.shape)
    assert (degrees_to_use + 1) ** 2 <= table.shape[-2], table.shape
    assert dirs.shape[1] == table.shape[0] and dirs.shape[-1] == 3 and table.shape[-1] == 3
    if masks is not None:
        assert masks.shape == dirs.shape[:-1], masks.shape
        masks = masks.contiguous()
    return _SphericalHarmonics.apply(degrees_to_use, dirs.contiguous(), table.contiguous(), masks, True)


def _unexpanded_table(coeffs: Tensor) -> Optional[Tensor]:
    """If `coeffs` [C,N,K,3] is a stride-0 broadcast of a [N,K,3] tensor, return that view."""
    if coeffs.dim() == 4 and coeffs.shape[0] > 1 and coeffs.stride(0) == 0:
        return coeffs[0]
    return None


class _SphericalHarmonics(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sh_degree: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor], broadcast: bool):
        _check_cuda(dirs, coeffs, masks)
        _f32(dirs), _f32(coeffs)
        lib = get_lib()
        K = coeffs.shape[-2]
        n_elems = dirs.numel() // 3
        n_rows = coeffs.numel() // (K * 3)
        colors = torch.zeros_like(dirs)
        if masks is not None and masks.dtype != torch.bool:
            masks = masks != 0
        if n_elems:
            native("sh_fwd", lib, dirs.device, n_elems, n_rows, K, sh_degree, _ptr(dirs), _ptr(coeffs), _ptr(masks),
                                           _ptr(colors))
        ctx.save_for_backward(dirs, coeffs, masks)
        ctx.sh_degree = sh_degree
        ctx.num_bases = K
        ctx.broadcast = broadcast
        return colors

    @staticmethod
    def backward(ctx, v_colors: Tensor):
        dirs, coeffs, masks = ctx.saved_tensors
        lib = get_lib()
        K = ctx.num_bases
        n_elems = dirs.numel() // 3
        n_rows = coeffs.numel() // (K * 3)
        compute_v_dirs = ctx.needs_input_grad[1]
        v_colors = v_colors.contiguous()
        v_coeffs = torch.empty(dirs.shape[:-1] + (K, 3), device=dirs.device, dtype=dirs.dtype)
        v_dirs = torch.empty_like(dirs) if compute_v_dirs else None
        if n_elems:
            native("sh_bwd", lib, dirs.device, n_elems, n_rows, K, ctx.sh_degree, _ptr(dirs), _ptr(coeffs), _ptr(masks),
                                           _ptr(v_colors), _ptr(v_coeffs), _ptr(v_dirs))
        if ctx.broadcast:
            v_coeffs = v_coeffs[0] if v_coeffs.shape[0] == 1 else v_coeffs.sum(dim=0)
        if not ctx.needs_input_grad[2]:
            v_coeffs = None
        return None, v_dirs, v_coeffs, None, None


def camera_centers(viewmats: Tensor) -> Tensor:
    """`torch.inverse(viewmats)[:, :3, 3]` (G/rendering.py:370, 382) as one tiny kernel
    (general 4x4 inverse, evaluated in double); no gradient."""
    viewmats = viewmats.detach().contiguous()
    _check_cuda(viewmats)
    _f32(viewmats)
    lib = get_lib()
    out = torch.empty((viewmats.shape[0], 3), device=viewmats.device, dtype=torch.float32)
    if viewmats.shape[0]:
        native("camera_centers", lib, viewmats.device, viewmats.shape[0], _ptr(viewmats), _ptr(out))
    return out


def _camera_parallel_colour_bwd(cp, means, campos, colors, v_colors, C, outs, v_means, run, run_peer=None):
    """The colour-cotangent exchange of the camera-parallel mode.  Peer mode (`camera_parallel(peer=
    PeerExchange)`): the masked cotangents are published into this rank's symmetric block, a flag
    barrier orders the ranks, and the kernel (`run_peer`) reads all W blocks in place over NVLink.
    Immediate mode: all-gather, then
    the kernel (`run`) sums the coefficient gradient over ALL cameras and the direction gradient over
    this rank's cameras; returns v_means.  Deferred mode (`camera_parallel(defer=True)`): the
    all-gather is only STARTED here (async) and the kernel launch is handed to
    `camera_parallel.finish()`, which queues it behind the arena all-reduce so that the two overlap;
    the direction gradient is then summed over all cameras on every rank (already global: it is added
    after the all-reduce) and this node returns no means gradient."""
    import torch.distributed as dist

    W, r = dist.get_world_size(cp), dist.get_rank(cp)
    N = means.shape[0]
    # every rank contributes a block of Cm camera slots.  Equal shards (the default, checked by
    # distributed.shard_cameras) have Cm == C; with `n_cameras_global` given to camera_parallel, shards
    # may differ by one camera (ids rank, rank + W, ...): short ranks pad with zero cotangents, which
    # add nothing to the sums.
    Cg = _CAMERA_PARALLEL.get("n_cameras_global")
    Cm = C if Cg is None else (Cg + W - 1) // W
    assert Cg is None or C == len(range(r, Cg, W)), (C, Cg, r, W)
    deferred = _CAMERA_PARALLEL.get("deferred")
    if deferred is not None:
        # deferred mode returns gradient tensors that are only filled later: that is only sound
        # when they ARE the arena views (gradient sink active) and nothing has to be added to them
        sink_ok = all(o is None or any(o.data_ptr() == d.data_ptr() for d in _GRAD_SINK.values()) for o in outs)
        fresh = all(lf.grad is None for lf in _CAMERA_PARALLEL.get("leaves", ()))
        if not (sink_ok and fresh):
            deferred = None
    peer = _CAMERA_PARALLEL.get("peer")
    if peer is not None and run_peer is not None and peer.fits(N, Cm) and peer.group is cp:
        slot = peer.next_slot()
        native("peer_publish_cotangents", get_lib(), means.device, C, N, Cm, peer.hdr, _ptr(campos.contiguous()),
               _ptr(colors), _ptr(v_colors), _ptr(peer.slot_view(slot)))  # colors None: v_colors is pre-masked
        peer.barrier()
        if deferred is None:
            run_peer(peer.bases_dev, 4 * peer.slot_off[slot], Cm, peer.hdr, outs, v_means, r * Cm, r * Cm + C, W * Cm)
            return v_means
        # overlapped: the colour kernel (NVLink reads of the peers' cotangents + the SH gradient write) runs on a
        # side stream next to the projection backward and the arena all-reduce; its direction gradient covers
        # THIS rank's cameras and is added to the means segment in camera_parallel.finish(), ahead of that
        # segment's all-reduce
        main, side = torch.cuda.current_stream(means.device), _side_stream(means.device)
        alias = tuple(None if o is None else o.detach() for o in outs)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            vm = torch.empty_like(means) if v_means is not None else None
            run_peer(peer.bases_dev, 4 * peer.slot_off[slot], Cm, peer.hdr, alias, vm, r * Cm, r * Cm + C, W * Cm)
            done = torch.cuda.Event()
            done.record(side)

        def join(all_cameras: bool, _keep=(colors, v_colors)):
            cur = torch.cuda.current_stream(means.device)
            cur.wait_event(done)
            if vm is not None:
                vm.record_stream(cur)
            return vm

        join.local = True  # the returned direction gradient is this rank's share, not the global sum
        deferred.append((means, join))
        return None
    # [C,N,3], zero where invisible or clamped (colors None: the caller masked it already)
    g_local = v_colors if colors is None else torch.where(colors > 0, v_colors, torch.zeros_like(v_colors))
    campos_c = campos.contiguous()
    if Cm != C:
        g_local = torch.cat([g_local, g_local.new_zeros((Cm - C, N, 3))])
        campos_c = torch.cat([campos_c, campos_c.new_zeros((Cm - C, 3))])
    g_all = torch.empty((W * Cm, N, 3), device=means.device, dtype=torch.float32)
    campos_all = torch.empty((W * Cm, 3), device=means.device, dtype=torch.float32)
    if deferred is None:
        dist.all_gather_into_tensor(g_all, g_local, group=cp)
        dist.all_gather_into_tensor(campos_all, campos_c, group=cp)
        run(campos_all, g_all, outs, v_means, r * Cm, r * Cm + C, W * Cm)
        return v_means
    works = [dist.all_gather_into_tensor(g_all, g_local, group=cp, async_op=True),
             dist.all_gather_into_tensor(campos_all, campos_c, group=cp, async_op=True)]
    # the kernel writes through fresh aliases: autograd adopts the returned gradient tensors without
    # a copy only while nothing else holds the same tensor object
    alias = tuple(None if o is None else o.detach() for o in outs)
    want_means = v_means is not None

    def finish(all_cameras: bool, _keep=(g_local, campos_c)):  # _keep: inputs of the in-flight all-gathers
        for w in works:
            w.wait()
        vm = torch.empty_like(means) if want_means else None
        lo, hi = (0, W * Cm) if all_cameras else (r * Cm, r * Cm + C)
        run(campos_all, g_all, alias, vm, lo, hi, W * Cm)
        return vm

    deferred.append((means, finish))
    return None


def sh_view_colors(sh_degree: int, means: Tensor, viewmats: Tensor, coeffs: Tensor, radii: Tensor) -> Tensor:
    """The colour stage of `rasterization()` for the unpacked layout, fused:

        dirs = means[None] - inverse(viewmats)[:, None, :3, 3]
        colors = clamp_min(spherical_harmonics(deg, dirs, coeffs, masks=radii > 0) + 0.5, 0)

    (G/rendering.py:368-392).  means [N,3], coeffs [N,K,3] or [C,N,K,3], radii [C,N] int32 ->
    colors [C,N,3].  Differentiable w.r.t. means and coeffs (not viewmats: callers that
    optimise poses use the unfused operators)."""
    C, N = radii.shape
    assert means.shape == (N, 3), means.shape
    assert coeffs.shape[-1] == 3 and (coeffs.shape[:-2] == (N,) or coeffs.shape[:-2] == (C, N)), coeffs.shape
    assert (sh_degree + 1) ** 2 <= coeffs.shape[-2], coeffs.shape
    campos = camera_centers(viewmats)
    return _ShViewColors.apply(sh_degree, means.contiguous(), campos, coeffs.contiguous(), radii.contiguous())


class _ShViewColors(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sh_degree, means, campos, coeffs, radii):
        _check_cuda(means, campos, coeffs, radii)
        _f32(means), _f32(coeffs)
        lib = get_lib()
        C, N = radii.shape
        K = coeffs.shape[-2]
        per_view = int(coeffs.dim() == 4)
        colors = torch.empty((C, N, 3), device=means.device, dtype=torch.float32)
        if C * N:
            native("sh_colors_fwd", lib, means.device, C, N, K, sh_degree, per_view, _ptr(means), _ptr(campos),
                   _ptr(coeffs), _ptr(radii), _ptr(colors))
        ctx.save_for_backward(means, campos, coeffs, radii, colors)
        ctx.sh_degree = sh_degree
        # where the coefficient gradient ends up (camera-parallel exchange: only when that is a known
        # set of leaves can the gradient be declared "already summed over ranks")
        ctx.coeff_leaves = _leaf_sources(coeffs)
        return colors

    @staticmethod
    def backward(ctx, v_colors):
        means, campos, coeffs, radii, colors = ctx.saved_tensors
        lib = get_lib()
        C, N = radii.shape
        K = coeffs.shape[-2]
        per_view = int(coeffs.dim() == 4)
        v_coeffs = _grad_out(coeffs)
        v_means = torch.empty_like(means) if ctx.needs_input_grad[1] else None
        v_colors = v_colors.contiguous()
        cp = _CAMERA_PARALLEL.get("group", None) if _CAMERA_PARALLEL else None
        if cp is not None and not per_view and N and ctx.coeff_leaves is not None:
            # camera-parallel exchange (splat_one_b200/distributed.py): the coefficient gradient
            # of camera c is the outer product B(dir_c) x v_rgb_c, so ranks all-gather their masked
            # colour cotangents (3 floats per Gaussian and camera) and every rank evaluates the
            # sum over ALL cameras itself, instead of all-reducing 3K floats per Gaussian
            def run(campos_all, g_all, v_out, v_means_out, lo, hi, WC):
                native("sh_colors_bwd", lib, means.device, WC, N, K, ctx.sh_degree, 0, _ptr(means), _ptr(campos_all),
                       _ptr(coeffs), None, None, _ptr(g_all), _ptr(v_out[0]), _ptr(v_means_out), lo, hi)

            def run_peer(bases, off_bytes, cams_per_block, hdr, v_out, v_means_out, lo, hi, WC):
                native("sh_colors_bwd_peer", lib, means.device, WC, N, K, ctx.sh_degree, _ptr(means), _ptr(coeffs),
                       bases, off_bytes, cams_per_block, hdr, _ptr(v_out[0]), _ptr(v_means_out), lo, hi)

            _CAMERA_PARALLEL["leaves"] = ctx.coeff_leaves
            v_means = _camera_parallel_colour_bwd(cp, means, campos, colors, v_colors, C, (v_coeffs,), v_means, run,
                                                  run_peer)
            # the gradient is global: mark the LEAVES it flows into (the table itself, or sh0 / shN behind
            # a torch.cat), never the data_ptr of a temporary
            for lf in ctx.coeff_leaves:
                _CAMERA_PARALLEL["reduced"].add(lf.data_ptr())
        elif N:
            native("sh_colors_bwd", lib, means.device, C, N, K, ctx.sh_degree, per_view, _ptr(means), _ptr(campos),
                   _ptr(coeffs), _ptr(radii), _ptr(colors), _ptr(v_colors), _ptr(v_coeffs),
                   _ptr(v_means), 0, C)
        return None, v_means, None, (v_coeffs if ctx.needs_input_grad[3] else None), None


def sh_view_colors_split(sh_degree: int, means: Tensor, viewmats: Tensor, sh0: Optional[Tensor], rest: Tensor,
                         radii: Tensor) -> Tensor:
    """`sh_view_colors` for ONE shared table given as the two tensors splat_one optimises,
    `sh0 [N,1,3]` and `rest = shN [N,K-1,3]` — i.e. `sh_view_colors(deg, means, viewmats,
    torch.cat([sh0, shN], 1), radii)` (R/utils/gsplat_utils/gsplat_trainer.py:474) without the
    concatenation and its split backward.  With `sh0=None`, `rest` is the whole [N,K,3] table.
    Coefficient rows are staged through shared memory (csrc/sh.cu, staged kernels)."""
    C, N = radii.shape
    assert means.shape == (N, 3), means.shape
    K = rest.shape[1] + (0 if sh0 is None else 1)
    assert rest.dim() == 3 and rest.shape[0] == N and rest.shape[2] == 3, rest.shape
    assert sh0 is None or sh0.shape == (N, 1, 3), sh0.shape
    assert (sh_degree + 1) ** 2 <= K, (sh_degree, K)
    campos = camera_centers(viewmats)
    return _ShViewColorsStaged.apply(sh_degree, means.contiguous(), campos,
                                     None if sh0 is None else sh0.contiguous(), rest.contiguous(), radii.contiguous())


def staged_colors_supported(K: int, split: bool) -> bool:
    """The staged kernels keep 128 coefficient rows per block in shared memory."""
    return get_lib().b200splat_sh_colors_staged_smem_bytes(K, int(split)) <= 200 * 1024


class _ShViewColorsStaged(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sh_degree, means, campos, sh0, rest, radii):
        _check_cuda(means, campos, sh0, rest, radii)
        _f32(means), _f32(sh0), _f32(rest)
        lib = get_lib()
        C, N = radii.shape
        K = rest.shape[1] + (0 if sh0 is None else 1)
        colors = torch.empty((C, N, 3), device=means.device, dtype=torch.float32)
        if C * N:
            native("sh_colors_staged_fwd", lib, means.device, C, N, K, sh_degree, _ptr(means), _ptr(campos), _ptr(sh0),
                   _ptr(rest), _ptr(radii), _ptr(colors))
        ctx.save_for_backward(means, campos, sh0, rest, radii, colors)
        ctx.sh_degree = sh_degree
        ctx.coeff_leaves = None
        if (sh0 is None or sh0.is_leaf) and rest.is_leaf:
            ctx.coeff_leaves = [rest] if sh0 is None else [sh0, rest]
        return colors

    @staticmethod
    def backward(ctx, v_colors):
        means, campos, sh0, rest, radii, colors = ctx.saved_tensors
        lib = get_lib()
        C, N = radii.shape
        K = rest.shape[1] + (0 if sh0 is None else 1)
        v_sh0 = _grad_out(sh0) if sh0 is not None else None
        v_rest = _grad_out(rest)
        v_means = torch.empty_like(means) if ctx.needs_input_grad[1] else None
        v_colors = v_colors.contiguous()
        cp = _CAMERA_PARALLEL.get("group", None) if _CAMERA_PARALLEL else None
        if cp is not None and N and ctx.coeff_leaves is not None:
            # camera-parallel exchange, as in _ShViewColors.backward
            outs = (v_sh0, v_rest) if sh0 is not None else (None, v_rest)

            def run(campos_all, g_all, v_out, v_means_out, lo, hi, WC):
                native("sh_colors_staged_bwd", lib, means.device, WC, N, K, ctx.sh_degree, _ptr(means),
                       _ptr(campos_all), _ptr(sh0), _ptr(rest), None, None, _ptr(g_all), _ptr(v_out[0]), _ptr(v_out[1]),
                       _ptr(v_means_out), lo, hi)

            def run_peer(bases, off_bytes, cams_per_block, hdr, v_out, v_means_out, lo, hi, WC):
                native("sh_colors_staged_bwd_peer", lib, means.device, WC, N, K, ctx.sh_degree, _ptr(means), _ptr(sh0),
                       _ptr(rest), bases, off_bytes, cams_per_block, hdr, _ptr(v_out[0]), _ptr(v_out[1]),
                       _ptr(v_means_out), lo, hi)

            _CAMERA_PARALLEL["leaves"] = ctx.coeff_leaves
            v_means = _camera_parallel_colour_bwd(cp, means, campos, colors, v_colors, C, outs, v_means, run, run_peer)
            for lf in ctx.coeff_leaves:
                _CAMERA_PARALLEL["reduced"].add(lf.data_ptr())
        elif N:
            native("sh_colors_staged_bwd", lib, means.device, C, N, K, ctx.sh_degree, _ptr(means), _ptr(campos),
                   _ptr(sh0), _ptr(rest), _ptr(radii), _ptr(colors), _ptr(v_colors), _ptr(v_sh0), _ptr(v_rest),
                   _ptr(v_means), 0, C)
        return (None, v_means, None, (v_sh0 if sh0 is not None and ctx.needs_input_grad[3] else None),
                (v_rest if ctx.needs_input_grad[4] else None), None)


def sh_view_colors_packed(sh_degree: int, means: Tensor, viewmats: Tensor, coeffs: Tensor, camera_ids: Tensor,
                          gaussian_ids: Tensor) -> Tensor:
    """The colour stage of `rasterization()` for the packed (COO) layout, fused:

        dirs = means[gaussian_ids] - inverse(viewmats)[camera_ids, :3, 3]
        shs = coeffs[gaussian_ids]            (or coeffs[camera_ids, gaussian_ids])
        colors = clamp_min(spherical_harmonics(deg, dirs, shs) + 0.5, 0)

    (G/rendering.py:370-392) -> colors [nnz,3], without the [nnz,K,3] gather copy and, in the
    backward, without the sort-based index_put of its gradient.  Differentiable w.r.t. means
    and coeffs."""
    C, N = viewmats.shape[0], means.shape[0]
    assert means.shape == (N, 3), means.shape
    assert coeffs.shape[-1] == 3 and (coeffs.shape[:-2] == (N,) or coeffs.shape[:-2] == (C, N)), coeffs.shape
    assert (sh_degree + 1) ** 2 <= coeffs.shape[-2], coeffs.shape
    assert camera_ids.shape == gaussian_ids.shape and camera_ids.dim() == 1
    campos = camera_centers(viewmats)
    return _ShViewColorsPacked.apply(sh_degree, means.contiguous(), campos, coeffs.contiguous(),
                                     camera_ids.contiguous(), gaussian_ids.contiguous())


class _ShViewColorsPacked(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sh_degree, means, campos, coeffs, camera_ids, gaussian_ids):
        _check_cuda(means, campos, coeffs, camera_ids, gaussian_ids)
        _f32(means), _f32(coeffs)
        if camera_ids.dtype != torch.int64 or gaussian_ids.dtype != torch.int64:
            raise RuntimeError("b200splat: camera_ids / gaussian_ids must be int64")
        lib = get_lib()
        C, N, nnz = campos.shape[0], means.shape[0], gaussian_ids.shape[0]
        K = coeffs.shape[-2]
        per_view = int(coeffs.dim() == 4)
        colors = torch.empty((nnz, 3), device=means.device, dtype=torch.float32)
        if nnz:
            native("sh_colors_packed_fwd", lib, means.device, nnz, C, N, K, sh_degree, per_view, _ptr(means),
                   _ptr(campos), _ptr(coeffs), _ptr(camera_ids), _ptr(gaussian_ids), _ptr(colors))
        ctx.save_for_backward(means, campos, coeffs, camera_ids, gaussian_ids, colors)
        ctx.sh_degree = sh_degree
        ctx.coeff_leaves = _leaf_sources(coeffs)
        return colors

    @staticmethod
    def backward(ctx, v_colors):
        means, campos, coeffs, camera_ids, gaussian_ids, colors = ctx.saved_tensors
        lib = get_lib()
        C, N, nnz = campos.shape[0], means.shape[0], gaussian_ids.shape[0]
        K = coeffs.shape[-2]
        per_view = int(coeffs.dim() == 4)
        cp = _CAMERA_PARALLEL.get("group", None) if _CAMERA_PARALLEL else None
        if cp is not None and not per_view and N and ctx.coeff_leaves is not None \
                and _CAMERA_PARALLEL.get("deferred") is None:
            # camera-parallel exchange for the packed layout (config E): the masked cotangents are scattered
            # into the dense [C,N,3] layout of the un-packed exchange (12 B per Gaussian and camera; zero =
            # invisible) and the same kernels sum the coefficient gradient over ALL ranks' cameras, instead of
            # all-reducing 3K floats per Gaussian (1.15 GB at 6 M Gaussians, K = 16)
            g = torch.where(colors > 0, v_colors, torch.zeros_like(v_colors))
            g_dense = torch.zeros((C, N, 3), device=means.device, dtype=torch.float32)
            if nnz:
                g_dense[camera_ids, gaussian_ids] = g
            v_coeffs = _grad_out(coeffs)
            v_means = torch.empty_like(means) if ctx.needs_input_grad[1] else None

            def run(campos_all, g_all, v_out, v_means_out, lo, hi, WC):
                native("sh_colors_bwd", lib, means.device, WC, N, K, ctx.sh_degree, 0, _ptr(means), _ptr(campos_all),
                       _ptr(coeffs), None, None, _ptr(g_all), _ptr(v_out[0]), _ptr(v_means_out), lo, hi)

            def run_peer(bases, off_bytes, cams_per_block, hdr, v_out, v_means_out, lo, hi, WC):
                native("sh_colors_bwd_peer", lib, means.device, WC, N, K, ctx.sh_degree, _ptr(means), _ptr(coeffs),
                       bases, off_bytes, cams_per_block, hdr, _ptr(v_out[0]), _ptr(v_means_out), lo, hi)

            _CAMERA_PARALLEL["leaves"] = ctx.coeff_leaves
            v_means = _camera_parallel_colour_bwd(cp, means, campos, None, g_dense, C, (v_coeffs,), v_means, run, run_peer)
            for lf in ctx.coeff_leaves:
                _CAMERA_PARALLEL["reduced"].add(lf.data_ptr())
            return None, v_means, None, (v_coeffs if ctx.needs_input_grad[3] else None), None, None
        v_coeffs = torch.zeros_like(coeffs)
        v_means = torch.zeros_like(means) if ctx.needs_input_grad[1] else None
        if nnz:
            native("sh_colors_packed_bwd", lib, means.device, nnz, C, N, K, ctx.sh_degree, per_view, _ptr(means),
                   _ptr(campos), _ptr(coeffs), _ptr(camera_ids), _ptr(gaussian_ids), _ptr(colors),
                   _ptr(v_colors.contiguous()), _ptr(v_coeffs), _ptr(v_means))
        return None, v_means, None, (v_coeffs if ctx.needs_input_grad[3] else None), None, None


def sh_view_colors_packed_split(sh_degree: int, means: Tensor, viewmats: Tensor, sh0: Tensor, shN: Tensor,
                                camera_ids: Tensor, gaussian_ids: Tensor) -> Tensor:
    """`sh_view_colors_packed(deg, means, viewmats, torch.cat([sh0, shN], 1), camera_ids, gaussian_ids)`
    (R/utils/gsplat_utils/gsplat_trainer.py:474 + G/rendering.py:370-392) without the concatenation
    and its split backward: the COO colour kernels read `sh0 [N,1,3]` / `shN [N,K-1,3]` in place."""
    N = means.shape[0]
    assert sh0.shape == (N, 1, 3) and shN.dim() == 3 and shN.shape[0] == N and shN.shape[2] == 3, (sh0.shape, shN.shape)
    assert (sh_degree + 1) ** 2 <= 1 + shN.shape[1], (sh_degree, shN.shape)
    campos = camera_centers(viewmats)
    return _ShViewColorsPackedSplit.apply(sh_degree, means.contiguous(), campos, sh0.contiguous(), shN.contiguous(),
                                          camera_ids.contiguous(), gaussian_ids.contiguous())


class _ShViewColorsPackedSplit(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sh_degree, means, campos, sh0, shN, camera_ids, gaussian_ids):
        _check_cuda(means, campos, sh0, shN, camera_ids, gaussian_ids)
        _f32(means), _f32(sh0), _f32(shN)
        if camera_ids.dtype != torch.int64 or gaussian_ids.dtype != torch.int64:
            raise RuntimeError("b200splat: camera_ids / gaussian_ids must be int64")
        lib = get_lib()
        C, N, nnz = campos.shape[0], means.shape[0], gaussian_ids.shape[0]
        K = 1 + shN.shape[1]
        colors = torch.empty((nnz, 3), device=means.device, dtype=torch.float32)
        if nnz:
            native("sh_colors_packed_split_fwd", lib, means.device, nnz, C, N, K, sh_degree, _ptr(means), _ptr(campos),
                   _ptr(sh0), _ptr(shN), _ptr(camera_ids), _ptr(gaussian_ids), _ptr(colors))
        ctx.save_for_backward(means, campos, sh0, shN, camera_ids, gaussian_ids, colors)
        ctx.sh_degree = sh_degree
        return colors

    @staticmethod
    def backward(ctx, v_colors):
        means, campos, sh0, shN, camera_ids, gaussian_ids, colors = ctx.saved_tensors
        lib = get_lib()
        C, N, nnz = campos.shape[0], means.shape[0], gaussian_ids.shape[0]
        K = 1 + shN.shape[1]
        v_sh0, v_shN = torch.zeros_like(sh0), torch.zeros_like(shN)
        v_means = torch.zeros_like(means) if ctx.needs_input_grad[1] else None
        if nnz:
            native("sh_colors_packed_split_bwd", lib, means.device, nnz, C, N, K, ctx.sh_degree, _ptr(means),
                   _ptr(campos), _ptr(sh0), _ptr(shN), _ptr(camera_ids), _ptr(gaussian_ids), _ptr(colors),
                   _ptr(v_colors.contiguous()), _ptr(v_sh0), _ptr(v_shN), _ptr(v_means))
        return (None, v_means, None, (v_sh0 if ctx.needs_input_grad[3] else None),
                (v_shN if ctx.needs_input_grad[4] else None), None, None)


class _GatherRows(torch.autograd.Function):
    """`x[ids]` along dim 0 (the packed-mode gathers of G/rendering.py:327-329, 366) with an
    atomic `index_add_` backward instead of autograd's sort-based index_put."""

    @staticmethod
    def forward(ctx, x, ids):
        ctx.save_for_backward(ids)
        ctx.n = x.shape[0]
        return x.index_select(0, ids)

    @staticmethod
    def backward(ctx, v):
        (ids,) = ctx.saved_tensors
        out = torch.zeros((ctx.n,) + v.shape[1:], device=v.device, dtype=v.dtype)
        out.index_add_(0, ids, v.contiguous())
        return out, None


def gather_rows(x: Tensor, ids: Tensor) -> Tensor:
    return _GatherRows.apply(x, ids)


# ----------------------------------------------------------------------------------------
# projection (a2, a3, a4)
# ----------------------------------------------------------------------------------------
def fully_fused_projection(
    means: Tensor,  # [N, 3]
    covars: Optional[Tensor],  # [N, 6] or None
    quats: Optional[Tensor],  # [N, 4] or None
    scales: Optional[Tensor],  # [N, 3] or None
    viewmats: Tensor,  # [C, 4, 4]
    Ks: Tensor,  # [C, 3, 3]
    width: int,
    height: int,
    eps2d: float = 0.3,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    packed: bool = False,
    sparse_grad: bool = False,
    calc_compensations: bool = False,
    camera_model: Literal["pinhole", "ortho", "fisheye", "spherical"] = "pinhole",
    _dense_means_grad: bool = False,
) -> Tuple[Tensor, ...]:
    """Projects Gaussians to 2D (G/cuda/_wrapper.py:203-339).

    `_dense_means_grad` (private, used by `rasterization()`): with sparse_grad, return the
    gradient of `means` dense.  When `means` also feeds the view-dependent colours, autograd
    has to add that dense gradient to this one anyway (and the reference's result is dense);
    adding dense + sparse costs ~13 ms at 6 M Gaussians, dense + dense does not.

    packed=False: (radii [C,N] int32, means2d [C,N,2], depths [C,N], conics [C,N,3],
    compensations [C,N] | None).  packed=True: (camera_ids, gaussian_ids [nnz] int64, radii,
    means2d, depths, conics, compensations) in (camera, gaussian) row-major order.
    """
    C = viewmats.size(0)
    N = means.size(0)
    assert means.size() == (N, 3), means.size()
    assert viewmats.size() == (C, 4, 4), viewmats.size()
    assert Ks.size() == (C, 3, 3), Ks.size()
    means = means.contiguous()
    if covars is not None:
        assert covars.size() == (N, 6), covars.size()
        covars = covars.contiguous()
    else:
        assert quats is not None, "covars or quats is required"
        assert scales is not None, "covars or scales is required"
        assert quats.size() == (N, 4), quats.size()
        assert scales.size() == (N, 3), scales.size()
        quats = quats.contiguous()
        scales = scales.contiguous()
    if sparse_grad:
        assert packed, "sparse_grad is only supported when packed is True"
    if camera_model not in CAMERA_MODELS:
        raise AttributeError(f"CameraModelType has no member {camera_model.upper()!r}")

    viewmats = viewmats.contiguous()
    Ks = Ks.contiguous()
    if packed:
        return _FullyFusedProjectionPacked.apply(
            means, covars, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane,
            radius_clip, sparse_grad, calc_compensations, camera_model, _dense_means_grad)
    return _FullyFusedProjection.apply(
        means, covars, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane,
        radius_clip, calc_compensations, camera_model)


class _FullyFusedProjection(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, covars, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane,
                radius_clip, calc_compensations, camera_model="pinhole"):
        _check_cuda(means, covars, quats, scales, viewmats, Ks)
        for t in (means, covars, quats, scales, viewmats, Ks):
            _f32(t)
        lib = get_lib()
        quats = _aligned16(quats)
        C, N = viewmats.shape[0], means.shape[0]
        dev = means.device
        # one allocation: radii | means2d | depths | conics (| compensations) as views of a flat buffer;
        # the kernel writes EVERY row (zeros for culled entries, which the reference leaves uninitialised,
        # CS/fully_fused_projection_fwd.cu:256-260), so no fill pass is needed
        n = C * N
        flat = torch.empty((n * (8 if calc_compensations else 7),), device=dev, dtype=torch.float32)
        # means2d first: it is read and written as float2 (8-byte aligned for every C*N, odd ones included)
        means2d = flat[:2 * n].view(C, N, 2)
        radii = flat[2 * n:3 * n].view(torch.int32).view(C, N)
        depths = flat[3 * n:4 * n].view(C, N)
        conics = flat[4 * n:7 * n].view(C, N, 3)
        compensations = flat[7 * n:8 * n].view(C, N) if calc_compensations else None
        if C and N:
            native("projection_fwd", lib, dev, C, N, _ptr(means), _ptr(covars), _ptr(quats), _ptr(scales), _ptr(viewmats), _ptr(Ks),
                    width, height, eps2d, near_plane, far_plane, radius_clip, CAMERA_MODELS[camera_model],
                    _ptr(radii), _ptr(means2d), _ptr(depths), _ptr(conics), _ptr(compensations))
        ctx.save_for_backward(means, covars, quats, scales, viewmats, Ks, radii, conics, compensations)
        ctx.width, ctx.height, ctx.eps2d = width, height, eps2d
        ctx.camera_model = CAMERA_MODELS[camera_model]
        ctx.mark_non_differentiable(radii)
        return radii, means2d, depths, conics, compensations

    @staticmethod
    def backward(ctx, v_radii, v_means2d, v_depths, v_conics, v_compensations):
        means, covars, quats, scales, viewmats, Ks, radii, conics, compensations = ctx.saved_tensors
        lib = get_lib()
        C, N = viewmats.shape[0], means.shape[0]
        dev = means.device
        if compensations is None:
            v_compensations = None
        elif v_compensations is not None:
            v_compensations = v_compensations.contiguous()
        # `means` also receives a gradient from the colour stage, so autograd sums two tensors and
        # its sink (if any) cannot be written here; the other three have a single producer
        v_means = torch.empty_like(means)
        v_covars = _grad_out(covars) if covars is not None else None
        v_quats = _grad_out(quats) if covars is None else None
        v_scales = _grad_out(scales) if covars is None else None
        v_viewmats = torch.zeros_like(viewmats) if ctx.needs_input_grad[4] else None
        sink = _STRATEGY_SINK.get("state") if _STRATEGY_SINK else None
        if N and sink is not None and sink.accepts(C, N, ctx.width, ctx.height):
            # DefaultStrategy._update_state folded into this kernel (splat_one_b200/strategy.py)
            sx, sy, inv = sink.scales(C)
            native("projection_bwd_state", lib, dev, C, N, _ptr(means), _ptr(covars), _ptr(quats), _ptr(scales),
                   _ptr(viewmats), _ptr(Ks), ctx.width, ctx.height, ctx.eps2d, ctx.camera_model, _ptr(radii),
                   _ptr(conics), _ptr(compensations), _ptr(v_means2d.contiguous()), _ptr(v_depths.contiguous()),
                   _ptr(v_conics.contiguous()), _ptr(v_compensations), _ptr(v_means), _ptr(v_covars), _ptr(v_quats),
                   _ptr(v_scales), _ptr(v_viewmats), sx, sy, inv, _ptr(sink.grad2d), _ptr(sink.count),
                   _ptr(sink.radii))
            sink.updates += 1
        elif N:
            native("projection_bwd", lib, dev, C, N, _ptr(means), _ptr(covars), _ptr(quats), _ptr(scales), _ptr(viewmats), _ptr(Ks),
                    ctx.width, ctx.height, ctx.eps2d, ctx.camera_model, _ptr(radii), _ptr(conics),
                    _ptr(compensations), _ptr(v_means2d.contiguous()), _ptr(v_depths.contiguous()),
                    _ptr(v_conics.contiguous()), _ptr(v_compensations), _ptr(v_means), _ptr(v_covars),
                    _ptr(v_quats), _ptr(v_scales), _ptr(v_viewmats))
        need = ctx.needs_input_grad
        return (v_means if need[0] else None, v_covars if need[1] else None, v_quats if need[2] else None,
                v_scales if need[3] else None, v_viewmats if need[4] else None,
                None, None, None, None, None, None, None, None, None, None)


class _FullyFusedProjectionPacked(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, covars, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane,
                radius_clip, sparse_grad, calc_compensations, camera_model="pinhole", dense_means_grad=False):
        _check_cuda(means, covars, quats, scales, viewmats, Ks)
        for t in (means, covars, quats, scales, viewmats, Ks):
            _f32(t)
        lib = get_lib()
        quats = _aligned16(quats)
        C, N = viewmats.shape[0], means.shape[0]
        dev = means.device
        cm = CAMERA_MODELS[camera_model]
        nnz = 0
        block_accum = None
        if C and N:
            block_accum = torch.empty((C * ((N + 255) // 256),), device=dev, dtype=torch.int32)
            nnz_dev = torch.empty((1,), device=dev, dtype=torch.int32)
            native("projection_packed_count", lib, dev, C, N, _ptr(means), _ptr(covars), _ptr(quats), _ptr(scales), _ptr(viewmats), _ptr(Ks),
                    width, height, eps2d, near_plane, far_plane, radius_clip, cm, _ptr(block_accum),
                    _ptr(nnz_dev))
            nnz = int(_read_back(nnz_dev)[0])  # the one host sync (CS/...packed_fwd.cu:352-353)
        indptr = torch.zeros((C + 1,), device=dev, dtype=torch.int32)
        camera_ids = torch.empty((nnz,), device=dev, dtype=torch.int64)
        gaussian_ids = torch.empty((nnz,), device=dev, dtype=torch.int64)
        radii = torch.empty((nnz,), device=dev, dtype=torch.int32)
        means2d = torch.empty((nnz, 2), device=dev, dtype=torch.float32)
        depths = torch.empty((nnz,), device=dev, dtype=torch.float32)
        conics = torch.empty((nnz, 3), device=dev, dtype=torch.float32)
        compensations = torch.zeros((nnz,), device=dev, dtype=torch.float32) if calc_compensations else None
        if nnz:
            native("projection_packed_fill", lib, dev, C, N, _ptr(means), _ptr(covars), _ptr(quats), _ptr(scales), _ptr(viewmats), _ptr(Ks),
                    width, height, eps2d, near_plane, far_plane, radius_clip, cm, _ptr(block_accum),
                    _ptr(indptr), _ptr(camera_ids), _ptr(gaussian_ids), _ptr(radii), _ptr(means2d), _ptr(depths),
                    _ptr(conics), _ptr(compensations))
        ctx.save_for_backward(camera_ids, gaussian_ids, means, covars, quats, scales, viewmats, Ks, conics,
                              compensations)
        ctx.width, ctx.height, ctx.eps2d = width, height, eps2d
        ctx.sparse_grad = sparse_grad
        ctx.dense_means_grad = dense_means_grad
        ctx.camera_model = cm
        ctx.indptr = indptr
        ctx.mark_non_differentiable(camera_ids, gaussian_ids, radii)
        return camera_ids, gaussian_ids, radii, means2d, depths, conics, compensations

    @staticmethod
    def backward(ctx, v_camera_ids, v_gaussian_ids, v_radii, v_means2d, v_depths, v_conics, v_compensations):
        (camera_ids, gaussian_ids, means, covars, quats, scales, viewmats, Ks, conics,
         compensations) = ctx.saved_tensors
        lib = get_lib()
        C, N, nnz = viewmats.shape[0], means.shape[0], camera_ids.shape[0]
        dev = means.device
        sparse_grad = ctx.sparse_grad
        if compensations is None:
            v_compensations = None
        elif v_compensations is not None:
            v_compensations = v_compensations.contiguous()
        rows = nnz if sparse_grad else N
        v_means = torch.zeros((rows, 3), device=dev, dtype=torch.float32)
        v_covars = torch.zeros((rows, 6), device=dev, dtype=torch.float32) if covars is not None else None
        v_quats = torch.zeros((rows, 4), device=dev, dtype=torch.float32) if covars is None else None
        v_scales = torch.zeros((rows, 3), device=dev, dtype=torch.float32) if covars is None else None
        v_viewmats = torch.zeros_like(viewmats) if ctx.needs_input_grad[4] else None
        if nnz:
            native("projection_packed_bwd", lib, dev, C, N, nnz, _ptr(means), _ptr(covars), _ptr(quats), _ptr(scales), _ptr(viewmats), _ptr(Ks),
                    ctx.width, ctx.height, ctx.eps2d, ctx.camera_model, _ptr(camera_ids), _ptr(gaussian_ids),
                    _ptr(conics), _ptr(compensations), _ptr(v_means2d.contiguous()), _ptr(v_depths.contiguous()),
                    _ptr(v_conics.contiguous()), _ptr(v_compensations), int(sparse_grad), _ptr(v_means),
                    _ptr(v_covars), _ptr(v_quats), _ptr(v_scales), _ptr(v_viewmats))
        need = ctx.needs_input_grad

        def _coo(values: Optional[Tensor], like: Optional[Tensor]):
            # G/cuda/_wrapper.py:1163-1203
            if values is None or not sparse_grad:
                return values
            return torch.sparse_coo_tensor(indices=gaussian_ids[None], values=values, size=like.size(),
                                           is_coalesced=len(viewmats) == 1, check_invariants=False)

        if sparse_grad and ctx.dense_means_grad and need[0]:
            g_means = torch.zeros_like(means).index_add_(0, gaussian_ids, v_means)
        else:
            g_means = _coo(v_means, means) if need[0] else None
        return (g_means,
                _coo(v_covars, covars) if need[1] else None,
                _coo(v_quats, quats) if need[2] else None,
                _coo(v_scales, scales) if need[3] else None,
                v_viewmats if need[4] else None,
                None, None, None, None, None, None, None, None, None, None, None)


# ----------------------------------------------------------------------------------------
# tile intersection (a6) and offset encode (a7)
# ----------------------------------------------------------------------------------------
@torch.no_grad()
def _isect_tiles_begin(
    means2d: Tensor,  # [C, N, 2] or [nnz, 2]
    radii: Tensor,  # [C, N] or [nnz]
    depths: Tensor,  # [C, N] or [nnz]
    tile_size: int,
    tile_width: int,
    tile_height: int,
    sort: bool = True,
    packed: bool = False,
    n_cameras: Optional[int] = None,
    camera_ids: Optional[Tensor] = None,
    gaussian_ids: Optional[Tensor] = None,
    want_offsets: bool = False,
    concurrent: bool = False,
):
    """isect_tiles in two halves: this call validates, allocates and queues everything that does
    not need `n_isects` (count, scan, depth order, the read-back) and returns `finish`, which waits
    for `n_isects`, allocates the outputs and queues the rest.  With `want_offsets` `finish()` also
    returns the [C, tile_height, tile_width] offsets of isect_offset_encode, which the depth-first
    path produces in its final pass for free.  `concurrent`: run the first half on a side stream
    (see below) so that work the caller queues between the two calls overlaps it."""
    if packed:
        nnz = means2d.size(0)
        assert means2d.shape == (nnz, 2), means2d.size()
        assert radii.shape == (nnz,), radii.size()
        assert depths.shape == (nnz,), depths.size()
        assert camera_ids is not None, "camera_ids is required if packed is True"
        assert gaussian_ids is not None, "gaussian_ids is required if packed is True"
        assert n_cameras is not None, "n_cameras is required if packed is True"
        camera_ids = camera_ids.contiguous()
        gaussian_ids = gaussian_ids.contiguous()
        C, N = n_cameras, 0
    else:
        C, N, _ = means2d.shape
        assert means2d.shape == (C, N, 2), means2d.size()
        assert radii.shape == (C, N), radii.size()
        assert depths.shape == (C, N), depths.size()
        nnz = 0

    means2d = _aligned8(means2d.contiguous())
    radii = radii.contiguous()
    depths = depths.contiguous()
    _check_cuda(means2d, radii, depths, camera_ids, gaussian_ids)
    if means2d.dtype != torch.float32 or depths.dtype != torch.float32:
        # the reference dispatches half/bf16/double and up-casts (CS/isect_tiles.cu:169-173)
        means2d, depths = means2d.float(), depths.float()
    if radii.dtype != torch.int32:
        raise RuntimeError(f"b200splat: radii must be int32, got {radii.dtype}")
    lib = get_lib()
    dev = means2d.device
    n_elems = nnz if packed else C * N
    n_tiles = tile_width * tile_height
    tile_n_bits = int(n_tiles).bit_length()  # == floor(log2(n_tiles)) + 1, CS/isect_tiles.cu:156
    cam_n_bits = int(C).bit_length()
    assert tile_n_bits + cam_n_bits <= 32, (tile_n_bits, cam_n_bits)

    tiles_per_gauss = torch.empty(radii.shape, device=dev, dtype=torch.int32)
    depth_first = sort and not _FORCE_GENERIC_SORT
    depth_ws, depth_sel = None, ctypes.c_int(0)
    cum_tiles = n_isects_dev = None
    side = ready = None
    if n_elems:
        n_isects_dev = torch.empty((2,), device=dev, dtype=torch.int64)
        ws, ws_bytes = None, 0
        if not depth_first:
            # generic path: running sums in input order for isect_fill (CS/isect_tiles.cu:200)
            cum_tiles = torch.empty((n_elems,), device=dev, dtype=torch.int64)
            ws_bytes = lib.b200splat_scan_workspace_bytes(n_elems)
            ws = torch.empty((ws_bytes,), device=dev, dtype=torch.uint8)
        dws_bytes = 0
        if depth_first:
            dws_bytes = lib.b200splat_isect_depth_order_workspace_bytes(n_elems)
            depth_ws = torch.empty((dws_bytes,), device=dev, dtype=torch.uint8)
        main = torch.cuda.current_stream(dev)
        if concurrent and depth_first:
            # everything up to the read-back runs on a side stream: the caller's next kernels on the
            # current stream (the colour stage) execute concurrently with the count + depth-order
            # phase.  All buffers were allocated above, BEFORE this event, so the side stream never
            # touches a block whose previous use on the current stream is still pending.
            side = _side_stream(dev)
            fork = torch.cuda.Event()
            fork.record(main)
            side.wait_event(fork)
        with torch.cuda.stream(side if side is not None else main):
            native("isect_count", lib, dev, int(packed), C, N, nnz, _ptr(means2d), _ptr(radii), _ptr(depths),
                   tile_size, tile_width, tile_height, _ptr(tiles_per_gauss), _ptr(cum_tiles), _ptr(n_isects_dev),
                   _ptr(ws), ws_bytes)
            # the read-back is queued right behind the count ...
            pending = _start_read_back(n_isects_dev, None, dev)
            if depth_first:
                # ... and phase 1 of the depth-first ordering, which needs only n_elems, behind it: the
                # host's round trip for n_isects overlaps ~0.1 ms of device work
                native("isect_depth_order", lib, dev, n_elems, _ptr(depths), _ptr(tiles_per_gauss), _ptr(depth_ws),
                       dws_bytes, ctypes.byref(depth_sel))
                if side is not None:
                    join = torch.cuda.Event()
                    join.record(side)
    else:
        pending = None
        ws = None
    # buffers touched by the first half stay referenced until finish() has joined it: with
    # `concurrent` they are used on the side stream, which the caching allocator does not track
    keep_alive = (ws, n_isects_dev, depth_ws)

    @torch.no_grad()
    def finish(_keep=keep_alive):
        n_isects, neg_depth = 0, False
        if pending is not None:
            n_isects, neg = pending()  # the one host sync (CS/isect_tiles.cu:201)
            neg_depth = bool(neg)
        if side is not None:
            torch.cuda.current_stream(dev).wait_event(join)
        return _isect_finish(lib, dev, packed, C, N, nnz, camera_ids, means2d, radii, depths, tiles_per_gauss, cum_tiles,
                             depth_ws, depth_sel.value, depth_first, neg_depth, n_isects, sort, tile_size, tile_width,
                             tile_height, tile_n_bits, cam_n_bits, want_offsets)

    return finish


def _isect_finish(lib, dev, packed, C, N, nnz, camera_ids, means2d, radii, depths, tiles_per_gauss, cum_tiles, depth_ws,
                  depth_sel, depth_first, neg_depth, n_isects, sort, tile_size, tile_width, tile_height, tile_n_bits,
                  cam_n_bits, want_offsets):
    isect_ids = torch.empty((n_isects,), device=dev, dtype=torch.int64)
    flatten_ids = torch.empty((n_isects,), device=dev, dtype=torch.int32)
    if n_isects:
        if depth_first and not neg_depth:
            # depth-first ordering (csrc/sort.cu): bit-identical to fill + full-key sort
            ws_bytes = lib.b200splat_isect_tile_order_workspace_bytes(n_isects)
            ws = torch.empty((ws_bytes,), device=dev, dtype=torch.uint8)
            offsets = (torch.empty((C, tile_height, tile_width), device=dev, dtype=torch.int32)
                       if want_offsets else None)
            native("isect_tile_order", lib, dev, int(packed), C, N, nnz, _ptr(camera_ids), _ptr(means2d), _ptr(radii),
                   _ptr(depths), _ptr(depth_ws), depth_sel, n_isects, tile_size, tile_width, tile_height,
                   _ptr(isect_ids), _ptr(flatten_ids), _ptr(offsets), _ptr(ws), ws_bytes)
            return tiles_per_gauss, isect_ids, flatten_ids, offsets
        if cum_tiles is None:
            # the depth-first first half produced only the total; inputs outside its contract (depths
            # with the sign bit set) fall back to fill + full-key sort, which needs the running sums
            n_elems = nnz if packed else C * N
            cum_tiles = torch.empty((n_elems,), device=dev, dtype=torch.int64)
            ws_bytes = lib.b200splat_scan_workspace_bytes(n_elems)
            ws = torch.empty((ws_bytes,), device=dev, dtype=torch.uint8)
            scratch = torch.empty((2,), device=dev, dtype=torch.int64)
            native("isect_count", lib, dev, int(packed), C, N, nnz, _ptr(means2d), _ptr(radii), _ptr(depths),
                   tile_size, tile_width, tile_height, _ptr(tiles_per_gauss), _ptr(cum_tiles), _ptr(scratch),
                   _ptr(ws), ws_bytes)
        native("isect_fill", lib, dev, int(packed), C, N, nnz, _ptr(camera_ids), _ptr(means2d), _ptr(radii),
               _ptr(depths), _ptr(cum_tiles), tile_size, tile_width, tile_height, _ptr(isect_ids), _ptr(flatten_ids))
        if sort:
            isect_ids_alt = torch.empty_like(isect_ids)
            flatten_ids_alt = torch.empty_like(flatten_ids)
            ws_bytes = lib.b200splat_sort_workspace_bytes(n_isects)
            ws = torch.empty((max(ws_bytes, 1),), device=dev, dtype=torch.uint8)
            selector = ctypes.c_int(0)
            native("isect_sort", lib, dev, n_isects, 32 + tile_n_bits + cam_n_bits, _ptr(isect_ids), _ptr(flatten_ids),
                   _ptr(isect_ids_alt), _ptr(flatten_ids_alt), _ptr(ws), ws_bytes, ctypes.byref(selector))
            if selector.value == 1:
                isect_ids, flatten_ids = isect_ids_alt, flatten_ids_alt
    return tiles_per_gauss, isect_ids, flatten_ids, None


@torch.no_grad()
def isect_tiles(
    means2d: Tensor,  # [C, N, 2] or [nnz, 2]
    radii: Tensor,  # [C, N] or [nnz]
    depths: Tensor,  # [C, N] or [nnz]
    tile_size: int,
    tile_width: int,
    tile_height: int,
    sort: bool = True,
    packed: bool = False,
    n_cameras: Optional[int] = None,
    camera_ids: Optional[Tensor] = None,
    gaussian_ids: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Tensor]:
    """Maps projected Gaussians to intersecting tiles (G/cuda/_wrapper.py:342-413).

    Returns (tiles_per_gauss int32 [C,N]|[nnz], isect_ids int64 [n_isects],
    flatten_ids int32 [n_isects]); bit-exact with the reference given the same inputs.
    """
    return _isect_tiles_begin(means2d, radii, depths, tile_size, tile_width, tile_height, sort, packed, n_cameras,
                              camera_ids, gaussian_ids)()[:3]


@torch.no_grad()
def isect_tiles_and_offsets(means2d, radii, depths, tile_size, tile_width, tile_height, packed=False,
                            n_cameras=None, camera_ids=None, gaussian_ids=None):
    """`isect_tiles(...)` followed by `isect_offset_encode(...)` (G/rendering.py:497-510) as one
    operator: returns (tiles_per_gauss, isect_ids, flatten_ids, isect_offsets)."""
    return isect_tiles_and_offsets_begin(means2d, radii, depths, tile_size, tile_width, tile_height, packed, n_cameras,
                                         camera_ids, gaussian_ids, concurrent=False)()


@torch.no_grad()
def isect_tiles_and_offsets_begin(means2d, radii, depths, tile_size, tile_width, tile_height, packed=False,
                                  n_cameras=None, camera_ids=None, gaussian_ids=None, concurrent=False):
    """First half of `isect_tiles_and_offsets`; returns a callable producing its result.  What
    `rasterization()` queues between the two halves (the colour stage) runs concurrently with the
    count / depth-order phase when `concurrent`."""
    fin = _isect_tiles_begin(means2d, radii, depths, tile_size, tile_width, tile_height, True, packed, n_cameras,
                             camera_ids, gaussian_ids, want_offsets=True, concurrent=concurrent)

    def finish():
        tpg, ids, flat, offs = fin()
        if offs is None:
            C = n_cameras if packed else means2d.shape[0]
            offs = isect_offset_encode(ids, C, tile_width, tile_height)
        return tpg, ids, flat, offs

    return finish


@torch.no_grad()
def isect_offset_encode(isect_ids: Tensor, n_cameras: int, tile_width: int, tile_height: int) -> Tensor:
    """Encodes intersection ids to offsets [C, tile_height, tile_width] int32
    (G/cuda/_wrapper.py:416-433)."""
    isect_ids = isect_ids.contiguous()
    _check_cuda(isect_ids)
    if isect_ids.dtype != torch.int64:
        raise RuntimeError(f"b200splat: isect_ids must be int64, got {isect_ids.dtype}")
    lib = get_lib()
    dev = isect_ids.device
    offsets = torch.empty((n_cameras, tile_height, tile_width), device=dev, dtype=torch.int32)
    if offsets.numel():
        native("isect_offset_encode", lib, dev, isect_ids.numel(), _ptr(isect_ids), n_cameras, tile_width,
                                                    tile_height, _ptr(offsets))
    return offsets


# ----------------------------------------------------------------------------------------
# rasterization (a8, a9)
# ----------------------------------------------------------------------------------------
def rasterize_to_pixels(
    means2d: Tensor,  # [C, N, 2] or [nnz, 2]
    conics: Tensor,  # [C, N, 3] or [nnz, 3]
    colors: Tensor,  # [C, N, channels] or [nnz, channels]
    opacities: Tensor,  # [C, N] or [nnz]
    image_width: int,
    image_height: int,
    tile_size: int,
    isect_offsets: Tensor,  # [C, tile_height, tile_width]
    flatten_ids: Tensor,  # [n_isects]
    backgrounds: Optional[Tensor] = None,  # [C, channels]
    masks: Optional[Tensor] = None,  # [C, tile_height, tile_width]
    packed: bool = False,
    absgrad: bool = False,
) -> Tuple[Tensor, Tensor]:
    """Rasterizes Gaussians to pixels (G/cuda/_wrapper.py:436-568).

    Returns (render_colors [C,H,W,channels], render_alphas [C,H,W,1]).
    """
    C = isect_offsets.size(0)
    if packed:
        nnz = means2d.size(0)
        assert means2d.shape == (nnz, 2), means2d.shape
        assert conics.shape == (nnz, 3), conics.shape
        assert colors.shape[0] == nnz, colors.shape
        assert opacities.shape == (nnz,), opacities.shape
    else:
        N = means2d.size(1)
        assert means2d.shape == (C, N, 2), means2d.shape
        assert conics.shape == (C, N, 3), conics.shape
        assert colors.shape[:2] == (C, N), colors.shape
        assert opacities.shape == (C, N), opacities.shape
    if backgrounds is not None:
        assert backgrounds.shape == (C, colors.shape[-1]), backgrounds.shape
        backgrounds = backgrounds.contiguous()
    if masks is not None:
        assert masks.shape == isect_offsets.shape, masks.shape
        masks = masks.contiguous()

    channels = colors.shape[-1]
    if channels > 513 or channels == 0:
        raise ValueError(f"Unsupported number of color channels: {channels}")

    tile_height, tile_width = isect_offsets.shape[1:3]
    assert (
        tile_height * tile_size >= image_height
    ), f"Assert Failed: {tile_height} * {tile_size} >= {image_height}"
    assert (
        tile_width * tile_size >= image_width
    ), f"Assert Failed: {tile_width} * {tile_size} >= {image_width}"

    def _one(colors_, backgrounds_):
        return _RasterizeToPixels.apply(
            means2d.contiguous(), conics.contiguous(), colors_.contiguous(), opacities.contiguous(),
            backgrounds_, masks, image_width, image_height, tile_size, isect_offsets.contiguous(),
            flatten_ids.contiguous(), absgrad)

    if channels <= _MAX_NATIVE_CHANNELS:
        # the kernels take the real channel count: no zero-padding copy as in :497-541
        return _one(colors, backgrounds)
    # 34..513 channels: render in native-width chunks (the reference pads to 64..513-wide
    # template instances instead); alphas are identical for every chunk.
    outs, alphas = [], None
    for s in range(0, channels, 32):
        rc, ra = _one(colors[..., s:s + 32], None if backgrounds is None else backgrounds[..., s:s + 32].contiguous())
        outs.append(rc)
        alphas = ra if alphas is None else alphas
    return torch.cat(outs, dim=-1), alphas


class _RasterizeToPixels(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means2d, conics, colors, opacities, backgrounds, masks, width, height, tile_size,
                isect_offsets, flatten_ids, absgrad):
        _check_cuda(means2d, conics, colors, opacities, backgrounds, masks, isect_offsets, flatten_ids)
        for t in (means2d, conics, colors, opacities, backgrounds):
            _f32(t)
        if isect_offsets.dtype != torch.int32 or flatten_ids.dtype != torch.int32:
            raise RuntimeError("b200splat: isect_offsets and flatten_ids must be int32")
        if masks is not None and masks.dtype != torch.bool:
            raise RuntimeError("b200splat: masks must be a bool tensor")
        lib = get_lib()
        dev = means2d.device
        C, tile_height, tile_width = isect_offsets.shape
        channels = colors.shape[-1]
        n_gauss = means2d.numel() // 2
        n_isects = flatten_ids.numel()
        means2d_a = _aligned16(means2d) if means2d.data_ptr() % 8 else means2d
        render_colors = torch.empty((C, height, width, channels), device=dev, dtype=torch.float32)
        render_alphas = torch.empty((C, height, width, 1), device=dev, dtype=torch.float32)
        last_ids = torch.empty((C, height, width), device=dev, dtype=torch.int32)
        if masks is not None:
            # masked tiles never write alpha/last_ids (CS/...fwd.cu:71-77); give them defined values
            render_alphas.zero_()
            last_ids.zero_()
        # warp-per-tile fast path (tile_size 16, <= 4 channels): one packed 48-byte record per
        # Gaussian, shared by the forward and the backward kernel
        records = quad_masks = None
        rec_bytes = 0 if _FORCE_GENERIC_RASTER else lib.b200splat_rasterize_records_bytes(n_gauss, channels, tile_size)
        if rec_bytes and n_isects:
            records = torch.empty((rec_bytes // 4,), device=dev, dtype=torch.float32)
            native("rasterize_pack", lib, dev, n_gauss, channels, _ptr(means2d_a), _ptr(conics), _ptr(colors),
                   _ptr(opacities), _ptr(records))
            if any(ctx.needs_input_grad[:5]):
                # one byte per list entry: the forward's quad culling decisions, reused by the backward
                quad_masks = torch.empty((n_isects,), device=dev, dtype=torch.uint8)
        if render_colors.numel():
            native("rasterize_fwd", lib, dev, C, n_gauss, n_isects, channels, _ptr(means2d_a), _ptr(conics), _ptr(colors), _ptr(opacities),
                    _ptr(backgrounds), _ptr(masks), width, height, tile_size, tile_width, tile_height,
                    _ptr(isect_offsets), _ptr(flatten_ids), _ptr(records), _ptr(quad_masks), _ptr(render_colors),
                    _ptr(render_alphas), _ptr(last_ids))
        ctx.save_for_backward(means2d, conics, colors, opacities, backgrounds, masks, isect_offsets, flatten_ids,
                              render_alphas, last_ids, records, quad_masks)
        ctx.width, ctx.height, ctx.tile_size, ctx.absgrad = width, height, tile_size, absgrad
        return render_colors, render_alphas

    @staticmethod
    def backward(ctx, v_render_colors: Tensor, v_render_alphas: Tensor):
        (means2d, conics, colors, opacities, backgrounds, masks, isect_offsets, flatten_ids, render_alphas,
         last_ids, records, quad_masks) = ctx.saved_tensors
        lib = get_lib()
        dev = means2d.device
        C, tile_height, tile_width = isect_offsets.shape
        channels = colors.shape[-1]
        n_gauss = means2d.numel() // 2
        n_isects = flatten_ids.numel()
        v_render_colors = v_render_colors.contiguous()
        v_render_alphas = v_render_alphas.contiguous()
        # one allocation, one fill for all accumulated gradients
        per = 2 + 3 + channels + 1 + (2 if ctx.absgrad else 0)
        flat = torch.zeros((n_gauss * per,), device=dev, dtype=torch.float32)
        o = 0
        v_means2d = flat[o:o + 2 * n_gauss].view_as(means2d); o += 2 * n_gauss
        v_conics = flat[o:o + 3 * n_gauss].view_as(conics); o += 3 * n_gauss
        v_colors = flat[o:o + channels * n_gauss].view_as(colors); o += channels * n_gauss
        v_opacities = flat[o:o + n_gauss].view_as(opacities); o += n_gauss
        v_means2d_abs = flat[o:o + 2 * n_gauss].view_as(means2d) if ctx.absgrad else None
        means2d_a = _aligned16(means2d) if means2d.data_ptr() % 8 else means2d
        if n_isects and render_alphas.numel():
            native("rasterize_bwd", lib, dev, C, n_gauss, n_isects, channels, _ptr(means2d_a), _ptr(conics), _ptr(colors), _ptr(opacities),
                    _ptr(backgrounds), _ptr(masks), ctx.width, ctx.height, ctx.tile_size, tile_width, tile_height,
                    _ptr(isect_offsets), _ptr(flatten_ids), _ptr(records), _ptr(quad_masks), _ptr(render_alphas),
                    _ptr(last_ids), _ptr(v_render_colors), _ptr(v_render_alphas), _ptr(v_means2d_abs), _ptr(v_means2d),
                    _ptr(v_conics), _ptr(v_colors), _ptr(v_opacities))
        if ctx.absgrad:
            means2d.absgrad = v_means2d_abs  # G/cuda/_wrapper.py:1005-1006
        v_backgrounds = None
        if ctx.needs_input_grad[4]:
            v_backgrounds = (v_render_colors * (1.0 - render_alphas).float()).sum(dim=(1, 2))
        return (v_means2d, v_conics, v_colors, v_opacities, v_backgrounds,
                None, None, None, None, None, None, None)


# ----------------------------------------------------------------------------------------
# rasterize_to_indices_in_range + accumulate
# ----------------------------------------------------------------------------------------
@torch.no_grad()
def rasterize_to_indices_in_range(
    range_start: int,
    range_end: int,
    transmittances: Tensor,  # [C, image_height, image_width]
    means2d: Tensor,  # [C, N, 2]
    conics: Tensor,  # [C, N, 3]
    opacities: Tensor,  # [C, N]
    image_width: int,
    image_height: int,
    tile_size: int,
    isect_offsets: Tensor,  # [C, tile_height, tile_width]
    flatten_ids: Tensor,  # [n_isects]
) -> Tuple[Tensor, Tensor, Tensor]:
    """Rasterizes a batch of Gaussians to images but only returns the indices
    (G/cuda/_wrapper.py:571-643; kernel CS/rasterize_to_indices_in_range.cu).

    `[range_start, range_end)` counts batches of tile_size² list entries per tile, front to
    back.  Returns (gaussian_ids, pixel_ids, camera_ids), int64 [M], grouped by pixel in
    (camera, row, column) order and front to back inside a pixel."""
    C, N, _ = means2d.shape
    assert conics.shape == (C, N, 3), conics.shape
    assert opacities.shape == (C, N), opacities.shape
    assert isect_offsets.shape[0] == C, isect_offsets.shape
    tile_height, tile_width = isect_offsets.shape[1:3]
    assert (
        tile_height * tile_size >= image_height
    ), f"Assert Failed: {tile_height} * {tile_size} >= {image_height}"
    assert (
        tile_width * tile_size >= image_width
    ), f"Assert Failed: {tile_width} * {tile_size} >= {image_width}"
    assert transmittances.shape == (C, image_height, image_width), transmittances.shape
    transmittances = transmittances.contiguous()
    means2d, conics, opacities = _aligned8(means2d.contiguous()), conics.contiguous(), opacities.contiguous()
    isect_offsets, flatten_ids = isect_offsets.contiguous(), flatten_ids.contiguous()
    _check_cuda(transmittances, means2d, conics, opacities, isect_offsets, flatten_ids)
    for t in (transmittances, means2d, conics, opacities):
        _f32(t)
    if isect_offsets.dtype != torch.int32 or flatten_ids.dtype != torch.int32:
        raise RuntimeError("b200splat: isect_offsets and flatten_ids must be int32")
    lib = get_lib()
    dev = means2d.device
    n_pix = C * image_height * image_width
    n_isects = flatten_ids.numel()
    # clamp the Python ints (callers pass e.g. 1e10 for "to the end") into uint32
    rs = int(min(max(range_start, 0), 0xFFFFFFFF))
    re_ = int(min(max(range_end, 0), 0xFFFFFFFF))
    n_elems = 0
    if n_pix and n_isects and N:
        cnts = torch.empty((n_pix,), device=dev, dtype=torch.int32)
        cum = torch.empty((n_pix,), device=dev, dtype=torch.int64)
        total = torch.empty((1,), device=dev, dtype=torch.int64)
        ws_bytes = lib.b200splat_scan_workspace_bytes(n_pix)
        ws = torch.empty((ws_bytes,), device=dev, dtype=torch.uint8)
        common = (rs, re_, C, N, n_isects, _ptr(means2d), _ptr(conics), _ptr(opacities), image_width, image_height,
                  tile_size, tile_width, tile_height, _ptr(isect_offsets), _ptr(flatten_ids), _ptr(transmittances))
        native("raster_indices_count", lib, dev, *common, _ptr(cnts), _ptr(cum), _ptr(total), _ptr(ws), ws_bytes)
        n_elems = int(total.item())  # the one host sync (CS/rasterize_to_indices_in_range.cu:263)
    gaussian_ids = torch.empty((n_elems,), device=dev, dtype=torch.int64)
    out_indices = torch.empty((n_elems,), device=dev, dtype=torch.int64)
    if n_elems:
        native("raster_indices_fill", lib, dev, *common, _ptr(cnts), _ptr(cum), _ptr(gaussian_ids), _ptr(out_indices))
    out_pixel_ids = out_indices % (image_width * image_height)
    out_camera_ids = out_indices // (image_width * image_height)
    return gaussian_ids, out_pixel_ids, out_camera_ids


def accumulate(
    means2d: Tensor,  # [C, N, 2]
    conics: Tensor,  # [C, N, 3]
    opacities: Tensor,  # [C, N]
    colors: Tensor,  # [C, N, channels]
    gaussian_ids: Tensor,  # [M]
    pixel_ids: Tensor,  # [M]
    camera_ids: Tensor,  # [M]
    image_width: int,
    image_height: int,
) -> Tuple[Tensor, Tensor]:
    """Alpha compositing of the listed (Gaussian, pixel, camera) intersections in plain
    PyTorch, differentiable by autograd (G/cuda/_torch_impl.py:485-572).

    The reference delegates the two segment operations to the third-party `nerfacc` package
    (`render_weight_from_alpha`, `accumulate_along_rays`; un-pinned git dependency, not
    vendored); here they are torch ops: transmittance = segmented exclusive product of
    (1 - alpha) over the entries of a pixel (float64 log-space prefix sums, so cancellation
    across millions of entries stays below fp32 resolution), accumulation = `index_add`.
    Like nerfacc, entries of one pixel must be contiguous and front to back — the order
    `rasterize_to_indices_in_range` returns.

    Returns (renders [C,H,W,channels], alphas [C,H,W,1])."""
    C, N = means2d.shape[:2]
    channels = colors.shape[-1]
    pixel_ids_x = pixel_ids % image_width
    pixel_ids_y = pixel_ids // image_width
    pixel_coords = torch.stack([pixel_ids_x, pixel_ids_y], dim=-1) + 0.5  # [M, 2]
    deltas = pixel_coords - means2d[camera_ids, gaussian_ids]  # [M, 2]
    c = conics[camera_ids, gaussian_ids]  # [M, 3]
    sigmas = 0.5 * (c[:, 0] * deltas[:, 0] ** 2 + c[:, 2] * deltas[:, 1] ** 2) + c[:, 1] * deltas[:, 0] * deltas[:, 1]
    alphas = torch.clamp_max(opacities[camera_ids, gaussian_ids] * torch.exp(-sigmas), 0.999)

    indices = camera_ids * image_height * image_width + pixel_ids
    total_pixels = C * image_height * image_width
    M = indices.numel()
    if M:
        log1m = torch.log1p(-alphas.double())
        incl = torch.cumsum(log1m, dim=0)
        excl = incl - log1m
        first = torch.ones_like(indices, dtype=torch.bool)
        first[1:] = indices[1:] != indices[:-1]
        # prefix sum just before the first entry of each pixel's segment, broadcast to the segment
        seg = torch.cumsum(first.long(), dim=0) - 1
        base = excl[first][seg]
        trans = torch.exp(excl - base).to(alphas.dtype)
        weights = alphas * trans
    else:
        weights = alphas
    renders = torch.zeros((total_pixels, channels), device=means2d.device, dtype=colors.dtype)
    renders = renders.index_add(0, indices, weights[:, None] * colors[camera_ids, gaussian_ids])
    accs = torch.zeros((total_pixels,), device=means2d.device, dtype=colors.dtype).index_add(0, indices, weights)
    return (renders.reshape(C, image_height, image_width, channels),
            accs.reshape(C, image_height, image_width, 1))


# ----------------------------------------------------------------------------------------
# un-fused projection chain
# ----------------------------------------------------------------------------------------
def quat_scale_to_covar_preci(
    quats: Tensor,  # [N, 4],
    scales: Tensor,  # [N, 3],
    compute_covar: bool = True,
    compute_preci: bool = True,
    triu: bool = False,
) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    """Converts quaternions and scales to covariance and precision matrices
    (G/cuda/_wrapper.py:76-107).  [N,6] upper triangles if `triu`, else [N,3,3]."""
    assert quats.dim() == 2 and quats.size(1) == 4, quats.size()
    assert scales.dim() == 2 and scales.size(1) == 3, scales.size()
    quats = quats.contiguous()
    scales = scales.contiguous()
    covars, precis = _QuatScaleToCovarPreci.apply(quats, scales, compute_covar, compute_preci, triu)
    return covars if compute_covar else None, precis if compute_preci else None


class _QuatScaleToCovarPreci(torch.autograd.Function):
    @staticmethod
    def forward(ctx, quats, scales, compute_covar=True, compute_preci=True, triu=False):
        _check_cuda(quats, scales)
        _f32(quats), _f32(scales)
        lib = get_lib()
        quats = _aligned16(quats)
        N = quats.shape[0]
        shape = (N, 6) if triu else (N, 3, 3)
        # a skipped output is an empty tensor, like the reference's (CS/quat_scale_to_covar_preci_fwd.cu)
        covars = torch.empty(shape if compute_covar else (0,), device=quats.device, dtype=torch.float32)
        precis = torch.empty(shape if compute_preci else (0,), device=quats.device, dtype=torch.float32)
        if N:
            native("quat_scale_to_covar_preci_fwd", lib, quats.device, N, _ptr(quats), _ptr(scales), int(triu),
                   _ptr(covars) if compute_covar else None, _ptr(precis) if compute_preci else None)
        ctx.save_for_backward(quats, scales)
        ctx.compute_covar, ctx.compute_preci, ctx.triu = compute_covar, compute_preci, triu
        return covars, precis

    @staticmethod
    def backward(ctx, v_covars, v_precis):
        quats, scales = ctx.saved_tensors
        lib = get_lib()
        N = quats.shape[0]
        if ctx.compute_covar and v_covars.is_sparse:
            v_covars = v_covars.to_dense()
        if ctx.compute_preci and v_precis.is_sparse:
            v_precis = v_precis.to_dense()
        v_covars = v_covars.contiguous() if ctx.compute_covar else None
        v_precis = v_precis.contiguous() if ctx.compute_preci else None
        v_quats = torch.empty_like(quats)
        v_scales = torch.empty_like(scales)
        if N:
            native("quat_scale_to_covar_preci_bwd", lib, quats.device, N, _ptr(quats), _ptr(scales), _ptr(v_covars),
                   _ptr(v_precis), int(ctx.triu), _ptr(v_quats), _ptr(v_scales))
        return v_quats, v_scales, None, None, None


def persp_proj(means: Tensor, covars: Tensor, Ks: Tensor, width: int, height: int) -> Tuple[Tensor, Tensor]:
    """DEPRECATED alias of `proj(..., camera_model="pinhole")` (G/cuda/_wrapper.py:110-139)."""
    import warnings

    warnings.warn("persp_proj is deprecated and will be removed in a future release. "
                  "Use proj with ortho=False instead.", DeprecationWarning)
    return proj(means, covars, Ks, width, height, "pinhole")


def proj(
    means: Tensor,  # [C, N, 3]
    covars: Tensor,  # [C, N, 3, 3]
    Ks: Tensor,  # [C, 3, 3]
    width: int,
    height: int,
    camera_model: Literal["pinhole", "ortho", "fisheye", "spherical"] = "pinhole",
) -> Tuple[Tensor, Tensor]:
    """Projection of camera-space Gaussians (G/cuda/_wrapper.py:142-172):
    returns (means2d [C,N,2], covars2d [C,N,2,2])."""
    C, N, _ = means.shape
    assert means.shape == (C, N, 3), means.size()
    assert covars.shape == (C, N, 3, 3), covars.size()
    assert Ks.shape == (C, 3, 3), Ks.size()
    if camera_model not in CAMERA_MODELS:
        raise AttributeError(f"CameraModelType has no member {camera_model.upper()!r}")
    return _Proj.apply(means.contiguous(), covars.contiguous(), Ks.contiguous(), width, height, camera_model)


class _Proj(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, covars, Ks, width, height, camera_model="pinhole"):
        _check_cuda(means, covars, Ks)
        _f32(means), _f32(covars), _f32(Ks)
        lib = get_lib()
        C, N = means.shape[:2]
        means2d = torch.empty((C, N, 2), device=means.device, dtype=torch.float32)
        covars2d = torch.empty((C, N, 2, 2), device=means.device, dtype=torch.float32)
        if C * N:
            native("proj_fwd", lib, means.device, C, N, _ptr(means), _ptr(covars), _ptr(Ks), width, height,
                   CAMERA_MODELS[camera_model], _ptr(means2d), _ptr(covars2d))
        ctx.save_for_backward(means, covars, Ks)
        ctx.width, ctx.height, ctx.camera_model = width, height, CAMERA_MODELS[camera_model]
        return means2d, covars2d

    @staticmethod
    def backward(ctx, v_means2d, v_covars2d):
        means, covars, Ks = ctx.saved_tensors
        lib = get_lib()
        C, N = means.shape[:2]
        v_means = torch.empty_like(means)
        v_covars = torch.empty_like(covars)
        if C * N:
            native("proj_bwd", lib, means.device, C, N, _ptr(means), _ptr(covars), _ptr(Ks), ctx.width, ctx.height,
                   ctx.camera_model, _ptr(_aligned16(v_means2d.contiguous())), _ptr(_aligned16(v_covars2d.contiguous())),
                   _ptr(v_means), _ptr(v_covars))
        return v_means, v_covars, None, None, None, None


def world_to_cam(
    means: Tensor,  # [N, 3]
    covars: Tensor,  # [N, 3, 3]
    viewmats: Tensor,  # [C, 4, 4]
) -> Tuple[Tensor, Tensor]:
    """Transforms Gaussians from world to camera coordinates (G/cuda/_wrapper.py:175-200):
    returns (means_c [C,N,3], covars_c [C,N,3,3])."""
    C = viewmats.size(0)
    N = means.size(0)
    assert means.size() == (N, 3), means.size()
    assert covars.size() == (N, 3, 3), covars.size()
    assert viewmats.size() == (C, 4, 4), viewmats.size()
    return _WorldToCam.apply(means.contiguous(), covars.contiguous(), viewmats.contiguous())


class _WorldToCam(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, covars, viewmats):
        _check_cuda(means, covars, viewmats)
        _f32(means), _f32(covars), _f32(viewmats)
        lib = get_lib()
        C, N = viewmats.shape[0], means.shape[0]
        means_c = torch.empty((C, N, 3), device=means.device, dtype=torch.float32)
        covars_c = torch.empty((C, N, 3, 3), device=means.device, dtype=torch.float32)
        if C * N:
            native("world_to_cam_fwd", lib, means.device, C, N, _ptr(means), _ptr(covars), _ptr(viewmats),
                   _ptr(means_c), _ptr(covars_c))
        ctx.save_for_backward(means, covars, viewmats)
        return means_c, covars_c

    @staticmethod
    def backward(ctx, v_means_c, v_covars_c):
        means, covars, viewmats = ctx.saved_tensors
        lib = get_lib()
        C, N = viewmats.shape[0], means.shape[0]
        need = ctx.needs_input_grad
        v_means = torch.empty_like(means) if need[0] else None
        v_covars = torch.empty_like(covars) if need[1] else None
        v_viewmats = torch.zeros_like(viewmats) if need[2] else None
        if C * N:
            native("world_to_cam_bwd", lib, means.device, C, N, _ptr(means), _ptr(covars), _ptr(viewmats),
                   _ptr(v_means_c.contiguous()), _ptr(v_covars_c.contiguous()), _ptr(v_means), _ptr(v_covars),
                   _ptr(v_viewmats))
        elif N:
            v_means = torch.zeros_like(means) if need[0] else None
            v_covars = torch.zeros_like(covars) if need[1] else None
        return v_means, v_covars, v_viewmats


# ----------------------------------------------------------------------------------------
# optimizer / densifier-side operators
# ----------------------------------------------------------------------------------------
def selective_adam_update(
    param: Tensor,
    param_grad: Tensor,
    exp_avg: Tensor,
    exp_avg_sq: Tensor,
    tiles_touched: Tensor,
    lr: float,
    b1: float,
    b2: float,
    eps: float,
    N: int,
    M: int,
) -> None:
    """In-place Adam update of the Gaussians flagged in `tiles_touched` [N] (bool)
    (G/cuda/_wrapper.py:19-34, kernel CS/adam.cu:16-44)."""
    _check_cuda(param, param_grad, exp_avg, exp_avg_sq, tiles_touched)
    for t in (param, param_grad, exp_avg, exp_avg_sq):
        _f32(t)
    if tiles_touched.dtype != torch.bool:
        raise RuntimeError(f"b200splat: tiles_touched must be bool, got {tiles_touched.dtype}")
    assert param.numel() == N * M and param_grad.numel() == N * M, (param.shape, N, M)
    assert exp_avg.numel() == N * M and exp_avg_sq.numel() == N * M and tiles_touched.numel() == N
    if N * M:
        native("selective_adam_update", get_lib(), param.device, _ptr(param), _ptr(param_grad), _ptr(exp_avg),
               _ptr(exp_avg_sq), _ptr(tiles_touched), float(lr), float(b1), float(b2), float(eps), N, M)


def compute_relocation(
    opacities: Tensor,  # [N]
    scales: Tensor,  # [N, 3]
    ratios: Tensor,  # [N]
    binoms: Tensor,  # [n_max, n_max]
) -> Tuple[Tensor, Tensor]:
    """New opacities / scales of relocated Gaussians, Eq. (9) of "3D Gaussian Splatting as
    Markov Chain Monte Carlo" (G/relocation.py:10-55, kernel CS/compute_relocation.cu:6-39).
    Like the reference, clamps `ratios` to [1, n_max] IN PLACE before converting to int."""
    N = opacities.shape[0]
    n_max, _ = binoms.shape
    assert scales.shape == (N, 3), scales.shape
    assert ratios.shape == (N,), ratios.shape
    opacities = opacities.contiguous()
    scales = scales.contiguous()
    ratios.clamp_(min=1, max=n_max)
    ratios = ratios.int().contiguous()
    binoms = binoms.contiguous()
    _check_cuda(opacities, scales, ratios, binoms)
    _f32(opacities), _f32(scales), _f32(binoms)
    new_opacities = torch.empty_like(opacities)
    new_scales = torch.empty_like(scales)
    if N:
        native("compute_relocation", get_lib(), opacities.device, N, _ptr(opacities), _ptr(scales), _ptr(ratios),
               _ptr(binoms), int(n_max), _ptr(new_opacities), _ptr(new_scales))
    return new_opacities, new_scales
