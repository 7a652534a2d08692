"""Camera-sharded data parallelism for the rasterization path.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Gaussians are
replicated, rank r renders cameras r, r+W, r+2W, ... of the global batch with the
single-GPU path unchanged, and the parameter gradients are summed with ONE all-reduce
over a flat fp32 arena [means | quats | scales | opacities | sh] (59·N floats at K=16).
This replaces — by design — the reference's Gaussian-sharded all-to-all mode
(`distributed=True`, G/rendering.py:279-294, 394-478; helpers in G/distributed.py).

The helpers work on any backend (gloo on CPU for the tests, nccl on the GPUs).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


class camera_parallel:
    """Context manager for forward + backward of one camera-sharded step.

    While active, the backward of the fused colour stage does not leave a per-rank SH
    coefficient gradient to be all-reduced (3K floats per Gaussian, 81 % of the payload at
    K = 16).  The coefficient gradient of one camera is the outer product of the SH basis at
    that camera's view direction with the colour cotangent, so ranks ALL-GATHER their masked
    colour cotangents (3 floats per Gaussian and camera) and camera centres, and every rank
    sums the outer products over all cameras in one kernel: the SH gradient comes out of
    `backward()` already global.  `reduced_ptrs` lists the parameters (by data_ptr) this
    happened for; `GradArena.all_reduce(skip_ptrs=...)` / `allreduce_mixed_gradients(skip_ptrs=...)`
    then leave them out.  Covered colour stages: the fused un-packed one (whole table, `torch.cat([sh0,
    shN])` of leaves, or the split `(sh0, shN)` pair) and the packed one with a whole table (its masked
    cotangents are scattered into the dense [C,N,3] layout first); per-view tables, the packed split pair
    and pose-gradient runs fall back to the per-rank gradient, which is then reduced like the others."""

    def __init__(self, group=None, defer: bool = False, n_cameras_global: Optional[int] = None,
                 peer: Optional["PeerExchange"] = None):
        """`peer`: a `PeerExchange` — the colour cotangents are then published into this rank's symmetric
        buffer and the colour backward kernel reads the other ranks' blocks IN PLACE over NVLink (no
        all-gather, no gathered copy).  With `defer=True` that kernel runs on a side stream, overlapped with the
        projection backward and the arena all-reduce, and `finish()` joins it.
        `n_cameras_global`: size of the global camera batch when it is NOT a multiple of the world
        size (shards then differ by one camera, `shard_cameras(..., allow_uneven=True)`); leave None for
        equal shards.  `defer=True` needs the gradient sink of the arena to be active around `backward()`
        and `p.grad is None` for the SH parameters (`zero_grad(set_to_none=True)`); where that does not
        hold the colour stage silently uses the immediate order."""
        self.group = group if group is not None else dist.group.WORLD
        self.reduced_ptrs = set()
        self.defer = defer
        self.n_cameras_global = n_cameras_global
        self.peer = peer
        self._deferred = []  # (means parameter, finish(all_cameras) -> v_means) of the colour stages

    def __enter__(self):
        from . import wrapper

        self._active = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        if self._active:
            wrapper._CAMERA_PARALLEL["group"] = self.group
            wrapper._CAMERA_PARALLEL["reduced"] = self.reduced_ptrs
            if self.defer:
                wrapper._CAMERA_PARALLEL["deferred"] = self._deferred
            if self.n_cameras_global is not None:
                wrapper._CAMERA_PARALLEL["n_cameras_global"] = int(self.n_cameras_global)
            if self.peer is not None:
                wrapper._CAMERA_PARALLEL["peer"] = self.peer
        return self

    def __exit__(self, *exc):
        from . import wrapper

        wrapper._CAMERA_PARALLEL.clear()
        return False

    def finish(self, arena: "GradArena", average: bool = False) -> None:
        """Deferred mode (`defer=True`), call after `backward()` instead of
        `arena.gather_from_params(); arena.all_reduce(skip_ptrs=...)`:

            backward:  ... raster bwd -> [colour node: START all-gather of cotangents] -> projection bwd
            finish:    all-reduce of the arena (async, NCCL stream)  ||  colour backward kernel over ALL
                       cameras (this stream, behind the all-gather)  -> means.grad += direction gradient

        so the all-gather overlaps the projection backward and the all-reduce overlaps the colour
        backward.  The direction (means) gradient of the colour stage is summed over all cameras on
        every rank, i.e. it is already global and is added AFTER the all-reduce.  Reduced gradients
        live in the arena views (`arena.scatter_to_params()` re-attaches them)."""
        arena.gather_from_params()
        W = dist.get_world_size(self.group)
        if arena.peer is not None and self._deferred and all(getattr(fin, "local", False) for _, fin in self._deferred):
            # peer exchange, overlapped: the colour kernel of each deferred node is already running on a side
            # stream (NVLink reads of the peers' cotangents; direction gradient of THIS rank's cameras only).
            #   high-priority stream: all-reduce of everything but the means segment  || colour kernel
            #   main stream:          join, means += local direction gradient, all-reduce of the means segment
            from .wrapper import _side_stream

            dev = arena.flat.device
            main, hp = torch.cuda.current_stream(dev), _side_stream(dev, priority=-1)
            late = {m.data_ptr() for m, _ in self._deferred if any(p is m for p in arena.params)}
            hp.wait_stream(main)
            with torch.cuda.stream(hp):
                arena.all_reduce(group=self.group, skip_ptrs=set(self.reduced_ptrs) | late)
            for m, fin in self._deferred:
                vm = fin(True)  # main stream waits for the side stream's colour kernel
                if vm is None:
                    continue
                for p, v in zip(arena.params, arena.views):
                    if p is m:
                        v.add_(vm)
                        break
                else:  # not an arena parameter: reduce and accumulate on the tensor itself
                    dist.all_reduce(vm, group=self.group)
                    m.grad = vm if m.grad is None else m.grad + vm
            self._deferred.clear()
            main.wait_stream(hp)  # the two all-reduce kernels share the handshake flags: strictly ordered
            if late:
                arena.all_reduce(group=self.group, only_ptrs=late)
            if average:
                for v in arena.views:
                    v.div_(W)
            return
        works = arena.all_reduce(group=self.group, async_op=True, skip_ptrs=self.reduced_ptrs)
        pending = [(m, fin(True)) for m, fin in self._deferred]
        self._deferred.clear()
        for w in works or ():
            w.wait()
        if average:
            for p, v in zip(arena.params, arena.views):
                v.div_(W)
        for m, vm in pending:
            if vm is None:
                continue
            if average:
                vm = vm / W
            for p, v in zip(arena.params, arena.views):
                if p is m:
                    v.add_(vm)
                    break
            else:  # not an arena parameter: accumulate on the tensor itself
                m.grad = vm if m.grad is None else m.grad + vm


def bind_host_to_gpu(device_index: int) -> Optional[List[int]]:
    """Pin the calling process to the CPU cores closest to GPU `device_index` (NVML's ideal CPU set: the
    cores of the NUMA node the GPU's PCIe root hangs off).  Call it BEFORE allocating pinned host buffers:
    page-locked memory is placed on the node of the thread that allocates it, and with eight ranks per host
    the per-step image traffic (33 MB each way per rank at 1080p) otherwise crosses the socket interconnect
    for half of the GPUs.  Returns the CPU list, or None where NVML / the affinity call is unavailable
    (containers with a fixed cpuset): nothing is changed then."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def shard_cameras(viewmats: Tensor, Ks: Tensor, rank: Optional[int] = None,
                  world_size: Optional[int] = None, allow_uneven: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    """Cameras owned by `rank`: indices rank, rank+W, ...  Returns (viewmats, Ks, global ids).
    The colour-cotangent exchange of `camera_parallel` all-gathers equal blocks: a global batch that
    is not a multiple of the world size must be announced (`allow_uneven=True` here and
    `camera_parallel(n_cameras_global=C)`), otherwise it is an error."""
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    if viewmats.shape[0] % world_size != 0 and not allow_uneven:
        raise ValueError(f"{viewmats.shape[0]} cameras do not divide over {world_size} ranks: pass allow_uneven=True "
                         f"and camera_parallel(n_cameras_global={viewmats.shape[0]})")
    ids = torch.arange(rank, viewmats.shape[0], world_size, device=viewmats.device)
    return viewmats[ids].contiguous(), Ks[ids].contiguous(), ids


def arena_layout(params: Sequence[Tensor]) -> Tuple[List[int], int]:
    """Offsets (in floats) of the parameters' gradient segments in a flat arena, and its length;
    every segment starts on a 16-byte boundary."""
    offs, n = [], 0
    for p in params:
        assert p.dtype == torch.float32, "arena holds fp32 gradients"
        offs.append(n)
        n += (p.numel() + 3) // 4 * 4
    return offs, n


class PeerExchange:
    """Symmetric (peer-mapped) device memory for the camera-parallel gradient exchange, and the two
    exchanges as this library's own kernels over it (csrc/peer.cu) instead of NCCL collectives:

    * colour cotangents: each rank PUBLISHES its masked cotangents and camera centres into its own
      block; after a flag barrier the colour backward kernel of every rank reads all W blocks in place
      over NVLink (`b200splat_sh_colors*_bwd_peer`) — the all-gather and its W-fold buffer disappear
      into the consumer's loads.  Two slots alternate between steps, so one barrier per step suffices
      (a slot is rewritten only after the barrier of the following step, which every rank reaches after
      its reads of that slot);
    * arena: `GradArena(params, peer=this)` places the flat gradient arena in the same buffer and
      all-reduces it with the two-shot kernel (`b200splat_peer_allreduce_f32`; switch-side reduction
      through the multicast mapping when the fabric offers one).

    Layout of the buffer (floats): [ flags | slot 0 | slot 1 | arena ], slot = {campos [Cm,3], pad to
    hdr, cotangents [Cm,N,3]}.  Construction is COLLECTIVE over `group` (symmetric allocation +
    rendezvous, `torch.distributed._symmetric_memory` is only the allocator/handle exchange).  Raises
    if the platform cannot map peer memory; callers fall back to the NCCL path (`camera_parallel()`
    without `peer`, `GradArena(params)`).  The buffer is sized for ONE (n_gaussians, cams_per_rank): after
    densification changes the number of Gaussians build a new exchange (and arena) on every rank — a step
    whose sizes do not match (`fits()`) silently takes the NCCL all-gather for the cotangents instead.
    All ranks must issue the same sequence of exchange calls (the handshake flags are positional); a rank
    that never arrives makes its peers trap after 20 s instead of hanging."""

    def __init__(self, n_gaussians: int, cams_per_rank: int = 1, group=None, arena_floats: int = 0,
                 device: Optional[torch.device] = None, use_multicast: Optional[bool] = None):
        import torch.distributed._symmetric_memory as symm

        from ._lib import get_lib

        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.N, self.Cm = int(n_gaussians), int(cams_per_rank)
        self.hdr, self.slot_floats, self.flag_floats, self.total_floats = self.layout(
            self.N, self.Cm, self.world, get_lib().b200splat_peer_flag_bytes(self.world))
        self.arena_floats = (int(arena_floats) + 3) // 4 * 4
        self.slot_off = [self.flag_floats, self.flag_floats + self.slot_floats]
        self.arena_off = self.flag_floats + 2 * self.slot_floats
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.buf = symm.empty(self.total_floats + self.arena_floats, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, self.group)
        self.bases_dev = int(self.handle.buffer_ptrs_dev)
        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        # switch-side reduction pays from 4 ranks on (at 2 the multicast path also routes a rank's own copy
        # through the switch: measured 0.157 vs 0.093 ms for 44 MB, tools/peer_bench.py)
        self.multicast_available = mc != 0
        if use_multicast is None:
            use_multicast = self.world >= 4
        self.multicast_base = mc if use_multicast else 0
        self._k = 0
        # the zeroed flags must be in place everywhere before the first handshake
        torch.cuda.synchronize(device)
        dist.barrier(self.group)

    @staticmethod
    def layout(n_gaussians: int, cams_per_rank: int, world: int, flag_bytes: int) -> Tuple[int, int, int, int]:
        """(hdr_floats, slot_floats, flag_floats, floats before the arena)."""
        hdr = (3 * cams_per_rank + 3) // 4 * 4
        slot = (hdr + 3 * cams_per_rank * n_gaussians + 3) // 4 * 4
        flags = (flag_bytes // 4 + 3) // 4 * 4
        return hdr, slot, flags, flags + 2 * slot

    def fits(self, n_gaussians: int, cams_per_rank: int) -> bool:
        return n_gaussians == self.N and cams_per_rank == self.Cm

    def next_slot(self) -> int:
        self._k += 1
        return self._k & 1

    def slot_view(self, slot: int) -> Tensor:
        o = self.slot_off[slot]
        return self.buf[o:o + self.slot_floats]

    def arena_view(self, n: int) -> Tensor:
        assert n <= self.arena_floats, f"PeerExchange was built with arena_floats={self.arena_floats} < {n}"
        return self.buf[self.arena_off:self.arena_off + n]

    def barrier(self) -> None:
        """Stream-ordered meeting of all ranks (one-warp kernel, flags in the symmetric buffers)."""
        from .wrapper import get_lib, native

        native("peer_barrier", get_lib(), self.buf.device, self.world, self.rank, self.bases_dev, 0)

    def all_reduce_(self, offset_floats: int, n_floats: int) -> None:
        """In-place SUM over ranks of buf[offset : offset + n] (a multiple of 4 floats, 16-byte aligned)."""
        from .wrapper import get_lib, native

        native("peer_allreduce_f32", get_lib(), self.buf.device, self.world, self.rank, self.bases_dev,
               self.multicast_base, 4 * offset_floats, n_floats, self.bases_dev, 0)


class GradArena:
    """Flat, 16-byte-segment-aligned fp32 buffer holding the gradients of a fixed list of
    parameters, so a step needs exactly one all-reduce launch.  With `peer` (a `PeerExchange` built
    with `arena_floats >= arena_layout(params)[1]`) the arena lives in symmetric memory and
    `all_reduce` is this library's two-shot kernel over NVLink instead of the NCCL collective."""

    def __init__(self, params: Sequence[Tensor], peer: Optional[PeerExchange] = None):
        self.params = list(params)
        self.offsets, n = arena_layout(self.params)
        self.peer = peer
        if peer is not None:
            self.flat = peer.arena_view(n)
            self.flat.zero_()
        else:
            self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, self.params)]

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def sink(self):
        """Context manager for the backward pass: gradients that the CUDA kernels fully overwrite
        are produced directly inside the arena (see `wrapper.gradient_sink`), so
        `gather_from_params` has nothing to copy for them."""
        from .wrapper import gradient_sink

        return gradient_sink(zip(self.params, self.views))

    def gather_from_params(self, zero_missing: bool = True) -> None:
        """Copy p.grad (dense or sparse COO) of every parameter into the arena (skipping the
        ones that already live there)."""
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is not None and not g.is_sparse and g.data_ptr() == v.data_ptr() and g.shape == v.shape:
                continue
            if g is None:
                if zero_missing:
                    v.zero_()
                continue
            if g.is_sparse:
                v.zero_()
                g = g.coalesce()
                v.index_add_(0, g.indices()[0], g.values())
            else:
                v.copy_(g)

    def all_reduce(self, group=None, average: bool = False, async_op: bool = False, skip_ptrs=(), only_ptrs=None):
        """SUM over ranks (optionally / world_size); a no-op outside a process group.
        Parameters whose data_ptr() is in `skip_ptrs` already hold a global gradient
        (`camera_parallel`) and are left out: the arena is reduced as the maximal contiguous runs
        of the remaining segments (one collective when the skipped parameter is the last one).
        `only_ptrs`: reduce just the segments of these parameters."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        runs, start = [], None
        for p, o in zip(self.params, self.offsets):
            if p.data_ptr() in skip_ptrs or (only_ptrs is not None and p.data_ptr() not in only_ptrs):
                if start is not None:
                    runs.append((start, o))
                    start = None
            elif start is None:
                start = o
        if start is not None:
            runs.append((start, self.flat.numel()))
        if self.peer is not None and (group is None or group is self.peer.group):
            W = self.peer.world
            for a, b in runs:
                self.peer.all_reduce_(self.peer.arena_off + a, b - a)  # current stream; nothing to wait for
                if average:
                    self.flat[a:b].div_(W)
            if average:
                for p, v in zip(self.params, self.views):
                    if p.data_ptr() in skip_ptrs:
                        v.div_(W)
            return [] if async_op else None
        works = []
        for a, b in runs:
            works.append(dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=group, async_op=async_op))
            if average and not async_op:
                self.flat[a:b].div_(dist.get_world_size(group))
        if average and not async_op:
            for p, v in zip(self.params, self.views):
                if p.data_ptr() in skip_ptrs:
                    v.div_(dist.get_world_size(group))
        return works if async_op else (works[-1] if works else None)

    def scatter_to_params(self) -> None:
        """Point every p.grad at its arena view (no copy)."""
        for p, v in zip(self.params, self.views):
            p.grad = v


def allreduce_gradients(params: Sequence[Tensor], arena: Optional[GradArena] = None, group=None,
                        average: bool = False) -> GradArena:
    """Pack p.grad of `params` into one arena, all-reduce it, and re-attach the views."""
    arena = arena or GradArena(params)
    arena.gather_from_params()
    arena.all_reduce(group=group, average=average)
    arena.scatter_to_params()
    return arena


def rasterization_dp(render_fn, params: Dict[str, Tensor], viewmats: Tensor, Ks: Tensor, width: int, height: int,
                     rank: Optional[int] = None, world_size: Optional[int] = None, **kwargs):
    """Render this rank's share of a global camera batch.

    `render_fn` is `splat_one_b200.rasterization` (or the oracle's, in CPU tests);
    `params` holds means/quats/scales/opacities/colors.  Returns (colors, alphas, meta,
    global camera ids of the local batch).  Intersection ids use LOCAL camera indices, as in
    the reference's own distributed mode (G/rendering.py:417-425)."""
    vm, k, ids = shard_cameras(viewmats, Ks, rank, world_size)
    rc, ra, meta = render_fn(params["means"], params["quats"], params["scales"], params["opacities"],
                             params["colors"], vm, k, width, height, **kwargs)
    return rc, ra, meta, ids


def sparse_allreduce(grads: Sequence[Tensor], group=None, dense_threshold: float = 0.4,
                     counts_out: Optional[list] = None) -> List[Tensor]:
    """SUM over ranks of sparse per-Gaussian gradients (packed mode with `sparse_grad=True`, config E).

    `grads`: sparse COO tensors of this rank that share ONE index vector — the `gaussian_ids` of the
    packed projection, [1, nnz] — with value rows [nnz, d_i] (the layout of the reference's sparse
    gradients, G/cuda/_wrapper.py:1163-1203).  Ranks exchange `(gaussian_ids, value rows)` with ONE
    padded all-gather (8 + 4·sum(d_i) bytes per visible Gaussian and rank) and every rank returns the
    concatenation as un-coalesced COO tensors: identical on all ranks, and equal to the all-reduced dense
    gradient once coalesced.  When the visible fraction sum_r(nnz_r) / N exceeds `dense_threshold` the
    rows are scattered into dense tensors and all-reduced instead (the gathered rows would then outweigh
    the dense payload); the result is dense in that case.  Outside a process group: returned unchanged."""
    grads = list(grads)
    if not grads or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return grads
    W = dist.get_world_size(group)
    g0 = grads[0]
    assert all(g.is_sparse and g.shape[0] == g0.shape[0] for g in grads), "sparse COO gradients over the same rows"
    ids = g0._indices()[0].contiguous()
    nnz, n_rows, dev = ids.numel(), g0.shape[0], ids.device
    for g in grads[1:]:
        assert g._indices().shape == g0._indices().shape, "gradients must share gaussian_ids"
    widths = [g._values().numel() // max(nnz, 1) if nnz else int(torch.tensor(g.shape[1:]).prod()) for g in grads]
    counts = torch.empty(W, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, torch.tensor([nnz], dtype=torch.int64, device=dev), group=group)
    counts = counts.tolist()  # one small host sync, like the n_isects / nnz read-backs of the path
    if counts_out is not None:
        counts_out[:] = counts
    total = sum(counts)
    if total > dense_threshold * n_rows:
        out = []
        for g in grads:
            d = torch.zeros(g.shape, dtype=g.dtype, device=dev)
            if nnz:
                d.index_add_(0, ids, g._values())
            dist.all_reduce(d, op=dist.ReduceOp.SUM, group=group)
            out.append(d)
        return out
    m = max(counts)
    D = sum(widths)
    ids_pad = torch.zeros(m, dtype=torch.int64, device=dev)
    rows_pad = torch.zeros((m, D), dtype=torch.float32, device=dev)
    if nnz:
        ids_pad[:nnz] = ids
        rows_pad[:nnz] = torch.cat([g._values().reshape(nnz, -1) for g in grads], dim=1)
    ids_all = torch.empty(W * m, dtype=torch.int64, device=dev)
    rows_all = torch.empty((W * m, D), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(ids_all, ids_pad, group=group)
    dist.all_gather_into_tensor(rows_all, rows_pad, group=group)
    keep = torch.cat([torch.arange(r * m, r * m + c, device=dev) for r, c in enumerate(counts)]) if total else \
        torch.zeros(0, dtype=torch.int64, device=dev)
    ids_cat, rows_cat = ids_all[keep], rows_all[keep]
    out, o = [], 0
    for g, w in zip(grads, widths):
        vals = rows_cat[:, o:o + w].reshape((total,) + tuple(g.shape[1:])).contiguous()
        out.append(torch.sparse_coo_tensor(ids_cat[None], vals, size=g.shape, is_coalesced=False,
                                           check_invariants=False))
        o += w
    return out


def allreduce_mixed_gradients(params: Sequence[Tensor], arena: Optional[GradArena] = None, group=None,
                              dense_threshold: float = 0.4, skip_ptrs=()) -> None:
    """Gradient exchange of a camera-sharded step whose parameters carry a MIX of dense and sparse
    gradients (packed mode, `sparse_grad=True`): the sparse ones (sharing `gaussian_ids`) go through
    `sparse_allreduce`, the dense ones through the flat arena.  `p.grad` of every parameter holds the
    global sum afterwards (sparse stays sparse below the threshold).  Inside `camera_parallel` the packed
    colour stage exchanges cotangents like the un-packed one (scattered to the dense [C,N,3] layout), so pass
    `skip_ptrs=cp.reduced_ptrs`."""
    params = list(params)
    sparse = [p for p in params if p.grad is not None and p.grad.is_sparse]
    # parameters whose gradient is already global (`camera_parallel.reduced_ptrs`: the SH table of the colour
    # exchange, 81 % of the bytes) are neither copied nor reduced
    dense = [p for p in params if p.grad is not None and not p.grad.is_sparse and p.data_ptr() not in skip_ptrs]
    if sparse:
        for p, g in zip(sparse, sparse_allreduce([p.grad for p in sparse], group=group, dense_threshold=dense_threshold)):
            p.grad = g
    if dense:
        if arena is None or [id(p) for p in arena.params] != [id(p) for p in dense]:
            arena = GradArena(dense)
        arena.gather_from_params()
        arena.all_reduce(group=group, skip_ptrs=skip_ptrs)
        arena.scatter_to_params()
