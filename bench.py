#!/usr/bin/env python
"""Benchmark of the rasterization hot path (BASELINE.json metric):
rendered Mpix/s, forward + backward, 1 M Gaussians / SH degree 3 / one 1920x1080 pinhole
camera per GPU ("config B"), on 1/2/4/8 B200 with camera-sharded data parallelism.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one `rasterization()` forward + `backward()` with fixed random cotangents on
one synthetic scene, plus — for N > 1 — the all-reduce of the parameter
gradients.  Prints ONE JSON line (see the keys below).  `--impl reference` times the CPU
restatement of the reference (oracle/, the reference's native code being CUDA-only and
its PyTorch `_torch_impl` having no runnable rasterizer) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_GAUSS = 1_000_000
WIDTH, HEIGHT = 1920, 1080
SH_DEGREE = 3
WORKLOAD = "configB: 1M Gaussians, SH3, 1920x1080 pinhole, packed=False, fwd+bwd"
# CPU arm: the FULL config B (a step of the oracle port takes a few seconds on the host cores);
# what is bounded is the number of steps (CPU_BUDGET_S of wall clock per run)
CPU_SAMPLE = dict(n=N_GAUSS, width=WIDTH, height=HEIGHT)
CPU_BUDGET_S = 200.0


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): an NVML
    polling thread (10 ms period) in this process; `nvidia-smi -lms` as the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        import threading

        self.sm, self.mx, self.reasons, self.power = [], [], set(), []
        self._stop = threading.Event()
        self._thread = None
        self._smi = None
        try:
            import pynvml as nv

            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons

            def poll():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        self.mx.append(float(mx))
                        r = int(get_reasons(h))
                        for name, bit in bits.items():
                            if r & bit:
                                self.reasons.add(name)
                        self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                    except Exception:
                        pass
                    self._stop.wait(0.01)

            self._thread = threading.Thread(target=poll, daemon=True)
            self._thread.start()
        except Exception:
            self._thread = None
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            try:
                self._smi = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                              "-lms", "100", "-i", str(gpu_index)], stdout=self.f,
                                             stderr=subprocess.DEVNULL)
            except Exception:
                self._smi = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        elif self._smi is not None:
            self._smi.terminate()
            try:
                self._smi.wait(timeout=5)
            except Exception:
                self._smi.kill()
            self.f.flush()
            self.f.seek(0)
            for line in self.f.read().splitlines():
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    self.sm.append(float(c[1]))
                    self.mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   c[5:9]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            try:
                os.unlink(self.f.name)
            except OSError:
                pass
        if self.sm:
            out = {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": max(self.mx), "reasons": sorted(self.reasons),
                   "samples": len(self.sm)}
            if self.power:
                out["power_w"] = round(statistics.median(self.power), 1)
        return out


# ----------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference, bounded sample
# ----------------------------------------------------------------------------------------
def cpu_reference_step(scene, params, cot):
    from oracle import raster_ref as RC
    from oracle import torch_ref as O

    for p in params:
        p.grad = None
    rc, ra, meta = O.rasterization(*params, scene["viewmats"], scene["Ks"], scene["width"], scene["height"],
                                   sh_degree=SH_DEGREE, packed=False, raster_fn=RC.rasterize_to_pixels)
    if cot is None:
        g = torch.Generator().manual_seed(123)
        cot = (torch.randn(rc.shape, generator=g), torch.randn(ra.shape, generator=g))
    ((rc * cot[0]).sum() + (ra * cot[1]).sum()).backward()
    return cot, meta


def cpu_baseline(steps: int, warmup: int, budget_s: float = CPU_BUDGET_S):
    """Time the oracle port on the host cores, on the same scene and sizes as the GPU arm (config B).
    `steps` / `warmup` are honoured as long as the run fits `budget_s` of wall clock (the first step is
    timed to decide).  Returns (Mpix/s, ms/step, cores, sample str, steps done, warm-ups done)."""
    from oracle import raster_ref as RC
    from splat_one_b200 import synthetic

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    RC.build()
    scene = synthetic.pinhole_scene(CPU_SAMPLE["n"], CPU_SAMPLE["width"], CPU_SAMPLE["height"], seed=42,
                                    sh_degree=SH_DEGREE)
    params = [scene[k].clone().requires_grad_() for k in ("means", "quats", "scales", "opacities", "sh")]
    t_start = time.perf_counter()
    cot, _ = cpu_reference_step(scene, params, None)  # first warm-up (builds the cotangents), timed for the budget
    t_first = time.perf_counter() - t_start
    fit = max(1, int((budget_s - t_first) / max(t_first, 1e-3)))
    warm_done = 1 + max(0, min(warmup - 1, fit - 1))
    for _ in range(warm_done - 1):
        cot, _ = cpu_reference_step(scene, params, cot)
    steps = max(1, min(steps, fit - (warm_done - 1)))
    t0 = time.perf_counter()
    for _ in range(steps):
        cot, meta = cpu_reference_step(scene, params, cot)
    dt = (time.perf_counter() - t0) / steps
    mpix = CPU_SAMPLE["width"] * CPU_SAMPLE["height"] / dt / 1e6
    sample = (f"full config B: {CPU_SAMPLE['n']} Gaussians @ {CPU_SAMPLE['width']}x{CPU_SAMPLE['height']} "
              f"(n_isects={meta['flatten_ids'].numel()}), oracle port: torch-CPU projection/SH/autograd + "
              f"numpy isect + C raster fwd/bwd, {steps} steps after {warm_done} warm-ups")
    return mpix, dt * 1e3, cores, sample, steps, warm_done


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mpix, ms, cores, sample, steps, warm = cpu_baseline(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "rendered Mpix/s fwd+bwd", "value": mpix, "unit": "Mpix/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
def train_step_report(S, scene, dev, reps: int = 30, warm: int = 5):
    """Extra, not the headline metric: one TRAINING step of the f4 row on the
    same scene — raw parameters -> splat_activations -> split-SH rasterization -> fused L1+SSIM ->
    backward to the raw parameters (R/utils/gsplat_utils/gsplat_trainer.py:446-497, 624-628),
    timed with CUDA events after warm-up."""
    raw = {
        "means": scene["means"], "quats": scene["quats"], "scales": torch.log(scene["scales"]),
        "opacities": torch.logit(scene["opacities"].clamp(1e-4, 1 - 1e-4)),
        "sh0": scene["sh"][:, :1].contiguous(), "shN": scene["sh"][:, 1:].contiguous(),
    }
    P = {k: v.to(dev).requires_grad_() for k, v in raw.items()}
    c2w = torch.inverse(scene["viewmats"][:1]).to(dev)
    Ks = scene["Ks"][:1].to(dev)
    pixels = torch.rand(1, HEIGHT, WIDTH, 3, generator=torch.Generator().manual_seed(7)).to(dev)

    def step():
        for p in P.values():
            p.grad = None
        rc, _, _ = S.rasterize_splats(P, c2w, Ks, WIDTH, HEIGHT, sh_degree=SH_DEGREE, packed=False)
        loss = S.l1_ssim_loss(rc, pixels, 0.2)
        loss.backward()
        return loss

    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        loss = step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    return {"ms_per_step": ms, "Mpix_per_s": HEIGHT * WIDTH / (ms * 1e-3) / 1e6, "steps": reps, "loss": loss.item(),
            "what": "raw params -> exp/sigmoid (1 kernel) -> rasterization(colors=(sh0, shN)) -> fused L1+SSIM "
                    "-> backward; config B scene, random target image"}



def reference_cuda_leg(S, scene, dev, ms_ours: float, reps: int = 20, warm: int = 5):
    """Extra, labelled leg (outside the timed region, never on the product path): the reference's OWN
    CUDA kernels — oracle/_ref, the fork's gsplat extension built for sm_100a by oracle/build_ref.py —
    chained the way G/rendering.py chains them (raw pybind calls: without the reference's Python and
    autograd overhead, which favours the reference), on the same scene, same GPU, same run.
    None when oracle/_ref was not built."""
    try:
        from oracle import ref_cuda
    except Exception:
        return None
    R = ref_cuda.load()
    if R is None:
        return {"unavailable": "oracle/_ref/gsplat_ref_csrc.so not built (python oracle/build_ref.py)"}
    P = {k: scene[k].to(dev) for k in ("means", "quats", "scales", "opacities", "sh")}
    P["viewmats"], P["Ks"] = scene["viewmats"][:1].to(dev), scene["Ks"][:1].to(dev)
    g = torch.Generator().manual_seed(1000)
    vc = torch.randn(1, HEIGHT, WIDTH, 3, generator=g).to(dev)
    va = torch.randn(1, HEIGHT, WIDTH, 1, generator=g).to(dev)
    for _ in range(warm):
        ref = ref_cuda.reference_chain(R, P, WIDTH, HEIGHT, "pinhole", vc, va)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ref = ref_cuda.reference_chain(R, P, WIDTH, HEIGHT, "pinhole", vc, va)
    b.record()
    torch.cuda.synchronize()
    ms_ref = a.elapsed_time(b) / reps
    # parity of this very run (the asserting version lives in tests/test_gpu_fullsize_parity.py)
    A = [P[k].clone().requires_grad_() for k in ("means", "quats", "scales", "opacities", "sh")]
    rc, ra, _ = S.rasterization(*A, P["viewmats"], P["Ks"], WIDTH, HEIGHT, sh_degree=SH_DEGREE, packed=False)
    torch.autograd.backward([rc, ra], [vc, va])
    err = (rc.detach() - ref["image"]).abs()
    gerr = {k: float((p.grad - ref["grads"][k]).abs().max() / (ref["grads"][k].abs().max() + 1e-30))
            for k, p in zip(("means", "quats", "scales", "opacities", "sh"), A)}
    return {"ms_per_step": ms_ref, "Mpix_per_s": HEIGHT * WIDTH / (ms_ref * 1e-3) / 1e6, "steps": reps,
            "speedup_of_this_library": ms_ref / ms_ours, "image_max_abs_err": float(err.max()),
            "image_frac_outside_1e-4": float((err > 1e-4).float().mean()), "grad_max_rel_err": gerr,
            "what": "reference CUDA (oracle/_ref, sm_100a build of the fork's kernels) chained as G/rendering.py does, "
                    "raw pybind calls, same scene / GPU / run; device-resident like `value`"}


# N > 1: which gradient exchange the timed loop uses.  "peer" (default) = this library's own kernels over
# NVLink peer memory (splat_one_b200.distributed.PeerExchange, csrc/peer.cu), falling back to "nccl" (all-gather
# + all-reduce library collectives) if the platform cannot map symmetric memory; the other one is A/B-timed.
EXCHANGE = os.environ.get("B200SPLAT_DP_EXCHANGE", "peer")


def make_peer(n_gaussians, cams_per_rank, params, world, note=None, with_arena=True):
    """PeerExchange sized for `params`, or None (with the reason in note["error"]) where unavailable."""
    if world <= 1 or EXCHANGE == "nccl":
        return None
    from splat_one_b200.distributed import PeerExchange, arena_layout

    try:
        return PeerExchange(n_gaussians, cams_per_rank, arena_floats=arena_layout(params)[1] if with_arena else 0,
                            use_multicast={"1": True, "0": False}.get(os.environ.get("B200SPLAT_DP_MULTICAST", "")))
    except Exception as e:  # no peer mapping on this platform: the NCCL exchange is used and the line says so
        if note is not None:
            note["error"] = f"{type(e).__name__}: {e}"[:300]
        return None


def exchange_kind(peer):
    if peer is None:
        return "nccl: all-gather of colour cotangents + all-reduce of the arena"
    return ("peer kernels over NVLink symmetric memory: cotangents read in place by the colour backward + two-shot "
            + ("multimem (switch-reduced)" if peer.multicast_base else "peer load/store") + " arena all-reduce")


def dp_parity_check(S, dist, world, rank, dev):
    """Outside the timed region, N > 1: gradients of the camera-sharded step (each rank one camera,
    colour-cotangent all-gather + arena all-reduce) against the SAME batch rendered by one process
    (C = N cameras on this GPU), 40 k Gaussians at 320x240.  Returns the max over ranks and parameters of
    max|dp - batch| / max|batch|."""
    from splat_one_b200 import synthetic
    from splat_one_b200.distributed import GradArena, camera_parallel

    W_, H_, N_ = 320, 240, 40000
    sc = synthetic.pinhole_scene(N_, W_, H_, seed=7, n_cameras=world)
    names = ("means", "quats", "scales", "opacities", "sh")
    g = torch.Generator().manual_seed(3)
    vc_all = torch.randn(world, H_, W_, 3, generator=g).to(dev)
    va_all = torch.randn(world, H_, W_, 1, generator=g).to(dev)

    def grads(cams, dp):
        P = [sc[k].to(dev).requires_grad_() for k in names]
        rc, ra, _ = S.rasterization(*P, sc["viewmats"][cams].to(dev), sc["Ks"][cams].to(dev), W_, H_,
                                    sh_degree=SH_DEGREE, packed=False)
        if dp:
            peer = make_peer(N_, 1, P, world)
            arena = GradArena(P, peer=peer)
            with arena.sink(), camera_parallel(peer=peer) as cp:
                torch.autograd.backward([rc, ra], [vc_all[cams], va_all[cams]])
            arena.gather_from_params()
            arena.all_reduce(skip_ptrs=cp.reduced_ptrs)
            arena.scatter_to_params()
        else:
            torch.autograd.backward([rc, ra], [vc_all[cams], va_all[cams]])
        return [p.grad.detach().clone() for p in P]

    g_dp, g_ref = grads([rank], True), grads(list(range(world)), False)
    worst = max(float((a - b).abs().max() / (b.abs().max() + 1e-30)) for a, b in zip(g_dp, g_ref))
    t = torch.tensor([worst], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def extra_config(S, dist, name, world, rank, dev, steps: int = 10, warm: int = 3):
    """The other multi-GPU configs of BASELINE.json as extra keys (not the headline): config C = 3 M
    Gaussians, a batch of `world` 1080p cameras (one per rank), dense gradient exchange; config E = 6 M
    Gaussians, 3840x2160, packed + sparse gradients, exchanged as (gaussian_ids, rows) with the dense
    fallback (splat_one_b200.distributed.allreduce_mixed_gradients)."""
    from splat_one_b200 import synthetic
    from splat_one_b200.distributed import GradArena, allreduce_mixed_gradients, camera_parallel

    if name == "C":
        n, W_, H_, kw = 3_000_000, 1920, 1080, dict(packed=False)
    else:
        n, W_, H_, kw = 6_000_000, 3840, 2160, dict(packed=True, sparse_grad=True)
    sc = synthetic.pinhole_scene(n, W_, H_, seed=43 if name == "C" else 45, n_cameras=world)
    names = ("means", "quats", "scales", "opacities", "sh")
    P = [sc[k].to(dev).requires_grad_() for k in names]
    vm, Ks = sc["viewmats"][rank::world].to(dev), sc["Ks"][rank::world].to(dev)
    g = torch.Generator().manual_seed(2000 + rank)
    vc = torch.randn(1, H_, W_, 3, generator=g).to(dev)
    va = torch.randn(1, H_, W_, 1, generator=g).to(dev)
    # C: cotangent slots + the arena in symmetric memory; E (packed + sparse): cotangent slots only
    peer = make_peer(n, 1, P, world, with_arena=name == "C")
    arena = GradArena(P, peer=peer) if (world > 1 and name == "C") else None
    counts = []

    def step():
        for p in P:
            p.grad = None
        rc, ra, meta = S.rasterization(*P, vm, Ks, W_, H_, sh_degree=SH_DEGREE, **kw)
        if arena is not None:
            with arena.sink(), camera_parallel(peer=peer) as cp:
                torch.autograd.backward([rc, ra], [vc, va])
            arena.gather_from_params()
            arena.all_reduce(skip_ptrs=cp.reduced_ptrs)
        elif world > 1:
            # packed colour stage: cotangent exchange (scattered to the dense layout) instead of the 1.15 GB SH
            # all-reduce; sparse (gaussian_ids, rows) exchange or its dense fallback for the projection gradients
            with camera_parallel(peer=peer) as cp:
                torch.autograd.backward([rc, ra], [vc, va])
            allreduce_mixed_gradients(P, skip_ptrs=cp.reduced_ptrs)
        else:
            torch.autograd.backward([rc, ra], [vc, va])
        return meta

    for _ in range(warm):
        meta = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        meta = step()
    b.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / steps
    out = {"ms_per_step": ms, "Mpix_per_s": world * H_ * W_ / (ms * 1e-3) / 1e6, "steps": steps, "warmup": warm,
           "n_gaussians": n, "image": f"{W_}x{H_}", "cameras": world, "n_isects_rank0": int(meta["flatten_ids"].numel()),
           "mode": ("unpacked, dense gradients; exchange = " + exchange_kind(peer)) if name == "C" else
                   "packed, sparse projection gradients: (gaussian_ids, rows) all-gather with dense fallback above 0.4 "
                   "visible; SH gradient through the colour-cotangent exchange = " + exchange_kind(peer)}
    del P, sc, arena, peer
    torch.cuda.empty_cache()
    return out


# A/B switch: overlapped gradient exchange (camera_parallel(defer=True) + finish()).  Measured at N = 2:
# 1.731 ms/step against 1.701 for the plain order (the collectives and the colour backward compete for
# HBM), so the plain order stays the default; not measured at N = 8.
DEFER = os.environ.get("B200SPLAT_DP_DEFER", "0") == "1"


def run_gpu(args):
    import torch.distributed as dist

    import splat_one_b200 as S
    from splat_one_b200 import synthetic, wrapper
    from splat_one_b200.distributed import GradArena, camera_parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # host side of a rank next to its GPU (pinned buffers are allocated below): matters for e2e at N = 8
    from splat_one_b200.distributed import bind_host_to_gpu
    host_cpus = bind_host_to_gpu(local_rank) if os.environ.get("B200SPLAT_NO_BIND", "0") != "1" else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # replicated Gaussians (same seed on every rank), one camera per rank (distinct poses)
    scene = synthetic.pinhole_scene(N_GAUSS, WIDTH, HEIGHT, seed=42, sh_degree=SH_DEGREE, n_cameras=world)
    names = ("means", "quats", "scales", "opacities", "sh")
    params = [scene[k].to(dev).requires_grad_() for k in names]
    vm_host = scene["viewmats"][rank::world].contiguous().pin_memory()
    K_host = scene["Ks"][rank::world].contiguous().pin_memory()
    vm, Ks = vm_host.to(dev), K_host.to(dev)
    C_local = vm.shape[0]
    g = torch.Generator().manual_seed(1000 + rank)
    vc_host = torch.randn(C_local, HEIGHT, WIDTH, 3, generator=g).pin_memory()
    va_host = torch.randn(C_local, HEIGHT, WIDTH, 1, generator=g).pin_memory()
    vc, va = vc_host.to(dev), va_host.to(dev)
    peer_note = {}
    peer = make_peer(N_GAUSS, C_local, params, world, peer_note)
    arena_nccl = GradArena(params) if world > 1 else None
    arena_peer = GradArena(params, peer=peer) if peer is not None else None

    def step(defer=DEFER, use_peer=True):
        for p in params:
            p.grad = None
        rc, ra, meta = S.rasterization(*params, vm, Ks, WIDTH, HEIGHT, sh_degree=SH_DEGREE, packed=False)
        skip = ()
        px = peer if use_peer else None
        arena = arena_peer if px is not None else arena_nccl
        if arena is not None:
            # SH / quats / scales gradients are produced inside the arena; the SH gradient comes
            # out already summed over ranks (colour-cotangent exchange, distributed.py)
            with arena.sink(), camera_parallel(defer=defer, peer=px) as cp:
                torch.autograd.backward([rc, ra], [vc, va])
            skip = cp.reduced_ptrs
        else:
            torch.autograd.backward([rc, ra], [vc, va])
        if arena is not None:
            prof = wrapper.profiler
            if defer:
                # all-gather overlapped the projection backward; the arena all-reduce overlaps the colour
                # backward kernel (splat_one_b200/distributed.py camera_parallel.finish)
                if prof.enabled:
                    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                    ev[0].record()
                cp.finish(arena)
                if prof.enabled:
                    ev[1].record()
                    prof.events.setdefault("grad_exchange_finish", []).append((ev[0], ev[1]))
                return rc, ra, meta
            if prof.enabled:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                ev[0].record()
            arena.gather_from_params()
            if prof.enabled:
                ev[1].record()
            arena.all_reduce(skip_ptrs=skip)
            if prof.enabled:
                ev[2].record()
                prof.events.setdefault("grad_gather", []).append((ev[0], ev[1]))
                prof.events.setdefault("grad_allreduce", []).append((ev[1], ev[2]))
        return rc, ra, meta

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        rc, ra, meta = step()
    barrier()

    # ---- timed region: device-resident inputs --------------------------------------------
    # inside the timed region only the two raster kernels (the dominant ones, `roofline` /
    # `raster_stages`) are bracketed with CUDA events; every native call is counted.  The full
    # per-stage table is measured in a separate pass below, outside the timed region.
    wrapper.profiler.reset()
    wrapper.profiler.enabled = True
    wrapper.profiler.only = {"rasterize_fwd", "rasterize_bwd"}
    sampler = ClockSampler(local_rank) if rank == 0 and os.environ.get('B200SPLAT_NO_SAMPLER', '0') != '1' else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        rc, ra, meta = step()
    e1.record()
    barrier()
    wrapper.profiler.enabled = False
    clocks = sampler.stop() if sampler else None
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    # A/B of the other exchanges (N > 1, outside `value`): a few steps each with the NCCL collectives in
    # plain order and in the overlapped (deferred) order
    ms_ab = {}
    if world > 1:
        n_ab = max(min(args.steps, 30), 5)
        for label, kw in (("peer_plain", dict(defer=False, use_peer=True)), ("peer_overlapped", dict(defer=True, use_peer=True)),
                          ("nccl_plain", dict(defer=False, use_peer=False)), ("nccl_deferred", dict(defer=True, use_peer=False))):
            if kw["use_peer"] == (peer is not None) and kw["defer"] == DEFER:
                continue  # that is the timed loop itself
            if kw["use_peer"] and peer is None:
                continue
            for _ in range(3):
                step(**kw)
            barrier()
            e0.record()
            for _ in range(n_ab):
                step(**kw)
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ab[label] = t.item() / n_ab
    launches = wrapper.profiler.launches()
    timed_stages = wrapper.profiler.summary_ms()
    # full stage table: a few more steps with every native call bracketed (not part of `value`)
    wrapper.profiler.reset()
    wrapper.profiler.only = None
    wrapper.profiler.enabled = True
    for _ in range(min(args.steps, 20)):
        step()
    torch.cuda.synchronize()
    wrapper.profiler.enabled = False
    stages = wrapper.profiler.summary_ms()
    stages.update(timed_stages)  # the raster kernels keep their in-region timings

    # realised sizes for the roofline denominators
    V = int((meta["radii"] > 0).sum())
    I = int(meta["flatten_ids"].numel())
    P = C_local * HEIGHT * WIDTH
    n_pairs = None

    # ---- e2e: host buffers in, image out, copies inside the timed region -----------------
    out_c_host = torch.empty(C_local, HEIGHT, WIDTH, 3).pin_memory()
    out_a_host = torch.empty(C_local, HEIGHT, WIDTH, 1).pin_memory()
    gnorm_host = torch.empty(1).pin_memory()

    h2d_stream = torch.cuda.Stream(device=dev)
    d2h_stream = torch.cuda.Stream(device=dev)
    # double-buffered device cotangents: the copy for step k+1 is issued during step k (a data
    # loader's prefetch), H2D and D2H run on their own streams (PCIe is full duplex)
    cot = [(torch.empty_like(vc), torch.empty_like(va)) for _ in range(2)]
    cot_ready = [torch.cuda.Event() for _ in range(2)]
    cot_free = [torch.cuda.Event() for _ in range(2)]
    state = {"k": 0}

    def prefetch(i):
        with torch.cuda.stream(h2d_stream):
            h2d_stream.wait_event(cot_free[i])  # the backward that last read this buffer is done
            cot[i][0].copy_(vc_host, non_blocking=True)
            cot[i][1].copy_(va_host, non_blocking=True)
            cot_ready[i].record(h2d_stream)

    def e2e_step(image_d2h=False):
        """camera H2D -> forward -> (next step's cotangent H2D [|| image D2H]) -> backward -> grad-norm D2H.
        Every step copies its camera and cotangent images host->device and reads its result (the
        gradient norm, as a trainer reads its loss) device->host; with `image_d2h` the rendered image
        and alpha also go back to the host every step (what a viewer does).  The big copies run on
        side streams and overlap kernels."""
        main = torch.cuda.current_stream(dev)
        i = state["k"] & 1
        state["k"] += 1
        vm_d = vm_host.to(dev, non_blocking=True)
        K_d = K_host.to(dev, non_blocking=True)
        for p in params:
            p.grad = None
        rc_, ra_, _ = S.rasterization(*params, vm_d, K_d, WIDTH, HEIGHT, sh_degree=SH_DEGREE, packed=False)
        fwd_done = torch.cuda.Event()
        fwd_done.record(main)
        if image_d2h:
            with torch.cuda.stream(d2h_stream):  # the rendered image leaves while the backward runs
                d2h_stream.wait_event(fwd_done)
                out_c_host.copy_(rc_.detach(), non_blocking=True)
                out_a_host.copy_(ra_.detach(), non_blocking=True)
                rc_.record_stream(d2h_stream)
                ra_.record_stream(d2h_stream)
        prefetch(i ^ 1)  # cotangents of the NEXT step
        main.wait_event(cot_ready[i])
        vc_d, va_d = cot[i]
        px = peer
        arena = arena_peer if px is not None else arena_nccl
        if arena is not None:
            with arena.sink(), camera_parallel(defer=DEFER, peer=px) as cp:
                torch.autograd.backward([rc_, ra_], [vc_d, va_d])
            if DEFER:
                cp.finish(arena)
            else:
                arena.gather_from_params()
                arena.all_reduce(skip_ptrs=cp.reduced_ptrs)
        else:
            torch.autograd.backward([rc_, ra_], [vc_d, va_d])
        cot_free[i].record(main)
        gnorm_host.copy_(params[0].grad.norm().reshape(1), non_blocking=True)

    def e2e_drain():
        main = torch.cuda.current_stream(dev)
        main.wait_stream(h2d_stream)
        main.wait_stream(d2h_stream)

    cot_free[0].record(torch.cuda.current_stream(dev))
    cot_free[1].record(torch.cuda.current_stream(dev))
    prefetch(0)
    # headline e2e: the step's RESULT — the rendered image and alpha, plus the gradient norm — goes back
    # to pinned host memory every step
    for _ in range(2):
        e2e_step(True)
    e2e_drain()
    barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_step(True)
    e2e_drain()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item() / args.steps
    h2d = vm_host.numel() * 4 + K_host.numel() * 4 + vc_host.numel() * 4 + va_host.numel() * 4
    d2h = out_c_host.numel() * 4 + out_a_host.numel() * 4 + 4
    # the same loop reading back only the gradient norm (what a trainer reads per step)
    n_small = max(args.steps // 2, 1)
    e2e_step(False)
    e2e_drain()
    barrier()
    e0.record()
    for _ in range(n_small):
        e2e_step(False)
    e2e_drain()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_small_ms = t.item() / n_small

    # host-link probe (outside every timed region): the e2e payloads alone, all ranks at once, one direction
    # at a time — what the host's PCIe path gives each rank when N ranks copy simultaneously
    def link_probe(direction):
        reps = 10
        barrier()
        e0.record()
        for _ in range(reps):
            if direction == "h2d":
                cot[0][0].copy_(vc_host, non_blocking=True)
                cot[0][1].copy_(va_host, non_blocking=True)
            else:
                out_c_host.copy_(vc, non_blocking=True)
                out_a_host.copy_(va, non_blocking=True)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return (vc_host.numel() + va_host.numel()) * 4 * reps / (t.item() * 1e-3) / 1e9

    host_link = {"h2d_GBps_per_rank": round(link_probe("h2d"), 2), "d2h_GBps_per_rank": round(link_probe("d2h"), 2),
                 "what": f"33 MB pinned copies, {world} rank(s) at once, min over ranks"}

    # ---- extra legs, all outside the timed regions -------------------------------------------
    dp_parity = dp_parity_check(S, dist, world, rank, dev) if world > 1 else None
    extras = {}
    if not args.no_extra_configs:
        # free the config-B scene first: config E needs several GB
        for name in ("C", "E"):
            try:
                extras[name] = extra_config(S, dist, name, world, rank, dev)
            except Exception as ex:  # never lose the headline line to an extra
                extras[name] = {"error": repr(ex)[:200]}

    if rank == 0:
        peak, peak_src = _peaks()
        # dominant kernel = the longest native call per step
        native_stages = {k: v for k, v in stages.items() if not k.startswith("grad_")}
        dom = max(native_stages.items(), key=lambda kv: kv[1]["total_ms"])[0] if native_stages else None
        N, C = N_GAUSS, C_local
        alg_bytes = {  # SURVEY.md §8(d), per launch
            "projection_fwd": 40 * N + 4 * C * N + 24 * V,
            "sh_fwd": 216 * V,
            "sh_colors_fwd": 216 * V + 4 * C * N,
            "sh_colors_bwd": 228 * V + 192 * N + 4 * C * N,
            "isect_sorted": 16 * V + 8 * C * N + 12 * I + 24 * I,
            "isect_depth_order": 16 * C * N + 12 * C * N,   # depth keys + order out, counts gathered + scanned
            "isect_tile_order": 16 * V + 12 * I + 24 * I,   # expand reads, ids + flatten_ids out, one pair sort
            "grad_gather": 2 * 236 * N,
            "grad_allreduce": 236 * N,
            "rasterize_pack": 36 * C * N + 48 * C * N,
            "isect_count": 4 * C * N + 8 * V + 4 * C * N + 8 * C * N,
            "isect_fill": 16 * V + 8 * C * N + 12 * I,
            "isect_sort": 24 * I,
            "isect_offset_encode": 8 * I + 4 * (meta["isect_offsets"].numel()),
            "rasterize_fwd": 40 * I + 20 * P,
            "rasterize_bwd": 40 * I + 24 * P + 72 * V,
            "sh_bwd": 228 * V + 192 * N,
            "projection_bwd": 40 * N + 4 * C * N + 36 * V + 40 * N,
        }
        roof = None
        if dom is not None:
            dur = stages[dom]["avg_ms"] * 1e-3
            ach = alg_bytes.get(dom, 0) / dur / 1e9
            traffic, traffic_src = None, None
            try:  # dram bytes per launch from the committed ncu --set full capture of the same kernel
                with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                    t = json.load(f).get(dom)
                if t:
                    traffic, traffic_src = t["traffic_bytes"], t["source"]
            except Exception:
                pass
            roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "algorithmic_bytes": alg_bytes.get(dom, 0), "avg_launch_ms": stages[dom]["avg_ms"]}
        # the north star's "raster stages": forward + backward kernels together
        raster = None
        if "rasterize_fwd" in stages and "rasterize_bwd" in stages:
            t_r = (stages["rasterize_fwd"]["avg_ms"] + stages["rasterize_bwd"]["avg_ms"]) * 1e-3
            b_r = alg_bytes["rasterize_fwd"] + alg_bytes["rasterize_bwd"]
            raster = {"kernels": "rasterize_fwd + rasterize_bwd", "algorithmic_bytes": b_r, "ms": t_r * 1e3,
                      "achieved": b_r / t_r / 1e9, "peak": peak, "unit": "GB/s", "frac": b_r / t_r / 1e9 / peak,
                      "bound": "fp32 issue (ncu: DRAM throughput ~3 % of peak, issue slots 69-78 % active; "
                               "profiles/r1_e_summary.md)"}
        stage_report = {k: {"avg_ms": round(v["avg_ms"], 4),
                            "GBps": round(alg_bytes.get(k, 0) / (v["avg_ms"] * 1e-3) / 1e9, 1) if v["avg_ms"] > 0 else None}
                        for k, v in stages.items()}
        train = ref_cuda_leg = None
        if world == 1:
            train = train_step_report(S, scene, dev)
            ref_cuda_leg = reference_cuda_leg(S, scene, dev, ms_step)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            mpix, ms, cores, sample, _, _ = cpu_baseline(3, 1, budget_s=40.0)
            cpu = {"value": mpix, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": "rendered Mpix/s fwd+bwd", "value": world * C_local * HEIGHT * WIDTH / (ms_step * 1e-3) / 1e6,
            "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_gaussians": N_GAUSS, "cameras_per_gpu": C_local,
                       "visible_pairs": V, "n_isects": I, "tiles_per_visible_gaussian": round(I / max(V, 1), 3),
                       "listed_pairs_per_pixel": round(I * 256 / max(P, 1), 1),
                       "l2_policy": "per-step working set ~1 GB > 126 MB L2",
                       "parallelism": f"camera-sharded dp{world}" + (
                           "; gradient exchange (12 MB/rank colour cotangents, 44 MB arena) = " + exchange_kind(peer)
                           + (", colour backward overlapped with projection backward + all-reduce" if DEFER else "")
                           if world > 1 else "")},
            "clocks": clocks,
            "e2e": {"value": world * C_local * HEIGHT * WIDTH / (e2e_ms * 1e-3) / 1e6, "unit": "Mpix/s",
                    "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "every step: camera + cotangent images H2D from pinned memory (cotangents prefetched one "
                            "step ahead on an H2D stream), rasterization()+backward, and the step's result — rendered "
                            "image + alpha (D2H stream) and the gradient norm — copied to pinned host memory; "
                            "Gaussians stay resident (model state)",
                    "grad_norm_only": {"value": world * C_local * HEIGHT * WIDTH / (e2e_small_ms * 1e-3) / 1e6,
                                       "ms_per_step": e2e_small_ms, "d2h_bytes_per_step": 4,
                                       "what": "same, reading back only the gradient norm (trainer-style)"}},
            "gpu_launches": launches,
            "roofline": roof,
            "raster_stages": raster,
            "stages": stage_report,
            "train_step": train,
            "cpu_baseline": cpu,
            "reference_cuda": ref_cuda_leg,
            "dp_parity_max_rel": dp_parity,
            "host_link": host_link,
            "host_binding": (f"{len(host_cpus)} cores nearest to the GPU (NVML affinity)" if host_cpus else "none"),
            "exchange": {"timed": ("peer" if peer is not None else "nccl") + ("_overlapped" if DEFER else "_plain"),
                         "multicast": bool(peer is not None and peer.multicast_base),
                         "peer_unavailable": peer_note.get("error"),
                         "ms_per_step_other": ms_ab} if world > 1 else None,
            "config_C": extras.get("C"),
            "config_E": extras.get("E"),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the config C / E extra keys")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
