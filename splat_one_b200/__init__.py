"""splat_one_b200 — B200-native (sm_100a) rasterization hot path behind the gsplat API
that inuex35/splat_one calls (`gsplat.rasterization` and the five operators under it).

    from splat_one_b200 import rasterization            # == gsplat.rendering.rasterization
    from splat_one_b200 import (fully_fused_projection, isect_tiles, isect_offset_encode,
                                rasterize_to_pixels, spherical_harmonics)

The compute lives in libb200splat.so (hand-written CUDA, C ABI in include/b200splat.h);
there is no CPU / PyTorch fallback.
"""
from .optimizers import SelectiveAdam
from .rendering import rasterization
from .step import l1_ssim_loss, rasterize_splats, splat_activations
from .strategy import strategy_state_sink, update_state as update_strategy_state
from .wrapper import (
    accumulate,
    compute_relocation,
    fully_fused_projection,
    isect_offset_encode,
    isect_tiles,
    persp_proj,
    proj,
    quat_scale_to_covar_preci,
    rasterize_to_indices_in_range,
    rasterize_to_pixels,
    selective_adam_update,
    world_to_cam,
    spherical_harmonics,
    spherical_harmonics_table,
)

__version__ = "0.1.0"

__all__ = [
    "rasterization",
    "fully_fused_projection",
    "isect_tiles",
    "isect_offset_encode",
    "rasterize_to_pixels",
    "rasterize_to_indices_in_range",
    "accumulate",
    "quat_scale_to_covar_preci",
    "proj",
    "persp_proj",
    "world_to_cam",
    "selective_adam_update",
    "compute_relocation",
    "SelectiveAdam",
    "strategy_state_sink",
    "update_strategy_state",
    "spherical_harmonics",
    "spherical_harmonics_table",
    "rasterize_splats",
    "splat_activations",
    "l1_ssim_loss",
]
