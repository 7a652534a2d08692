/*
 * b200splat.h — C ABI of libb200splat.so, the sm_100a replacement for the native
 * entry points of the gsplat fork used by inuex35/splat_one.
 *
 * Every function is `extern "C"`, takes plain device pointers + explicit sizes + a
 * `cudaStream_t` (passed as `void*`), owns no memory and keeps no global state
 * (re-entrant: splat_one calls from a training thread and the Qt thread at once,
 * SURVEY.md §7 H7).  All floating tensors are fp32, row-major, densely packed.
 * Return value: 0 on success, non-zero on failure; the message is available from
 * `b200splat_last_error()` (thread-local).
 *
 * Path shorthand for the citations:  CS/ = /root/reference/submodules/gsplat/gsplat/cuda/csrc/
 * Each entry point names the reference pybind11 symbol (CS/ext.cpp:11-56) and the
 * C++ signature (CS/bindings.h) it replaces.
 *
 * Data-dependent sizes (`nnz`, `n_isects`) use two-phase calls: a `*_count` call writes
 * the total into device memory, the host reads one integer, allocates, and calls
 * `*_fill`.
 */
#ifndef B200SPLAT_H
#define B200SPLAT_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define B200SPLAT_API __attribute__((visibility("default")))
#else
#define B200SPLAT_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* CS/bindings.h:34-40 `enum CameraModelType` (pybind: CS/ext.cpp:4-9). */
enum b200splat_camera_model {
    B200SPLAT_PINHOLE = 0,
    B200SPLAT_ORTHO = 1,
    B200SPLAT_FISHEYE = 2,
    B200SPLAT_SPHERICAL = 3
};

/* ABI version of this header (bumped on any signature change). */
B200SPLAT_API int b200splat_abi_version(void);
/* Thread-local message of the last failing call on this thread ("" if none). */
B200SPLAT_API const char *b200splat_last_error(void);
/* Compile-time facts, for diagnostics: returns "sm_100a". */
B200SPLAT_API const char *b200splat_arch(void);

/* Copy n_words 32-bit words from device memory to PINNED host memory from inside a kernel (no
 * DMA engine): the read-back of n_isects / nnz between the `*_count` and `*_fill` halves of the
 * two-phase calls (replaces the blocking `.item()` of CS/isect_tiles.cu:201 and
 * CS/fully_fused_projection_packed_fwd.cu:352-353).  The host waits on an event recorded behind
 * this call on the same stream. */
B200SPLAT_API int b200splat_copy_small(const void *src, void *dst_pinned_host, uint32_t n_words, void *stream);

/* ------------------------------------------------------------------------------------
 * a2  fully_fused_projection_fwd        CS/bindings.h:95-116, kernel
 *     CS/fully_fused_projection_fwd.cu:20-216.
 * covars [N,6] XOR (quats [N,4], scales [N,3]).  compensations may be NULL.
 * Outputs [C,N,...]; EVERY entry is written: culled pairs get radii = 0 and zeros in the other outputs
 * (the reference leaves those uninitialised), so the caller needs no zero-fill pass.
 * ---------------------------------------------------------------------------------- */
B200SPLAT_API int b200splat_projection_fwd(
    uint32_t C, uint32_t N,
    const float *means, const float *covars, const float *quats, const float *scales,
    const float *viewmats, const float *Ks,
    uint32_t image_width, uint32_t image_height,
    float eps2d, float near_plane, float far_plane, float radius_clip,
    int camera_model,
    int32_t *radii, float *means2d, float *depths, float *conics, float *compensations,
    void *stream);

/* a3  fully_fused_projection_bwd        CS/bindings.h:118-146, kernel
 *     CS/fully_fused_projection_bwd.cu:20-271.
 * v_means/v_covars/v_quats/v_scales are OVERWRITTEN (no pre-zeroing needed): one thread
 * owns one Gaussian and loops over the C cameras, so there are no atomics and the
 * result is deterministic.  v_viewmats [C,4,4] (optional, may be NULL) must be zeroed by
 * the caller; it is accumulated with block-reduced atomics.
 * compensations / v_compensations may be NULL. */
B200SPLAT_API int b200splat_projection_bwd(
    uint32_t C, uint32_t N,
    const float *means, const float *covars, const float *quats, const float *scales,
    const float *viewmats, const float *Ks,
    uint32_t image_width, uint32_t image_height, float eps2d, int camera_model,
    const int32_t *radii, const float *conics, const float *compensations,
    const float *v_means2d, const float *v_depths, const float *v_conics,
    const float *v_compensations,
    float *v_means, float *v_covars, float *v_quats, float *v_scales, float *v_viewmats,
    void *stream);

/* f3  DefaultStrategy._update_state     G/strategy/default.py:239-262 (running densification
 *     statistics), folded into the projection backward: besides the gradients of a3, for every Gaussian n
 *       state_grad2d[n] += sum_c [radii[c,n] > 0] * |(v_means2d[c,n].x * grad_scale_x, .y * grad_scale_y)|
 *       state_count[n]  += sum_c [radii[c,n] > 0]
 *       state_radii[n]   = max(state_radii[n], max_c radii[c,n] / max_wh)      (state_radii may be NULL)
 *     with grad_scale = (W/2, H/2) * n_cameras and max_wh = max(W,H) as in :220-226, :258-261.
 *     v_means2d is the cotangent this call receives, i.e. the `means2d.grad` the reference retains. */
B200SPLAT_API int b200splat_projection_bwd_state(
    uint32_t C, uint32_t N,
    const float *means, const float *covars, const float *quats, const float *scales,
    const float *viewmats, const float *Ks,
    uint32_t image_width, uint32_t image_height, float eps2d, int camera_model,
    const int32_t *radii, const float *conics, const float *compensations,
    const float *v_means2d, const float *v_depths, const float *v_conics,
    const float *v_compensations,
    float *v_means, float *v_covars, float *v_quats, float *v_scales, float *v_viewmats,
    float grad_scale_x, float grad_scale_y, float max_wh,
    float *state_grad2d, float *state_count, float *state_radii, void *stream);

/* The same update as one stand-alone kernel, for the cases the fused form does not cover: packed layout
 * (gaussian_ids != NULL: grads [nnz,2], radii [nnz], atomics per visible pair) and `absgrad` statistics
 * (grads = means2d.absgrad).  Unpacked: grads [C,N,2], radii [C,N], nnz ignored. */
B200SPLAT_API int b200splat_strategy_update_state(
    uint32_t C, uint32_t N, uint32_t nnz, const int64_t *gaussian_ids,
    const float *grads, const int32_t *radii,
    float grad_scale_x, float grad_scale_y, float max_wh,
    float *state_grad2d, float *state_count, float *state_radii, void *stream);

/* ------------------------------------------------------------------------------------
 * a4  fully_fused_projection_packed_fwd  CS/bindings.h:254-278, kernel
 *     CS/fully_fused_projection_packed_fwd.cu:20-267 (two launches + cumsum + .item()).
 * Phase 1 (`_count`): block_cnts[C*blocks_per_row] (blocks_per_row = ceil(N/256)),
 *   then an inclusive scan into block_accum (same length, int32) and the total into
 *   *nnz_out (device int32).  Host reads nnz_out.
 * Phase 2 (`_fill`): writes the COO outputs in (camera, gaussian) row-major order and
 *   indptr[C+1].
 * ---------------------------------------------------------------------------------- */
B200SPLAT_API int b200splat_projection_packed_count(
    uint32_t C, uint32_t N,
    const float *means, const float *covars, const float *quats, const float *scales,
    const float *viewmats, const float *Ks,
    uint32_t image_width, uint32_t image_height,
    float eps2d, float near_plane, float far_plane, float radius_clip, int camera_model,
    int32_t *block_accum, /* [C*ceil(N/256)] out: inclusive scan of per-block counts */
    int32_t *nnz_out,     /* [1] device */
    void *stream);

B200SPLAT_API int b200splat_projection_packed_fill(
    uint32_t C, uint32_t N,
    const float *means, const float *covars, const float *quats, const float *scales,
    const float *viewmats, const float *Ks,
    uint32_t image_width, uint32_t image_height,
    float eps2d, float near_plane, float far_plane, float radius_clip, int camera_model,
    const int32_t *block_accum,
    int32_t *indptr, int64_t *camera_ids, int64_t *gaussian_ids,
    int32_t *radii, float *means2d, float *depths, float *conics, float *compensations,
    void *stream);

/* a4  fully_fused_projection_packed_bwd  CS/bindings.h:280-310, kernel
 *     CS/fully_fused_projection_packed_bwd.cu:20-310.
 * sparse_grad != 0: v_* are [nnz,·] rows, overwritten.  sparse_grad == 0: v_* are
 * [N,·], must be zeroed by the caller, accumulated with atomics (rows of one Gaussian
 * from different cameras are not adjacent in COO order). v_viewmats as in a3. */
B200SPLAT_API int b200splat_projection_packed_bwd(
    uint32_t C, uint32_t N, uint32_t nnz,
    const float *means, const float *covars, const float *quats, const float *scales,
    const float *viewmats, const float *Ks,
    uint32_t image_width, uint32_t image_height, float eps2d, int camera_model,
    const int64_t *camera_ids, const int64_t *gaussian_ids,
    const float *conics, const float *compensations,
    const float *v_means2d, const float *v_depths, const float *v_conics,
    const float *v_compensations,
    int sparse_grad,
    float *v_means, float *v_covars, float *v_quats, float *v_scales, float *v_viewmats,
    void *stream);

/* ------------------------------------------------------------------------------------
 * a5  compute_sh_fwd / compute_sh_bwd    CS/bindings.h:235-249, kernels
 *     CS/compute_sh_fwd.cu:12-38, CS/compute_sh_bwd.cu:13-49, math
 *     CS/spherical_harmonics.cuh:17-366.
 * n_elems directions; coeffs has `n_coeff_rows` rows of [K,3]; element e reads row
 * (e % n_coeff_rows) — n_coeff_rows == n_elems is the reference layout, n_coeff_rows ==
 * N with n_elems == C*N evaluates a [N,K,3] table for C cameras without materialising
 * the `expand(C,...)` copy (rendering.py:386, _wrapper.py:71-73).
 * masks (uint8/bool, may be NULL): masked elements are left untouched in `colors`.
 * Unlike the reference (legacy default stream, CS/compute_sh_fwd.cu:58-61) the kernels
 * run on `stream`.
 * bwd: v_coeffs [n_elems,K,3] is fully OVERWRITTEN (zeros for masked rows and for bases
 * above the active degree); v_dirs (may be NULL) [n_elems,3] fully overwritten. */
B200SPLAT_API int b200splat_sh_fwd(
    uint32_t n_elems, uint32_t n_coeff_rows, uint32_t K, uint32_t degrees_to_use,
    const float *dirs, const float *coeffs, const uint8_t *masks,
    float *colors, void *stream);

B200SPLAT_API int b200splat_sh_bwd(
    uint32_t n_elems, uint32_t n_coeff_rows, uint32_t K, uint32_t degrees_to_use,
    const float *dirs, const float *coeffs, const uint8_t *masks,
    const float *v_colors,
    float *v_coeffs, float *v_dirs, void *stream);

/* Fused colour stage of rasterization() for the unpacked layout (G/rendering.py:368-392:
 * torch.inverse(viewmats)[:, :3, 3]; dirs = means - campos; spherical_harmonics(...,
 * masks = radii > 0); clamp_min(colors + 0.5, 0)) as one kernel per direction.
 *  camera_centers: campos [C,3] = inverse(viewmats)[:, :3, 3] (general 4x4).
 *  sh_colors_fwd : colors [C,N,3] (zeros where radii <= 0).  per_view != 0: coeffs is
 *                  [C,N,K,3], else one [N,K,3] table shared by the cameras.
 *  sh_colors_bwd : v_coeffs ([N,K,3] or [C,N,K,3]) and v_means [N,3] (may be NULL), both
 *                  fully overwritten; sums over cameras are done in registers.  Only cameras
 *                  in [means_cam_begin, means_cam_end) contribute to v_means.  radii == NULL
 *                  and colors == NULL: v_colors is pre-masked (zero for invisible Gaussians
 *                  and clamped channels) — the form in which camera-parallel ranks exchange
 *                  their colour cotangents instead of all-reducing the K-times larger
 *                  coefficient gradient (splat_one_b200/distributed.py). */
B200SPLAT_API int b200splat_camera_centers(uint32_t C, const float *viewmats, float *campos, void *stream);

B200SPLAT_API int b200splat_sh_colors_fwd(
    uint32_t C, uint32_t N, uint32_t K, uint32_t degrees_to_use, int per_view,
    const float *means, const float *campos, const float *coeffs, const int32_t *radii,
    float *colors, void *stream);

B200SPLAT_API int b200splat_sh_colors_bwd(
    uint32_t C, uint32_t N, uint32_t K, uint32_t degrees_to_use, int per_view,
    const float *means, const float *campos, const float *coeffs, const int32_t *radii,
    const float *colors, const float *v_colors,
    float *v_coeffs, float *v_means,
    uint32_t means_cam_begin, uint32_t means_cam_end, void *stream);

/* Packed (COO) variants of the fused colour stage (packed=True branch of G/rendering.py:
 * 370-392: `means[gaussian_ids] - campos[camera_ids]`, `colors[gaussian_ids]`, SH, +0.5, clamp).
 * Rows i = 0..nnz-1 name (camera_ids[i], gaussian_ids[i]) (int64, as the packed projection
 * emits them).  colors [nnz,3].  bwd: v_coeffs ([N,K,3] or [C,N,K,3]) and v_means [N,3] (may be
 * NULL) must be ZEROED by the caller; rows of invisible Gaussians stay zero, rows seen by
 * several cameras are summed with atomics (direct stores when C == 1 or per_view). */
B200SPLAT_API int b200splat_sh_colors_packed_fwd(
    uint32_t nnz, uint32_t C, uint32_t N, uint32_t K, uint32_t degrees_to_use, int per_view,
    const float *means, const float *campos, const float *coeffs,
    const int64_t *camera_ids, const int64_t *gaussian_ids,
    float *colors, void *stream);

B200SPLAT_API int b200splat_sh_colors_packed_bwd(
    uint32_t nnz, uint32_t C, uint32_t N, uint32_t K, uint32_t degrees_to_use, int per_view,
    const float *means, const float *campos, const float *coeffs,
    const int64_t *camera_ids, const int64_t *gaussian_ids,
    const float *colors, const float *v_colors,
    float *v_coeffs, float *v_means, void *stream);

/* ------------------------------------------------------------------------------------
 * a6  isect_tiles                         CS/bindings.h:148-160, kernel
 *     CS/isect_tiles.cu:17-105, host :107-307.
 * Phase 1 (`_count`): tiles_per_gauss [n_elems] int32, cum_tiles [n_elems] int64
 *   (inclusive scan) and n_isects_out (device int64[2]: [0] = n_isects, [1] = 1 if some
 *   visible depth has its sign bit set, which sign-extends into the tile/camera fields of
 *   the reference key and is routed to the generic sort; depths may be NULL).
 *   cum_tiles may be NULL: only tiles_per_gauss and the total n_isects are produced then (no
 *   scan; the depth-first ordering below scans the counts itself, in depth order).
 * Phase 2 (`_fill`): unsorted keys `cam | tile | depth bits` + flat indices.
 * Phase 3 (`_sort`): stable LSD radix sort of (key,value) on bits [0, end_bit) over a
 *   pair of ping-pong buffers (the reference's cub::DoubleBuffer, CS/isect_tiles.cu:262-299;
 *   here the library's own onesweep passes, csrc/sort.cu — no CUB);
 *   `*selector_out` says which buffer holds the result.  Workspace size from
 *   `_sort_workspace_bytes`.
 * packed != 0: n_elems = nnz and camera_ids [nnz] int64 gives the camera of each row;
 * packed == 0: n_elems = C*N, camera = idx / N.
 * ---------------------------------------------------------------------------------- */
B200SPLAT_API int b200splat_isect_count(
    int packed, uint32_t C, uint32_t N, uint32_t nnz,
    const float *means2d, const int32_t *radii, const float *depths,
    uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
    int32_t *tiles_per_gauss, int64_t *cum_tiles, int64_t *n_isects_out,
    void *scan_workspace, size_t scan_workspace_bytes,
    void *stream);

B200SPLAT_API size_t b200splat_scan_workspace_bytes(uint64_t n_elems);

B200SPLAT_API int b200splat_isect_fill(
    int packed, uint32_t C, uint32_t N, uint32_t nnz,
    const int64_t *camera_ids,
    const float *means2d, const int32_t *radii, const float *depths,
    const int64_t *cum_tiles,
    uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
    int64_t *isect_ids, int32_t *flatten_ids,
    void *stream);

B200SPLAT_API size_t b200splat_sort_workspace_bytes(uint64_t n_isects);

B200SPLAT_API int b200splat_isect_sort(
    uint64_t n_isects, uint32_t end_bit,
    int64_t *keys_a, int32_t *vals_a,   /* input, clobbered */
    int64_t *keys_b, int32_t *vals_b,   /* alternate buffers */
    void *workspace, size_t workspace_bytes,
    int *selector_out,                  /* HOST int: 0 => result in *_a, 1 => in *_b */
    void *stream);

/* Phases 2+3 in one call for non-negative depths — the B200 path used by
 * rasterization(): Gaussians are ordered by depth first (n_elems 32-bit keys), expanded
 * into their tiles in that order, and only the cam|tile bits are sorted at intersection
 * scale; the output is bit-identical to `_fill` + `_sort` (see csrc/sort.cu).
 * `offsets` (optional, may be NULL): [C*n_tiles] int32, the result of
 * `b200splat_isect_offset_encode` on the produced ids, written by the same final pass
 * (requires n_isects > 0; the caller zero-fills for n_isects == 0). */
B200SPLAT_API size_t b200splat_isect_sorted_workspace_bytes(uint64_t n_elems, uint64_t n_isects);

B200SPLAT_API int b200splat_isect_sorted(
    int packed, uint32_t C, uint32_t N, uint32_t nnz,
    const int64_t *camera_ids,
    const float *means2d, const int32_t *radii, const float *depths,
    const int32_t *tiles_per_gauss, uint64_t n_isects,
    uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
    int64_t *isect_ids, int32_t *flatten_ids, int32_t *offsets,
    void *workspace, size_t workspace_bytes,
    void *stream);

/* The two phases of b200splat_isect_sorted as separate calls.  Phase 1 (depth order of the
 * (camera, Gaussian) pairs + scan of their tile counts in that order) needs only n_elems, so
 * a host binding launches it BEFORE it reads n_isects back from b200splat_isect_count and the
 * round trip of that one host sync (CS/isect_tiles.cu:201) is hidden behind device work.
 * `selector_out` (host) says which half of the ping-pong buffers holds the order; pass it and
 * the same workspace to phase 2 as `depth_workspace` / `depth_selector`. */
B200SPLAT_API size_t b200splat_isect_depth_order_workspace_bytes(uint64_t n_elems);

B200SPLAT_API int b200splat_isect_depth_order(
    uint64_t n_elems, const float *depths, const int32_t *tiles_per_gauss,
    void *workspace, size_t workspace_bytes, int *selector_out, void *stream);

B200SPLAT_API size_t b200splat_isect_tile_order_workspace_bytes(uint64_t n_isects);

B200SPLAT_API int b200splat_isect_tile_order(
    int packed, uint32_t C, uint32_t N, uint32_t nnz, const int64_t *camera_ids,
    const float *means2d, const int32_t *radii, const float *depths,
    const void *depth_workspace, int depth_selector, uint64_t n_isects,
    uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
    int64_t *isect_ids, int32_t *flatten_ids, int32_t *offsets,
    void *workspace, size_t workspace_bytes, void *stream);

/* a7  isect_offset_encode                 CS/bindings.h:162-167, kernel
 *     CS/isect_tiles.cu:309-355.  offsets [C*n_tiles] int32 fully written
 *     (all zeros when n_isects == 0). */
B200SPLAT_API int b200splat_isect_offset_encode(
    uint64_t n_isects, const int64_t *isect_ids,
    uint32_t C, uint32_t tile_width, uint32_t tile_height,
    int32_t *offsets, void *stream);

/* ------------------------------------------------------------------------------------
 * a8  rasterize_to_pixels_fwd            CS/bindings.h:169-185, kernel
 *     CS/rasterize_to_pixels_fwd.cu:16-186.
 * n_gauss = C*N (unpacked) or nnz (packed): number of rows of means2d/conics/colors/
 * opacities.  channels: any 1..33 (the Python wrapper chunks above that).
 * backgrounds [C,channels] / masks [C,tile_h,tile_w] (uint8/bool) may be NULL.
 *
 * `records` (optional, may be NULL): a packed 48-byte-per-Gaussian table built by
 * `b200splat_rasterize_pack` (size from `b200splat_rasterize_records_bytes`, which is 0
 * when the fast path does not apply: it needs tile_size == 16 and channels <= 4).  With
 * records the warp-per-tile kernels run (same results); without, the generic kernels.
 * The same table serves the forward and the backward call.
 *
 * `quad_masks` (optional, may be NULL; only used together with `records`): [n_isects] bytes of
 * scratch.  The forward call stores, for every (tile, Gaussian) list entry it stages, which 8x8
 * quads of the tile the Gaussian can reach with alpha >= 1/255; the backward call of the same
 * (records, offsets, flatten_ids) reads them instead of repeating the exact rectangle test.
 * ---------------------------------------------------------------------------------- */
B200SPLAT_API size_t b200splat_rasterize_records_bytes(uint32_t n_gauss, uint32_t channels, uint32_t tile_size);

B200SPLAT_API int b200splat_rasterize_pack(
    uint32_t n_gauss, uint32_t channels,
    const float *means2d, const float *conics, const float *colors, const float *opacities,
    void *records, void *stream);

B200SPLAT_API int b200splat_rasterize_fwd(
    uint32_t C, uint32_t n_gauss, uint64_t n_isects, uint32_t channels,
    const float *means2d, const float *conics, const float *colors, const float *opacities,
    const float *backgrounds, const uint8_t *masks,
    uint32_t image_width, uint32_t image_height,
    uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
    const int32_t *tile_offsets, const int32_t *flatten_ids,
    const void *records, uint8_t *quad_masks,
    float *render_colors, float *render_alphas, int32_t *last_ids,
    void *stream);

/* a9  rasterize_to_pixels_bwd            CS/bindings.h:187-216, kernel
 *     CS/rasterize_to_pixels_bwd.cu:16-277.
 * v_means2d, v_conics, v_colors, v_opacities (and v_means2d_abs if not NULL) must be
 * zeroed by the caller; they are accumulated with reduced atomics. */
B200SPLAT_API int b200splat_rasterize_bwd(
    uint32_t C, uint32_t n_gauss, uint64_t n_isects, uint32_t channels,
    const float *means2d, const float *conics, const float *colors, const float *opacities,
    const float *backgrounds, const uint8_t *masks,
    uint32_t image_width, uint32_t image_height,
    uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
    const int32_t *tile_offsets, const int32_t *flatten_ids,
    const void *records, const uint8_t *quad_masks,
    const float *render_alphas, const int32_t *last_ids,
    const float *v_render_colors, const float *v_render_alphas,
    float *v_means2d_abs, float *v_means2d, float *v_conics, float *v_colors,
    float *v_opacities,
    void *stream);

/* ------------------------------------------------------------------------------------
 * f1  rasterize_to_indices_in_range      CS/bindings.h (rasterize_to_indices_in_range_tensor),
 *     kernel + host CS/rasterize_to_indices_in_range.cu:17-307; Python
 *     gsplat/cuda/_wrapper.py:571-643.
 * Lists, per pixel in (camera, row, column) order and front to back, the Gaussians of the
 * list batches [range_start, range_end) (a batch = tile_size^2 entries of the tile's list)
 * that pass the alpha test before the pixel's transmittance (starting from
 * `transmittances` [C,H,W]) would drop to <= 1e-4.  Two-phase: `_count` writes per-pixel
 * counts `chunk_cnts` [C*H*W] int32, their inclusive prefix sums `chunk_cum` [C*H*W]
 * int64 and the total `*n_elems_out` (device); the host reads the total, allocates
 * `gaussian_ids` / `pixel_ids` [n_elems] int64 and calls `_fill` with the same arguments.
 * pixel_ids are global: camera * H * W + row * W + column; gaussian_ids are in [0, N).
 * ---------------------------------------------------------------------------------- */
B200SPLAT_API int b200splat_raster_indices_count(
    uint32_t range_start, uint32_t range_end, uint32_t C, uint32_t N, uint64_t n_isects,
    const float *means2d, const float *conics, const float *opacities,
    uint32_t image_width, uint32_t image_height,
    uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
    const int32_t *tile_offsets, const int32_t *flatten_ids, const float *transmittances,
    int32_t *chunk_cnts, int64_t *chunk_cum, int64_t *n_elems_out,
    void *scan_workspace, size_t scan_workspace_bytes, void *stream);

B200SPLAT_API int b200splat_raster_indices_fill(
    uint32_t range_start, uint32_t range_end, uint32_t C, uint32_t N, uint64_t n_isects,
    const float *means2d, const float *conics, const float *opacities,
    uint32_t image_width, uint32_t image_height,
    uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
    const int32_t *tile_offsets, const int32_t *flatten_ids, const float *transmittances,
    const int32_t *chunk_cnts, const int64_t *chunk_cum,
    int64_t *gaussian_ids, int64_t *pixel_ids, void *stream);

/* ------------------------------------------------------------------------------------
 * f2  the un-fused exported operators of the projection chain (gsplat/cuda/_wrapper.py:76-200).
 *
 * quat_scale_to_covar_preci_fwd/bwd   CS/bindings.h (quat_scale_to_covar_preci_*_tensor),
 *     kernels CS/quat_scale_to_covar_preci_{fwd,bwd}.cu.  quats [N,4] (wxyz, un-normalised,
 *     16-byte aligned), scales [N,3]; covars / precis: [N,6] upper triangle if triu else
 *     [N,3,3]; either may be NULL (not computed / no cotangent).
 * world_to_cam_fwd/bwd                CS/world_to_cam_{fwd,bwd}.cu.  means [N,3], covars
 *     [N,3,3], viewmats [C,4,4] -> means_c [C,N,3], covars_c [C,N,3,3].  bwd: v_means /
 *     v_covars are fully overwritten (no zero-init needed); v_viewmats [C,4,4] is
 *     accumulated with atomics and must be zeroed by the caller; any of the three may be
 *     NULL, and so may either cotangent.
 * proj_fwd/bwd                        CS/proj_{fwd,bwd}.cu.  means [C,N,3], covars [C,N,3,3],
 *     Ks [C,3,3] -> means2d [C,N,2], covars2d [C,N,2,2] for the given camera model.
 * ---------------------------------------------------------------------------------- */
B200SPLAT_API int b200splat_quat_scale_to_covar_preci_fwd(
    uint32_t N, const float *quats, const float *scales, int triu,
    float *covars, float *precis, void *stream);

B200SPLAT_API int b200splat_quat_scale_to_covar_preci_bwd(
    uint32_t N, const float *quats, const float *scales,
    const float *v_covars, const float *v_precis, int triu,
    float *v_quats, float *v_scales, void *stream);

B200SPLAT_API int b200splat_world_to_cam_fwd(
    uint32_t C, uint32_t N, const float *means, const float *covars, const float *viewmats,
    float *means_c, float *covars_c, void *stream);

B200SPLAT_API int b200splat_world_to_cam_bwd(
    uint32_t C, uint32_t N, const float *means, const float *covars, const float *viewmats,
    const float *v_means_c, const float *v_covars_c,
    float *v_means, float *v_covars, float *v_viewmats, void *stream);

B200SPLAT_API int b200splat_proj_fwd(
    uint32_t C, uint32_t N, const float *means, const float *covars, const float *Ks,
    uint32_t image_width, uint32_t image_height, int camera_model,
    float *means2d, float *covars2d, void *stream);

B200SPLAT_API int b200splat_proj_bwd(
    uint32_t C, uint32_t N, const float *means, const float *covars, const float *Ks,
    uint32_t image_width, uint32_t image_height, int camera_model,
    const float *v_means2d, const float *v_covars2d,
    float *v_means, float *v_covars, void *stream);

/* ------------------------------------------------------------------------------------
 * f3  optimizer / densifier-side kernels fed by the path's outputs.
 *
 * selective_adam_update   CS/adam.cu:16-82 (gsplat/cuda/_wrapper.py:19-34): Adam moments
 *     and step for the M elements of every Gaussian whose `visible` byte (bool) is set;
 *     no bias correction, exactly as the reference.  param/exp_avg/exp_avg_sq [N*M] are
 *     updated in place.
 * compute_relocation      CS/compute_relocation.cu:6-70 (gsplat/relocation.py:10-55):
 *     ratios int32 [N] already clamped to [1, n_max]; binoms fp32 [n_max, n_max], n_max <= 64.
 * ---------------------------------------------------------------------------------- */
B200SPLAT_API int b200splat_selective_adam_update(
    float *param, const float *param_grad, float *exp_avg, float *exp_avg_sq,
    const uint8_t *visible, float lr, float b1, float b2, float eps,
    uint32_t N, uint32_t M, void *stream);

B200SPLAT_API int b200splat_compute_relocation(
    uint32_t N, const float *opacities, const float *scales, const int32_t *ratios,
    const float *binoms, int n_max, float *new_opacities, float *new_scales, void *stream);

/* ------------------------------------------------------------------------------------
 * f4  the training step either side of rasterization()
 *     (R/utils/gsplat_utils/gsplat_trainer.py; R = splat_one checkout).
 *
 * sh_colors_staged_fwd/bwd   the fused colour stage of rasterization() (rendering.py:368-392,
 *     as b200splat_sh_colors_fwd/bwd with per_view = 0) for ONE shared coefficient table
 *     given either whole (`sh0 == NULL`, `rest` = coeffs [N,K,3]) or as the two tensors
 *     splat_one optimises (`sh0` [N,1,3], `rest` = shN [N,K-1,3]) — which removes the
 *     `torch.cat([sh0, shN], 1)` of gsplat_trainer.py:474 and its split backward.  Coefficient
 *     rows move through shared memory (coalesced 128-bit global accesses); dynamic shared
 *     memory per block = b200splat_sh_colors_staged_smem_bytes(K, split).  bwd conventions
 *     (radii / colors == NULL: pre-masked cotangents; means camera window) as sh_colors_bwd;
 *     v_sh0 [N,1,3] / v_rest are fully overwritten.
 * splat_activations_fwd/bwd  scales = exp(scales_raw) [N,3], opacities = sigmoid(opacities_raw)
 *     [N] (gsplat_trainer.py:458-459) and their VJPs (either v_*_raw may be NULL).
 * l1_ssim_fwd/bwd            loss = l1_loss(img, target) (1 - lambda) + (1 - ssim) lambda with
 *     ssim = fused_ssim(img, target, padding="valid") (gsplat_trainer.py:624-628; fused_ssim is
 *     an un-vendored third-party dependency: standard SSIM, 11x11 Gaussian window sigma 1.5,
 *     C1 = 0.01^2, C2 = 0.03^2, zero-padded statistics, mean over interior pixels).  img /
 *     target / v_img / d_* are [C,H,W,3] fp32 (the layout rasterization() returns; no NCHW
 *     copy).  fwd writes out3 = {loss, l1 mean, ssim mean} (device) and, when d_mu != NULL,
 *     the three derivative maps the backward convolves; v_loss: device scalar or NULL (= 1).
 *     workspace: b200splat_l1_ssim_workspace_bytes(C, H, W) bytes.
 * ---------------------------------------------------------------------------------- */
B200SPLAT_API size_t b200splat_sh_colors_staged_smem_bytes(uint32_t K, int split);

B200SPLAT_API int b200splat_sh_colors_staged_fwd(
    uint32_t C, uint32_t N, uint32_t K, uint32_t degrees_to_use,
    const float *means, const float *campos, const float *sh0, const float *rest,
    const int32_t *radii, float *colors, void *stream);

B200SPLAT_API int b200splat_sh_colors_staged_bwd(
    uint32_t C, uint32_t N, uint32_t K, uint32_t degrees_to_use,
    const float *means, const float *campos, const float *sh0, const float *rest,
    const int32_t *radii, const float *colors, const float *v_colors,
    float *v_sh0, float *v_rest, float *v_means,
    uint32_t means_cam_begin, uint32_t means_cam_end, void *stream);

/* ------------------------------------------------------------------------------------
 * e   camera-parallel peer exchange (no counterpart in the reference, whose
 *     multi-GPU mode is Gaussian-sharded: G/rendering.py:397-478).  The colour backward of C =
 *     W * cams_per_block cameras reads the pre-masked colour cotangents and camera centres of camera
 *     block b IN PLACE from the memory of rank b over NVLink: peer_bases is a DEVICE array of W base
 *     addresses mapped into this process (torch symmetric memory: buffer_ptrs_dev); block b lives at
 *     peer_bases[b] + offset_bytes = { campos [cams_per_block][3] fp32, padding up to hdr_floats
 *     floats, v_colors [cams_per_block][N][3] fp32 }.  The caller orders the peers' writes before this
 *     call (symmetric-memory barrier on the same stream).  Otherwise as the non-peer calls with
 *     radii == colors == NULL: coefficient gradient summed over all C cameras, direction gradient over
 *     cameras [means_cam_begin, means_cam_end).
 * ---------------------------------------------------------------------------------- */
B200SPLAT_API int b200splat_sh_colors_bwd_peer(
    uint32_t C, uint32_t N, uint32_t K, uint32_t degrees_to_use,
    const float *means, const float *coeffs,
    const void *peer_bases, uint64_t offset_bytes, uint32_t cams_per_block, uint32_t hdr_floats,
    float *v_coeffs, float *v_means,
    uint32_t means_cam_begin, uint32_t means_cam_end, void *stream);

B200SPLAT_API int b200splat_sh_colors_staged_bwd_peer(
    uint32_t C, uint32_t N, uint32_t K, uint32_t degrees_to_use,
    const float *means, const float *sh0, const float *rest,
    const void *peer_bases, uint64_t offset_bytes, uint32_t cams_per_block, uint32_t hdr_floats,
    float *v_sh0, float *v_rest, float *v_means,
    uint32_t means_cam_begin, uint32_t means_cam_end, void *stream);

/* The exchange itself, as kernels over the same symmetric buffers (csrc/peer.cu).  flag_bases /
 * peer_bases: DEVICE arrays of W mapped base addresses (one symmetric buffer per rank, the same
 * offsets valid in each); the flag region (b200splat_peer_flag_bytes(W) bytes at flag_offset_bytes,
 * zero-initialised once by the caller) carries the release/acquire handshakes.  All ranks must issue
 * the same sequence of these calls.
 *   publish:   campos [C,3] and the cotangents masked by colors > 0 ([C,N,3]; colors == NULL: already
 *              masked) -> this rank's block
 *              (layout above; camera slots C..cams_per_block-1 are zero-filled);
 *   barrier:   returns (in stream order) once every rank has reached it: the peers' blocks are readable;
 *   allreduce: in-place SUM over ranks of n_floats fp32 at offset_bytes of every rank's buffer, two-shot
 *              (rank r reduces and re-broadcasts slice r).  multicast_base != 0: the NVSwitch multicast
 *              alias of the W buffers (multimem.ld_reduce / multimem.st); 0: plain peer loads and stores. */
B200SPLAT_API size_t b200splat_peer_flag_bytes(uint32_t world);
B200SPLAT_API int b200splat_peer_publish_cotangents(
    uint32_t C, uint32_t N, uint32_t cams_per_block, uint32_t hdr_floats,
    const float *campos, const float *colors, const float *v_colors, float *block, void *stream);
B200SPLAT_API int b200splat_peer_barrier(
    uint32_t world, uint32_t rank, const void *flag_bases, uint64_t flag_offset_bytes, void *stream);
B200SPLAT_API int b200splat_peer_allreduce_f32(
    uint32_t world, uint32_t rank, const void *peer_bases, uint64_t multicast_base,
    uint64_t offset_bytes, uint64_t n_floats, const void *flag_bases, uint64_t flag_offset_bytes,
    void *stream);

/* viewmats = inverse(camtoworlds) [C,4,4] (gsplat_trainer.py:483), adjugate in double, no
 * host synchronisation (torch.linalg.inv syncs to report singular inputs). */
B200SPLAT_API int b200splat_invert_4x4(uint32_t C, const float *mats, float *out, void *stream);

/* Packed (COO) colour stage for the split table: as b200splat_sh_colors_packed_fwd/bwd with
 * per_view = 0, coefficients read from sh0 [N,1,3] / shN [N,K-1,3].  v_sh0 / v_shN accumulate:
 * zero-fill them first. */
B200SPLAT_API int b200splat_sh_colors_packed_split_fwd(
    uint32_t nnz, uint32_t C, uint32_t N, uint32_t K, uint32_t degrees_to_use,
    const float *means, const float *campos, const float *sh0, const float *shN,
    const int64_t *camera_ids, const int64_t *gaussian_ids, float *colors, void *stream);

B200SPLAT_API int b200splat_sh_colors_packed_split_bwd(
    uint32_t nnz, uint32_t C, uint32_t N, uint32_t K, uint32_t degrees_to_use,
    const float *means, const float *campos, const float *sh0, const float *shN,
    const int64_t *camera_ids, const int64_t *gaussian_ids, const float *colors,
    const float *v_colors, float *v_sh0, float *v_shN, float *v_means, void *stream);

B200SPLAT_API int b200splat_splat_activations_fwd(
    uint32_t N, const float *scales_raw, const float *opacities_raw,
    float *scales, float *opacities, void *stream);

B200SPLAT_API int b200splat_splat_activations_bwd(
    uint32_t N, const float *scales, const float *opacities,
    const float *v_scales, const float *v_opacities,
    float *v_scales_raw, float *v_opacities_raw, void *stream);

B200SPLAT_API size_t b200splat_l1_ssim_workspace_bytes(uint32_t C, uint32_t H, uint32_t W);

B200SPLAT_API int b200splat_l1_ssim_fwd(
    uint32_t C, uint32_t H, uint32_t W, const float *img, const float *target, float ssim_lambda,
    float *d_mu, float *d_xx, float *d_xy, float *out3,
    void *workspace, size_t workspace_bytes, void *stream);

B200SPLAT_API int b200splat_l1_ssim_bwd(
    uint32_t C, uint32_t H, uint32_t W, const float *img, const float *target, float ssim_lambda,
    const float *d_mu, const float *d_xx, const float *d_xy, const float *v_loss,
    float *v_img, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SPLAT_H */
