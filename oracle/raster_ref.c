/* ORACLE (test infrastructure, not product code) — plain-C restatement of the tile
 * rasterizer of inuex35/splat_one's gsplat fork.
 *
 *   raster_fwd_ref : CS/rasterize_to_pixels_fwd.cu:16-186 (per pixel, front to back)
 *   raster_bwd_ref : CS/rasterize_to_pixels_bwd.cu:16-277 (per pixel, back to front,
 *                    closed-form gradients :192-241)
 * CS = /root/reference/submodules/gsplat/gsplat/cuda/csrc.
 *
 * One scalar loop per pixel in fp32 (expf from libm instead of the GPU's ex2.approx);
 * gradients are accumulated in double so that the oracle's own summation error is
 * negligible next to the tolerance.  `margin` (optional) receives, per pixel, the smallest
 * relative distance of any evaluated pair to one of the two decision thresholds
 * (alpha = 1/255, next_T = 1e-4): pixels with a tiny margin are the ones where a
 * different-but-valid rounding may flip a skip/stop decision.
 *
 * Parity status: UNPINNED against executed reference output (the reference has no
 * CPU-runnable rasterizer); cross-checked against oracle/torch_ref.py and autograd.
 *
 * Build: see oracle/Makefile (gcc -O2 -pthread -shared -fPIC; tiles are spread over the host
 * cores with a small pthread work queue).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ---- minimal pthread "parallel for" over tiles (libgomp is not in this image) ------- */
typedef void (*tile_fn)(int64_t tile, void *ctx);
typedef struct { tile_fn fn; void *ctx; int64_t n; int64_t next; } pf_job;
static void *pf_worker(void *arg) {
    pf_job *j = (pf_job *)arg;
    for (;;) {
        int64_t t0 = __atomic_fetch_add(&j->next, 8, __ATOMIC_RELAXED);
        if (t0 >= j->n) break;
        int64_t t1 = t0 + 8 < j->n ? t0 + 8 : j->n;
        for (int64_t t = t0; t < t1; ++t) j->fn(t, j->ctx);
    }
    return NULL;
}
static int g_threads = 0;
void raster_ref_set_threads(int n) { g_threads = n; }
int raster_ref_get_threads(void) {
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)(n > 256 ? 256 : n) : 1;
}
static void parallel_for(int64_t n, tile_fn fn, void *ctx) {
    int nt = raster_ref_get_threads();
    pf_job job = {fn, ctx, n, 0};
    if (nt <= 1 || n <= 8) { pf_worker(&job); return; }
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < nt - 1; ++i)
        if (pthread_create(&th[started], NULL, pf_worker, &job) == 0) ++started;
    pf_worker(&job);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
}

#define ALPHA_MAX 0.999f
#define ALPHA_MIN (1.f / 255.f)
#define T_EPS 1e-4f

static inline float fminf_(float a, float b) { return a < b ? a : b; }

typedef struct {
    uint32_t C; int64_t n_gauss; int64_t n_isects; uint32_t channels;
    const float *means2d, *conics, *colors, *opacities, *backgrounds; const uint8_t *masks;
    uint32_t W, H, ts, tw, th; const int32_t *offsets, *flatten_ids;
    float *out_colors, *out_alphas; int32_t *last_ids_out; float *margin;
    const float *render_alphas; const int32_t *last_ids; const float *v_render_colors, *v_render_alphas;
    double *v_means2d_abs, *v_means2d, *v_conics, *v_colors, *v_opacities;
} rctx;

static void fwd_tile(int64_t tl, void *vctx) {
    const rctx *x = (const rctx *)vctx;
    const uint32_t C = x->C, channels = x->channels, W = x->W, H = x->H, ts = x->ts, tw = x->tw, th = x->th;
    const int64_t n_isects = x->n_isects;
    const float *means2d = x->means2d, *conics = x->conics, *colors = x->colors, *opacities = x->opacities;
    const float *backgrounds = x->backgrounds; const uint8_t *masks = x->masks;
    const int32_t *offsets = x->offsets, *flatten_ids = x->flatten_ids;
    float *out_colors = x->out_colors, *out_alphas = x->out_alphas, *margin = x->margin;
    int32_t *last_ids = x->last_ids_out;
    const int64_t n_tiles = (int64_t)tw * th;
    const int64_t total_tiles = (int64_t)C * n_tiles;
    {
        const uint32_t cam = (uint32_t)(tl / n_tiles);
        const int64_t tid = tl % n_tiles;
        const uint32_t ty = (uint32_t)(tid / tw), tx = (uint32_t)(tid % tw);
        const int32_t rs = offsets[tl];
        const int32_t re = (tl == total_tiles - 1) ? (int32_t)n_isects : offsets[tl + 1];
        const float *bg = backgrounds ? backgrounds + (size_t)cam * channels : NULL;
        float *pix_out = (float *)malloc(sizeof(float) * channels);
        for (uint32_t ly = 0; ly < ts; ++ly) {
            for (uint32_t lx = 0; lx < ts; ++lx) {
                const uint32_t i = ty * ts + ly, j = tx * ts + lx;
                if (i >= H || j >= W) continue;
                const size_t pix = ((size_t)cam * H + i) * W + j;
                if (masks && !masks[tl]) {
                    for (uint32_t k = 0; k < channels; ++k) out_colors[pix * channels + k] = bg ? bg[k] : 0.f;
                    if (margin) margin[pix] = 1.f;
                    continue;
                }
                const float px = (float)j + 0.5f, py = (float)i + 0.5f;
                float T = 1.f, mrg = 1.f;
                int32_t cur = 0;
                for (uint32_t k = 0; k < channels; ++k) pix_out[k] = 0.f;
                for (int32_t idx = rs; idx < re; ++idx) {
                    const int32_t g = flatten_ids[idx];
                    const float dx = means2d[2 * (size_t)g] - px, dy = means2d[2 * (size_t)g + 1] - py;
                    const float ca = conics[3 * (size_t)g], cb = conics[3 * (size_t)g + 1], cc = conics[3 * (size_t)g + 2];
                    const float sigma = 0.5f * (ca * dx * dx + cc * dy * dy) + cb * dx * dy;
                    const float alpha = fminf_(ALPHA_MAX, opacities[g] * expf(-sigma));
                    const float ma = fabsf(alpha * 255.f - 1.f);
                    if (sigma >= 0.f && ma < mrg) mrg = ma;
                    if (sigma < 0.f || alpha < ALPHA_MIN) continue;
                    const float next_T = T * (1.f - alpha);
                    const float mt = fabsf(next_T * 1e4f - 1.f);
                    if (mt < mrg) mrg = mt;
                    if (next_T <= T_EPS) break;
                    const float vis = alpha * T;
                    const float *c = colors + (size_t)g * channels;
                    for (uint32_t k = 0; k < channels; ++k) pix_out[k] += c[k] * vis;
                    cur = idx;
                    T = next_T;
                }
                out_alphas[pix] = 1.f - T;
                for (uint32_t k = 0; k < channels; ++k)
                    out_colors[pix * channels + k] = bg ? (pix_out[k] + T * bg[k]) : pix_out[k];
                last_ids[pix] = cur;
                if (margin) margin[pix] = mrg;
            }
        }
        free(pix_out);
    }
}

void raster_fwd_ref(uint32_t C, int64_t n_isects, uint32_t channels, const float *means2d, const float *conics,
                    const float *colors, const float *opacities, const float *backgrounds, const uint8_t *masks,
                    uint32_t W, uint32_t H, uint32_t ts, uint32_t tw, uint32_t th, const int32_t *offsets,
                    const int32_t *flatten_ids, float *out_colors, float *out_alphas, int32_t *last_ids,
                    float *margin) {
    rctx x;
    memset(&x, 0, sizeof(x));
    x.C = C; x.n_isects = n_isects; x.channels = channels; x.means2d = means2d; x.conics = conics;
    x.colors = colors; x.opacities = opacities; x.backgrounds = backgrounds; x.masks = masks;
    x.W = W; x.H = H; x.ts = ts; x.tw = tw; x.th = th; x.offsets = offsets; x.flatten_ids = flatten_ids;
    x.out_colors = out_colors; x.out_alphas = out_alphas; x.last_ids_out = last_ids; x.margin = margin;
    parallel_for((int64_t)C * tw * th, fwd_tile, &x);
}

static inline void atomic_add_d(double *p, double v) {
    uint64_t *ip = (uint64_t *)p;
    uint64_t old = __atomic_load_n(ip, __ATOMIC_RELAXED), neu;
    do {
        double d;
        memcpy(&d, &old, 8);
        d += v;
        memcpy(&neu, &d, 8);
    } while (!__atomic_compare_exchange_n(ip, &old, neu, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}

static void bwd_tile(int64_t tl, void *vctx) {
    const rctx *x = (const rctx *)vctx;
    const uint32_t C = x->C, channels = x->channels, W = x->W, H = x->H, ts = x->ts, tw = x->tw, th = x->th;
    const int64_t n_isects = x->n_isects;
    const float *means2d = x->means2d, *conics = x->conics, *colors = x->colors, *opacities = x->opacities;
    const float *backgrounds = x->backgrounds; const uint8_t *masks = x->masks;
    const int32_t *offsets = x->offsets, *flatten_ids = x->flatten_ids, *last_ids = x->last_ids;
    const float *render_alphas = x->render_alphas, *v_render_colors = x->v_render_colors;
    const float *v_render_alphas = x->v_render_alphas;
    double *v_means2d_abs = x->v_means2d_abs, *v_means2d = x->v_means2d, *v_conics = x->v_conics;
    double *v_colors = x->v_colors, *v_opacities = x->v_opacities;
    const int64_t n_tiles = (int64_t)tw * th;
    const int64_t total_tiles = (int64_t)C * n_tiles;
    {
        if (masks && !masks[tl]) return;
        const uint32_t cam = (uint32_t)(tl / n_tiles);
        const int64_t tid = tl % n_tiles;
        const uint32_t ty = (uint32_t)(tid / tw), tx = (uint32_t)(tid % tw);
        const int32_t rs = offsets[tl];
        const int32_t re = (tl == total_tiles - 1) ? (int32_t)n_isects : offsets[tl + 1];
        const float *bg = backgrounds ? backgrounds + (size_t)cam * channels : NULL;
        float *buffer = (float *)malloc(sizeof(float) * channels);
        for (uint32_t ly = 0; ly < ts; ++ly) {
            for (uint32_t lx = 0; lx < ts; ++lx) {
                const uint32_t i = ty * ts + ly, j = tx * ts + lx;
                if (i >= H || j >= W) continue;
                const size_t pix = ((size_t)cam * H + i) * W + j;
                const float px = (float)j + 0.5f, py = (float)i + 0.5f;
                const float T_final = 1.f - render_alphas[pix];
                float T = T_final;
                const int32_t bin_final = last_ids[pix];
                const float *v_c = v_render_colors + pix * channels;
                const float v_a = v_render_alphas[pix];
                for (uint32_t k = 0; k < channels; ++k) buffer[k] = 0.f;
                for (int32_t idx = re - 1; idx >= rs; --idx) {
                    if (idx > bin_final) continue;
                    const int32_t g = flatten_ids[idx];
                    const float dx = means2d[2 * (size_t)g] - px, dy = means2d[2 * (size_t)g + 1] - py;
                    const float ca = conics[3 * (size_t)g], cb = conics[3 * (size_t)g + 1], cc = conics[3 * (size_t)g + 2];
                    const float opac = opacities[g];
                    const float sigma = 0.5f * (ca * dx * dx + cc * dy * dy) + cb * dx * dy;
                    const float vis = expf(-sigma);
                    const float alpha = fminf_(ALPHA_MAX, opac * vis);
                    if (sigma < 0.f || alpha < ALPHA_MIN) continue;
                    const float ra = 1.f / (1.f - alpha);
                    T *= ra;
                    const float fac = alpha * T;
                    const float *c = colors + (size_t)g * channels;
                    float v_alpha = 0.f;
                    for (uint32_t k = 0; k < channels; ++k) {
                        atomic_add_d(v_colors + (size_t)g * channels + k, (double)(fac * v_c[k]));
                        v_alpha += (c[k] * T - buffer[k] * ra) * v_c[k];
                    }
                    v_alpha += T_final * ra * v_a;
                    if (bg) {
                        float accum = 0.f;
                        for (uint32_t k = 0; k < channels; ++k) accum += bg[k] * v_c[k];
                        v_alpha += -T_final * ra * accum;
                    }
                    if (opac * vis <= ALPHA_MAX) {
                        const float v_sigma = -opac * vis * v_alpha;
                        atomic_add_d(v_conics + 3 * (size_t)g, (double)(0.5f * v_sigma * dx * dx));
                        atomic_add_d(v_conics + 3 * (size_t)g + 1, (double)(v_sigma * dx * dy));
                        atomic_add_d(v_conics + 3 * (size_t)g + 2, (double)(0.5f * v_sigma * dy * dy));
                        const float vx = v_sigma * (ca * dx + cb * dy), vy = v_sigma * (cb * dx + cc * dy);
                        atomic_add_d(v_means2d + 2 * (size_t)g, (double)vx);
                        atomic_add_d(v_means2d + 2 * (size_t)g + 1, (double)vy);
                        if (v_means2d_abs) {
                            atomic_add_d(v_means2d_abs + 2 * (size_t)g, (double)fabsf(vx));
                            atomic_add_d(v_means2d_abs + 2 * (size_t)g + 1, (double)fabsf(vy));
                        }
                        atomic_add_d(v_opacities + g, (double)(vis * v_alpha));
                    }
                    for (uint32_t k = 0; k < channels; ++k) buffer[k] += c[k] * fac;
                }
            }
        }
        free(buffer);
    }
}

/* Gradients are written as double arrays (zeroed here). v_means2d_abs may be NULL. */
void raster_bwd_ref(uint32_t C, int64_t n_gauss, int64_t n_isects, uint32_t channels, const float *means2d,
                    const float *conics, const float *colors, const float *opacities, const float *backgrounds,
                    const uint8_t *masks, uint32_t W, uint32_t H, uint32_t ts, uint32_t tw, uint32_t th,
                    const int32_t *offsets, const int32_t *flatten_ids, const float *render_alphas,
                    const int32_t *last_ids, const float *v_render_colors, const float *v_render_alphas,
                    double *v_means2d_abs, double *v_means2d, double *v_conics, double *v_colors,
                    double *v_opacities) {
    memset(v_means2d, 0, sizeof(double) * 2 * n_gauss);
    memset(v_conics, 0, sizeof(double) * 3 * n_gauss);
    memset(v_colors, 0, sizeof(double) * channels * n_gauss);
    memset(v_opacities, 0, sizeof(double) * n_gauss);
    if (v_means2d_abs) memset(v_means2d_abs, 0, sizeof(double) * 2 * n_gauss);
    rctx x;
    memset(&x, 0, sizeof(x));
    x.C = C; x.n_gauss = n_gauss; x.n_isects = n_isects; x.channels = channels; x.means2d = means2d;
    x.conics = conics; x.colors = colors; x.opacities = opacities; x.backgrounds = backgrounds; x.masks = masks;
    x.W = W; x.H = H; x.ts = ts; x.tw = tw; x.th = th; x.offsets = offsets; x.flatten_ids = flatten_ids;
    x.render_alphas = render_alphas; x.last_ids = last_ids; x.v_render_colors = v_render_colors;
    x.v_render_alphas = v_render_alphas; x.v_means2d_abs = v_means2d_abs; x.v_means2d = v_means2d;
    x.v_conics = v_conics; x.v_colors = v_colors; x.v_opacities = v_opacities;
    parallel_for((int64_t)C * tw * th, bwd_tile, &x);
}
