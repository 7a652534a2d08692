// raster_common.cuh — shared pieces of the tile rasterizer (a8/a9).
#pragma once
#include "common.cuh"

namespace b2s {

constexpr float kAlphaMax = 0.999f;          // CS/rasterize_to_pixels_fwd.cu:146
constexpr float kAlphaMin = 1.f / 255.f;     // :147
constexpr float kTransmittanceEps = 1e-4f;   // :152

// channel counts with a compiled kernel; a request for `channels` runs the smallest
// instance >= channels with the real stride (no padding copy, unlike _wrapper.py:497-541).
static inline int pick_cdim(uint32_t channels) {
    static const int dims[] = {1, 2, 3, 4, 5, 8, 9, 16, 17, 32, 33};
    for (int d : dims)
        if ((int)channels <= d) return d;
    return -1;
}

struct TileCoord {
    uint32_t cam, tile_id, tile_lin;
    uint32_t i, j;       // pixel row / column
    bool inside;
};

__device__ __forceinline__ TileCoord tile_coord(uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                                                uint32_t W, uint32_t H) {
    TileCoord t;
    const uint32_t n_tiles = tile_width * tile_height;
    t.tile_lin = blockIdx.x;
    t.cam = t.tile_lin / n_tiles;
    t.tile_id = t.tile_lin - t.cam * n_tiles;
    const uint32_t ty = t.tile_id / tile_width, tx = t.tile_id - ty * tile_width;
    const uint32_t tr = threadIdx.x;
    const uint32_t ly = tr / tile_size, lx = tr - ly * tile_size;
    t.i = ty * tile_size + ly;
    t.j = tx * tile_size + lx;
    t.inside = (ly < tile_size) && (t.i < H) && (t.j < W);
    return t;
}

}  // namespace b2s
