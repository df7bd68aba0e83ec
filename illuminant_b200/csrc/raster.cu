// "Next" row N2: ParticleSystem.Render -- the rasterisation of RenderColor / RenderData into a render
// target (Illuminant/Shaders/RasterizeParticleSystem.fx, Particles/ParticleSystem.cs:876-1039).
//
// The reference draws one instanced quad per particle, chunk after chunk, and lets the output-merger blend them in draw
// order.  Here the same ORDERED result is produced tile by tile:
//   1. raster_count_kernel   one thread per particle: vertex-shader work (rotated quad in pixel space), number of 16x16 pixel
//                            tiles its bounding box touches (0 for dead / off-screen particles);
//   2. exclusive scan (CUB)  -> where each particle's (tile, particle) pairs start; the pairs are emitted IN DRAW ORDER;
//   3. raster_emit_kernel    writes the pairs;
//   4. stable radix sort (CUB) of the pairs by tile id: each tile's run keeps the draw order;
//   5. raster_ranges_kernel  run boundaries per tile;
//   6. raster_shade_kernel   one CTA per tile, one thread per pixel: the tile's quads are staged through shared memory 256 at
//                            a time (each thread rebuilds one quad from the 48 B of particle state), every pixel walks the
//                            batch in order -- coverage, pixel shader, blend in fp32 registers -- and stores its texel once.
// The reference reads and writes the target once per covered quad-pixel; here every pixel is written once.  Coverage and blend
// arithmetic uses individually rounded IEEE operations (x-ops, ilb_device.cuh) so that the CPU oracle makes bit-identical
// coverage decisions; CUB provides the scan and the sort (plumbing), the vertex / pixel work is hand-written.
#include <cuda_fp16.h>

#include <algorithm>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "ilb_internal.h"

namespace {

#include "ilb_bezier.cuh"

constexpr int RTILE = 16;          // tile edge in pixels; one CTA of RTILE*RTILE threads per tile
constexpr int RBATCH = RTILE * RTILE;

struct RasterParams {
    const float4 *P, *RD, *RC;  // PositionAndLife, RenderData, RenderColor of the live chunks (draw order = array order)
    unsigned total;
    int W, H, tilesX, tilesY;
    float4 globalColor, region, sizeFactorAndPosition, scale, zConfiguration, options, texelAndSize, animRateRotZToY;
    ilb_bezier1 rounding;
    float vpx, vpy, vsx, vsy;
    const uchar4* tex;
    int texW, texH, filter, blend;
    void* target;
    int fmt, clear;
    float4 clearColor;
    unsigned *counts, *offsets;     // per particle (+1)
    unsigned *keys, *vals;          // sorted pairs
    unsigned *tileStart, *tileEnd;  // per tile
    unsigned long long* pairTotal;  // 64-bit sum of counts (the 32-bit scan may wrap when quads are huge: checked before use)
};

struct Sprite {           // one quad in pixel space + the per-quad varyings of VS_PosVelAttr
    float cx, cy;         // centre
    float m00, m01, m10, m11;  // pixel offset -> unit coordinates (u, v) of the quad
    float r, g, b, a;     // RenderColor
    float frameU, frameV; // frameTexCoord
    float rounding;       // clamp(RoundingPowerFromLife(life), 0.001, 1)
    float valid;
};

// VS_PosVelAttr (RasterizeParticleSystem.fx:62-150).  Returns false for particles that draw nothing.  x0..y1: the pixel
// bounding box (conservative, clamped to the target).
// FULL = false: geometry only (the binning kernels need neither the colour nor the per-quad varyings).
template <bool FULL>
ILB_DEV bool makeSprite(const RasterParams& R, unsigned i, Sprite& s, int& x0, int& y0, int& x1, int& y1) {
    const float4 position = __ldg(R.P + i);
    const float life = position.w;
    if (!(life > 0.0f)) return false;  // :75-79 (StippleFactor == 1: StippleReject never rejects)
    const float4 renderData = __ldg(R.RD + i);
    const float angle = fmodf(renderData.y, xmul(2.0f, ILB_PI));  // :81-82
    const float zf = fmaxf(0.0f, xadd(1.0f, xmul(position.z, R.zConfiguration.x)));  // :85
    const float sx = xmul(xmul(xmul(renderData.x, R.texelAndSize.z), R.sizeFactorAndPosition.x), zf);  // :84
    const float sy = xmul(xmul(xmul(renderData.x, R.texelAndSize.w), R.sizeFactorAndPosition.y), zf);
    const float sn = dm_sinf(angle), cs = dm_cosf(angle);
    // displayXyz (:91-93), minus the viewport position, times the viewport scale (:97-107)
    const float dispx = xadd(xmul(position.x, R.scale.x), R.sizeFactorAndPosition.z);
    const float dispy = xadd(xmul(xsub(position.y, xmul(position.z, R.animRateRotZToY.w)), R.scale.y), R.sizeFactorAndPosition.w);
    s.cx = xmul(xsub(dispx, R.vpx), R.vsx);
    s.cy = xmul(xsub(dispy, R.vpy), R.vsy);
    // the quad's axes in pixel space: rotatedCorner (:44-59) * Scale * ViewportScale for the unit corners (1,0) and (0,1)
    const float kx = xmul(R.scale.x, R.vsx), ky = xmul(R.scale.y, R.vsy);
    const float ax = xmul(xmul(cs, sx), kx), ay = xmul(xmul(sn, sx), ky);
    const float bx = -xmul(xmul(sn, sy), kx), by = xmul(xmul(cs, sy), ky);
    const float det = xsub(xmul(ax, by), xmul(ay, bx));
    if (!(fabsf(det) > 0.0f)) return false;  // zero-area quad (also NaN)
    s.m00 = xdiv(by, det); s.m01 = xdiv(-bx, det); s.m10 = xdiv(-ay, det); s.m11 = xdiv(ax, det);
    const float ex = xadd(fabsf(ax), fabsf(bx)), ey = xadd(fabsf(ay), fabsf(by));
    // pixel x can be covered only if its centre x + 0.5 lies within cx -+ ex: x in [cx - ex - 0.5, cx + ex - 0.5]; 1/64 px of slack
    // covers the rounding of the sums (one ulp at 16384 is 1/1024 px).  Conservative is all this has to be: coverage itself is
    // decided per pixel by the exact (u, v) test.
    const float fx0 = xsub(xsub(s.cx, ex), 0.515625f), fx1 = xsub(xadd(s.cx, ex), 0.484375f);
    const float fy0 = xsub(xsub(s.cy, ey), 0.515625f), fy1 = xsub(xadd(s.cy, ey), 0.484375f);
    if (!(fx0 <= fx1) || !(fy0 <= fy1) || isinf(fx0) || isinf(fx1) || isinf(fy0) || isinf(fy1)) return false;
    if (fx1 < 0.0f || fy1 < 0.0f || fx0 > (float)(R.W - 1) || fy0 > (float)(R.H - 1)) return false;  // off-screen
    x0 = (int)floorf(fmaxf(fx0, 0.0f)); x1 = (int)ceilf(fminf(fx1, (float)(R.W - 1)));
    y0 = (int)floorf(fmaxf(fy0, 0.0f)); y1 = (int)ceilf(fminf(fy1, (float)(R.H - 1)));
    if (!FULL) return true;
    const float4 color = __ldg(R.RC + i);
    s.r = color.x; s.g = color.y; s.b = color.z; s.a = color.w;
    s.frameU = 0.0f; s.frameV = 0.0f;
    if (R.filter != ILB_TEXTURE_NONE) {  // animation frame of the sprite sheet (:114-141)
        const float tsx = xsub(R.region.z, R.region.x), tsy = xsub(R.region.w, R.region.y);
        const float fcx = floorf(xdiv(1.0f, tsx)), fcy = floorf(xdiv(1.0f, tsy));
        float fix = floorf(xmul(fabsf(R.animRateRotZToY.x), life)), fiy = floorf(xmul(fabsf(R.animRateRotZToY.y), life));
        const float maxAngleX = xdiv(xmul(2.0f, ILB_PI), fcx), maxAngleY = xdiv(xmul(2.0f, ILB_PI), fcy);
        const float ffvx = floorf(xadd(xdiv(angle, maxAngleX), 0.5f)), ffvy = floorf(xadd(xdiv(angle, maxAngleY), 0.5f));  // round()
        fiy = xadd(fiy, floorf(renderData.w));
        if (R.options.z != 0.0f) fix = xadd(fix, ffvx);
        if (R.options.w != 0.0f) fiy = xadd(fiy, ffvy);
        fix = fmodf(fmaxf(fix, 0.0f), fcx);
        fiy = fminf(fmaxf(fiy, 0.0f), xsub(fcy, 1.0f));
        if (R.animRateRotZToY.x < 0.0f) fix = xsub(xsub(fcx, fix), 1.0f);
        if (R.animRateRotZToY.y < 0.0f) fiy = xsub(xsub(fcy, fiy), 1.0f);
        s.frameU = xmul(fix, tsx); s.frameV = xmul(fiy, tsy);
    }
    const float roundingPower = evaluateBezier1(R.rounding, life);  // :146-149
    s.rounding = fminf(fmaxf(roundingPower, 0.001f), 1.0f);
    s.valid = 1.0f;
    return true;
}

ILB_DEV void tileRange(const RasterParams& R, int x0, int y0, int x1, int y1, int& tx0, int& ty0, int& tx1, int& ty1) {
    tx0 = x0 / RTILE; ty0 = y0 / RTILE; tx1 = x1 / RTILE; ty1 = y1 / RTILE;
}

__global__ void __launch_bounds__(256) raster_count_kernel(const __grid_constant__ RasterParams R) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned n = 0;
    if (i < R.total) {
        Sprite s;
        int x0, y0, x1, y1;
        if (makeSprite<false>(R, i, s, x0, y0, x1, y1)) {
            int tx0, ty0, tx1, ty1;
            tileRange(R, x0, y0, x1, y1, tx0, ty0, tx1, ty1);
            n = (unsigned)(tx1 - tx0 + 1) * (unsigned)(ty1 - ty0 + 1);
        }
    }
    if (i <= R.total) R.counts[i] = n;  // counts[total] = 0: the scan's last element is the number of pairs
    unsigned long long sum = n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(R.pairTotal, sum);
}

__global__ void __launch_bounds__(256) raster_emit_kernel(const __grid_constant__ RasterParams R) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R.total) return;
    if (R.counts[i] == 0) return;
    Sprite s;
    int x0, y0, x1, y1;
    if (!makeSprite<false>(R, i, s, x0, y0, x1, y1)) return;
    int tx0, ty0, tx1, ty1;
    tileRange(R, x0, y0, x1, y1, tx0, ty0, tx1, ty1);
    unsigned o = R.offsets[i];
    for (int ty = ty0; ty <= ty1; ty++)
        for (int tx = tx0; tx <= tx1; tx++, o++) {
            R.keys[o] = (unsigned)(ty * R.tilesX + tx);
            R.vals[o] = i;
        }
}

__global__ void __launch_bounds__(256) raster_ranges_kernel(const unsigned* __restrict__ keys, unsigned n, unsigned* tileStart, unsigned* tileEnd) {
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned k = keys[j];
    if (j == 0 || keys[j - 1] != k) tileStart[k] = j;
    if (j == n - 1 || keys[j + 1] != k) tileEnd[k] = j + 1;
}

// ---- pixel shader pieces (RasterizeParticleSystem.fx:152-254) ---------------------------------------------------------
ILB_DEV f4 spriteTexel(const RasterParams& R, int x, int y) {
    x = min(max(x, 0), R.texW - 1);
    y = min(max(y, 0), R.texH - 1);
    const uchar4 t = __ldg(R.tex + (size_t)y * R.texW + x);
    return mk4(xdiv((float)t.x, 255.0f), xdiv((float)t.y, 255.0f), xdiv((float)t.z, 255.0f), xdiv((float)t.w, 255.0f));
}
ILB_DEV f4 spriteSample(const RasterParams& R, float u, float v) {  // BitmapSampler / BitmapPointSampler (:28-43), one mip level
    if (R.filter == ILB_TEXTURE_POINT) return spriteTexel(R, (int)floorf(xmul(u, (float)R.texW)), (int)floorf(xmul(v, (float)R.texH)));
    const float fx = xsub(xmul(u, (float)R.texW), 0.5f), fy = xsub(xmul(v, (float)R.texH), 0.5f);
    const float x0 = floorf(fx), y0 = floorf(fy);
    const float tx = xsub(fx, x0), ty = xsub(fy, y0);
    const f4 top = xlerp4(spriteTexel(R, (int)x0, (int)y0), spriteTexel(R, (int)x0 + 1, (int)y0), tx);
    const f4 bottom = xlerp4(spriteTexel(R, (int)x0, (int)y0 + 1), spriteTexel(R, (int)x0 + 1, (int)y0 + 1), tx);
    return xlerp4(top, bottom, ty);
}
ILB_DEV float computeCircularAlpha(const RasterParams& R, float u, float v, float rounding) {  // :152-163
    if (R.options.x == 0.0f) return 1.0f;
    const float distance = xsqrt(xadd(xmul(u, u), xmul(v, v)));
    const float power = fmaxf(rounding, 0.01f);
    const float divisor = fmaxf(saturatef(xsub(1.0f, power)), 0.001f);
    const float distanceFromEdge = xdiv(saturatef(xsub(distance, power)), divisor);
    const float powDistanceFromEdge = powf(distanceFromEdge, power);
    return saturatef(xsub(1.0f, powDistanceFromEdge));
}

ILB_DEV f4 loadTarget(const void* base, int fmt, size_t i) {
    if (fmt == ILB_FORMAT_FLOAT4) return mk4(reinterpret_cast<const float4*>(base)[i]);
    if (fmt == ILB_FORMAT_HALF4) {
        const uint2 v = reinterpret_cast<const uint2*>(base)[i];
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
        return mk4(lo.x, lo.y, hi.x, hi.y);
    }
    const uint32_t v = reinterpret_cast<const uint32_t*>(base)[i];
    return mk4(xdiv((float)(v & 255u), 255.0f), xdiv((float)((v >> 8) & 255u), 255.0f), xdiv((float)((v >> 16) & 255u), 255.0f), xdiv((float)(v >> 24), 255.0f));
}
ILB_DEV void storeTarget(void* base, int fmt, size_t i, f4 c) {
    if (fmt == ILB_FORMAT_FLOAT4) {
        reinterpret_cast<float4*>(base)[i] = to_float4(c);
    } else if (fmt == ILB_FORMAT_HALF4) {
        const __half2 lo = __floats2half2_rn(c.x, c.y), hi = __floats2half2_rn(c.z, c.w);
        uint2 v;
        v.x = *reinterpret_cast<const uint32_t*>(&lo);
        v.y = *reinterpret_cast<const uint32_t*>(&hi);
        reinterpret_cast<uint2*>(base)[i] = v;
    } else {
        const uint32_t r = (uint32_t)(saturatef(c.x) * 255.0f + 0.5f), g = (uint32_t)(saturatef(c.y) * 255.0f + 0.5f);
        const uint32_t b = (uint32_t)(saturatef(c.z) * 255.0f + 0.5f), a = (uint32_t)(saturatef(c.w) * 255.0f + 0.5f);
        reinterpret_cast<uint32_t*>(base)[i] = r | (g << 8) | (b << 16) | (a << 24);
    }
}

// One CTA per 16x16 tile.  Each of the 8 warps owns an 8x4 pixel block of the tile (lane -> (lane & 7, lane >> 3)), so that a
// quad whose pixel bounding box misses the block is rejected by one warp-uniform test instead of 32 per-lane (u, v) evaluations:
// quads are typically a few pixels wide, while the tile's list holds every quad that touches any of its 256 pixels.
__global__ void __launch_bounds__(RBATCH) raster_shade_kernel(const __grid_constant__ RasterParams R) {
    __shared__ Sprite batch[RBATCH];
    __shared__ int4 boxes[RBATCH];  // pixel bounding box (x0, y0, x1, y1) of each staged quad; x1 < x0 for quads that draw nothing
    const int tile = blockIdx.x;
    const int tx = tile % R.tilesX, ty = tile / R.tilesX;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bx0 = tx * RTILE + (warp & 1) * 8, by0 = ty * RTILE + (warp >> 1) * 4;  // the warp's 8x4 block
    const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
    const bool inside = px < R.W && py < R.H;
    const size_t pi = (size_t)py * (size_t)R.W + (size_t)px;
    f4 acc = mk4(R.clearColor);
    if (inside && !R.clear) acc = loadTarget(R.target, R.fmt, pi);
    const unsigned start = R.tileStart[tile], end = R.tileEnd[tile];  // start == end: no quad touches the tile
    const float pcx = xadd((float)px, 0.5f), pcy = xadd((float)py, 0.5f);
    for (unsigned base = start; base < end; base += RBATCH) {
        const unsigned n = min((unsigned)RBATCH, end - base);
        __syncthreads();  // the previous batch is consumed
        if (threadIdx.x < n) {
            Sprite s;
            int x0 = 0, y0 = 0, x1 = -1, y1 = -1;
            if (!makeSprite<true>(R, R.vals[base + threadIdx.x], s, x0, y0, x1, y1)) { x0 = 0; x1 = -1; }
            batch[threadIdx.x] = s;
            boxes[threadIdx.x] = make_int4(x0, y0, x1, y1);
        }
        __syncthreads();
        // each lane tests one staged quad's box against the warp's block, one ballot per 32 quads, and the warp visits only the set
        // bits -- in ascending order, so the draw order holds (measured on 8 M small quads: 4.34 -> 2.99 ms per render against a
        // warp-uniform test per quad)
        for (unsigned k0 = 0; k0 < n; k0 += 32) {
            bool hit = false;
            if (k0 + lane < n) {
                const int4 b = boxes[k0 + lane];
                hit = !(b.z < bx0 || b.x > bx0 + 7 || b.w < by0 || b.y > by0 + 3);
            }
            unsigned hits = __ballot_sync(0xFFFFFFFFu, hit);
            while (hits) {
                const unsigned k = k0 + (unsigned)(__ffs(hits) - 1);
                hits &= hits - 1;
            const Sprite& s = batch[k];
            const float dx = xsub(pcx, s.cx), dy = xsub(pcy, s.cy);
            const float u = xadd(xmul(dx, s.m00), xmul(dy, s.m01)), v = xadd(xmul(dx, s.m10), xmul(dy, s.m11));
            if (!(u >= -1.0f && u < 1.0f && v >= -1.0f && v < 1.0f)) continue;
            // PS_NoTexture / PS_Texture / PS_TexturePoint; (1 / 512) in the shader is an integer division: the thresholds are 0
            f4 result = mk4(s.r, s.g, s.b, s.a);
            if (R.filter != ILB_TEXTURE_NONE) {
                if (s.a > 0.0f) {
                    const float ccx = xadd(xdiv(u, 2.0f), 0.5f), ccy = xadd(xdiv(v, 2.0f), 0.5f);
                    const float tu = xadd(xlerp(R.region.x, R.region.z, ccx), s.frameU), tv = xadd(xlerp(R.region.y, R.region.w, ccy), s.frameV);
                    result = xmul4(result, spriteSample(R, tu, tv));
                    result = xmul4(result, mk4(R.globalColor));
                }
            } else {
                result = xmul4(result, mk4(R.globalColor));
            }
            result = xscale4(result, computeCircularAlpha(R, u, v, s.rounding));
            if (!(result.w > 0.0f)) continue;  // discard
            if (R.blend == ILB_BLEND_ALPHA) {
                const float k1 = xsub(1.0f, result.w);
                acc = xadd4(result, xscale4(acc, k1));
            } else if (R.blend == ILB_BLEND_ADDITIVE) {
                acc = xadd4(xscale4(result, result.w), acc);
            } else {
                acc = result;
            }
            }  // quad
        }
    }
    if (inside) storeTarget(R.target, R.fmt, pi, acc);
}

// ---- multi-GPU ParticleSystem.Render: composite of per-rank layers ---------------------------------------------------------
// Chunks are sharded over the ranks in contiguous ranges, i.e. in DRAW ORDER, so rank k's chunks come
// after rank k - 1's.  Every rank renders its chunks over a TRANSPARENT float4 layer with ilb_particles_render_device; the image
// the reference would draw is the layers composited in rank order: premultiplied "over" is associative (AlphaBlend), additive
// blending is a sum.  This kernel composites rows [row_begin, row_end) -- each rank takes one band -- reading the band of EVERY
// rank's layer through its peer mapping (P2P loads over NVLink) and storing the finished texels into EVERY rank's target (peer
// stores): the reduce-scatter and the all-gather of the image in one launch, no NCCL call on the data path.
constexpr int MAX_LAYERS = 8;
struct CompositeParams {
    const float4* layers[MAX_LAYERS];
    void* targets[MAX_LAYERS];
    int nlayers, ntargets, W, row_begin, row_end, fmt, blend, clear;
    float4 clearColor;
};
__global__ void __launch_bounds__(256) raster_composite_kernel(const __grid_constant__ CompositeParams C) {
    const size_t n = (size_t)C.W * (size_t)(C.row_end - C.row_begin), base = (size_t)C.W * (size_t)C.row_begin;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t pi = base + i;
        f4 acc = C.clear ? mk4(C.clearColor) : loadTarget(C.targets[0], C.fmt, pi);
        for (int k = 0; k < C.nlayers; k++) {   // rank order == draw order
            const f4 l = mk4(__ldcs(C.layers[k] + pi));
            if (C.blend == ILB_BLEND_ALPHA) acc = xadd4(l, xscale4(acc, xsub(1.0f, l.w)));
            else acc = xadd4(l, acc);
        }
        for (int t = 0; t < C.ntargets; t++) storeTarget(C.targets[t], C.fmt, pi, acc);
    }
}

int reserveU32(ilb_ctx* ctx, unsigned** p, size_t* cap, size_t count) {
    return ilb_reserve(ctx, reinterpret_cast<void**>(p), cap, std::max<size_t>(count, 4) * sizeof(unsigned), false);
}

}  // namespace

void ilb_raster_release(ilb_psys* ps) {
    for (int i = 0; i < ILB_RASTER_BUFFERS; i++)
        if (ps->raster[i]) { cudaFree(ps->raster[i]); ps->raster[i] = nullptr; ps->raster_capacity[i] = 0; }
}

int ilb_raster_launch(ilb_psys* ps, const ilb_particle_render* r, const void* d_texture, void* d_target) {
    ilb_ctx* ctx = ps->ctx;
    if (!r || !d_target) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null argument");
    if (r->width <= 0 || r->height <= 0 || r->width > 32768 || r->height > 32768) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad target size %dx%d", r->width, r->height);
    if (r->target_format != ILB_FORMAT_FLOAT4 && r->target_format != ILB_FORMAT_HALF4 && r->target_format != ILB_FORMAT_RGBA8)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad target format %d", r->target_format);
    if (r->blend < ILB_BLEND_ALPHA || r->blend > ILB_BLEND_OPAQUE) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad blend %d", r->blend);
    if (r->texture_filter < ILB_TEXTURE_NONE || r->texture_filter > ILB_TEXTURE_LINEAR) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad texture filter %d", r->texture_filter);
    if (r->texture_filter != ILB_TEXTURE_NONE && (!d_texture || r->texture_width < 1 || r->texture_height < 1))
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "textured material without a texture");
    if (r->StippleFactor != 1.0f) return ilb_fail(ctx, ILB_ERR_UNSUPPORTED, "StippleFactor != 1 (StippleReject lives in the un-vendored sq/Fracture)");
    if (r->RenderingOptions.y >= 0.5f) return ilb_fail(ctx, ILB_ERR_UNSUPPORTED, "DitheredOpacity (Dither64 lives in the un-vendored sq/Fracture DitherCommon.fxh)");

    RasterParams R;
    memset(&R, 0, sizeof(R));
    R.P = ps->buf[0]; R.RD = ps->buf[4]; R.RC = ps->buf[3];
    R.total = (unsigned)((size_t)ps->live_chunks * ps->per_chunk);
    R.W = r->width; R.H = r->height;
    R.tilesX = (R.W + RTILE - 1) / RTILE; R.tilesY = (R.H + RTILE - 1) / RTILE;
    auto f4 = [](const ilb_float4& v) { return make_float4(v.x, v.y, v.z, v.w); };
    R.globalColor = f4(r->GlobalColor); R.region = f4(r->BitmapTextureRegion); R.sizeFactorAndPosition = f4(r->SizeFactorAndPosition);
    R.scale = f4(r->Scale); R.zConfiguration = f4(r->ZConfiguration); R.options = f4(r->RenderingOptions);
    R.texelAndSize = f4(r->TexelAndSize); R.animRateRotZToY = f4(r->AnimationRateAndRotationAndZToY);
    R.rounding = r->RoundingPowerFromLife;
    R.vpx = r->ViewportPosition[0]; R.vpy = r->ViewportPosition[1]; R.vsx = r->ViewportScale[0]; R.vsy = r->ViewportScale[1];
    R.tex = reinterpret_cast<const uchar4*>(d_texture); R.texW = r->texture_width; R.texH = r->texture_height;
    R.filter = r->texture_filter; R.blend = r->blend;
    R.target = d_target; R.fmt = r->target_format; R.clear = r->clear; R.clearColor = f4(r->ClearColor);

    const size_t tiles = (size_t)R.tilesX * R.tilesY;
    int rc;
    unsigned** buf = reinterpret_cast<unsigned**>(ps->raster);
    if ((rc = reserveU32(ctx, &buf[0], &ps->raster_capacity[0], (size_t)R.total + 1))) return rc;  // counts
    if ((rc = reserveU32(ctx, &buf[1], &ps->raster_capacity[1], (size_t)R.total + 1))) return rc;  // offsets
    if ((rc = reserveU32(ctx, &buf[2], &ps->raster_capacity[2], tiles * 2))) return rc;            // tileStart, tileEnd
    R.counts = buf[0]; R.offsets = buf[1]; R.tileStart = buf[2]; R.tileEnd = buf[2] + tiles;
    ILB_CUDA(ctx, cudaMemsetAsync(R.tileStart, 0, tiles * 2 * sizeof(unsigned), ctx->stream));

    if ((rc = ilb_reserve(ctx, &ps->raster[8], &ps->raster_capacity[8], 16, false))) return rc;    // 64-bit pair total
    R.pairTotal = reinterpret_cast<unsigned long long*>(ps->raster[8]);
    ILB_CUDA(ctx, cudaMemsetAsync(R.pairTotal, 0, sizeof(unsigned long long), ctx->stream));
    unsigned pairs = 0;
    if (R.total > 0) {
        raster_count_kernel<<<(R.total + 1 + 255) / 256, 256, 0, ctx->stream>>>(R);
        ctx->launches++;
        ILB_CUDA(ctx, cudaGetLastError());
        size_t scanBytes = 0;
        ILB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, R.counts, R.offsets, (int)(R.total + 1), ctx->stream));
        if ((rc = ilb_reserve(ctx, &ps->raster[3], &ps->raster_capacity[3], std::max<size_t>(scanBytes, 16), false))) return rc;  // CUB temp storage
        ILB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ps->raster[3], scanBytes, R.counts, R.offsets, (int)(R.total + 1), ctx->stream));
        unsigned long long pairs64 = 0;
        ILB_CUDA(ctx, cudaMemcpyAsync(&pairs64, R.pairTotal, sizeof(pairs64), cudaMemcpyDeviceToHost, ctx->stream));
        ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (pairs64 > (1ull << 30))
            return ilb_fail(ctx, ILB_ERR_OUT_OF_MEMORY, "%llu quad/tile pairs: particles this large are outside the rasteriser's budget", pairs64);
        pairs = (unsigned)pairs64;
    }
    if (pairs > 0) {
        if ((rc = reserveU32(ctx, &buf[4], &ps->raster_capacity[4], pairs))) return rc;  // unsorted keys
        if ((rc = reserveU32(ctx, &buf[5], &ps->raster_capacity[5], pairs))) return rc;  // unsorted values
        if ((rc = reserveU32(ctx, &buf[6], &ps->raster_capacity[6], pairs))) return rc;  // sorted keys
        if ((rc = reserveU32(ctx, &buf[7], &ps->raster_capacity[7], pairs))) return rc;  // sorted values
        R.keys = buf[4]; R.vals = buf[5];
        raster_emit_kernel<<<(R.total + 255) / 256, 256, 0, ctx->stream>>>(R);
        ctx->launches++;
        ILB_CUDA(ctx, cudaGetLastError());
        int bits = 1;
        while (((size_t)1 << bits) < tiles) bits++;
        size_t sortBytes = 0;
        ILB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, buf[4], buf[6], buf[5], buf[7], (int)pairs, 0, bits, ctx->stream));
        if ((rc = ilb_reserve(ctx, &ps->raster[3], &ps->raster_capacity[3], std::max<size_t>(sortBytes, 16), false))) return rc;
        ILB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ps->raster[3], sortBytes, buf[4], buf[6], buf[5], buf[7], (int)pairs, 0, bits, ctx->stream));
        R.keys = buf[6]; R.vals = buf[7];
        raster_ranges_kernel<<<(pairs + 255) / 256, 256, 0, ctx->stream>>>(R.keys, pairs, R.tileStart, R.tileEnd);
        ctx->launches++;
        ILB_CUDA(ctx, cudaGetLastError());
    }
    raster_shade_kernel<<<(unsigned)tiles, RBATCH, 0, ctx->stream>>>(R);
    ctx->launches++;
    ILB_CUDA(ctx, cudaGetLastError());
    return ILB_OK;
}

int ilb_raster_composite(ilb_ctx* ctx, const void* const* d_layers, int layer_count, int width, int height, int row_begin, int row_end,
                         int blend, int target_format, const ilb_float4* clear_color, void* const* d_targets, int target_count) {
    if (!d_layers || !d_targets || layer_count < 1 || layer_count > MAX_LAYERS || target_count < 1 || target_count > MAX_LAYERS)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "1..%d layers and targets", MAX_LAYERS);
    if (width <= 0 || height <= 0 || row_begin < 0 || row_end > height || row_begin > row_end)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad rows [%d,%d) of %dx%d", row_begin, row_end, width, height);
    if (blend != ILB_BLEND_ALPHA && blend != ILB_BLEND_ADDITIVE)
        return ilb_fail(ctx, ILB_ERR_UNSUPPORTED, "only AlphaBlend and Additive layers composite (an Opaque layer has no coverage mask)");
    if (target_format != ILB_FORMAT_FLOAT4 && target_format != ILB_FORMAT_HALF4 && target_format != ILB_FORMAT_RGBA8)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad target format %d", target_format);
    if (row_begin == row_end) return ILB_OK;
    CompositeParams C;
    memset(&C, 0, sizeof(C));
    for (int k = 0; k < layer_count; k++) {
        if (!d_layers[k]) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null layer %d", k);
        C.layers[k] = reinterpret_cast<const float4*>(d_layers[k]);
    }
    for (int t = 0; t < target_count; t++) {
        if (!d_targets[t]) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null target %d", t);
        C.targets[t] = d_targets[t];
    }
    C.nlayers = layer_count; C.ntargets = target_count; C.W = width; C.row_begin = row_begin; C.row_end = row_end;
    C.fmt = target_format; C.blend = blend; C.clear = clear_color ? 1 : 0;
    if (clear_color) C.clearColor = make_float4(clear_color->x, clear_color->y, clear_color->z, clear_color->w);
    const size_t n = (size_t)width * (size_t)(row_end - row_begin);
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 16);
    raster_composite_kernel<<<grid, 256, 0, ctx->stream>>>(C);
    ctx->launches++;
    ILB_CUDA(ctx, cudaGetLastError());
    return ILB_OK;
}
