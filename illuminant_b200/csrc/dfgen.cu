// Distance-field generation for analytic obstructions ("next" row N1 of SURVEY.md section 8f) -- the producer of the
// Rgba64 atlas both hot paths read.  Replaces the reference's per-obstruction instanced quads with MAX blending
// (Lighting/LightingRenderer.DistanceField.cs:347-400, Shaders/DistanceFunction.fx:15-48): one thread owns one
// atlas texel (4 packed z-slices), loops the obstruction list once and stores 8 bytes.
#include <cstring>

#include "ilb_internal.h"
#include "ilb_shapes.cuh"

namespace {

struct DFGenParams {
    uint2* tex;
    const uint2* base;  // static field of a DynamicDistanceField (same atlas geometry) or nullptr
    int tw, th, slice_w, slice_h, slice_count, columns, physical;
    float maxEnc, zOffset, depth, invX, invY;
    const ilb_obstruction* obs;
    int count;
};

constexpr int GEN_TILE = 16;

__global__ void __launch_bounds__(GEN_TILE * GEN_TILE) df_generate_kernel(const __grid_constant__ DFGenParams P) {
    __shared__ uint16_t s_list[GEN_TILE * GEN_TILE];
    __shared__ int s_warpCount[8];
    const int p = blockIdx.z;  // physical slice
    const int tid = threadIdx.y * GEN_TILE + threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int x = blockIdx.x * GEN_TILE + threadIdx.x, y = blockIdx.y * GEN_TILE + threadIdx.y;
    const bool valid = (x < P.slice_w) && (y < P.slice_h);
    const float wx = (float)x * P.invX, wy = (float)y * P.invY;  // getPositionXy DistanceFunction.fx:29-32
    // tile bounds in world space for obstruction culling
    const float tx0 = (float)(blockIdx.x * GEN_TILE) * P.invX, tx1 = (float)(blockIdx.x * GEN_TILE + GEN_TILE - 1) * P.invX;
    const float ty0 = (float)(blockIdx.y * GEN_TILE) * P.invY, ty1 = (float)(blockIdx.y * GEN_TILE + GEN_TILE - 1) * P.invY;
    float sliceZ[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {  // SliceIndexToZ LightingRenderer.DistanceField.cs:32-35
        const float s = ((float)(3 * p + k) / fmaxf(1.0f, (float)P.slice_count));
        sliceZ[k] = (s * P.depth) + P.zOffset;
    }
    float best[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // ClearDistanceField.fx:27-39
    const int cellX = (p % P.columns) * P.slice_w, cellY = (p / P.columns) * P.slice_h;
    if (P.base && valid) {  // dynamic field: the slice starts as a copy of the static texture (LightingRenderer.DistanceField.cs:113-118)
        const uint2 b = __ldg(P.base + (size_t)(cellY + y) * (size_t)P.tw + (size_t)(cellX + x));
        const float k = 1.0f / 65535.0f;
        best[0] = xmul(u16lo(b.x), k); best[1] = xmul(u16hi(b.x), k); best[2] = xmul(u16lo(b.y), k); best[3] = xmul(u16hi(b.y), k);
    }
    for (int base = 0; base < P.count; base += GEN_TILE * GEN_TILE) {
        const int oi = base + tid;
        bool keep = false;
        if (oi < P.count) {
            const ilb_obstruction& o = P.obs[oi];
            // quad extent of DistanceFunctionVertexShader (DistanceFunction.fx:15-27)
            const float msize = fmaxf(fmaxf(fabsf(o.size[0]), fabsf(o.size[1])), fabsf(o.size[2])) + P.maxEnc + 4.0f;
            keep = (o.center[0] + msize >= tx0) && (o.center[0] - msize <= tx1) && (o.center[1] + msize >= ty0) &&
                   (o.center[1] - msize <= ty1);
        }
        const unsigned ballot = __ballot_sync(0xFFFFFFFFu, keep);
        if (lane == 0) s_warpCount[warp] = __popc(ballot);
        __syncthreads();
        int offset = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const int c = s_warpCount[w];
            if (w < warp) offset += c;
            total += c;
        }
        if (keep) s_list[offset + __popc(ballot & ((1u << lane) - 1u))] = (uint16_t)tid;
        __syncthreads();
        for (int k = 0; k < total; k++) {
            const ilb_obstruction& o = P.obs[base + (int)s_list[k]];
            const float msize = fmaxf(fmaxf(fabsf(o.size[0]), fabsf(o.size[1])), fabsf(o.size[2])) + P.maxEnc + 4.0f;
            if (fabsf(wx - o.center[0]) > msize || fabsf(wy - o.center[1]) > msize) continue;
            const f3 c = mk3(o.center[0], o.center[1], o.center[2]), sz = mk3(o.size[0], o.size[1], o.size[2]);
            const f4 rot = mk4(o.rotation[0], o.rotation[1], o.rotation[2], o.rotation[3]);
#pragma unroll
            for (int z = 0; z < 4; z++) {
                const float d = evaluateByTypeId(o.type, mk3(wx, wy, sliceZ[z]), c, sz, rot);
                const float e = ILB_DISTANCE_ZERO - (d / P.maxEnc);  // encodeDistance DistanceFieldCommon.fxh:264-266
                best[z] = fmaxf(best[z], e);                         // BlendFunction.Max, LoadMaterials.cs:171-175
            }
        }
        __syncthreads();
    }
    if (valid) {
        uint32_t q[4];
#pragma unroll
        for (int z = 0; z < 4; z++) q[z] = (uint32_t)floorf(saturatef(best[z]) * 65535.0f + 0.5f);  // UNORM16 store
        const int ox = (p % P.columns) * P.slice_w, oy = (p / P.columns) * P.slice_h;
        P.tex[(size_t)(oy + y) * (size_t)P.tw + (size_t)(ox + x)] = make_uint2(q[0] | (q[1] << 16), q[2] | (q[3] << 16));
    }
}

}  // namespace

int ilb_dfgen_launch(ilb_ctx* ctx, uint2* tex, const uint2* base, int tw, int th, int slice_w, int slice_h, int slice_count,
                     const ilb_df_uniforms* u, const ilb_obstruction* obs, int count) {
    DFGenParams P;
    memset(&P, 0, sizeof(P));
    P.tex = tex; P.base = base; P.tw = tw; P.th = th; P.slice_w = slice_w; P.slice_h = slice_h; P.slice_count = slice_count;
    P.columns = (int)u->TextureSliceCount.x;
    P.physical = (slice_count + 2) / 3;
    if (P.columns < 1 || slice_w < 1 || slice_h < 1 || slice_count < 1)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad distance-field geometry");
    const int rows = (P.physical + P.columns - 1) / P.columns;
    if (P.columns * slice_w > tw || rows * slice_h > th)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "atlas %dx%d too small for %d slices of %dx%d in %d columns", tw, th, P.physical, slice_w, slice_h, P.columns);
    P.maxEnc = u->Extent.w; P.zOffset = u->ConeAndMisc.y; P.depth = u->Extent.z;
    P.invX = u->ConeAndMisc.w; P.invY = u->StepAndMisc2.w;
    ilb_obstruction* d_obs = nullptr;
    if (count > 0) {
        ILB_CUDA(ctx, cudaMallocAsync(&d_obs, sizeof(ilb_obstruction) * (size_t)count, ctx->stream));
        ILB_CUDA(ctx, cudaMemcpyAsync(d_obs, obs, sizeof(ilb_obstruction) * (size_t)count, cudaMemcpyHostToDevice, ctx->stream));
    }
    P.obs = d_obs; P.count = count;
    if (base) ILB_CUDA(ctx, cudaMemcpyAsync(tex, base, sizeof(uint2) * (size_t)tw * (size_t)th, cudaMemcpyDeviceToDevice, ctx->stream));
    else ILB_CUDA(ctx, cudaMemsetAsync(tex, 0, sizeof(uint2) * (size_t)tw * (size_t)th, ctx->stream));
    const dim3 grid((slice_w + GEN_TILE - 1) / GEN_TILE, (slice_h + GEN_TILE - 1) / GEN_TILE, P.physical);
    df_generate_kernel<<<grid, dim3(GEN_TILE, GEN_TILE), 0, ctx->stream>>>(P);
    ctx->launches++;
    ILB_CUDA(ctx, cudaGetLastError());
    if (d_obs) ILB_CUDA(ctx, cudaFreeAsync(d_obs, ctx->stream));
    // obs is caller-owned host memory read by an async copy: finish before returning
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ILB_OK;
}
