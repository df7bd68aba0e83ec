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
    const ilb_height_volume* volumes;  // height volumes (DistanceField.fx) and their packed edges
    const float4* edges;
    int nvolumes;
    int first_physical;                // the launch covers physical slices [first_physical, first_physical + gridDim.z)
};

// ---- Shaders/DistanceField.fx: signed distance to an extruded polygon ------------------------------------------------
// sdPolygonInit / sdPolygonVertex (un-vendored sq/Fracture SDF2D.fxh) restated from the published sdPolygon they implement:
// d = squared distance to the closest edge, s flips for every edge the ray from p towards +x crosses; the shader passes
// (vi, vj) = (edge end, edge start) (DistanceField.fx:88,94).  The crossing test compares two products exactly (x-ops): a
// fused multiply-add on one side could flip the sign for points on an edge's supporting line.
ILB_DEV void sdPolygonVertex(f2 p, f2 vi, f2 vj, float& d, float& s) {
    const f2 e = vj - vi, w = p - vi;
    const float t = fminf(fmaxf(dot2(w, e) / dot2(e, e), 0.0f), 1.0f);
    const f2 b = w - e * t;
    d = fminf(d, dot2(b, b));
    const bool c0 = p.y >= vi.y, c1 = p.y < vj.y, c2 = xmul(e.x, w.y) > xmul(e.y, w.x);
    if ((c0 && c1 && c2) || (!c0 && !c1 && !c2)) s = -s;
}
ILB_DEV float computeDistanceZ(float sliceZ, float z0, float z1) {  // :47-55
    if (sliceZ >= z0) return (sliceZ <= z1) ? fmaxf(sliceZ - z1, z0 - sliceZ) : (sliceZ - z1);
    return z0 - sliceZ;
}
ILB_DEV float finalEval(float z, float z0, float z1, float distanceSq, float sign) {  // :57-74, PolygonXyBias 1.5 (:13)
    const float distanceZ = computeDistanceZ(z, z0, z1);
    const float distanceXy = (sqrtf(distanceSq) * sign) + 1.5f;
    if (distanceXy <= 0.0f) return (distanceZ <= 0.0f) ? (distanceXy + distanceZ) : distanceZ;
    return fmaxf(distanceXy, 0.0f) + fmaxf(distanceZ, 0.0f);
}

constexpr int GEN_TILE = 16;

__global__ void __launch_bounds__(GEN_TILE * GEN_TILE) df_generate_kernel(const __grid_constant__ DFGenParams P) {
    __shared__ uint16_t s_list[GEN_TILE * GEN_TILE];
    __shared__ int s_warpCount[8];
    const int p = P.first_physical + blockIdx.z;  // physical slice
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
    // height volumes, in list order (RenderDistanceFieldHeightVolumes, LightingRenderer.DistanceField.cs:185-260): the quad is the
    // polygon's bounds expanded by DistanceLimit = 520 (LightingRenderer.cs:316), MAX-blended like the analytic obstructions
    if (valid) {
        for (int v = 0; v < P.nvolumes; v++) {
            const ilb_height_volume& hv = P.volumes[v];
            if (wx < hv.bounds[0] - 520.0f || wx > hv.bounds[2] + 520.0f || wy < hv.bounds[1] - 520.0f || wy > hv.bounds[3] + 520.0f) continue;
            const float4* e = P.edges + hv.first_edge;
            const f2 xy = mk2(wx, wy);
            float4 edge = __ldg(e);
            float d = dot2(xy - mk2(edge.z, edge.w), xy - mk2(edge.z, edge.w)), sgn = 1.0f;   // sdPolygonInit
            sdPolygonVertex(xy, mk2(edge.z, edge.w), mk2(edge.x, edge.y), d, sgn);
            for (int j = 1; j < hv.edge_count; j++) {
                edge = __ldg(e + j);
                sdPolygonVertex(xy, mk2(edge.z, edge.w), mk2(edge.x, edge.y), d, sgn);
            }
            const float z0 = hv.z_base, z1 = hv.z_base + hv.height;
#pragma unroll
            for (int z = 0; z < 4; z++) best[z] = fmaxf(best[z], ILB_DISTANCE_ZERO - (finalEval(sliceZ[z], z0, z1, d, sgn) / P.maxEnc));
        }
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
                     const ilb_df_uniforms* u, const ilb_obstruction* obs, int count, const ilb_height_volume* volumes, int volume_count,
                     const ilb_float4* edges, int edge_count, int first_physical, int physical_count) {
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
    if (first_physical < 0 || physical_count < 0 || first_physical + physical_count > P.physical)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "physical slices [%d,+%d) outside [0,%d)", first_physical, physical_count, P.physical);
    if (volume_count < 0 || edge_count < 0 || (volume_count > 0 && (!volumes || !edges)))
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null height-volume arrays");
    for (int v = 0; v < volume_count; v++)
        if (volumes[v].edge_count < 1 || volumes[v].first_edge < 0 || volumes[v].first_edge + volumes[v].edge_count > edge_count)
            return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "height volume %d: edges [%d,+%d) outside [0,%d)", v, volumes[v].first_edge, volumes[v].edge_count, edge_count);
    if (physical_count == 0) return ILB_OK;
    P.maxEnc = u->Extent.w; P.zOffset = u->ConeAndMisc.y; P.depth = u->Extent.z;
    P.invX = u->ConeAndMisc.w; P.invY = u->StepAndMisc2.w;
    P.first_physical = first_physical;
    // one staging allocation for the three host arrays (obstructions, volumes, edges), released on every path
    const size_t obsBytes = sizeof(ilb_obstruction) * (size_t)count, volBytes = sizeof(ilb_height_volume) * (size_t)volume_count;
    const size_t obsPad = (obsBytes + 15) & ~(size_t)15, volPad = (volBytes + 15) & ~(size_t)15, edgeBytes = sizeof(float4) * (size_t)edge_count;
    char* d_stage = nullptr;
    cudaError_t e = cudaSuccess;
    if (obsPad + volPad + edgeBytes > 0) {
        e = cudaMallocAsync(reinterpret_cast<void**>(&d_stage), obsPad + volPad + edgeBytes, ctx->stream);
        if (e != cudaSuccess) return ilb_cuda_fail(ctx, e, "allocate distance-field generation inputs");
        if (obsBytes) e = cudaMemcpyAsync(d_stage, obs, obsBytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess && volBytes) e = cudaMemcpyAsync(d_stage + obsPad, volumes, volBytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess && edgeBytes) e = cudaMemcpyAsync(d_stage + obsPad + volPad, edges, edgeBytes, cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e == cudaSuccess) {
        P.obs = reinterpret_cast<const ilb_obstruction*>(d_stage); P.count = count;
        P.volumes = reinterpret_cast<const ilb_height_volume*>(d_stage + obsPad); P.nvolumes = volume_count;
        P.edges = reinterpret_cast<const float4*>(d_stage + obsPad + volPad);
        // the kernel writes every texel of the slices it covers (clear value = 0 or the static texel), nothing else
        const dim3 grid((slice_w + GEN_TILE - 1) / GEN_TILE, (slice_h + GEN_TILE - 1) / GEN_TILE, physical_count);
        df_generate_kernel<<<grid, dim3(GEN_TILE, GEN_TILE), 0, ctx->stream>>>(P);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (d_stage) cudaFreeAsync(d_stage, ctx->stream);
    // the inputs are caller-owned host memory read by asynchronous copies: finish before returning
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return ilb_cuda_fail(ctx, e, "distance-field generation");
    return ILB_OK;
}
