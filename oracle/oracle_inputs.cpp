// TEST INFRASTRUCTURE -- CPU oracle for the INPUT PRODUCERS next to the hot paths:
// analytic-obstruction distance-field rasterisation and the G-buffer texel encoding.
// Only tests/, smoke() and bench.py's cpu_baseline / --impl reference legs may call this.
// PARITY UNPINNED by reference outputs; restated from the shaders cited below (relative to Illuminant/).
#include <omp.h>

#include <cstring>

#include "../include/illuminant_b200.h"
#include "hlsl.hpp"
#include "oracle.h"

using namespace hlsl;

extern "C" {

void orc_detmath(int function, const float* x, float* out, long n) {
    for (long i = 0; i < n; i++) out[i] = function == 0 ? dm_sinf(x[i]) : (function == 1 ? dm_cosf(x[i]) : dm_acosf(x[i]));
}

// RenderDistanceField for analytic obstructions: per physical slice p the texel (r,g,b,a) holds encoded
// distances at z = SliceIndexToZ(3p..3p+3) (Lighting/LightingRenderer.DistanceField.cs:32-35, :347-400);
// slice cleared to 0 (Shaders/ClearDistanceField.fx:27-39); each obstruction's quad covers
// center.xy +- (max|size| + MaximumEncodedDistance + 4) (Shaders/DistanceFunction.fx:15-27) and is MAX-blended
// (LoadMaterials.cs:171-175) with encodeDistance(evaluateX(...)) (DistanceFunction.fx:34-48,
// DistanceFieldCommon.fxh:264-266); the render target is UNORM16 (round-to-nearest, saturating).
// `base` (nullable): the static field of a DynamicDistanceField; the slice is then "cleared" to the static texel
// (ClearDistanceField.fx:27-39 with ClearTexture = StaticTexture, LightingRenderer.DistanceField.cs:113-118) instead of 0.
int orc_generate_distance_field(uint16_t* out, const uint16_t* base, int tw, int th, int slice_w, int slice_h, int slice_count,
                                const ilb_df_uniforms* u, const ilb_obstruction* obs, int count, int nthreads) {
    if (base) memcpy(out, base, sizeof(uint16_t) * 4 * (size_t)tw * th);
    else memset(out, 0, sizeof(uint16_t) * 4 * (size_t)tw * th);
    return orc_update_distance_field_slices(out, base, tw, th, slice_w, slice_h, slice_count, u, obs, count, nullptr, 0, nullptr, 0, 0,
                                            (slice_count + 2) / 3, nthreads);
}

}  // extern "C"

namespace {
// ---- Shaders/DistanceField.fx (height volumes) ---------------------------------------------------------------------
// sdPolygonInit / sdPolygonVertex: sq/Fracture SDF2D.fxh is not vendored; restated from the published definition it
// implements (Quilez, sdPolygon): d = squared distance to the closest edge, s flips for every edge that the ray from p
// towards +x crosses.  vi / vj are passed as (b, a) by the shader (DistanceField.fx:88,94): vi = edge end, vj = edge start.
void sdPolygonVertex(float2 p, float2 vi, float2 vj, float& d, float& s) {
    float2 e = vj - vi, w = p - vi;
    float2 b = w - e * clamp(dot(w, e) / dot(e, e), 0.0f, 1.0f);
    d = fminf(d, dot(b, b));
    bool c0 = p.y >= vi.y, c1 = p.y < vj.y, c2 = (e.x * w.y) > (e.y * w.x);
    if ((c0 && c1 && c2) || (!c0 && !c1 && !c2)) s = -s;
}
void sdPolygonInit(float2 p, float2 vi, float2 vj, float& d, float& s) {
    d = dot(p - vi, p - vi);
    s = 1.0f;
    sdPolygonVertex(p, vi, vj, d, s);
}
float computeDistanceZ(float sliceZ, float2 zRange) {  // :47-55
    if (sliceZ >= zRange.x) {
        if (sliceZ <= zRange.y) return fmaxf(sliceZ - zRange.y, zRange.x - sliceZ);
        return sliceZ - zRange.y;
    }
    return zRange.x - sliceZ;
}
float finalEval(float z, float2 zRange, float resultDistanceSq, float sign) {  // :57-74, PolygonXyBias 1.5 (:13)
    float distanceZ = computeDistanceZ(z, zRange);
    float distanceXy = (sqrtf(resultDistanceSq) * sign) + 1.5f;
    if (distanceXy <= 0) {
        if (distanceZ <= 0) return distanceXy + distanceZ;
        return distanceZ;
    }
    return fmaxf(distanceXy, 0.0f) + fmaxf(distanceZ, 0.0f);
}
void polygonDistance(const ilb_height_volume& hv, const ilb_float4* edges, float2 xy, float& d, float& s) {  // computeSliceDistances :76-93
    const ilb_float4* e = edges + hv.first_edge;
    sdPolygonInit(xy, float2(e[0].z, e[0].w), float2(e[0].x, e[0].y), d, s);
    for (int j = 1; j < hv.edge_count; j++) sdPolygonVertex(xy, float2(e[j].z, e[j].w), float2(e[j].x, e[j].y), d, s);
}
}  // namespace

extern "C" {

float orc_height_volume_distance(const ilb_height_volume* hv, const ilb_float4* edges, float x, float y, float z) {
    float d, s;
    polygonDistance(*hv, edges, float2(x, y), d, s);
    return finalEval(z, float2(hv->z_base, hv->z_base + hv->height), d, s);
}

int orc_update_distance_field_slices(uint16_t* out, const uint16_t* base, int tw, int th, int slice_w, int slice_h, int slice_count,
                                     const ilb_df_uniforms* u, const ilb_obstruction* obs, int count, const ilb_height_volume* volumes,
                                     int nvolumes, const ilb_float4* edges, int nedges, int first_physical, int physical_count, int nthreads) {
    if (nthreads > 0) omp_set_num_threads(nthreads);
    const int physical = (slice_count + 2) / 3;
    const int columns = (int)u->TextureSliceCount.x;
    const float maxEnc = u->Extent.w, zOffset = u->ConeAndMisc.y, depth = u->Extent.z;
    const float invX = u->ConeAndMisc.w, invY = u->StepAndMisc2.w;
    const float DISTANCE_ZERO = 192.0f / 255.0f;
    const float DistanceLimit = 520.0f;  // LightingRenderer.cs:316
    if (first_physical < 0 || physical_count < 0 || first_physical + physical_count > physical || columns < 1) return ILB_ERR_INVALID_ARGUMENT;
    for (int v = 0; v < nvolumes; v++)
        if (volumes[v].edge_count < 1 || volumes[v].first_edge < 0 || volumes[v].first_edge + volumes[v].edge_count > nedges) return ILB_ERR_INVALID_ARGUMENT;
    for (int p = first_physical; p < first_physical + physical_count; p++) {
        const int ox = (p % columns) * slice_w, oy = (p / columns) * slice_h;
        if (ox + slice_w > tw || oy + slice_h > th) return ILB_ERR_INVALID_ARGUMENT;
        float sliceZ[4];
        for (int k = 0; k < 4; k++) {
            float s = ((float)(3 * p + k) / fmaxf(1.0f, (float)slice_count));
            sliceZ[k] = (s * depth) + zOffset;
        }
#pragma omp parallel for schedule(dynamic, 8)
        for (int y = 0; y < slice_h; y++)
            for (int x = 0; x < slice_w; x++) {
                float wx = (float)x * invX, wy = (float)y * invY;  // getPositionXy, DistanceFunction.fx:29-32
                float best[4] = {0, 0, 0, 0};
                if (base) {
                    const uint16_t* b = base + 4 * ((size_t)(oy + y) * tw + (ox + x));
                    for (int k = 0; k < 4; k++) best[k] = (float)b[k] * (1.0f / 65535.0f);
                }
                for (int i = 0; i < count; i++) {
                    const ilb_obstruction& o = obs[i];
                    float msize = fmaxf(fmaxf(fabsf(o.size[0]), fabsf(o.size[1])), fabsf(o.size[2])) + maxEnc + 4;
                    if (fabsf(wx - o.center[0]) > msize || fabsf(wy - o.center[1]) > msize) continue;
                    for (int k = 0; k < 4; k++) {
                        float wp[3] = {wx, wy, sliceZ[k]};
                        float d = orc_evaluate_by_type_id(o.type, wp, o.center, o.size, o.rotation);
                        float e = DISTANCE_ZERO - (d / maxEnc);
                        best[k] = fmaxf(best[k], e);
                    }
                }
                for (int v = 0; v < nvolumes; v++) {  // RenderDistanceFieldHeightVolumes (LightingRenderer.DistanceField.cs:185-260)
                    const ilb_height_volume& hv = volumes[v];
                    if (wx < hv.bounds[0] - DistanceLimit || wx > hv.bounds[2] + DistanceLimit || wy < hv.bounds[1] - DistanceLimit ||
                        wy > hv.bounds[3] + DistanceLimit)
                        continue;
                    float d2, sgn;
                    polygonDistance(hv, edges, float2(wx, wy), d2, sgn);
                    float2 zRange(hv.z_base, hv.z_base + hv.height);
                    for (int k = 0; k < 4; k++) {
                        float e = DISTANCE_ZERO - (finalEval(sliceZ[k], zRange, d2, sgn) / maxEnc);   // encodeDistance
                        best[k] = fmaxf(best[k], e);
                    }
                }
                uint16_t* t = out + 4 * ((size_t)(oy + y) * tw + (ox + x));
                for (int k = 0; k < 4; k++) t[k] = (uint16_t)floorf(saturate(best[k]) * 65535.0f + 0.5f);
            }
    }
    return 0;
}

// encodeGBufferSample (Shaders/GBufferShaderCommon.fxh:10-35) with encodeNormalSpherical (EnvironmentCommon.fxh:34-40)
void orc_encode_gbuffer_sample(const float* normal, float relativeY, float z, int dead, int enableShadows,
                               int fullbright, float* out4) {
    if (dead) {
        out4[0] = 0; out4[1] = 0; out4[2] = -99999; out4[3] = -99999;
        return;
    }
    float3 n(normal[0], normal[1], normal[2]);
    float2 enc(0.0f);
    if (any(n)) {
        if (fabsf(n.x) < 0.0001f) n.x = 0.0001f;
        enc = (float2(atan2f(n.y, n.x) / PI, n.z) + 1.0f) * 0.5f;
    }
    out4[0] = enc.x; out4[1] = enc.y; out4[2] = relativeY;
    out4[3] = fullbright ? 99999.0f
                         : (((z + 1024.0f) / 1024.0f) * (enableShadows ? 1.0f : -1.0f)) + (enableShadows ? 0.0f : -1.0f);
}

// IEEE binary16 conversions (round-to-nearest-even), i.e. XNA HalfVector4 packing
void orc_float_to_half(const float* in, uint16_t* out, long n) {
    for (long i = 0; i < n; i++) {
        uint32_t x; memcpy(&x, &in[i], 4);
        uint32_t sign = (x >> 16) & 0x8000u, mant = x & 0x7FFFFFu;
        int exp = (int)((x >> 23) & 0xFF);
        uint16_t h;
        if (exp == 255) h = (uint16_t)(sign | 0x7C00u | (mant ? 0x200u : 0));
        else {
            int e = exp - 127 + 15;
            if (e >= 31) h = (uint16_t)(sign | 0x7C00u);
            else if (e <= 0) {
                if (e < -10) h = (uint16_t)sign;
                else {
                    mant |= 0x800000u;
                    int shift = 14 - e;
                    uint32_t hm = mant >> shift, rem = mant & ((1u << shift) - 1), half = 1u << (shift - 1);
                    if (rem > half || (rem == half && (hm & 1))) hm++;
                    h = (uint16_t)(sign | hm);
                }
            } else {
                uint32_t hm = mant >> 13, rem = mant & 0x1FFFu;
                uint32_t v = ((uint32_t)e << 10) | hm;
                if (rem > 0x1000u || (rem == 0x1000u && (hm & 1))) v++;
                h = (uint16_t)(sign | v);
            }
        }
        out[i] = h;
    }
}

void orc_half_to_float(const uint16_t* in, float* out, long n) {
    for (long i = 0; i < n; i++) {
        uint16_t h = in[i];
        uint32_t sign = (uint32_t)(h & 0x8000) << 16, exp = (h >> 10) & 0x1F, man = h & 0x3FF, bits;
        if (exp == 0) {
            if (man == 0) bits = sign;
            else {
                int e = -1;
                do { e++; man <<= 1; } while ((man & 0x400) == 0);
                bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FF) << 13);
            }
        } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
        else bits = sign | ((exp - 15 + 127) << 23) | (man << 13);
        memcpy(&out[i], &bits, 4);
    }
}

}  // extern "C"
