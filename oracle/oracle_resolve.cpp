// TEST INFRASTRUCTURE -- CPU oracle for illuminant_b200 (see oracle/README.md).  "Next" row N3 of SURVEY.md section 8f:
// the lightmap resolve (Illuminant/Shaders/Resolve.fx, HDR.fxh) and the luminance buffer behind TryComputeHistogram.
// Follows the reference shaders line by line; file:line citations are relative to the reference tree (Illuminant/...).
//
// PARITY UNPINNED (like the rest of the oracle): the reference holds no tests or fixtures for this path, and three helpers
// it calls live in the un-vendored, un-pinned sq/Fracture (Squared/RenderLib/Shaders): pSRGBToPLinear / pLinearToPSRGB
// (sRGBCommon.fxh) are restated here from the published IEC 61966-2-1 transfer functions applied to the un-premultiplied
// colour; ApplyDither (DitherCommon.fxh) is the identity at the handler's default Strength 0 (LightingRenderer.cs:1489-1494),
// the only value the boundary accepts.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "hlsl.hpp"
#include "oracle.h"

using namespace hlsl;

namespace {

// ---- sq/Fracture sRGBCommon.fxh (un-vendored): IEC 61966-2-1
float SRGBToLinear1(float s) { return (s <= 0.04045f) ? s / 12.92f : powf((s + 0.055f) / 1.055f, 2.4f); }
float LinearToSRGB1(float l) { return (l <= 0.0031308f) ? l * 12.92f : 1.055f * powf(l, 1.0f / 2.4f) - 0.055f; }
float4 pSRGBToPLinear(float4 c) {
    if (!(c.w > 0.0f)) return float4(0.0f, 0.0f, 0.0f, c.w);
    float3 s = c.xyz() / c.w;
    float3 l = float3(SRGBToLinear1(s.x), SRGBToLinear1(s.y), SRGBToLinear1(s.z));
    return float4(l * c.w, c.w);
}
float4 pLinearToPSRGB(float4 c) {
    if (!(c.w > 0.0f)) return float4(0.0f, 0.0f, 0.0f, c.w);
    float3 l = c.xyz() / c.w;
    float3 s = float3(LinearToSRGB1(l.x), LinearToSRGB1(l.y), LinearToSRGB1(l.z));
    return float4(s * c.w, c.w);
}

// ---- HDR.fxh
const float3 RGBToLuminance = float3(0.299f, 0.587f, 0.114f);  // HDR.fxh:10

float4 GammaCompress(const ilb_resolve& P, float4 color) {  // HDR.fxh:12-19
    float3 rgb = max(color.xyz() + P.Offset, float3(0.0f));
    float resultLuminance = dot(rgb, RGBToLuminance);
    float scaledLuminance = (resultLuminance * P.MiddleGray) / P.AverageLuminance;
    float compressedLuminance = (scaledLuminance * (1 + (scaledLuminance / P.MaximumLuminanceSquared))) / (1 + scaledLuminance);
    float rescaleFactor = compressedLuminance / resultLuminance;
    return float4(rgb * rescaleFactor, color.w);
}

const float kA = 0.15f, kB = 0.50f, kC = 0.10f, kD = 0.20f, kE = 0.02f, kF = 0.30f;  // HDR.fxh:25-30

float Uncharted2Tonemap1(float value) {  // HDR.fxh:32-38
    return ((value * (kA * value + kC * kB) + kD * kE) / (value * (kA * value + kB) + kD * kF)) - kE / kF;
}
float3 Uncharted2Tonemap(float3 rgb) {  // HDR.fxh:40-46
    return float3(Uncharted2Tonemap1(rgb.x), Uncharted2Tonemap1(rgb.y), Uncharted2Tonemap1(rgb.z));
}

float3 pow3(float3 v, float e) { return float3(powf(v.x, e), powf(v.y, e), powf(v.z, e)); }

// ---- Resolve.fx
float4 ResolveCommon(const ilb_resolve& P, float4 color) {  // Resolve.fx:30-45 (texel fetch done by the caller)
    float4 result = color * P.InverseScaleFactor;
    result.w = 1;
    return result;
}

float4 ResolveWithAlbedoCommon(const ilb_resolve& P, float4 light, float4 albedo) {  // Resolve.fx:47-68
    if (P.AlbedoIsSRGB != 0.0f) albedo = pSRGBToPLinear(albedo);
    light *= P.InverseScaleFactor * 2;
    float3 a = albedo.xyz();
    return float4(lerp(a, a * light.xyz(), saturate(light.w)), albedo.w);
}

float4 resolvePixel(const ilb_resolve& P, float4 light, const float* albedoTexel) {
    float4 result = albedoTexel ? ResolveWithAlbedoCommon(P, light, float4(albedoTexel[0], albedoTexel[1], albedoTexel[2], albedoTexel[3]))
                                : ResolveCommon(P, light);
    float3 rgb;
    switch (P.hdr_mode) {
        case ILB_HDR_GAMMA_COMPRESS:  // Resolve.fx:92-112, :160-181
            result = GammaCompress(P, result);
            break;
        case ILB_HDR_TONE_MAP: {  // Resolve.fx:114-137, :183-217
            float3 preToneMap = max(float3(0.0f), result.xyz() + P.Offset) * (P.ExposureMinusOne + 1);
            result = float4(Uncharted2Tonemap(preToneMap) / Uncharted2Tonemap1(P.WhitePoint), result.w);
            result = float4(pow3(result.xyz(), P.GammaMinusOne + 1), result.w);
            break;
        }
        default:  // Resolve.fx:70-90, :139-158
            rgb = max(float3(0.0f), result.xyz() + P.Offset);
            rgb *= (P.ExposureMinusOne + 1);
            rgb = pow3(rgb, P.GammaMinusOne + 1);
            result = float4(rgb, result.w);
            break;
    }
    if (P.ResolveToSRGB != 0.0f) result = pLinearToPSRGB(result);
    // ApplyDither(result.rgb, vpos): identity at DitheringStrength == 0
    return result;
}

}  // namespace

extern "C" int orc_resolve_lighting(const ilb_resolve* p, const float* lightmap, const float* albedo, float* out) {
    if (!p || !lightmap || !out) return -1;
    ilb_resolve P = *p;
    if (P.InverseScaleFactor == 0.0f) P.InverseScaleFactor = 1.0f;  // LightingRenderer.cs:1469-1473
    const long n = (long)P.width * P.height;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) {
        const float* l = lightmap + 4 * i;
        float4 r = resolvePixel(P, float4(l[0], l[1], l[2], l[3]), albedo ? albedo + 4 * i : nullptr);
        out[4 * i + 0] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
    }
    return 0;
}

namespace {
// LINEAR / CLAMP fetch at texture coordinates (u, v) of a w x h float4 texture: fp32 bilinear weights
float4 sampleLinearClamp(const float* tex, int w, int h, float u, float v) {
    const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    const float x0f = floorf(x), y0f = floorf(y);
    const float fx = x - x0f, fy = y - y0f;
    auto at = [&](int ix, int iy) {
        ix = std::min(std::max(ix, 0), w - 1);
        iy = std::min(std::max(iy, 0), h - 1);
        const float* t = tex + 4 * ((size_t)iy * w + ix);
        return float4(t[0], t[1], t[2], t[3]);
    };
    const int x0 = (int)x0f, y0 = (int)y0f;
    const float4 top = lerp(at(x0, y0), at(x0 + 1, y0), fx), bottom = lerp(at(x0, y0 + 1), at(x0 + 1, y0 + 1), fx);
    return lerp(top, bottom, fy);
}
}  // namespace

extern "C" int orc_resolve_lighting_placed(const ilb_resolve* p, const ilb_resolve_placement* place, const float* lightmap, const float* albedo,
                                           float* target) {
    if (!p || !place || !lightmap || !target) return -1;
    ilb_resolve P = *p;
    if (P.InverseScaleFactor == 0.0f) P.InverseScaleFactor = 1.0f;
    const float u0 = place->AlbedoRegion[0], v0 = place->AlbedoRegion[1], u1 = place->AlbedoRegion[2], v1 = place->AlbedoRegion[3];
    // the quad: the first texture's region in texels times Scale (BitmapDrawCall), at Position
    const float qw = (albedo ? (u1 - u0) * (float)place->albedo_width : (float)P.width) * place->Scale[0];
    const float qh = (albedo ? (v1 - v0) * (float)place->albedo_height : (float)P.height) * place->Scale[1];
    if (!(qw > 0.0f) || !(qh > 0.0f)) return 0;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < place->target_height; y++)
        for (int x = 0; x < place->target_width; x++) {
            const float tx = (((float)x + 0.5f) - place->Position[0]) / qw, ty = (((float)y + 0.5f) - place->Position[1]) / qh;
            if (!(tx >= 0.0f && tx < 1.0f && ty >= 0.0f && ty < 1.0f)) continue;
            // texCoord2 over the lightmap region (0, 0)-(1, 1) plus LightmapUVOffset, clamped to it (Resolve.fx:35-36, :52-53)
            const float lu = fminf(fmaxf(tx + P.LightmapUVOffset[0], 0.0f), 1.0f), lv = fminf(fmaxf(ty + P.LightmapUVOffset[1], 0.0f), 1.0f);
            const float4 light = sampleLinearClamp(lightmap, P.width, P.height, lu, lv);
            float a4[4];
            const float* ap = nullptr;
            if (albedo) {
                const float au = fminf(fmaxf(u0 + tx * (u1 - u0), u0), u1), av = fminf(fmaxf(v0 + ty * (v1 - v0), v0), v1);
                const float4 a = sampleLinearClamp(albedo, place->albedo_width, place->albedo_height, au, av);
                a4[0] = a.x; a4[1] = a.y; a4[2] = a.z; a4[3] = a.w;
                ap = a4;
            }
            const float4 r = resolvePixel(P, light, ap);
            float* o = target + 4 * ((size_t)y * place->target_width + x);
            o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
        }
    return 0;
}

// CalculateLuminancePixelShader (Resolve.fx:219-234) into the half-size SurfaceFormat.Single target of
// UpdateLuminanceBuffer (LightingRenderer.cs:855-898), then `level` box-filter mip steps.
extern "C" int orc_compute_luminance(const float* lightmap, int w, int h, int level, float* out) {
    if (!lightmap || !out || level < 0) return -1;
    const float3 RgbToGray = float3(0.299f, 0.587f, 0.144f);  // Resolve.fx:15 (sic)
    int lw = w / 2, lh = h / 2;
    if (lw <= 0 || lh <= 0) return -1;
    float* cur = new float[(size_t)lw * lh];
    for (int y = 0; y < lh; y++)
        for (int x = 0; x < lw; x++) {
            const float* t = lightmap + 4 * ((size_t)(2 * y + 1) * w + (2 * x + 1));
            float3 rgbScaled = float3(t[0], t[1], t[2]) * RgbToGray;
            cur[(size_t)y * lw + x] = (rgbScaled.x + rgbScaled.y + rgbScaled.z);
        }
    for (int k = 0; k < level; k++) {
        const int nw = lw / 2, nh = lh / 2;
        if (nw <= 0 || nh <= 0) { delete[] cur; return -1; }
        float* nxt = new float[(size_t)nw * nh];
        for (int y = 0; y < nh; y++)
            for (int x = 0; x < nw; x++) {
                const float* r0 = cur + (size_t)(2 * y) * lw + 2 * x;
                const float* r1 = r0 + lw;
                nxt[(size_t)y * nw + x] = ((r0[0] + r0[1]) + (r1[0] + r1[1])) * 0.25f;
            }
        delete[] cur;
        cur = nxt; lw = nw; lh = nh;
    }
    memcpy(out, cur, sizeof(float) * (size_t)lw * lh);
    delete[] cur;
    return 0;
}
