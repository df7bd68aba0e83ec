// TEST INFRASTRUCTURE -- CPU oracle for illuminant_b200 (see oracle/README.md).  "Next" row N3 of SURVEY.md section 8f:
// the lightmap resolve (Illuminant/Shaders/Resolve.fx, HDR.fxh) and the luminance buffer behind TryComputeHistogram.
// Follows the reference shaders line by line; file:line citations are relative to the reference tree (Illuminant/...).
//
// PARITY UNPINNED (like the rest of the oracle): the reference holds no tests or fixtures for this path, and three helpers
// it calls live in the un-vendored, un-pinned sq/Fracture (Squared/RenderLib/Shaders): pSRGBToPLinear / pLinearToPSRGB
// (sRGBCommon.fxh) are restated here from the published IEC 61966-2-1 transfer functions applied to the un-premultiplied
// colour; ApplyDither (DitherCommon.fxh) is the identity at the handler's default Strength 0 (LightingRenderer.cs:1489-1494),
// the only value the boundary accepts.
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
