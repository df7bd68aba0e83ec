// TEST INFRASTRUCTURE -- CPU oracle for illuminant_b200 (see oracle/README.md).  "Next" row N3 of SURVEY.md section 8f:
// the lightmap resolve (Illuminant/Shaders/Resolve.fx, HDR.fxh) and the luminance buffer behind TryComputeHistogram.
// Follows the reference shaders line by line; file:line citations are relative to the reference tree (Illuminant/...).
//
// PARITY UNPINNED (like the rest of the oracle): the reference holds no tests or fixtures for this path, and three helpers
// it calls live in the un-vendored, un-pinned sq/Fracture (Squared/RenderLib/Shaders): pSRGBToPLinear / pLinearToPSRGB
// (sRGBCommon.fxh) are restated here from the published IEC 61966-2-1 transfer functions applied to the un-premultiplied
// colour; ApplyDither (DitherCommon.fxh) is the identity at the handler's default Strength 0 (LightingRenderer.cs:1489-1494) and
// otherwise follows the CONVENTION stated at ilb_dithering in include/illuminant_b200.h; ReadLUT (LUTCommon.fxh) follows the
// convention stated at ilb_lut_blending.  Neither of the two can be checked against the reference's source.
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

// ---- ApplyDither: the convention of ilb_dithering (the reference's lives in the un-vendored sq/Fracture DitherCommon.fxh)
ilb_dithering g_dither = {0.0f, 255.0f, 0.0f, 1.0f, 0.0f, 1.0f};
float3 ApplyDither(const ilb_resolve& P, float3 rgb, int x, int y) {
    const float strength = (P.DitheringStrength != 0.0f) ? P.DitheringStrength : g_dither.Strength;
    if (strength == 0.0f) return rgb;
    const float unit = (g_dither.Unit != 0.0f) ? g_dither.Unit : 255.0f, invUnit = 1.0f / unit;
    const float band = (g_dither.BandSize != 0.0f) ? g_dither.BandSize : 1.0f;
    const float lo = g_dither.RangeMin, hi = (g_dither.RangeMax > g_dither.RangeMin) ? g_dither.RangeMax : 1.0f;
    const float f = fmodf(g_dither.FrameIndex, 4.0f) + 0.5f, ph = 23.0f * f / 17.0f, phase = ph - floorf(ph);
    const float s = (float)((2 * x + 7 * y) % 17) * (1.0f / 17.0f) + phase;
    const float t = (s - floorf(s)) * band;
    auto one = [&](float c) {
        const float c8 = c * unit;
        const float a = truncf(c8), b = ceilf(c8);
        const float q = (((c8 - a) >= t) ? b : a) * invUnit;
        return ((c >= lo) && (c <= hi)) ? c + strength * (q - c) : c;
    };
    return float3(one(rgb.x), one(rgb.y), one(rgb.z));
}

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
    return result;  // the callers apply ApplyDither(result.rgb, vpos)
}
float4 dithered(const ilb_resolve& P, float4 c, int x, int y) { return float4(ApplyDither(P, c.xyz(), x, y), c.w); }

}  // namespace

extern "C" int orc_resolve_lighting(const ilb_resolve* p, const float* lightmap, const float* albedo, float* out) {
    if (!p || !lightmap || !out) return -1;
    ilb_resolve P = *p;
    if (P.InverseScaleFactor == 0.0f) P.InverseScaleFactor = 1.0f;  // LightingRenderer.cs:1469-1473
    const long n = (long)P.width * P.height;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) {
        const float* l = lightmap + 4 * i;
        float4 r = dithered(P, resolvePixel(P, float4(l[0], l[1], l[2], l[3]), albedo ? albedo + 4 * i : nullptr), (int)(i % P.width), (int)(i / P.width));
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
            const float4 r = dithered(P, resolvePixel(P, light, ap), x, y);
            float* o = target + 4 * ((size_t)y * place->target_width + x);
            o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
        }
    return 0;
}

extern "C" void orc_set_dithering(const ilb_dithering* d) {
    const ilb_dithering defaults = {0.0f, 255.0f, 0.0f, 1.0f, 0.0f, 1.0f};
    g_dither = d ? *d : defaults;
}

namespace {
// ReadLUT: the convention of ilb_lut_blending (sq/Fracture LUTCommon.fxh is un-vendored).  `tex`: float4 texels
float3 ReadLUT(const float* tex, int res, int rows, float3 value, float offU, float offV) {
    const float resm1 = (float)(res - 1);
    const float blue = value.z * resm1;
    const float s0 = floorf(blue), s1 = fminf(s0 + 1.0f, resm1), w = blue - s0;
    const int tw = res * res, th = res * rows;
    const float invW = 1.0f / (float)tw, invH = 1.0f / (float)th;
    const float uIn = 0.5f + value.x * resm1, v = (0.5f + value.y * resm1) * invH + offV;
    const float3 a = sampleLinearClamp(tex, tw, th, (s0 * (float)res + uIn) * invW + offU, v).xyz();
    const float3 b = sampleLinearClamp(tex, tw, th, (s1 * (float)res + uIn) * invW + offU, v).xyz();
    return lerp(a, b, w);
}
const float3 RgbToGrayLUT = float3(0.299f, 0.587f, 0.144f);  // LUTResolve.fx:16 (sic)
}  // namespace

// LUTBlendedResolveWithAlbedoCommon (LUTResolve.fx:57-117) + LUTBlendedLightingResolveWithAlbedoPixelShader (:119-135), 1:1
extern "C" int orc_resolve_lighting_lut(const ilb_resolve* p, const ilb_lut_blending* lut, const float* dark, const float* bright,
                                        const float* lightmap, const float* albedoTex, float* out) {
    if (!p || !lut || !dark || !bright || !lightmap || !albedoTex || !out) return -1;
    if (p->hdr_mode != ILB_HDR_NONE) return -2;  // "LUT blending is not compatible with this type of lighting resolve"
    ilb_resolve P = *p;
    if (P.InverseScaleFactor == 0.0f) P.InverseScaleFactor = 1.0f;
    const float3 LUTLevels = float3(lut->DarkLevel, lut->NeutralBandSize, lut->BrightLevel);
    const long n = (long)P.width * P.height;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) {
        float4 light = float4(lightmap[4 * i], lightmap[4 * i + 1], lightmap[4 * i + 2], lightmap[4 * i + 3]);
        float4 albedo = float4(albedoTex[4 * i], albedoTex[4 * i + 1], albedoTex[4 * i + 2], albedoTex[4 * i + 3]);
        if (P.AlbedoIsSRGB != 0.0f) albedo = pSRGBToPLinear(albedo);
        light *= P.InverseScaleFactor * 2;
        float3 weight = light.xyz();
        float bandWidth = saturate(LUTLevels.z - LUTLevels.x);
        float neutralBandWidth = fminf(LUTLevels.y, bandWidth - 0.01f);
        bool hasNeutralBand = (neutralBandWidth > 0);
        bool normalize = !(lut->PerChannel != 0.0f) || hasNeutralBand;
        if (normalize) {
            weight *= RgbToGrayLUT;
            weight = float3(weight.x + weight.y + weight.z);
        }
        float3 a = saturate(albedo.xyz());
        float3 lutValue1 = ReadLUT(dark, lut->dark_resolution, lut->dark_row_count, a, lut->LUTOffsets[0], lut->LUTOffsets[1]);
        float3 lutValue2 = ReadLUT(bright, lut->bright_resolution, lut->bright_row_count, a, lut->LUTOffsets[2], lut->LUTOffsets[3]);
        float3 blendedValue;
        if (hasNeutralBand) {
            float transitionSize = (bandWidth - neutralBandWidth) * 0.5f;
            float v = weight.x - LUTLevels.x, v2 = v - transitionSize, v3 = v2 - neutralBandWidth;
            float3 val1 = lerp(lutValue1, a, saturate(v / transitionSize));
            blendedValue = lerp(val1, lutValue2, saturate(v3 / transitionSize));
        } else {
            if (LUTLevels.z > LUTLevels.x) {
                weight = weight - LUTLevels.x;
                weight = max(float3(0.0f), weight);
                weight = weight / (LUTLevels.z - LUTLevels.x);
                weight = saturate(weight);
            } else {  // HACK (:105-108)
                weight = weight - LUTLevels.x;
                weight = saturate(weight);
            }
            blendedValue = lutValue1 + weight * (lutValue2 - lutValue1);  // lerp with a per-channel weight
        }
        float4 result = float4(blendedValue * ((lut->LUTOnly != 0.0f) ? float3(1.0f) : light.xyz()), albedo.w);
        float3 rgb = max(float3(0.0f), result.xyz() + P.Offset);
        rgb *= (P.ExposureMinusOne + 1);
        rgb = pow3(rgb, P.GammaMinusOne + 1);
        result = float4(rgb, result.w);
        if (P.ResolveToSRGB != 0.0f) result = pLinearToPSRGB(result);
        result = dithered(P, result, (int)(i % P.width), (int)(i / P.width));
        out[4 * i + 0] = result.x; out[4 * i + 1] = result.y; out[4 * i + 2] = result.z; out[4 * i + 3] = result.w;
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
