// "Next" row N3: the lightmap resolve of Illuminant/Shaders/Resolve.fx + HDR.fxh as ONE streaming
// kernel, and the luminance buffer / mip chain behind RenderedLighting.TryComputeHistogram.
//
// resolve_kernel<MODE, ALBEDO, VEC>: one thread per group of four consecutive pixels (the screen-aligned 1:1 resolve has no
// dependence on the pixel's row, so the frame is a flat array).  HBM-bound by construction: per pixel it reads the lightmap
// texel (8 B HalfVector4) and the albedo texel (4 B Color) and writes one backbuffer texel (4 B Color) = 16 B (12 B without
// albedo); four pixels per thread make every access a 16- or 32-byte vector, loads are streaming (__ldcs: read once),
// stores too (__stcs).  Uniform branches skip pow() when Gamma == 1 and the sRGB transfer when it is off, otherwise the
// three accurate powf per pixel would make the kernel ALU-bound.
//
// The reference does this as a full-screen BitmapBatch draw with material {Screen,World}Space{,GammaCompressed,ToneMapped}
// LightingResolve{,WithAlbedo} (LightingRenderer.cs:1537-1645).
#include <cuda_fp16.h>

#include <algorithm>
#include <cstring>

#include "ilb_internal.h"

namespace {

struct ResolveParams {
    const void* lightmap;
    const void* albedo;
    void* out;
    unsigned long long n;  // pixels
    int lm_fmt, al_fmt, out_fmt;
    float invScale, invScale2;  // InverseScaleFactor, InverseScaleFactor * 2 (Resolve.fx:40, :61)
    float offset, exposure, gamma;
    float middleGray, averageLuminance, maxLumSq;
    float invWhiteScale;  // 1 / Uncharted2Tonemap1(WhitePoint), host-evaluated (HDR.fxh:32-38, Resolve.fx:131)
    int albedoIsSRGB, resolveToSRGB;
    // ApplyDither (convention of ilb_dithering): strength 0 = off
    float ditherStrength, ditherUnit, ditherInvUnit, ditherPhase, ditherBand, ditherMin, ditherMax;
    int width;  // row length, for the pixel coordinates the dither pattern needs
};

// ---- texel access ---------------------------------------------------------------------------------------------------
ILB_DEV f4 unpackHalf4(uint2 v) {
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    return mk4(lo.x, lo.y, hi.x, hi.y);
}
// UNORM8 -> float, c / 255, in two instructions per channel and without the conversion pipe: the byte-permute builds the bit
// pattern of 8388608 + c (0x4B0000cc) and one fused multiply-add computes RN((8388608 + c) * r - 8388608 * r) = RN(c * r),
// r = fl(1 / 255) (8388608 * r is exact).  RN(c * r) is within one ulp of the correctly rounded c / 255 (the oracle's decode),
// five orders of magnitude inside the resolve tolerance; and since |c * r * 255 - c| < 1e-4, packing an unchanged value returns
// the byte it came from (the pass-through properties of tests/test_gpu_resolve.py hold exactly).  This replaces a correctly
// rounded division in six instructions per channel: the tone-mapped resolve is bound by instruction issue, and a quarter of
// its instructions were these decodes.
template <int K>
ILB_DEV float unorm8ToFloat(uint32_t v) {
    const float r255 = 1.0f / 255.0f;
    return __fmaf_rn(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7540u | K)), r255, -8388608.0f * r255);
}
ILB_DEV f4 unpackRgba8(uint32_t v) { return mk4(unorm8ToFloat<0>(v), unorm8ToFloat<1>(v), unorm8ToFloat<2>(v), unorm8ToFloat<3>(v)); }
// the correctly rounded c / 255 (q = c * r, one Markstein correction): the luminance buffer is compared bit for bit
ILB_DEV f4 unpackRgba8Exact(uint32_t v) {
    const float r255 = 1.0f / 255.0f;
    return mk4(udiv((float)(v & 255u), 255.0f, r255), udiv((float)((v >> 8) & 255u), 255.0f, r255), udiv((float)((v >> 16) & 255u), 255.0f, r255),
               udiv((float)(v >> 24), 255.0f, r255));
}
// float -> UNORM8: floor(saturate(c) * 255 + 0.5), NaN -> 0 (saturatef).  The truncation is a round-toward-zero add of 2^23
// (the integer lands in the low mantissa bits: t < 256.5), again without the conversion pipe; three byte-permutes gather the bytes.
ILB_DEV uint32_t unorm8Bits(float c) { return __float_as_uint(__fadd_rz(saturatef(c) * 255.0f + 0.5f, 8388608.0f)); }
ILB_DEV uint32_t packRgba8(f4 c) {
    const uint32_t rg = __byte_perm(unorm8Bits(c.x), unorm8Bits(c.y), 0x3340u);   // bytes: r, g, *, *
    const uint32_t ba = __byte_perm(unorm8Bits(c.z), unorm8Bits(c.w), 0x3340u);
    return __byte_perm(rg, ba, 0x5410u);
}
ILB_DEV f4 loadTexel(const void* base, int fmt, unsigned long long i) {
    if (fmt == ILB_FORMAT_HALF4) return unpackHalf4(__ldcs(reinterpret_cast<const uint2*>(base) + i));
    if (fmt == ILB_FORMAT_FLOAT4) return mk4(__ldcs(reinterpret_cast<const float4*>(base) + i));
    return unpackRgba8(__ldcs(reinterpret_cast<const unsigned int*>(base) + i));
}
ILB_DEV void loadTexels4(const void* base, int fmt, unsigned long long g, f4 t[4]) {  // texels 4g .. 4g+3, 16-byte aligned base
    if (fmt == ILB_FORMAT_HALF4) {
        const uint4 a = __ldcs(reinterpret_cast<const uint4*>(base) + 2 * g), b = __ldcs(reinterpret_cast<const uint4*>(base) + 2 * g + 1);
        t[0] = unpackHalf4(make_uint2(a.x, a.y)); t[1] = unpackHalf4(make_uint2(a.z, a.w));
        t[2] = unpackHalf4(make_uint2(b.x, b.y)); t[3] = unpackHalf4(make_uint2(b.z, b.w));
    } else if (fmt == ILB_FORMAT_FLOAT4) {
#pragma unroll
        for (int k = 0; k < 4; k++) t[k] = mk4(__ldcs(reinterpret_cast<const float4*>(base) + 4 * g + k));
    } else {
        const uint4 a = __ldcs(reinterpret_cast<const uint4*>(base) + g);
        t[0] = unpackRgba8(a.x); t[1] = unpackRgba8(a.y); t[2] = unpackRgba8(a.z); t[3] = unpackRgba8(a.w);
    }
}

// ---- sq/Fracture sRGBCommon.fxh (un-vendored): IEC 61966-2-1 on the un-premultiplied colour, as in the oracle
ILB_DEV float srgbToLinear1(float s) { return (s <= 0.04045f) ? s / 12.92f : powf((s + 0.055f) / 1.055f, 2.4f); }
ILB_DEV float linearToSrgb1(float l) { return (l <= 0.0031308f) ? l * 12.92f : 1.055f * powf(l, 1.0f / 2.4f) - 0.055f; }
ILB_DEV f4 pSRGBToPLinear(f4 c) {
    if (!(c.w > 0.0f)) return mk4(0.0f, 0.0f, 0.0f, c.w);
    const f3 s = xyz(c) / c.w;
    return mk4(mk3(srgbToLinear1(s.x), srgbToLinear1(s.y), srgbToLinear1(s.z)) * c.w, c.w);
}
ILB_DEV f4 pLinearToPSRGB(f4 c) {
    if (!(c.w > 0.0f)) return mk4(0.0f, 0.0f, 0.0f, c.w);
    const f3 l = xyz(c) / c.w;
    return mk4(mk3(linearToSrgb1(l.x), linearToSrgb1(l.y), linearToSrgb1(l.z)) * c.w, c.w);
}

// ---- HDR.fxh
// value >= 0.  The curve is the difference of two nearly equal numbers near black (kD*kE / (kD*kF) - kE/kF = +7.45e-9 in IEEE
// fp32 at value == 0) and its SIGN decides whether a following pow(x, Gamma != 1) is a number or NaN.  In exact arithmetic the
// curve is >= 0 for value >= 0, so the quotient uses the fast division (with IEEE divisions the kernel is ALU-bound: 72 us per 4K
// frame, profiles/r1_n2n3_launches.csv) and the result is clamped at 0: within 1e-8 of the IEEE value at black, never NaN.
ILB_DEV float uncharted2Tonemap1(float value) {  // HDR.fxh:32-46
    const float kA = 0.15f, kB = 0.50f, kC = 0.10f, kD = 0.20f, kE = 0.02f, kF = 0.30f;
    return fmaxf(__fdividef(value * (kA * value + kC * kB) + kD * kE, value * (kA * value + kB) + kD * kF) - kE / kF, 0.0f);
}

ILB_DEV f3 pow3u(f3 v, float e) {  // pow(x, 1) == x exactly: skip the three powf when Gamma == 1 (uniform branch)
    if (e == 1.0f) return v;
    return mk3(powf(v.x, e), powf(v.y, e), powf(v.z, e));
}

// ApplyDither under the convention stated at ilb_dithering (the reference's lives in the un-vendored sq/Fracture): ordered
// dithering to multiples of 1 / Unit with a 17-periodic threshold, blended by Strength, inside [RangeMin, RangeMax]
ILB_DEV float ditherThreshold(const ResolveParams& P, int x, int y) {
    // individually rounded: the threshold is compared against, so the oracle must see the same bits
    const float s = xadd(xmul((float)((2 * x + 7 * y) % 17), 1.0f / 17.0f), P.ditherPhase);  // ditherPhase = frac(23 * ((FrameIndex mod 4) + 0.5) / 17)
    return xmul(xsub(s, floorf(s)), P.ditherBand);
}
ILB_DEV float dither1(const ResolveParams& P, float c, float t) {
    const float c8 = c * P.ditherUnit;
    const float a = truncf(c8), b = ceilf(c8);
    const float q = (((c8 - a) >= t) ? b : a) * P.ditherInvUnit;
    return ((c >= P.ditherMin) && (c <= P.ditherMax)) ? lerpf(c, q, P.ditherStrength) : c;
}
ILB_DEV f4 applyDither(const ResolveParams& P, f4 c, unsigned long long pixel) {
    if (P.ditherStrength == 0.0f) return c;  // uniform: the default
    const int y = (int)(pixel / (unsigned long long)P.width), x = (int)(pixel - (unsigned long long)y * (unsigned long long)P.width);
    const float t = ditherThreshold(P, x, y);
    return mk4(dither1(P, c.x, t), dither1(P, c.y, t), dither1(P, c.z, t), c.w);
}

template <int MODE, bool ALBEDO>
ILB_DEV f4 resolvePixel(const ResolveParams& P, f4 light, f4 albedo) {
    f4 result;
    if (ALBEDO) {  // ResolveWithAlbedoCommon, Resolve.fx:47-68
        if (P.albedoIsSRGB) albedo = pSRGBToPLinear(albedo);
        light = light * P.invScale2;
        const f3 a = xyz(albedo);
        result = mk4(lerp3(a, a * xyz(light), saturatef(light.w)), albedo.w);
    } else {  // ResolveCommon, Resolve.fx:30-45
        result = light * P.invScale;
        result.w = 1.0f;
    }
    if (MODE == ILB_HDR_GAMMA_COMPRESS) {  // GammaCompress, HDR.fxh:12-19
        const f3 rgb = max3(xyz(result) + P.offset, mk3(0.0f));
        const float resultLuminance = rgb.x * 0.299f + rgb.y * 0.587f + rgb.z * 0.114f;
        const float scaledLuminance = (resultLuminance * P.middleGray) / P.averageLuminance;
        const float compressedLuminance = (scaledLuminance * (1.0f + (scaledLuminance / P.maxLumSq))) / (1.0f + scaledLuminance);
        const float rescaleFactor = compressedLuminance / resultLuminance;  // 0 / 0 = NaN for black, like the shader
        result = mk4(rgb * rescaleFactor, result.w);
    } else if (MODE == ILB_HDR_TONE_MAP) {  // Resolve.fx:127-133
        const f3 pre = max3(mk3(0.0f), xyz(result) + P.offset) * P.exposure;
        const f3 tm = mk3(uncharted2Tonemap1(pre.x), uncharted2Tonemap1(pre.y), uncharted2Tonemap1(pre.z)) * P.invWhiteScale;
        result = mk4(pow3u(tm, P.gamma), result.w);
    } else {  // Resolve.fx:84-86
        f3 rgb = max3(mk3(0.0f), xyz(result) + P.offset);
        rgb = rgb * P.exposure;
        result = mk4(pow3u(rgb, P.gamma), result.w);
    }
    if (P.resolveToSRGB) result = pLinearToPSRGB(result);
    return result;  // the caller applies ApplyDither (it knows the pixel's coordinates)
}

#ifndef ILB_RESOLVE_MINBLOCKS
#define ILB_RESOLVE_MINBLOCKS 5   // 51 registers: 45.7 -> 43.6 us for the tone-mapped 4K resolve (6: 44.0, 8: 51.9 -- spills)
#endif
template <int MODE, bool ALBEDO, bool VEC>
__global__ void __launch_bounds__(256, ILB_RESOLVE_MINBLOCKS) resolve_kernel(const __grid_constant__ ResolveParams P) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long groups = P.n >> 2;
    if (VEC) for (unsigned long long g = tid; g < groups; g += stride) {
        f4 l[4], a[4], r[4];
        loadTexels4(P.lightmap, P.lm_fmt, g, l);
        if (ALBEDO) loadTexels4(P.albedo, P.al_fmt, g, a);
#pragma unroll
        for (int k = 0; k < 4; k++) r[k] = applyDither(P, resolvePixel<MODE, ALBEDO>(P, l[k], ALBEDO ? a[k] : mk4(0.0f)), 4 * g + k);
        if (P.out_fmt == ILB_FORMAT_RGBA8) {
            __stcs(reinterpret_cast<uint4*>(P.out) + g, make_uint4(packRgba8(r[0]), packRgba8(r[1]), packRgba8(r[2]), packRgba8(r[3])));
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) __stcs(reinterpret_cast<float4*>(P.out) + 4 * g + k, to_float4(r[k]));
        }
    }
    // tail (n % 4 pixels), or every pixel when a pointer is not 16-byte aligned
    for (unsigned long long i = (VEC ? (groups << 2) : 0ull) + tid; i < P.n; i += stride) {
        const f4 l = loadTexel(P.lightmap, P.lm_fmt, i);
        const f4 a = ALBEDO ? loadTexel(P.albedo, P.al_fmt, i) : mk4(0.0f);
        const f4 r = applyDither(P, resolvePixel<MODE, ALBEDO>(P, l, a), i);
        if (P.out_fmt == ILB_FORMAT_RGBA8) reinterpret_cast<uint32_t*>(P.out)[i] = packRgba8(r);
        else reinterpret_cast<float4*>(P.out)[i] = to_float4(r);
    }
}

// ---- scaled / offset resolve (ResolveLighting drawn as a quad, LightingRenderer.cs:1537-1645) ------------------------------
struct PlacedParams {
    ResolveParams R;
    int lw, lh, aw, ah, tw, th;     // lightmap, albedo, target sizes
    int x0, y0, x1, y1;             // target pixels whose centre can lie inside the quad (half-open)
    float px, py, qw, qh;           // quad position and size in target pixels
    float u0, v0, u1, v1;           // albedo region
    float uvox, uvoy;               // LightmapUVOffset
};
// LINEAR / CLAMP fetch, fp32 weights, individually rounded operations (the oracle's sampleLinearClamp)
ILB_DEV f4 sampleLinearClamp(const void* tex, int fmt, int w, int h, float u, float v) {
    const float x = xsub(xmul(u, (float)w), 0.5f), y = xsub(xmul(v, (float)h), 0.5f);
    const float x0f = floorf(x), y0f = floorf(y);
    const float fx = xsub(x, x0f), fy = xsub(y, y0f);
    const int xa = min(max((int)x0f, 0), w - 1), xb = min(max((int)x0f + 1, 0), w - 1);
    const int ya = min(max((int)y0f, 0), h - 1), yb = min(max((int)y0f + 1, 0), h - 1);
    const f4 t00 = loadTexel(tex, fmt, (unsigned long long)ya * w + xa), t10 = loadTexel(tex, fmt, (unsigned long long)ya * w + xb);
    const f4 t01 = loadTexel(tex, fmt, (unsigned long long)yb * w + xa), t11 = loadTexel(tex, fmt, (unsigned long long)yb * w + xb);
    return xlerp4(xlerp4(t00, t10, fx), xlerp4(t01, t11, fx), fy);
}
template <int MODE, bool ALBEDO>
__global__ void __launch_bounds__(256) resolve_placed_kernel(const __grid_constant__ PlacedParams P) {
    const int x = P.x0 + blockIdx.x * 32 + (threadIdx.x & 31), y = P.y0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= P.x1 || y >= P.y1) return;
    const float tx = xdiv(xsub(xadd((float)x, 0.5f), P.px), P.qw), ty = xdiv(xsub(xadd((float)y, 0.5f), P.py), P.qh);
    if (!(tx >= 0.0f && tx < 1.0f && ty >= 0.0f && ty < 1.0f)) return;     // the pixel centre is outside the quad
    const float lu = fminf(fmaxf(xadd(tx, P.uvox), 0.0f), 1.0f), lv = fminf(fmaxf(xadd(ty, P.uvoy), 0.0f), 1.0f);
    const f4 light = sampleLinearClamp(P.R.lightmap, P.R.lm_fmt, P.lw, P.lh, lu, lv);
    f4 albedo = mk4(0.0f);
    if (ALBEDO) {
        const float au = fminf(fmaxf(xadd(P.u0, xmul(tx, xsub(P.u1, P.u0))), P.u0), P.u1);
        const float av = fminf(fmaxf(xadd(P.v0, xmul(ty, xsub(P.v1, P.v0))), P.v0), P.v1);
        albedo = sampleLinearClamp(P.R.albedo, P.R.al_fmt, P.aw, P.ah, au, av);
    }
    const size_t i = (size_t)y * (size_t)P.tw + (size_t)x;
    const f4 r = applyDither(P.R, resolvePixel<MODE, ALBEDO>(P.R, light, albedo), i);  // P.R.width is the target's row length here
    if (P.R.out_fmt == ILB_FORMAT_RGBA8) reinterpret_cast<uint32_t*>(P.R.out)[i] = packRgba8(r);
    else reinterpret_cast<float4*>(P.R.out)[i] = to_float4(r);
}

// ---- LUT-blended resolve (LUTResolve.fx:57-135) ------------------------------------------------------------------------------
struct LutParams {
    ResolveParams R;
    const void* dark;
    const void* bright;
    int dres, bres, drows, brows;      // LUTResolutionsAndRowCounts
    float level0, neutral, level1;     // LUTLevels
    int perChannel, lutOnly;
    float off[4];                      // LUTOffsets
};
// LINEAR / CLAMP fetch of a ColorLUT texel (cached loads: every pixel reads the same few kilobytes)
ILB_DEV f3 sampleLutLinearClamp(const void* tex, int w, int h, float u, float v) {
    const float x = xsub(xmul(u, (float)w), 0.5f), y = xsub(xmul(v, (float)h), 0.5f);
    const float x0f = floorf(x), y0f = floorf(y);
    const float fx = xsub(x, x0f), fy = xsub(y, y0f);
    const int xa = min(max((int)x0f, 0), w - 1), xb = min(max((int)x0f + 1, 0), w - 1);
    const int ya = min(max((int)y0f, 0), h - 1), yb = min(max((int)y0f + 1, 0), h - 1);
    const unsigned int* t = reinterpret_cast<const unsigned int*>(tex);
    const f4 t00 = unpackRgba8(__ldg(t + ya * w + xa)), t10 = unpackRgba8(__ldg(t + ya * w + xb));
    const f4 t01 = unpackRgba8(__ldg(t + yb * w + xa)), t11 = unpackRgba8(__ldg(t + yb * w + xb));
    return xyz(xlerp4(xlerp4(t00, t10, fx), xlerp4(t01, t11, fx), fy));
}
// ReadLUT under the convention stated at ilb_lut_blending (sq/Fracture LUTCommon.fxh is un-vendored): strip layout, row 0
ILB_DEV f3 readLUT(const void* tex, int res, int rows, f3 value, float offU, float offV) {
    const float resm1 = (float)(res - 1);
    const float blue = value.z * resm1;
    const float s0 = floorf(blue), s1 = fminf(s0 + 1.0f, resm1), w = blue - s0;
    const int tw = res * res, th = res * rows;
    const float invW = 1.0f / (float)tw, invH = 1.0f / (float)th;
    const float uIn = 0.5f + value.x * resm1, v = (0.5f + value.y * resm1) * invH + offV;
    const f3 a = sampleLutLinearClamp(tex, tw, th, (s0 * (float)res + uIn) * invW + offU, v);
    const f3 b = sampleLutLinearClamp(tex, tw, th, (s1 * (float)res + uIn) * invW + offU, v);
    return lerp3(a, b, w);
}
ILB_DEV f4 lutResolvePixel(const LutParams& P, f4 light, f4 albedo) {  // LUTBlendedResolveWithAlbedoCommon :57-117 + the pixel shader :119-135
    if (P.R.albedoIsSRGB) albedo = pSRGBToPLinear(albedo);
    light = light * P.R.invScale2;
    f3 weight = xyz(light);
    const float bandWidth = saturatef(P.level1 - P.level0);
    const float neutralBandWidth = fminf(P.neutral, bandWidth - 0.01f);
    const bool hasNeutralBand = neutralBandWidth > 0.0f;
    if (!P.perChannel || hasNeutralBand) {  // RgbToGray, LUTResolve.fx:16 (0.144 sic)
        const float gray = weight.x * 0.299f + weight.y * 0.587f + weight.z * 0.144f;
        weight = mk3(gray);
    }
    const f3 a = mk3(saturatef(albedo.x), saturatef(albedo.y), saturatef(albedo.z));
    const f3 lut1 = readLUT(P.dark, P.dres, P.drows, a, P.off[0], P.off[1]), lut2 = readLUT(P.bright, P.bres, P.brows, a, P.off[2], P.off[3]);
    f3 blended;
    if (hasNeutralBand) {
        const float transitionSize = (bandWidth - neutralBandWidth) * 0.5f;
        const float v = weight.x - P.level0, v2 = v - transitionSize, v3 = v2 - neutralBandWidth;
        const f3 val1 = lerp3(lut1, a, saturatef(v / transitionSize));
        blended = lerp3(val1, lut2, saturatef(v3 / transitionSize));
    } else {
        if (P.level1 > P.level0) {
            weight = max3(weight - P.level0, mk3(0.0f)) / (P.level1 - P.level0);
            weight = mk3(saturatef(weight.x), saturatef(weight.y), saturatef(weight.z));
        } else {
            weight = weight - P.level0;
            weight = mk3(saturatef(weight.x), saturatef(weight.y), saturatef(weight.z));
        }
        blended = lut1 + weight * (lut2 - lut1);
    }
    f3 rgb = P.lutOnly ? blended : blended * xyz(light);
    rgb = max3(mk3(0.0f), rgb + P.R.offset) * P.R.exposure;
    f4 result = mk4(pow3u(rgb, P.R.gamma), albedo.w);
    if (P.R.resolveToSRGB) result = pLinearToPSRGB(result);
    return result;
}
__global__ void __launch_bounds__(256) resolve_lut_kernel(const __grid_constant__ LutParams P) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.R.n; i += stride) {
        const f4 l = loadTexel(P.R.lightmap, P.R.lm_fmt, i), a = loadTexel(P.R.albedo, P.R.al_fmt, i);
        const f4 r = applyDither(P.R, lutResolvePixel(P, l, a), i);
        if (P.R.out_fmt == ILB_FORMAT_RGBA8) reinterpret_cast<uint32_t*>(P.R.out)[i] = packRgba8(r);
        else reinterpret_cast<float4*>(P.R.out)[i] = to_float4(r);
    }
}

// ---- luminance ------------------------------------------------------------------------------------------------------
// CalculateLuminancePixelShader (Resolve.fx:219-234) drawn into the half-size Single target (LightingRenderer.cs:839-898):
// texel (x, y) point-samples lightmap texel (2x+1, 2y+1).  Individually rounded products / sums: bit-identical to the oracle.
__global__ void __launch_bounds__(256) luminance_level0_kernel(const void* lightmap, int fmt, int w, int lw, int lh, float* out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= lw || y >= lh) return;
    const unsigned long long i = (unsigned long long)(2 * y + 1) * (unsigned long long)w + (unsigned long long)(2 * x + 1);
    const f4 t = (fmt == ILB_FORMAT_RGBA8) ? unpackRgba8Exact(__ldcs(reinterpret_cast<const unsigned int*>(lightmap) + i)) : loadTexel(lightmap, fmt, i);
    out[(size_t)y * lw + x] = xadd(xadd(xmul(t.x, 0.299f), xmul(t.y, 0.587f)), xmul(t.z, 0.144f));  // 0.144: Resolve.fx:15 (sic)
}
// one mip step: 2x2 box filter, size floor(size / 2)
__global__ void __launch_bounds__(256) luminance_downsample_kernel(const float* in, int lw, float* out, int nw, int nh) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nw || y >= nh) return;
    const float* p0 = in + (size_t)(2 * y) * lw + 2 * x;
    const float* p1 = p0 + lw;
    out[(size_t)y * nw + x] = xmul(xadd(xadd(p0[0], p0[1]), xadd(p1[0], p1[1])), 0.25f);
}

float hostTonemap1(float value) {  // same expression as the device / oracle function, evaluated in fp32 (-ffp-contract=off)
    const float kA = 0.15f, kB = 0.50f, kC = 0.10f, kD = 0.20f, kE = 0.02f, kF = 0.30f;
    return ((value * (kA * value + kC * kB) + kD * kE) / (value * (kA * value + kB) + kD * kF)) - kE / kF;
}

template <int MODE, bool ALBEDO>
void launchResolve(ilb_ctx* ctx, const ResolveParams& P, bool vec, int grid) {
    if (vec) resolve_kernel<MODE, ALBEDO, true><<<grid, 256, 0, ctx->stream>>>(P);
    else resolve_kernel<MODE, ALBEDO, false><<<grid, 256, 0, ctx->stream>>>(P);
}

}  // namespace

static int fillResolveParams(ilb_ctx* ctx, const ilb_resolve* r, const void* d_lightmap, const void* d_albedo, void* d_output, bool placed,
                             ResolveParams* out) {
    if (r->width <= 0 || r->height <= 0) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad size %dx%d", r->width, r->height);
    if (r->lightmap_format != ILB_FORMAT_FLOAT4 && r->lightmap_format != ILB_FORMAT_HALF4 && r->lightmap_format != ILB_FORMAT_RGBA8)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad lightmap format %d", r->lightmap_format);
    if (d_albedo && r->albedo_format != ILB_FORMAT_FLOAT4 && r->albedo_format != ILB_FORMAT_RGBA8)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "albedo format must be RGBA8 or FLOAT4");
    if (r->output_format != ILB_FORMAT_FLOAT4 && r->output_format != ILB_FORMAT_RGBA8)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "output format must be RGBA8 or FLOAT4");
    if (r->hdr_mode < ILB_HDR_NONE || r->hdr_mode > ILB_HDR_TONE_MAP) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad hdr_mode %d", r->hdr_mode);
    if (!placed && (r->LightmapUVOffset[0] != 0.0f || r->LightmapUVOffset[1] != 0.0f))
        return ilb_fail(ctx, ILB_ERR_UNSUPPORTED, "LightmapUVOffset != 0 needs the scaled / offset resolve (ilb_resolve_lighting_placed)");
    ResolveParams& P = *out;
    memset(&P, 0, sizeof(P));
    P.lightmap = d_lightmap; P.albedo = d_albedo; P.out = d_output;
    P.n = (unsigned long long)r->width * (unsigned long long)r->height;
    P.lm_fmt = r->lightmap_format; P.al_fmt = r->albedo_format; P.out_fmt = r->output_format;
    const float inv = (r->InverseScaleFactor != 0.0f) ? r->InverseScaleFactor : 1.0f;  // LightingRenderer.cs:1469-1473
    P.invScale = inv; P.invScale2 = inv * 2;
    P.offset = r->Offset; P.exposure = r->ExposureMinusOne + 1; P.gamma = r->GammaMinusOne + 1;
    P.middleGray = r->MiddleGray; P.averageLuminance = r->AverageLuminance; P.maxLumSq = r->MaximumLuminanceSquared;
    P.invWhiteScale = 1.0f / hostTonemap1(r->WhitePoint);
    P.albedoIsSRGB = r->AlbedoIsSRGB != 0.0f; P.resolveToSRGB = r->ResolveToSRGB != 0.0f;
    // ApplyDither: the context's settings (ilb_set_dithering), Strength overridden by a non-zero DitheringStrength of this call
    const ilb_dithering& d = ctx->dither;
    P.ditherStrength = (r->DitheringStrength != 0.0f) ? r->DitheringStrength : d.Strength;
    P.ditherUnit = (d.Unit != 0.0f) ? d.Unit : 255.0f;
    P.ditherInvUnit = 1.0f / P.ditherUnit;
    P.ditherBand = (d.BandSize != 0.0f) ? d.BandSize : 1.0f;
    P.ditherMin = d.RangeMin;
    P.ditherMax = (d.RangeMax > d.RangeMin) ? d.RangeMax : 1.0f;
    {
        const float f = std::fmod(d.FrameIndex, 4.0f) + 0.5f, s = 23.0f * f / 17.0f;
        P.ditherPhase = s - std::floor(s);
    }
    P.width = r->width;
    return ILB_OK;
}

int ilb_resolve_placed_launch(ilb_ctx* ctx, const ilb_resolve* r, const ilb_resolve_placement* pl, const void* d_lightmap, const void* d_albedo,
                              void* d_target) {
    if (!pl || pl->target_width <= 0 || pl->target_height <= 0) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad placement");
    if (d_albedo && (pl->albedo_width <= 0 || pl->albedo_height <= 0)) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad albedo size");
    PlacedParams P;
    memset(&P, 0, sizeof(P));
    int rc = fillResolveParams(ctx, r, d_lightmap, d_albedo, d_target, true, &P.R);
    if (rc) return rc;
    P.R.width = pl->target_width;  // the dither pattern follows the target's pixel grid
    P.lw = r->width; P.lh = r->height; P.aw = pl->albedo_width; P.ah = pl->albedo_height; P.tw = pl->target_width; P.th = pl->target_height;
    P.u0 = pl->AlbedoRegion[0]; P.v0 = pl->AlbedoRegion[1]; P.u1 = pl->AlbedoRegion[2]; P.v1 = pl->AlbedoRegion[3];
    P.px = pl->Position[0]; P.py = pl->Position[1];
    P.qw = (d_albedo ? (P.u1 - P.u0) * (float)P.aw : (float)P.lw) * pl->Scale[0];
    P.qh = (d_albedo ? (P.v1 - P.v0) * (float)P.ah : (float)P.lh) * pl->Scale[1];
    P.uvox = r->LightmapUVOffset[0]; P.uvoy = r->LightmapUVOffset[1];
    if (!(P.qw > 0.0f) || !(P.qh > 0.0f)) return ILB_OK;   // an empty quad draws nothing
    // conservative pixel box of the quad (the kernel decides per pixel centre)
    auto lo = [](float v, int n) { return (int)std::fmin(std::fmax(std::floor(v - 1.0f), 0.0f), (float)n); };
    auto hi = [](float v, int n) { return (int)std::fmin(std::fmax(std::ceil(v + 1.0f), 0.0f), (float)n); };
    P.x0 = lo(P.px, P.tw); P.x1 = hi(P.px + P.qw, P.tw); P.y0 = lo(P.py, P.th); P.y1 = hi(P.py + P.qh, P.th);
    if (P.x1 <= P.x0 || P.y1 <= P.y0) return ILB_OK;
    const dim3 grid((P.x1 - P.x0 + 31) / 32, (P.y1 - P.y0 + 7) / 8);
    const bool albedo = d_albedo != nullptr;
#define ILB_PLACED(MODE, ALB) resolve_placed_kernel<MODE, ALB><<<grid, 256, 0, ctx->stream>>>(P)
    switch (r->hdr_mode * 2 + (albedo ? 1 : 0)) {
        case 0: ILB_PLACED(ILB_HDR_NONE, false); break;
        case 1: ILB_PLACED(ILB_HDR_NONE, true); break;
        case 2: ILB_PLACED(ILB_HDR_GAMMA_COMPRESS, false); break;
        case 3: ILB_PLACED(ILB_HDR_GAMMA_COMPRESS, true); break;
        case 4: ILB_PLACED(ILB_HDR_TONE_MAP, false); break;
        default: ILB_PLACED(ILB_HDR_TONE_MAP, true); break;
    }
#undef ILB_PLACED
    ctx->launches++;
    ILB_CUDA(ctx, cudaGetLastError());
    return ILB_OK;
}

int ilb_resolve_launch(ilb_ctx* ctx, const ilb_resolve* r, const void* d_lightmap, const void* d_albedo, void* d_output) {
    ResolveParams P;
    {
        const int rc = fillResolveParams(ctx, r, d_lightmap, d_albedo, d_output, false, &P);
        if (rc) return rc;
    }
    const bool vec = (((uintptr_t)d_lightmap | (uintptr_t)d_albedo | (uintptr_t)d_output) & 15u) == 0;
    // enough 256-thread CTAs for every group of four pixels, capped at 8 waves of 148 SMs x 8 resident CTAs (grid-stride)
    const unsigned long long work = vec ? std::max<unsigned long long>(P.n >> 2, 1) : P.n;
    const int grid = (int)std::min<unsigned long long>((work + 255) / 256, 148ull * 8 * 8);
    const bool albedo = d_albedo != nullptr;
    switch (r->hdr_mode * 2 + (albedo ? 1 : 0)) {
        case 0: launchResolve<ILB_HDR_NONE, false>(ctx, P, vec, grid); break;
        case 1: launchResolve<ILB_HDR_NONE, true>(ctx, P, vec, grid); break;
        case 2: launchResolve<ILB_HDR_GAMMA_COMPRESS, false>(ctx, P, vec, grid); break;
        case 3: launchResolve<ILB_HDR_GAMMA_COMPRESS, true>(ctx, P, vec, grid); break;
        case 4: launchResolve<ILB_HDR_TONE_MAP, false>(ctx, P, vec, grid); break;
        default: launchResolve<ILB_HDR_TONE_MAP, true>(ctx, P, vec, grid); break;
    }
    ctx->launches++;
    ILB_CUDA(ctx, cudaGetLastError());
    return ILB_OK;
}

int ilb_resolve_lut_launch(ilb_ctx* ctx, const ilb_resolve* r, const ilb_lut_blending* lut, const void* d_dark, const void* d_bright,
                           const void* d_lightmap, const void* d_albedo, void* d_output) {
    if (r->hdr_mode != ILB_HDR_NONE) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "LUT blending is not compatible with this type of lighting resolve");
    if (lut->dark_resolution < 2 || lut->bright_resolution < 2 || lut->dark_row_count < 1 || lut->bright_row_count < 1 ||
        lut->dark_resolution > 256 || lut->bright_resolution > 256)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad LUT geometry");
    LutParams P;
    memset(&P, 0, sizeof(P));
    const int rc = fillResolveParams(ctx, r, d_lightmap, d_albedo, d_output, false, &P.R);
    if (rc) return rc;
    P.dark = d_dark; P.bright = d_bright;
    P.dres = lut->dark_resolution; P.bres = lut->bright_resolution; P.drows = lut->dark_row_count; P.brows = lut->bright_row_count;
    P.level0 = lut->DarkLevel; P.neutral = lut->NeutralBandSize; P.level1 = lut->BrightLevel;
    P.perChannel = lut->PerChannel != 0.0f; P.lutOnly = lut->LUTOnly != 0.0f;
    for (int k = 0; k < 4; k++) P.off[k] = lut->LUTOffsets[k];
    const int grid = (int)std::min<unsigned long long>((P.R.n + 255) / 256, 148ull * 8 * 8);
    resolve_lut_kernel<<<grid, 256, 0, ctx->stream>>>(P);
    ctx->launches++;
    ILB_CUDA(ctx, cudaGetLastError());
    return ILB_OK;
}

int ilb_luminance_launch(ilb_ctx* ctx, int width, int height, int fmt, const void* d_lightmap, int level, float* out_host) {
    int lw = width / 2, lh = height / 2;
    if (level < 0 || lw <= 0 || lh <= 0 || (lw >> level) <= 0 || (lh >> level) <= 0)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "luminance level %d does not exist for a %dx%d lightmap", level, width, height);
    const size_t bytes0 = sizeof(float) * (size_t)lw * (size_t)lh;
    int rc = ilb_reserve(ctx, &ctx->d_luminance[0], &ctx->d_luminance_capacity[0], bytes0, false);
    if (rc) return rc;
    if (level > 0) {
        rc = ilb_reserve(ctx, &ctx->d_luminance[1], &ctx->d_luminance_capacity[1], std::max<size_t>(bytes0 / 4, 16), false);
        if (rc) return rc;
    }
    float* cur = reinterpret_cast<float*>(ctx->d_luminance[0]);
    float* nxt = reinterpret_cast<float*>(ctx->d_luminance[1]);
    luminance_level0_kernel<<<dim3((lw + 255) / 256, lh), 256, 0, ctx->stream>>>(d_lightmap, fmt, width, lw, lh, cur);
    ctx->launches++;
    ILB_CUDA(ctx, cudaGetLastError());
    for (int k = 0; k < level; k++) {
        const int nw = lw / 2, nh = lh / 2;
        luminance_downsample_kernel<<<dim3((nw + 255) / 256, nh), 256, 0, ctx->stream>>>(cur, lw, nxt, nw, nh);
        ctx->launches++;
        ILB_CUDA(ctx, cudaGetLastError());
        std::swap(cur, nxt);
        lw = nw; lh = nh;
    }
    ILB_CUDA(ctx, cudaMemcpyAsync(out_host, cur, sizeof(float) * (size_t)lw * (size_t)lh, cudaMemcpyDeviceToHost, ctx->stream));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ILB_OK;
}
