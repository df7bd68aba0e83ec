// TEST INFRASTRUCTURE -- CPU oracle for the LIGHTING hot path (L1-L11 of SURVEY.md section 8).
//
// A line-by-line fp32 restatement of the reference HLSL pixel shaders, in the reference's own
// multi-pass structure (one full pass per light, additive accumulate).  Only tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs may call this; the product never does.
// PARITY UNPINNED: the reference ships no tests or golden images for this path, and
// its runtime (HLSL/D3D + un-vendored sq/Fracture) cannot run here, so this oracle is pinned only by
// closed-form known answers (tests/test_oracle_kat.py), not by reference outputs.
//
// Citations are relative to /root/reference/Illuminant/.
#include <omp.h>

#include <cstring>
#include <vector>

#include "../include/illuminant_b200.h"
#include "hlsl.hpp"
#include "oracle.h"

using namespace hlsl;

namespace {

inline float4 f4(const ilb_float4& v) { return float4(v.x, v.y, v.z, v.w); }

// ---------------------------------------------------------------- distance field (L1, L2)
// Shaders/DistanceFieldCommon.fxh:189-364
struct DistanceField {
    const uint16_t* tex;
    int tw, th;
    float4 ConeAndMisc, TextureSliceAndTexelSize, StepAndMisc2, TextureSliceCount, Extent, Packed1;
    DistanceField(const uint16_t* t, int w, int h, const ilb_df_uniforms& u)
        : tex(t), tw(w), th(h), ConeAndMisc(f4(u.ConeAndMisc)),
          TextureSliceAndTexelSize(f4(u.TextureSliceAndTexelSize)), StepAndMisc2(f4(u.StepAndMisc2)),
          TextureSliceCount(f4(u.TextureSliceCount)), Extent(f4(u.Extent)), Packed1(f4(u.Packed1)) {}

    float getDistanceFieldZOffset() const { return ConeAndMisc.y; }       // :208-210
    float getMaximumEncodedDistance() const { return Extent.w; }          // :212-214
    float getStepLimit() const { return StepAndMisc2.x; }                 // :217-219
    float getMinStepSize() const { return Packed1.w; }                    // :221-224
    float getLongStepFactor() const { return StepAndMisc2.z; }            // :226-228
    float getMaxConeRadius() const { return ConeAndMisc.x; }              // :230-232
    float getConeGrowthFactor() const { return 1.0f; }                    // :234-237 (dead uniform)
    float getOcclusionToOpacityPower() const { return ConeAndMisc.z; }    // :239-241
    float getInvScaleFactorX() const { return ConeAndMisc.w; }
    float getInvScaleFactorY() const { return StepAndMisc2.w; }
    float2 getDistanceSliceSize() const { return TextureSliceAndTexelSize.xy(); }                        // :255-257
    float2 getDistanceTexelSize() const { return float2(TextureSliceAndTexelSize.z, TextureSliceAndTexelSize.w); }  // :259-261
    float getMaximumValidZ() const { return Packed1.z; }
    float getInvSliceCountXTimesOneThird() const { return Packed1.x; }
    float getZToSliceIndex() const { return Packed1.y; }

    static constexpr float DISTANCE_ZERO = 192.0f / 255.0f;               // :8
    float decodeDistance(float e) const { return (DISTANCE_ZERO - e) * getMaximumEncodedDistance(); }  // :268-270

    // UNORM16 texel fetch: value = c * (1/65535) (shared convention with the CUDA path)
    float4 texel(int ix, int iy) const {
        const uint16_t* p = tex + 4 * ((size_t)iy * (size_t)tw + (size_t)ix);
        const float k = 1.0f / 65535.0f;
        return float4((float)p[0] * k, (float)p[1] * k, (float)p[2] * k, (float)p[3] * k);
    }
    // sampler :273-281 -- MinMag LINEAR, AddressU WRAP, AddressV CLAMP, exact fp32 weights
    // (texel centres at integer+0.5, D3D10+/FNA addressing).
    float4 tex2Dlod(float2 uv) const {
        float x = uv.x * (float)tw - 0.5f, y = uv.y * (float)th - 0.5f;
        float x0f = floorf(x), y0f = floorf(y);
        float fx = x - x0f, fy = y - y0f;
        int x0 = (int)x0f, y0 = (int)y0f;
        int x1 = x0 + 1, y1 = y0 + 1;
        x0 %= tw; if (x0 < 0) x0 += tw;
        x1 %= tw; if (x1 < 0) x1 += tw;
        if (y0 < 0) y0 = 0; if (y0 > th - 1) y0 = th - 1;
        if (y1 < 0) y1 = 0; if (y1 > th - 1) y1 = th - 1;
        float4 top = lerp(texel(x0, y0), texel(x1, y0), fx);
        float4 bot = lerp(texel(x0, y1), texel(x1, y1), fx);
        return lerp(top, bot, fy);
    }

    float2 computeDistanceFieldSliceUv(float virtualSliceIndex) const {  // :303-311
        float columnIndex = floorf(virtualSliceIndex / 3);
        float rowIndexF = virtualSliceIndex * getInvSliceCountXTimesOneThird();
        float rowIndex = floorf(rowIndexF);
        return float2(columnIndex, rowIndex) * getDistanceSliceSize();
    }

    float sampleDistanceFieldEx(float3 position) const {  // :313-353
        position.z -= getDistanceFieldZOffset();
        float3 extent = Extent.xyz();
        float3 clampedPosition = clamp(position, float3(0.0f), extent);
        float3 distanceToVolume3 = -min(position, float3(0.0f)) + (max(position, extent) - extent);
        float distanceToVolume = length(distanceToVolume3);

        float slicePosition = min(clampedPosition.z, getMaximumValidZ()) * getZToSliceIndex();
        float virtualSliceIndex = floorf(slicePosition);

        float2 texelUv = clampedPosition.xy() * getDistanceTexelSize();
        float2 uv = computeDistanceFieldSliceUv(virtualSliceIndex) + texelUv;
        float4 packedSample = tex2Dlod(uv);

        float maskPatternIndex = fmod(virtualSliceIndex, 3);
        float subslice = slicePosition - virtualSliceIndex, blendedSample;
        if (maskPatternIndex >= 2)
            blendedSample = lerp(packedSample.z, packedSample.w, subslice);
        else if (maskPatternIndex >= 1)
            blendedSample = lerp(packedSample.y, packedSample.z, subslice);
        else
            blendedSample = lerp(packedSample.x, packedSample.y, subslice);

        float decodedDistance = decodeDistance(blendedSample);
        return decodedDistance + distanceToVolume;
    }
};

// ---------------------------------------------------------------- cone trace (L3)
// Shaders/ConeTrace.fxh
const float MIN_CONE_RADIUS = 0.33f;
const float MAX_STEP_RAMP_WINDOW = 2;
const float TRACE_INITIAL_OFFSET_PX = 0.5f;
const float FULLY_SHADOWED_THRESHOLD = 0.075f;
const float UNSHADOWED_THRESHOLD = 0.95f;
const float HACK_DISTANCE_OFFSET = 1.5f;
const float TRACE_END_MULTIPLIER = 100;

struct TraceState {
    float3 origin, direction;
    float3 data;  // position, length, visibility
};

void coneTraceInitialize(TraceState& state, float3 startPosition, float3 endPosition, float startOffset,
                         float lightRadius, bool startAtEnd) {  // :37-49
    float3 traceVector = (endPosition - startPosition);
    float traceLength = length(traceVector);
    state.origin = startPosition;
    state.direction = traceVector / traceLength;
    state.data.y = max(traceLength - lightRadius, 1);
    state.data.x = startAtEnd ? traceLength - startOffset : startOffset;
    state.data.z = 1.0f;
}

float coneTraceStep(const DistanceField& df, float4 config, float distanceToObstacle, float offset,
                    float& visibility) {  // :51-71
    float localSphereRadius = min((config.y * offset) + MIN_CONE_RADIUS, config.x);
    float localVisibility = ((distanceToObstacle + HACK_DISTANCE_OFFSET) / localSphereRadius);
    visibility = min(visibility, localVisibility);
    return max(abs(distanceToObstacle) * df.getLongStepFactor(), config.z);
}

float coneTraceAdvance(const DistanceField& df, TraceState& state, float4 config) {  // :73-82
    float sample = df.sampleDistanceFieldEx(state.origin + (state.direction * state.data.x));
    state.data.x += coneTraceStep(df, config, sample, state.data.x, state.data.z);
    return saturate(state.data.z - FULLY_SHADOWED_THRESHOLD) * saturate(state.data.y - state.data.x);
}

float coneTraceAdvanceEx(const DistanceField& df, TraceState& state, float4 config) {  // :84-96
    float sample = df.sampleDistanceFieldEx(state.origin + (state.direction * state.data.x));
    state.data.x = min(state.data.x + coneTraceStep(df, config, sample, state.data.x, state.data.z), state.data.y);
    return saturate(state.data.z - FULLY_SHADOWED_THRESHOLD) *
           saturate((state.data.y - state.data.x) * TRACE_END_MULTIPLIER);
}

float4 createTraceConfig(const DistanceField& df, float2 lightRamp, float2 coneGrowthFactorAndDistanceFalloff) {  // :122-139
    float maxRadius = clamp(lightRamp.x, MIN_CONE_RADIUS, df.getMaxConeRadius());
    float rampLength = max(lightRamp.y, 16);
    float radiusGrowthPerPixel = maxRadius / rampLength * coneGrowthFactorAndDistanceFalloff.x;
    return float4(maxRadius, radiusGrowthPerPixel, max(1, df.getMinStepSize()), coneGrowthFactorAndDistanceFalloff.y);
}

float traceFinalResult(const DistanceField& df, float visibility) {  // :182-188, LineLightCore.fxh:59-65
    return powf(saturate(saturate((visibility - FULLY_SHADOWED_THRESHOLD)) /
                         (UNSHADOWED_THRESHOLD - FULLY_SHADOWED_THRESHOLD)),
                df.getOcclusionToOpacityPower());
}

float coneTrace(const DistanceField& df, float3 lightCenter, float2 lightRamp,
                float2 coneGrowthFactorAndDistanceFalloff, float3 shadedPixelPosition, bool enable,
                int* stepsTaken = nullptr) {  // :141-191
    TraceState traceA;
    coneTraceInitialize(traceA, shadedPixelPosition, lightCenter, TRACE_INITIAL_OFFSET_PX, lightRamp.x, false);
    float4 config = createTraceConfig(df, lightRamp, coneGrowthFactorAndDistanceFalloff);

    float stepsRemaining = df.getStepLimit();
    float liveness = ((df.Extent.x > 0) && enable) ? 1.0f : 0.0f;
    float stepLiveness;
    int n = 0;
    while (liveness > 0) {
        stepsRemaining--;
        stepLiveness = coneTraceAdvance(df, traceA, config);
        liveness = stepsRemaining * stepLiveness;
        n++;
    }
    if (stepsTaken) *stepsTaken = n;

    if (stepsRemaining == 0) traceA.data.x = traceA.data.y;

    float stepWindowVisibility = stepsRemaining / MAX_STEP_RAMP_WINDOW;
    float visibility = min(traceA.data.z, stepWindowVisibility);
    float finalResult = traceFinalResult(df, visibility);
    return enable ? finalResult : 1.0f;
}

// ---------------------------------------------------------------- environment / G-buffer (L4)
struct Frame {
    float4 EnvironmentZAndScale, EnvironmentZToY, GBufferTexelSizeAndMisc;
    float GBufferViewportRelative;
    float2 ViewportPosition;
    const void* gbuffer;
    int gw, gh, gfmt;
    bool stencilCulling;

    float getGroundZ() const { return EnvironmentZAndScale.x; }   // EnvironmentCommon.fxh:9-31
    float getMaximumZ() const { return EnvironmentZAndScale.y; }
    float getZToYMultiplier() const { return EnvironmentZToY.x; }
    float getInvZToYMultiplier() const { return EnvironmentZToY.y; }
    float getLightOcclusion() const { return EnvironmentZToY.z; }
    float2 getEnvironmentRenderScale() const { return float2(EnvironmentZAndScale.z, EnvironmentZAndScale.w); }
    float2 GetViewportScale() const { return float2(GBufferTexelSizeAndMisc.z, GBufferTexelSizeAndMisc.w); }  // LightCommon.fxh:30-34
    float2 GetViewportPosition() const { return ViewportPosition; }
};

float halfToFloat(uint16_t h) {
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
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

// sampler LightCommon.fxh:36-44: POINT, CLAMP
float4 gbufferPoint(const Frame& fr, float2 uv) {
    int ix = (int)floorf(uv.x * (float)fr.gw), iy = (int)floorf(uv.y * (float)fr.gh);
    if (ix < 0) ix = 0; if (ix > fr.gw - 1) ix = fr.gw - 1;
    if (iy < 0) iy = 0; if (iy > fr.gh - 1) iy = fr.gh - 1;
    size_t i = (size_t)iy * fr.gw + ix;
    if (fr.gfmt == ILB_FORMAT_FLOAT4) {
        const float* p = (const float*)fr.gbuffer + 4 * i;
        return float4(p[0], p[1], p[2], p[3]);
    }
    const uint16_t* p = (const uint16_t*)fr.gbuffer + 4 * i;
    return float4(halfToFloat(p[0]), halfToFloat(p[1]), halfToFloat(p[2]), halfToFloat(p[3]));
}

float3 decodeNormalSpherical(float2 enc) {  // EnvironmentCommon.fxh:42-52
    if (any(enc)) {
        float2 ang = enc * 2 - 1;
        float2 scth(dm_sinf(ang.x * PI), dm_cosf(ang.x * PI));  // sincos()
        float2 scphi = float2(sqrtf(1.0f - ang.y * ang.y), ang.y);
        return float3(scth.y * scphi.x, scth.x * scphi.x, scphi.y);
    }
    return float3(0.0f);
}

const float GBUFFER_Z_SCALE = 1024, GBUFFER_Z_OFFSET = 1024;

// LightCommon.fxh:58-144
float3 sampleGBuffer(const Frame& fr, float2 screenPositionPx, float3& worldPosition, float3& normal,
                     bool& enableShadows, bool& fullbright, float4* rawSample = nullptr) {
    enableShadows = true;
    fullbright = false;
    float3 cameraPosition;
    if (any(fr.GBufferTexelSizeAndMisc.xy())) {
        float2 sourceXy = screenPositionPx;
        if (fr.GBufferViewportRelative != 0) {
            sourceXy = sourceXy / fr.GetViewportScale();
            sourceXy += fr.GetViewportPosition();
        }
        float2 uv = (sourceXy + 0.5f) * fr.GBufferTexelSizeAndMisc.xy();
        float4 sample = gbufferPoint(fr, uv);
        if (rawSample) *rawSample = sample;

        float relativeY = sample.z;
        float worldZ = sample.w;
        if (worldZ < 0) {
            worldZ += 1;
            worldZ = -worldZ;
            enableShadows = false;
        } else if (worldZ >= 9999) {
            worldZ = 0;
            enableShadows = false;
            fullbright = true;
        }
        worldZ *= GBUFFER_Z_SCALE;
        worldZ -= GBUFFER_Z_OFFSET;

        screenPositionPx = screenPositionPx / fr.getEnvironmentRenderScale();
        cameraPosition = float3(screenPositionPx, fr.getMaximumZ() + 0.01f);
        worldPosition = float3((screenPositionPx + float2(0, relativeY)) / fr.GetViewportScale() + fr.GetViewportPosition(),
                               worldZ);
        if (any(sample.xy()))
            normal = decodeNormalSpherical(sample.xy());
        else
            normal = float3(0, 0, 0);
    } else {
        if (rawSample) *rawSample = float4(0, 0, 0, 0);
        screenPositionPx = screenPositionPx / fr.getEnvironmentRenderScale();
        cameraPosition = float3(screenPositionPx, fr.getMaximumZ() + 0.01f);
        worldPosition = float3(screenPositionPx / fr.GetViewportScale() + fr.GetViewportPosition(), fr.getGroundZ());
        normal = float3(0, 0, 1);
    }
    return cameraPosition;
}

bool checkShadowFilter(float4 evenMoreLightProperties, bool enableShadows) {  // LightCommon.fxh:146-152
    float filter = evenMoreLightProperties.x;
    if (filter < 0) return false;
    return (filter > 0.5f) != enableShadows;
}

// ---------------------------------------------------------------- light response (L5, L6)
const float DOT_OFFSET = 0.15f, DOT_RAMP_RANGE = 0.15f;
const float DIRECTIONAL_DOT_OFFSET = 0.35f, DIRECTIONAL_DOT_RAMP_RANGE = 0.35f;
const float DOT_EXPONENT = 0.85f;

float computeNormalFactorEx(float3 lightNormal, float3 shadedPixelNormal, float offset, float range) {  // :154-165
    if (!any(shadedPixelNormal)) return 1;
    float d = dot(-lightNormal, shadedPixelNormal);
    return powf(saturate((d + offset) / range), DOT_EXPONENT);
}

float computeSphereLightOpacity(const Frame& fr, float3 shadedPixelPosition, float3 shadedPixelNormal,
                                float3 lightCenter, float4 lightProperties, float yDistanceFactor) {  // :173-210
    float lightRadius = lightProperties.x;
    float lightRampLength = lightProperties.y;
    float falloffMode = lightProperties.z;

    float3 distance3 = shadedPixelPosition - lightCenter;
    distance3.y *= yDistanceFactor;
    float distance = length(distance3);
    float distanceFactor = 1 - saturate((distance - lightRadius) / lightRampLength);

    if (fr.getLightOcclusion() > 0)
        distanceFactor *= 1 - saturate(distance3.z / fr.getLightOcclusion());

    float3 lightNormal = distance3 / distance;
    float normalFactor = computeNormalFactorEx(lightNormal, shadedPixelNormal, DOT_OFFSET, DOT_RAMP_RANGE);

    if (falloffMode >= 2) {
        distanceFactor = 1 - saturate(distance - lightRadius);
        normalFactor = 1;
    } else if (falloffMode >= 1) {
        distanceFactor *= distanceFactor;
    }
    return saturate((normalFactor * distanceFactor) + saturate(lightRadius - distance));
}

float CalcSphereLightSpecularity(float3 cameraPosition, float3 shadedPixelPosition, float3 shadedPixelNormal,
                                 float3 lightCenter, float power) {  // :212-222
    float3 lightDirection = shadedPixelPosition - lightCenter;
    float3 h = normalize(normalize(cameraPosition - shadedPixelPosition) - lightDirection);
    return powf(saturate(dot(h, shadedPixelNormal)), power);
}

float computeDirectionalLightOpacity(float4 lightDirection, float3 shadedPixelNormal) {  // :224-231
    if (lightDirection.w < 0.1f) return 1;
    return computeNormalFactorEx(lightDirection.xyz(), shadedPixelNormal, DIRECTIONAL_DOT_OFFSET,
                                 DIRECTIONAL_DOT_RAMP_RANGE);
}

float computeAO(const DistanceField& df, float3 shadedPixelPosition, float3 shadedPixelNormal,
                float4 moreLightProperties, bool visible) {  // AOCommon.fxh:1-20
    float aoRadius = moreLightProperties.x, aoOpacity = moreLightProperties.w;
    if ((aoRadius >= 0.5f) && (df.Extent.x > 0) && visible) {
        float distance = df.sampleDistanceFieldEx(shadedPixelPosition +
                                                  float3(0, 0, shadedPixelNormal.z * moreLightProperties.x));
        float clampedDistance = clamp(distance, 0, aoRadius);
        float result = 1 - saturate(clampedDistance / moreLightProperties.x);
        result *= result;
        result = 1 - result;
        return (1 - aoOpacity) + (result * aoOpacity);
    }
    return 1;
}

// ---------------------------------------------------------------- sphere light (L7)
// returns false when the fragment is discarded
bool SphereLightPixelCore(const Frame& fr, const DistanceField& df, float3 shadedPixelPosition,
                          float3 shadedPixelNormal, float3 lightCenter, float4 lightProperties,
                          float4 moreLightProperties, float& opacity, float* preTrace = nullptr, float* cone = nullptr) {  // SphereLightCore.fxh:58-158
    const float SELF_OCCLUSION_HACK = 1.6f;
    const float SHADOW_OPACITY_THRESHOLD = (0.75f / 255.0f);
    // prologue :58-80
    float distanceOpacity = computeSphereLightOpacity(fr, shadedPixelPosition, shadedPixelNormal, lightCenter,
                                                      lightProperties, moreLightProperties.z);
    bool visible = (distanceOpacity > 0) && (shadedPixelPosition.x > -9999);
    moreLightProperties.x *= max(0, shadedPixelNormal.z);
    if (!visible) return false;

    float aoOpacity = computeAO(df, shadedPixelPosition, shadedPixelNormal, moreLightProperties, visible);
    float preTraceOpacity = distanceOpacity * aoOpacity;

    bool traceShadows = visible && (lightProperties.w != 0) && (preTraceOpacity >= SHADOW_OPACITY_THRESHOLD);
    float coneOpacity = coneTrace(df, lightCenter, float2(lightProperties.x, lightProperties.y),
                                  float2(df.getConeGrowthFactor(), moreLightProperties.y),
                                  shadedPixelPosition + (SELF_OCCLUSION_HACK * shadedPixelNormal), traceShadows);
    opacity = preTraceOpacity * coneOpacity;  // epilogue :82-97
    if (preTrace) *preTrace = preTraceOpacity;  // SphereLightPixelCoreWithRamp (:160-199) hands both to its epilogue
    if (cone) *cone = coneOpacity;
    return true;
}

// ---- ramp textures (RampCommon.fxh): RampTextureSampler = LINEAR min / mag, U CLAMP, V WRAP; fp32 bilinear weights
struct RampTexture { const float* texels; int w, h; };
std::vector<RampTexture> g_ramps;  // id - 1 -> texture (orc_set_ramp_textures)
float4 SampleFromRamp2(const RampTexture& t, float2 xy) {  // RampCommon.fxh:19-21
    const float x = xy.x * (float)t.w - 0.5f, y = xy.y * (float)t.h - 0.5f;
    const float x0f = floorf(x), y0f = floorf(y);
    const float fx = x - x0f, fy = y - y0f;
    auto at = [&](int ix, int iy) {
        ix = std::min(std::max(ix, 0), t.w - 1);
        iy = ((iy % t.h) + t.h) % t.h;
        const float* q = t.texels + 4 * ((size_t)iy * t.w + ix);
        return float4(q[0], q[1], q[2], q[3]);
    };
    const int x0 = (int)x0f, y0 = (int)y0f;
    const float4 top = lerp(at(x0, y0), at(x0 + 1, y0), fx), bottom = lerp(at(x0, y0 + 1), at(x0 + 1, y0 + 1), fx);
    return lerp(top, bottom, fy);
}
// SphereLightPixelEpilogueWithRamp (SphereLightCore.fxh:99-119)
float3 SphereLightPixelEpilogueWithRamp(const RampTexture& ramp, float preTraceOpacity, float coneOpacity, float3 distanceToCenter,
                                        float4 evenMoreLightProperties) {
    float angle = atan2f(distanceToCenter.y, distanceToCenter.x);
    return SampleFromRamp2(ramp, float2(preTraceOpacity, (angle + evenMoreLightProperties.z) * evenMoreLightProperties.w)).xyz() * coneOpacity;
}

// Rasterised coverage of the sphere-light geometry: SphereLightVertexShader (SphereLightCore.fxh:13-56)
// over the 12-vertex cross of FillSphereBuffer (LightingRenderer.cs:636-656): three axis-aligned quads
// [1/7,6/7]x[0,1], [6/7,1]x[1/7,6/7], [0,1/7]x[1/7,6/7] of lerp(tl, br, w); vertices with w.y < 0.5 are
// shifted up by radiusOffset + zOffset (2.5D).  A pixel is covered when its centre is inside.
bool sphereLightCovers(const Frame& fr, const ilb_light_vertex& v, float px, float py) {
    float3 lightCenter(v.LightPosition1.x, v.LightPosition1.y, v.LightPosition1.z);
    float radius = v.LightProperties.x + v.LightProperties.y + 1;
    float deltaY = (radius) - (radius / v.MoreLightProperties.z);
    float3 radius3 = float3(radius, radius - (deltaY / 2.0f), 0);
    float3 tl = lightCenter - radius3, br = lightCenter + radius3;
    float radiusOffset = radius * fr.getInvZToYMultiplier();
    float zOffset = lightCenter.z * fr.getZToYMultiplier();
    // pixel centre -> world (inverse of :52-54 with an orthographic view transform)
    float2 s = fr.GetViewportScale() * fr.getEnvironmentRenderScale();
    float wx = (px + 0.5f) / s.x + fr.GetViewportPosition().x;
    float wy = (py + 0.5f) / s.y + fr.GetViewportPosition().y;
    const float cOne = 1.0f / 7.0f, mOne = 6.0f / 7.0f;
    auto X = [&](float w) { return lerp(tl.x, br.x, w); };
    auto Y = [&](float w) {
        float y = lerp(tl.y, br.y, w);
        if (w < 0.5f) { y -= radiusOffset; y -= zOffset; }
        return y;
    };
    auto inside = [&](float x0, float x1, float y0, float y1) {
        return (wx >= X(x0)) && (wx <= X(x1)) && (wy >= Y(y0)) && (wy <= Y(y1));
    };
    return inside(cOne, mOne, 0, 1) || inside(mOne, 1, cOne, mOne) || inside(0, cOne, cOne, mOne);
}

bool SphereLightPixelShader(const Frame& fr, const DistanceField& df, const ilb_light_vertex& v, float2 vpos,
                            float4& result, int rampTexture = 0) {  // SphereLight.fx:7-46; with a ramp texture :48-87
    float3 lightCenter(v.LightPosition1.x, v.LightPosition1.y, v.LightPosition1.z);
    float4 lightProperties = f4(v.LightProperties), moreLightProperties = f4(v.MoreLightProperties);
    float4 color = f4(v.Color1), specular = f4(v.Color2), evenMoreLightProperties = f4(v.EvenMoreLightProperties);

    float3 shadedPixelPosition, shadedPixelNormal;
    bool enableShadows, fullbright;
    float3 cameraPosition = sampleGBuffer(fr, vpos, shadedPixelPosition, shadedPixelNormal, enableShadows, fullbright);
    if (fullbright || checkShadowFilter(evenMoreLightProperties, enableShadows)) return false;
    lightProperties.w *= enableShadows ? 1.0f : 0.0f;

    float opacity, preTraceOpacity, coneOpacity;
    if (!SphereLightPixelCore(fr, df, shadedPixelPosition, shadedPixelNormal, lightCenter, lightProperties,
                              moreLightProperties, opacity, &preTraceOpacity, &coneOpacity))
        return false;
    float specularity = CalcSphereLightSpecularity(cameraPosition, shadedPixelPosition, shadedPixelNormal,
                                                   lightCenter, specular.w);
    if (rampTexture > 0 && rampTexture <= (int)g_ramps.size()) {  // SphereLightWithDistanceRampPixelShader SphereLight.fx:48-87
        float3 opacity3 = SphereLightPixelEpilogueWithRamp(g_ramps[rampTexture - 1], preTraceOpacity, coneOpacity,
                                                           shadedPixelPosition - lightCenter, evenMoreLightProperties);
        result = float4((color.xyz() * color.w * opacity3) + (specular.xyz() * specularity * opacity3), 1);
        return true;
    }
    float3 rgb = (color.xyz() * color.w * opacity) + (specular.xyz() * specularity * opacity);
    result = float4(rgb, 1);
    return true;
}

// ---------------------------------------------------------------- particle light (N4)
// ParticleLightVertexShader (ParticleLight.fx:53-68): one axis-aligned quad lerp(tl, br, corner), tl = center - radius,
// br = center + radius, tl.y -= radius * invZToY + center.z * zToY.  The LightVertex carries what the vertex shader hands
// to the pixel shader: LightPosition1 = particle position, LightProperties / MoreLightProperties = the template's,
// Color1 = unpremultiplied attribute colour * LightColor, Color2 = LightSpecularColor.
bool particleLightCovers(const Frame& fr, const ilb_light_vertex& v, float px, float py) {
    float3 lightCenter(v.LightPosition1.x, v.LightPosition1.y, v.LightPosition1.z);
    float radius = v.LightProperties.x + v.LightProperties.y + 1;
    float3 radius3 = float3(radius, radius, 0);
    float3 tl = lightCenter - radius3, br = lightCenter + radius3;
    tl.y -= radius * fr.getInvZToYMultiplier();
    tl.y -= lightCenter.z * fr.getZToYMultiplier();
    float2 s = fr.GetViewportScale() * fr.getEnvironmentRenderScale();
    float wx = (px + 0.5f) / s.x + fr.GetViewportPosition().x;
    float wy = (py + 0.5f) / s.y + fr.GetViewportPosition().y;
    return (wx >= tl.x) && (wx <= br.x) && (wy >= tl.y) && (wy <= br.y);
}

bool ParticleLightPixelShader(const Frame& fr, const DistanceField& df, const ilb_light_vertex& v, float2 vpos,
                              float4& result) {  // ParticleLight.fx:84-118
    float3 lightCenter(v.LightPosition1.x, v.LightPosition1.y, v.LightPosition1.z);
    float4 lightProperties = f4(v.LightProperties), moreLightProperties = f4(v.MoreLightProperties);
    float4 lightColor = f4(v.Color1), specular = f4(v.Color2);
    float3 shadedPixelPosition, shadedPixelNormal;
    bool enableShadows, fullbright;
    float3 cameraPosition = sampleGBuffer(fr, vpos, shadedPixelPosition, shadedPixelNormal, enableShadows, fullbright);
    if (fullbright) return false;
    lightProperties.w *= enableShadows ? 1.0f : 0.0f;
    float opacity;
    if (!SphereLightPixelCore(fr, df, shadedPixelPosition, shadedPixelNormal, lightCenter, lightProperties,
                              moreLightProperties, opacity))
        return false;
    float specularity = CalcSphereLightSpecularity(cameraPosition, shadedPixelPosition, shadedPixelNormal, lightCenter, specular.w);
    float3 rgb = (lightColor.xyz() * lightColor.w * opacity) + (specular.xyz() * specularity * opacity);
    result = float4(rgb, 1);
    return true;
}

// ---------------------------------------------------------------- directional light (L8)
bool DirectionalLightPixelCore(const DistanceField& df, float3 shadedPixelPosition, float3 shadedPixelNormal,
                               float4 lightDirection, float4 lightProperties, float4 moreLightProperties,
                               float& opacity) {  // DirectionalLight.fx:52-93 (useOpacityRamp = false)
    const float SELF_OCCLUSION_HACK = 1.5f;
    float lightOpacity = computeDirectionalLightOpacity(lightDirection, shadedPixelNormal);
    bool visible = (shadedPixelPosition.x > -9999);
    moreLightProperties.x *= max(0, shadedPixelNormal.z);
    float aoOpacity = computeAO(df, shadedPixelPosition, shadedPixelNormal, moreLightProperties, visible);
    lightOpacity *= aoOpacity;

    bool traceShadows = visible && (lightProperties.x != 0) && (lightOpacity >= 1 / 256.0f) && (lightDirection.w >= 0.1f);
    float3 fakeLightCenter = shadedPixelPosition - (lightDirection.xyz() * lightProperties.y);
    float2 fakeRamp = float2(lightProperties.z, moreLightProperties.y);
    lightOpacity *= coneTrace(df, fakeLightCenter, fakeRamp, float2(lightProperties.w, moreLightProperties.y),
                              shadedPixelPosition + (SELF_OCCLUSION_HACK * shadedPixelNormal), traceShadows);
    if (!visible) return false;  // clip(visible ? 1 : -1)
    opacity = lightOpacity;
    return true;
}

// DirectionalLightVertexShader (DirectionalLight.fx:19-37): quad = lerp(min, max, corner), corner in {0,1}^2
bool directionalLightCovers(const Frame& fr, const ilb_light_vertex& v, float px, float py) {
    float2 s = fr.GetViewportScale() * fr.getEnvironmentRenderScale();
    float wx = (px + 0.5f) / s.x + fr.GetViewportPosition().x;
    float wy = (py + 0.5f) / s.y + fr.GetViewportPosition().y;
    return (wx >= v.LightPosition1.x) && (wx <= v.LightPosition2.x) && (wy >= v.LightPosition1.y) &&
           (wy <= v.LightPosition2.y);
}

bool DirectionalLightPixelShader(const Frame& fr, const DistanceField& df, const ilb_light_vertex& v, float2 vpos,
                                 float4& result) {  // DirectionalLight.fx:95-127
    float4 lightDirection = f4(v.Color2), lightProperties = f4(v.LightProperties);
    float4 moreLightProperties = f4(v.MoreLightProperties), color = f4(v.Color1);
    float4 evenMoreLightProperties = f4(v.EvenMoreLightProperties);
    float3 shadedPixelPosition, shadedPixelNormal;
    bool enableShadows, fullbright;
    sampleGBuffer(fr, vpos, shadedPixelPosition, shadedPixelNormal, enableShadows, fullbright);
    if (fullbright || checkShadowFilter(evenMoreLightProperties, enableShadows)) return false;
    lightProperties.x *= enableShadows ? 1.0f : 0.0f;
    float opacity;
    if (!DirectionalLightPixelCore(df, shadedPixelPosition, shadedPixelNormal, lightDirection, lightProperties,
                                   moreLightProperties, opacity))
        return false;
    result = float4(color.xyz() * color.w * opacity, 1);
    return true;
}

// ---------------------------------------------------------------- line light (L9)
float3 closestPointOnLineSegment3(float3 a, float3 b, float3 pt, float& t) {  // DistanceFieldCommon.fxh:151-155
    float3 ab = b - a;
    t = saturate(dot(pt - a, ab) / dot(ab, ab));
    return a + t * ab;
}

float rectangleSolidAngle(float3 worldPos, float3 p0, float3 p1, float3 p2, float3 p3) {  // FBPBR.fxh:33-51
    float3 v0 = p0 - worldPos, v1 = p1 - worldPos, v2 = p2 - worldPos, v3 = p3 - worldPos;
    float3 n0 = normalize(cross(v0, v1));
    float3 n1 = normalize(cross(v1, v2));
    float3 n2 = normalize(cross(v2, v3));
    float3 n3 = normalize(cross(v3, v0));
    float g0 = dm_acosf(dot(-n0, n1));
    float g1 = dm_acosf(dot(-n1, n2));
    float g2 = dm_acosf(dot(-n2, n3));
    float g3 = dm_acosf(dot(-n3, n0));
    return g0 + g1 + g2 + g3 - 2 * PI;
}

float computeLineLightOpacity(float3 worldPos, float3 worldNormal, float3 P0, float3 P1, float4 lightProperties,
                              float3& spherePosition, float& u) {  // FBPBR.fxh:53-101
    float3 lightLeft = normalize(P1 - P0);
    float3 lightCenter = lerp(P0, P1, 0.5f);
    float lightRadius = lightProperties.x;

    spherePosition = closestPointOnLineSegment3(P0, P1, worldPos, u);
    float3 forward = normalize(spherePosition - worldPos);
    float3 up = cross(lightLeft, forward);
    float3 p0 = P0 + lightRadius * up;
    float3 p1 = P0 - lightRadius * up;
    float3 p2 = P1 - lightRadius * up;
    float3 p3 = P1 + lightRadius * up;
    float solidAngle = rectangleSolidAngle(worldPos, p0, p1, p2, p3);
    float illuminance = solidAngle * 0.2f *
                        (saturate(dot(normalize(p0 - worldPos), worldNormal)) +
                         saturate(dot(normalize(p1 - worldPos), worldNormal)) +
                         saturate(dot(normalize(p2 - worldPos), worldNormal)) +
                         saturate(dot(normalize(p3 - worldPos), worldNormal)) +
                         saturate(dot(normalize(lightCenter - worldPos), worldNormal)));
    float3 sphereUnormL = spherePosition - worldPos;
    float3 sphereL = normalize(sphereUnormL);
    float sqrSphereDistance = dot(sphereUnormL, sphereUnormL);
    float illuminanceSphere = PI * saturate(dot(sphereL, worldNormal)) * ((lightRadius * lightRadius) / sqrSphereDistance);
    illuminance = illuminance + illuminanceSphere;
    return saturate(illuminance);
}

float lineConeTrace(const DistanceField& df, float3 startPosition, float3 endPosition, float u, float2 lightRamp,
                    float2 coneGrowthFactorAndDistanceFalloff, float3 shadedPixelPosition, bool enable) {  // LineLightCore.fxh:17-68
    TraceState a, b, c;
    float3 delta = endPosition - startPosition;
    float deltaLength = length(delta);
    float offset = max(saturate((lightRamp.x + 1) / deltaLength), 0.03f);

    coneTraceInitialize(a, shadedPixelPosition, startPosition + saturate(u - offset) * delta, TRACE_INITIAL_OFFSET_PX, lightRamp.x, false);
    coneTraceInitialize(b, shadedPixelPosition, startPosition + u * delta, TRACE_INITIAL_OFFSET_PX, lightRamp.x, false);
    coneTraceInitialize(c, shadedPixelPosition, startPosition + saturate(u + offset) * delta, TRACE_INITIAL_OFFSET_PX, lightRamp.x, false);

    float4 config = createTraceConfig(df, lightRamp, coneGrowthFactorAndDistanceFalloff);
    float stepsRemaining = df.getStepLimit();
    float liveness = ((df.Extent.x > 0) && enable) ? 1.0f : 0.0f;
    while (liveness > 0) {
        float stepLiveness = coneTraceAdvanceEx(df, a, config) + coneTraceAdvanceEx(df, b, config) +
                             coneTraceAdvanceEx(df, c, config);
        stepsRemaining--;
        liveness = stepsRemaining * stepLiveness;
    }
    float stepWindowVisibility = stepsRemaining / MAX_STEP_RAMP_WINDOW;
    float visibility = min((a.data.z + b.data.z + c.data.z) / 3, stepWindowVisibility);
    float finalResult = traceFinalResult(df, visibility);
    return enable ? finalResult : 1.0f;
}

bool LineLightPixelCore(const DistanceField& df, float3 shadedPixelPosition, float3 shadedPixelNormal,
                        float3 startPosition, float3 endPosition, float& u, float4 lightProperties,
                        float4 moreLightProperties, float& opacity) {  // LineLightCore.fxh:70-120
    const float SELF_OCCLUSION_HACK = 1.5f;
    const float SHADOW_OPACITY_THRESHOLD = (0.75f / 255.0f);
    float4 coneLightProperties = lightProperties;
    float3 lightCenter;
    float distanceOpacity = computeLineLightOpacity(shadedPixelPosition, shadedPixelNormal, startPosition, endPosition,
                                                    lightProperties, lightCenter, u);
    bool visible = (distanceOpacity > 0) && (shadedPixelPosition.x > -9999);
    if (!visible) return false;  // clip
    moreLightProperties.x *= max(0, shadedPixelNormal.z);
    float aoOpacity = computeAO(df, shadedPixelPosition, shadedPixelNormal, moreLightProperties, visible);
    float preTraceOpacity = distanceOpacity * aoOpacity;
    bool traceShadows = visible && (lightProperties.w != 0) && (preTraceOpacity >= SHADOW_OPACITY_THRESHOLD);
    float coneOpacity = lineConeTrace(df, startPosition, endPosition, u, float2(coneLightProperties.x, coneLightProperties.y),
                                      float2(df.getConeGrowthFactor(), moreLightProperties.y),
                                      shadedPixelPosition + (SELF_OCCLUSION_HACK * shadedPixelNormal), traceShadows);
    opacity = preTraceOpacity * coneOpacity;
    return true;
}

bool LineLightPixelShader(const Frame& fr, const DistanceField& df, const ilb_light_vertex& v, float2 vpos,
                          float4& result) {  // LineLight.fx:7-42
    float3 startPosition(v.LightPosition1.x, v.LightPosition1.y, v.LightPosition1.z);
    float3 endPosition(v.LightPosition2.x, v.LightPosition2.y, v.LightPosition2.z);
    float4 lightProperties = f4(v.LightProperties), moreLightProperties = f4(v.MoreLightProperties);
    float4 startColor = f4(v.Color1), endColor = f4(v.Color2);
    float3 shadedPixelPosition, shadedPixelNormal;
    bool enableShadows, fullbright;
    sampleGBuffer(fr, vpos, shadedPixelPosition, shadedPixelNormal, enableShadows, fullbright);
    if (fullbright) return false;
    lightProperties.w *= enableShadows ? 1.0f : 0.0f;
    float u, opacity;
    if (!LineLightPixelCore(df, shadedPixelPosition, shadedPixelNormal, startPosition, endPosition, u, lightProperties,
                            moreLightProperties, opacity))
        return false;
    float4 color = lerp(startColor, endColor, u);
    result = float4(color.xyz() * color.w * opacity, 1);
    return true;
}

// LineLightVertexShader (LineLightCore.fxh:122-173): bounds = min/max(start,end) -/+ 9999
bool lineLightCovers(const Frame& fr, const ilb_light_vertex& v, float px, float py) {
    float2 s = fr.GetViewportScale() * fr.getEnvironmentRenderScale();
    float wx = (px + 0.5f) / s.x + fr.GetViewportPosition().x;
    float wy = (py + 0.5f) / s.y + fr.GetViewportPosition().y;
    float radius = v.LightProperties.x + v.LightProperties.y + 1;
    float x0 = min(v.LightPosition1.x, v.LightPosition2.x) - 9999, x1 = max(v.LightPosition1.x, v.LightPosition2.x) + 9999;
    float y0 = min(v.LightPosition1.y, v.LightPosition2.y) - 9999, y1 = max(v.LightPosition1.y, v.LightPosition2.y) + 9999;
    y0 -= radius * fr.getInvZToYMultiplier();
    y0 -= v.LightPosition1.z * fr.getZToYMultiplier();
    return (wx >= x0) && (wx <= x1) && (wy >= y0) && (wy <= y1);
}

// stencil mask of UpdateMaskFromGBuffer (Shaders/GBufferMask.fx:26-44)
bool stencilMaskPasses(const Frame& fr, float4 g) {
    float minW = -abs(fr.getMaximumZ()) - 1;
    float maxW = -abs(fr.getGroundZ()) - 1;
    return !((g.w >= 9999) || (g.w < minW) || ((g.w < 0) && (g.w > maxW)));
}

Frame makeFrame(const ilb_lighting_frame* f, const void* gbuffer, int gw, int gh, int gfmt) {
    Frame fr;
    fr.EnvironmentZAndScale = f4(f->EnvironmentZAndScale);
    fr.EnvironmentZToY = f4(f->EnvironmentZToY);
    fr.GBufferTexelSizeAndMisc = f4(f->GBufferTexelSizeAndMisc);
    fr.GBufferViewportRelative = f->GBufferViewportRelative;
    fr.ViewportPosition = float2(f->ViewportPosition[0], f->ViewportPosition[1]);
    fr.gbuffer = gbuffer;
    fr.gw = gw; fr.gh = gh; fr.gfmt = gfmt;
    fr.stencilCulling = f->stencil_culling != 0;
    if (!gbuffer) fr.GBufferTexelSizeAndMisc.x = fr.GBufferTexelSizeAndMisc.y = 0;
    return fr;
}

}  // namespace

// ================================================================== exported entry points

extern "C" {

// The ramp textures ilb_light_batch.ramp_texture refers to (id - 1 indexes the arrays): float4 texels, kept by reference.
void orc_set_ramp_textures(int count, const float* const* texels, const int* widths, const int* heights) {
    g_ramps.clear();
    for (int i = 0; i < count; i++) g_ramps.push_back(RampTexture{texels[i], widths[i], heights[i]});
}

// RenderLighting (Lighting/LightingRenderer.cs:917-1168) in the reference's multi-pass form:
// clear to ambient, then one full-band pass per light with additive blend (BlendState.Additive:
// rgb += src.rgb * src.a(=1), a += 1).  Output: fp32 float4, width*(row_end-row_begin).
int orc_render_lighting(const uint16_t* df_tex, int tw, int th, const void* gbuffer, int gw, int gh, int gfmt,
                        const ilb_lighting_frame* f, const ilb_light_batch* batches, int nbatches,
                        const ilb_light_vertex* verts, int nverts, float* out, int nthreads) {
    return orc_render_lighting_strided(df_tex, tw, th, gbuffer, gw, gh, gfmt, f, batches, nbatches, verts, nverts, 1, out, nthreads);
}

// The same passes over rows row_begin, row_begin + row_stride, ... < row_end only (a sample of rows spread over a frame
// whose cost is not uniform); `out` holds those rows packed: width * ceil((row_end - row_begin) / row_stride) float4.
int orc_render_lighting_strided(const uint16_t* df_tex, int tw, int th, const void* gbuffer, int gw, int gh, int gfmt,
                                const ilb_lighting_frame* f, const ilb_light_batch* batches, int nbatches,
                                const ilb_light_vertex* verts, int nverts, int row_stride, float* out, int nthreads) {
    if (row_stride < 1) return ILB_ERR_INVALID_ARGUMENT;
    const int W = f->width, r0 = f->row_begin, r1 = f->row_end;
    const int nrows = (r1 - r0 + row_stride - 1) / row_stride;
    Frame fr = makeFrame(f, gbuffer, gw, gh, gfmt);
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
    for (int yi = 0; yi < nrows; yi++)
        for (int x = 0; x < W; x++) {
            float* o = out + 4 * ((size_t)yi * W + x);
            o[0] = f->ClearColor.x; o[1] = f->ClearColor.y; o[2] = f->ClearColor.z; o[3] = f->ClearColor.w;
        }
    for (int b = 0; b < nbatches; b++) {
        const ilb_light_batch& batch = batches[b];
        DistanceField df(df_tex, tw, th, batch.df);
        if (!df_tex) df.Extent.x = 0;
        for (int i = 0; i < batch.vertex_count; i++) {
            int vi = batch.first_vertex + i;
            if (vi < 0 || vi >= nverts) return ILB_ERR_INVALID_ARGUMENT;
            const ilb_light_vertex& v = verts[vi];
#pragma omp parallel for collapse(2) schedule(dynamic, 256)
            for (int yi = 0; yi < nrows; yi++)
                for (int x = 0; x < W; x++) {
                    const int y = r0 + yi * row_stride;
                    float4 result;
                    bool lit = false;
                    float2 vpos((float)x, (float)y);
                    if (fr.stencilCulling && any(fr.GBufferTexelSizeAndMisc.xy())) {
                        float3 p, n; bool es, fb; float4 raw;
                        sampleGBuffer(fr, vpos, p, n, es, fb, &raw);
                        if (!stencilMaskPasses(fr, raw)) continue;
                    }
                    switch (batch.light_type) {
                        case ILB_LIGHT_SPHERE:
                            lit = sphereLightCovers(fr, v, (float)x, (float)y) && SphereLightPixelShader(fr, df, v, vpos, result, batch.ramp_texture);
                            break;
                        case ILB_LIGHT_DIRECTIONAL:
                            lit = directionalLightCovers(fr, v, (float)x, (float)y) && DirectionalLightPixelShader(fr, df, v, vpos, result);
                            break;
                        case ILB_LIGHT_LINE:
                            lit = lineLightCovers(fr, v, (float)x, (float)y) && LineLightPixelShader(fr, df, v, vpos, result);
                            break;
                        case ILB_LIGHT_PARTICLE:
                            lit = particleLightCovers(fr, v, (float)x, (float)y) && ParticleLightPixelShader(fr, df, v, vpos, result);
                            break;
                        default:
                            break;
                    }
                    if (lit) {
                        float* o = out + 4 * ((size_t)yi * W + x);
                        o[0] += result.x; o[1] += result.y; o[2] += result.z; o[3] += result.w;
                    }
                }
        }
    }
    return 0;
}

// UpdateLightProbes (Lighting/LightingRenderer.LightProbes.cs:49-110) with the probe pixel shaders
// SphereLightProbe.fx:19-44, DirectionalLight.fx:163-190, LineLightProbe.fx:23-48 (line lights are shaded
// as sphere lights at LightPosition1 -- reference quirk).  Output fp32 float4[nprobes], cleared to 0.
int orc_update_light_probes(const uint16_t* df_tex, int tw, int th, const ilb_lighting_frame* f,
                            const ilb_light_batch* batches, int nbatches, const ilb_light_vertex* verts, int nverts,
                            const ilb_float4* positions, const ilb_float4* normals, int nprobes, float* out) {
    Frame fr = makeFrame(f, nullptr, 0, 0, 0);
    for (int p = 0; p < nprobes; p++) out[4 * p] = out[4 * p + 1] = out[4 * p + 2] = out[4 * p + 3] = 0;
    for (int b = 0; b < nbatches; b++) {
        const ilb_light_batch& batch = batches[b];
        DistanceField df(df_tex, tw, th, batch.df);
        if (!df_tex) df.Extent.x = 0;
        for (int i = 0; i < batch.vertex_count; i++) {
            int vi = batch.first_vertex + i;
            if (vi < 0 || vi >= nverts) return ILB_ERR_INVALID_ARGUMENT;
            const ilb_light_vertex& v = verts[vi];
            for (int p = 0; p < nprobes; p++) {
                // sampleLightProbeBuffer (LightCommon.fxh:233-254)
                float4 positionSample = f4(positions[p]);
                float opacity = positionSample.w;
                if (opacity <= 0) continue;  // discard
                float3 shadedPixelPosition = positionSample.xyz();
                float4 n = f4(normals[p]);
                float3 shadedPixelNormal = n.xyz();
                float enableShadows = n.w;
                float4 lightProperties = f4(v.LightProperties), moreLightProperties = f4(v.MoreLightProperties);
                float4 color = f4(v.Color1);
                float coreOpacity;
                bool lit;
                if (batch.light_type == ILB_LIGHT_DIRECTIONAL) {
                    lightProperties.x *= enableShadows;
                    moreLightProperties.x = moreLightProperties.w = 0;
                    lit = DirectionalLightPixelCore(df, shadedPixelPosition, shadedPixelNormal, f4(v.Color2), lightProperties,
                                                    moreLightProperties, coreOpacity);
                } else if (batch.light_type == ILB_LIGHT_SPHERE || batch.light_type == ILB_LIGHT_LINE) {
                    float3 lightCenter(v.LightPosition1.x, v.LightPosition1.y, v.LightPosition1.z);
                    lightProperties.w *= enableShadows;
                    moreLightProperties.x = moreLightProperties.w = 0;
                    lit = SphereLightPixelCore(fr, df, shadedPixelPosition, shadedPixelNormal, lightCenter, lightProperties,
                                               moreLightProperties, coreOpacity);
                } else {
                    return ILB_ERR_INVALID_ARGUMENT;
                }
                if (!lit) continue;
                opacity *= coreOpacity;
                float3 rgb = color.xyz() * color.w * opacity;
                out[4 * p + 0] += rgb.x; out[4 * p + 1] += rgb.y; out[4 * p + 2] += rgb.z; out[4 * p + 3] += 1;
            }
        }
    }
    return 0;
}

// ---- unit entry points for known-answer tests
float orc_sample_distance_field(const uint16_t* df_tex, int tw, int th, const ilb_df_uniforms* u, float x, float y, float z) {
    DistanceField df(df_tex, tw, th, *u);
    return df.sampleDistanceFieldEx(float3(x, y, z));
}

float orc_cone_trace(const uint16_t* df_tex, int tw, int th, const ilb_df_uniforms* u, const float* lightCenter,
                     float radius, float rampLength, float growth, float distanceFalloff, const float* shadedPos,
                     int enable, int* steps) {
    DistanceField df(df_tex, tw, th, *u);
    if (!df_tex) df.Extent.x = 0;
    return coneTrace(df, float3(lightCenter[0], lightCenter[1], lightCenter[2]), float2(radius, rampLength),
                     float2(growth, distanceFalloff), float3(shadedPos[0], shadedPos[1], shadedPos[2]), enable != 0, steps);
}

float orc_sphere_light_opacity(const ilb_lighting_frame* f, const float* pos, const float* normal, const float* center,
                               const float* lightProperties, float yFactor) {
    Frame fr = makeFrame(f, nullptr, 0, 0, 0);
    return computeSphereLightOpacity(fr, float3(pos[0], pos[1], pos[2]), float3(normal[0], normal[1], normal[2]),
                                     float3(center[0], center[1], center[2]),
                                     float4(lightProperties[0], lightProperties[1], lightProperties[2], lightProperties[3]), yFactor);
}

void orc_decode_gbuffer(const ilb_lighting_frame* f, const void* gbuffer, int gw, int gh, int gfmt, int x, int y,
                        float* worldPos, float* normal, int* enableShadows, int* fullbright) {
    Frame fr = makeFrame(f, gbuffer, gw, gh, gfmt);
    float3 p, n; bool es, fb;
    sampleGBuffer(fr, float2((float)x, (float)y), p, n, es, fb);
    worldPos[0] = p.x; worldPos[1] = p.y; worldPos[2] = p.z;
    normal[0] = n.x; normal[1] = n.y; normal[2] = n.z;
    *enableShadows = es; *fullbright = fb;
}

}  // extern "C"
