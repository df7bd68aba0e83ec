// LIGHTING hot path: one fused tile kernel replaces the reference's
// "one rasterised instanced quad per light + additive ROP blend" (Lighting/LightingRenderer.cs:1149-1166).
//
//   per 16x16 tile:  decode the G-buffer once per pixel (L4)            -> registers
//                    reduce the tile's world-space AABB (warp shuffles) -> per-tile light culling
//                    ordered compaction of surviving lights (ballot)    -> shared-memory index list
//                    per pixel: loop lights in draw order, evaluate the sphere / directional / line
//                    response + AO + cone trace through the packed distance field (L2, L3, L5-L9)
//                    accumulate in fp32 registers, one lightmap store per pixel (L10)
//
// The per-light math keeps the operation order of the reference pixel shaders (file:line cited per function,
// relative to Illuminant/Shaders/); what changes is data movement: G-buffer and lightmap are touched once per
// pixel instead of once per light-pixel, the light list is culled per tile, and there is no blend unit.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include <cuda_fp16.h>

#include "ilb_internal.h"

namespace {

#ifndef ILB_TILE_H
#define ILB_TILE_H 16
#endif
#ifndef ILB_SPLIT_MARCH
#define ILB_SPLIT_MARCH 1   // rays that leave the field volume march in two loops (inside part, outside part), see coneTraceMarch
#endif
#ifndef ILB_SWIZZLE
#define ILB_SWIZZLE 16  // CTAs walk down bands of this many tile rows (column-major inside a band): 2-D locality in L1/L2
                        // (C4 frame: 1 -> 7.39 ms, 4 -> 7.41, 8 -> 7.32, 16 -> 7.28, 32 -> 7.26, 64 -> 7.28; 270-row bands are best at 16)
#endif
constexpr int TILE_W = 16, TILE_H = ILB_TILE_H, TILE_THREADS = TILE_W * TILE_H, TILE_WARPS = TILE_THREADS / 32;
constexpr int MAX_OUTPUTS = 8;
constexpr int ILB_LIGHT_PARTICLE_BIT = 8;  // DLight::type of a particle light (a bit for the TYPES masks; the public id, ILB_LIGHT_PARTICLE, is 3)
constexpr int ILB_LIGHT_RAMP_BIT = 16;     // DLight::type of a sphere light drawn with a ramp texture = ILB_LIGHT_SPHERE | ILB_LIGHT_RAMP_BIT
constexpr int ILB_LIGHT_TYPE_MASK = 15;

// One light, flattened on the host from (batch, LightVertex): the LightVertex fields (Vertices.cs:10-39) plus
// the batch's quality uniforms and the rasterised coverage of its quad.
struct __align__(16) DLight {
    float4 pos1, pos2;                       // LightPosition1, LightPosition2
    float4 props, more, evenMore;            // LightProperties, MoreLightProperties, EvenMoreLightProperties
    float4 color1, color2;                   // Color1, Color2
    float4 quality;                          // MaxConeRadius, OcclusionToOpacityPower, StepLimit, MinStepSize
    float longStep;                          // LongStepFactor
    int hasField;                            // batch.df.Extent.x > 0 && a field is bound
    int type;                                // ilb_light_type
    float rcpRamp;                           // RN(1 / LightProperties.y) for udiv(), 0 when that is not a safe normal number
    float4 covX, covY;                       // coverage edges, see coverage()
    int px0, py0, px1, py1;                  // conservative pixel bounds of the quad (inclusive)
};

// Per-light values of the line-light shaders that do not depend on the pixel, evaluated once on the host with the same
// IEEE fp32 operations (flattenLights): normalize(P1 - P0), lerp(P0, P1, 0.5), P1 - P0 and its squared / plain length,
// the reciprocal of the squared length for udiv(), and lineConeTrace's `offset`.  Indexed like the DLight array.
struct __align__(16) DLine {
    float4 left;    // lightLeft.xyz, RN(1 / dot(ab, ab)) for udiv() (0 when that is not a safe normal number)
    float4 center;  // lightCenter.xyz, offset = max(saturate((radius + 1) / |P1 - P0|), 0.03)
    float4 ab;      // (P1 - P0).xyz, dot(ab, ab)
};

// The frame's light records in the constant bank (frames of up to ILB_CONST_LIGHTS lights; larger lists -- particle lights --
// stay in global memory).  Every lane of a warp shades the same light at the same time, so a constant-bank read is a
// broadcast; what it buys over the global copy is REGISTERS: a field read with LDC c[3][index + offset] is re-read where it
// is used (after the cone trace: colour, specular, ...) instead of being loaded up front and kept -- or spilled -- across the
// march.  The culling phase (one light per THREAD) and the out-of-line IEEE fallback keep reading the global copy.
constexpr int ILB_CONST_LIGHTS = 256;
__constant__ DLight c_lights[ILB_CONST_LIGHTS];
__constant__ DLine c_lines[ILB_CONST_LIGHTS];

// One LightSource.RampTexture as the kernel samples it (float4 texels; RampCommon.fxh:5-12: LINEAR, U CLAMP, V WRAP)
struct RampTex { const float4* texels; int w, h, pad; };

struct LightingParams {
    DFGeometry df;
    float4 envZAndScale, envZToY, gbTexelAndMisc, clear;
    float gbViewportRelative, vpx, vpy;
    const void* gbuffer;
    int gw, gh, gfmt;
    const DLight* lights;
    const DLine* lines;   // per-light line constants (valid for line lights)
    const struct RampTex* ramps;  // ramp textures (SphereLightWithDistanceRamp); DLight::evenMore.y = 1-based index, 0 = none
    int nlights;
    int width, height, row_begin, row_end;
    int out_format, stencil;
    void* outs[MAX_OUTPUTS];
    int nouts;
    int out_row_base;  // row index of outs[] row 0 (row_begin for band buffers, 0 for full frames)
    const unsigned* tile_order;  // heaviest-first permutation of the tile indices (ILB_OPT_LIGHT_TILE_ORDER), or null
    int tiles_x, tiles_y;
    const float4* accum_in;  // fp32 sums of an earlier pass over this row band (nullptr: start from `clear`)
    float4* accum_out;       // leave the fp32 sums here instead of storing the lightmap (nullptr: final pass)
    // concurrent passes (light_accumulate_persistent_kernel): this pass's tile queue, and the per-tile arrival counters of
    // the two passes -- then accum_out is this pass's scratch and accum_in the other pass's
    unsigned* tile_counter;
    unsigned* tile_done;
};

struct Pixel {
    f3 pos, normal, camera;
    bool enableShadows, fullbright, maskOk;
};

// ---- cone trace (ConeTrace.fxh) --------------------------------------------------------------------------
// Template parameters used throughout the per-light code:
//   FIELD  0 = sample the Rgba64 atlas, 1 = sample the expanded planes (ilb_device.cuh); bit-identical results
//   FAST   true = sqrt / reciprocal through the deferred-guard forms (`bad` collects the range checks; when it comes
//          back set the caller re-evaluates the light with FAST = false), false = plain IEEE x-ops
#define MIN_CONE_RADIUS 0.33f
#define MAX_STEP_RAMP_WINDOW 2.0f
#define TRACE_INITIAL_OFFSET_PX 0.5f
#define FULLY_SHADOWED_THRESHOLD 0.075f
#define UNSHADOWED_THRESHOLD 0.95f
#define HACK_DISTANCE_OFFSET 1.5f
#define TRACE_END_MULTIPLIER 100.0f

struct TraceConfig {  // createTraceConfig :122-139 (+ the quality uniforms the step reads)
    float maxRadius, growth, minStep, longStep, stepLimit, power;
};

ILB_DEV TraceConfig makeTraceConfig(const DLight& L, float rampX, float rampY, float growthFactor) {
    TraceConfig c;
    c.maxRadius = clampf(rampX, MIN_CONE_RADIUS, L.quality.x);
    const float rampLength = fmaxf(rampY, 16.0f);
    c.growth = c.maxRadius / rampLength * growthFactor;
    c.minStep = fmaxf(1.0f, L.quality.w);
    c.longStep = L.longStep;
    c.stepLimit = L.quality.z;
    c.power = L.quality.y;
    return c;
}

struct Trace {  // TraceState :31-35
    f3 origin, direction;
    float t, len, vis;
    float tSafe;  // lineTraceMarch: see safeInsideLimit
};

// returns true when the marched interval t < len stays on the segment start -> end
template <bool FAST>
ILB_DEV bool traceInit(Trace& s, f3 start, f3 end, float lightRadius, Guard& bad) {  // coneTraceInitialize :37-49
    const f3 v = xsub3(end, start);
    const float l = tlength3<FAST>(v, bad);
    s.origin = start;
    s.direction = xdivs3(v, l);
    s.len = fmaxf(xsub(l, lightRadius), 1.0f);
    s.t = TRACE_INITIAL_OFFSET_PX;
    s.vis = 1.0f;
    return s.len <= l;
}

ILB_DEV float traceStep(const TraceConfig& c, float d, float offset, float& vis) {  // coneTraceStep :51-71
    const float localSphereRadius = fminf((c.growth * offset) + MIN_CONE_RADIUS, c.maxRadius);
    // smooth factor, and the divisor lies in [0.33, MaxConeRadius]: reciprocal + multiply without the range scaling of `/`
    const float localVisibility = __fdividef(d + HACK_DISTANCE_OFFSET, localSphereRadius);
    vis = fminf(vis, localVisibility);  // smooth: a few ulp in vis cannot change the result discontinuously
    return fmaxf(xmul(fabsf(d), c.longStep), c.minStep);  // exact: the step length moves the march
}

ILB_DEV float traceFinal(const TraceConfig& c, float visibility) {  // :182-188
    const float v = saturatef(saturatef(visibility - FULLY_SHADOWED_THRESHOLD) / (UNSHADOWED_THRESHOLD - FULLY_SHADOWED_THRESHOLD));
    if (c.power == 1.0f) return v;  // uniform per light: OcclusionToOpacityPower defaults to 1 (pow(v, 1) == v)
    return powf(v, c.power);
}

// Largest ray parameter up to which origin + direction * t provably stays inside the field volume (a conservative bound:
// 0.05 px of slack on every face covers the rounding of the per-sample position, which is below 1e-3 px for coordinates
// up to 8192).  The march then tests `t <= tSafe` (one compare) instead of the six-compare volume test per sample; beyond
// the bound it samples through the general path, which is correct everywhere, so the bits never depend on the bound.
ILB_DEV float safeInsideLimit(const DFGeometry& g, f3 o, f3 d) {
    const float m = 0.05f;
    const float oz = o.z - g.zOffset;
    if (!((o.x >= m) && (o.x <= g.ex - m) && (o.y >= m) && (o.y <= g.ey - m) && (oz >= m) && (oz <= g.ez - m))) return -1.0f;
    // a NaN direction (zero-length ray: its light-pixel is re-evaluated by the IEEE fallback) must not reach the short sampler,
    // whose texel addressing (floorBiased) is only defined for finite coordinates
    if (!((d.x == d.x) && (d.y == d.y) && (d.z == d.z))) return -1.0f;
    const float BIG = 3.0e38f;
    const float tx = (d.x > 0.0f) ? __fdividef((g.ex - m) - o.x, d.x) : ((d.x < 0.0f) ? __fdividef(m - o.x, d.x) : BIG);
    const float ty = (d.y > 0.0f) ? __fdividef((g.ey - m) - o.y, d.y) : ((d.y < 0.0f) ? __fdividef(m - o.y, d.y) : BIG);
    const float tz = (d.z > 0.0f) ? __fdividef((g.ez - m) - oz, d.z) : ((d.z < 0.0f) ? __fdividef(m - oz, d.z) : BIG);
    return fminf(fminf(tx, ty), tz) * (1.0f - 3.0e-5f);
}

// one step of the march (coneTraceAdvance :73-82 + the loop condition :168-176); returns the liveness of the trace
template <int FIELD, bool INSIDE>
ILB_DEV bool coneTraceMarchStep(const DFGeometry& g, const TraceConfig& c, Trace& a, float& stepsRemaining) {
    stepsRemaining -= 1.0f;
    const f3 sp = xadd3(a.origin, xscale3(a.direction, a.t));
    const float d = sampleFieldT<FIELD, INSIDE>(g, sp);
    a.t = xadd(a.t, traceStep(c, d, a.t, a.vis));
    // liveness = stepsRemaining * saturate(vis - 0.075) * saturate(len - t) > 0 (ConeTrace.fxh:168-176): a product of
    // non-negative factors that cannot underflow (a non-zero factor is at least one ulp of 0.075 resp. of t >= 0.5), so it
    // is positive exactly when every factor is
    return (stepsRemaining > 0.0f) && (a.vis > FULLY_SHADOWED_THRESHOLD) && (a.len > a.t);
}

template <int FIELD, bool INSIDE>
ILB_DEV void coneTraceMarch(const DFGeometry& g, const TraceConfig& c, Trace& a, float& stepsRemaining) {
    bool live = true;
#if ILB_SPLIT_MARCH
    if (!INSIDE) {
        // A ray that ends outside the volume still spends most of its samples inside it: there the clamp is the identity and the
        // distance-to-volume term is exactly 0, so the short sampler gives the same bits.  t only grows, so the march is two
        // loops -- the short sampler while t <= tSafe, the general one for the rest -- each with one copy of its sampler.
        const float tSafe = safeInsideLimit(g, a.origin, a.direction);
        while (live && (a.t <= tSafe)) live = coneTraceMarchStep<FIELD, true>(g, c, a, stepsRemaining);
    }
    while (live) live = coneTraceMarchStep<FIELD, INSIDE>(g, c, a, stepsRemaining);
#else
    // a ray that ends outside the volume still spends most of its samples inside it: there the clamp is the
    // identity and the distance-to-volume term is exactly 0, so the short sampler gives the same bits
    const float tSafe = INSIDE ? 0.0f : safeInsideLimit(g, a.origin, a.direction);
    while (live) {
        if (INSIDE || (a.t <= tSafe)) live = coneTraceMarchStep<FIELD, true>(g, c, a, stepsRemaining);
        else live = coneTraceMarchStep<FIELD, false>(g, c, a, stepsRemaining);
    }
#endif
}

template <int FIELD, bool FAST>
ILB_DEV float coneTrace(const DFGeometry& g, const DLight& L, f3 lightCenter, float rampX, float rampY,
                        float growthFactor, f3 shaded, bool enable, Guard& bad) {  // coneTrace :141-191
    // a disabled trace returns 1 whatever its state (ConeTrace.fxh:159,190): its set-up (a normalisation and three IEEE
    // divisions) is only run for the pixels that march
    if (!enable) return 1.0f;
    const TraceConfig c = makeTraceConfig(L, rampX, rampY, growthFactor);
    float stepsRemaining = c.stepLimit, vis = 1.0f;
    if (L.hasField) {
        Trace a;
        const bool onSegment = traceInit<FAST>(a, shaded, lightCenter, rampX, bad);
        // every sample lies on the segment shaded -> lightCenter (t < len <= |v|); the field volume is convex, so when both
        // ends are inside every sample is inside and the clamp / distance-to-volume work of the sampler is skipped
        const bool inside = onSegment && insideField(g, shaded) && insideField(g, lightCenter);
#if ILB_NO_INSIDE_PATH
        coneTraceMarch<FIELD, false>(g, c, a, stepsRemaining);
#else
        if (inside) coneTraceMarch<FIELD, true>(g, c, a, stepsRemaining);
        else coneTraceMarch<FIELD, false>(g, c, a, stepsRemaining);
#endif
        vis = a.vis;
    }
    return traceFinal(c, fminf(vis, stepsRemaining / MAX_STEP_RAMP_WINDOW));
}

template <int FIELD, bool INSIDE>
ILB_DEV float traceAdvanceEx(const DFGeometry& g, const TraceConfig& c, Trace& s) {  // coneTraceAdvanceEx :84-96
    const f3 sp = xadd3(s.origin, xscale3(s.direction, s.t));
    const float d = (INSIDE || (s.t <= s.tSafe)) ? sampleFieldT<FIELD, true>(g, sp) : sampleFieldT<FIELD, false>(g, sp);
    s.t = fminf(xadd(s.t, traceStep(c, d, s.t, s.vis)), s.len);
    // saturate(vis - 0.075) * saturate((len - t) * 100): only its sign is used (lineConeTrace sums three of these and
    // multiplies by stepsRemaining); positive exactly when both factors are (no underflow, see coneTraceMarch)
    return ((s.vis > FULLY_SHADOWED_THRESHOLD) && (s.len > s.t)) ? 1.0f : 0.0f;
}

// ---- light response (LightCommon.fxh, AOCommon.fxh) ---------------------------------------------------------
#define DOT_EXPONENT 0.85f

// Division by a per-light / per-shader constant whose reciprocal is at hand.  FAST: the three-instruction Markstein form (udiv),
// which equals div.rn for every FINITE dividend whose quotient does not overflow -- but turns an infinite dividend into NaN
// (inf * r - y * inf) where div.rn gives +-inf, and saturate() maps those to different ends of [0, 1].  Infinite operands only
// come from non-finite G-buffer texels (a HalfVector4 G-buffer stores its "dead" marker -99999 as -inf); they also trip the
// range guard of the square roots next to every one of these divisions, so the re-evaluation (FAST = false) is where they are
// seen, and it divides the IEEE way.
template <bool FAST>
ILB_DEV float ldiv(float x, float y, float r) { return FAST ? udiv(x, y, r) : xdivz(x, y); }

template <int RANGE_1000, bool FAST>  // offset == range == RANGE_1000 / 1000 (0.15 for sphere lights, 0.35 for directional lights)
ILB_DEV float normalFactorEx(f3 lightNormal, f3 n) {  // computeNormalFactorEx :154-165
    if (!any3(n)) return 1.0f;
    constexpr float range = (float)RANGE_1000 / 1000.0f;
    static_assert(RANGE_1000 == 150 || RANGE_1000 == 350, "the two call sites of the reference");
    const float d = xdot3(-lightNormal, n);  // exact: its sign decides `visible` (discard / alpha count)
    return powf(saturatef(ldiv<FAST>(xadd(d, range), range, 1.0f / range)), DOT_EXPONENT);
}

template <bool FAST>
ILB_DEV float sphereLightOpacity(float lightOcclusion, f3 p, f3 n, f3 center, float4 props, float yFactor, float rRamp, Guard& bad) {  // :173-210
    // x-ops: where this falloff reaches exactly 0 the fragment is discarded, which changes the lightmap's alpha count
    f3 d3 = xsub3(p, center);
    d3.y = xmul(d3.y, yFactor);
    const float distance = tlength3<FAST>(d3, bad);
    float distanceFactor = xsub(1.0f, saturatef(ldiv<FAST>(xsub(distance, props.x), props.y, rRamp)));
    if (lightOcclusion > 0.0f) distanceFactor = xmul(distanceFactor, xsub(1.0f, saturatef(xdiv(d3.z, lightOcclusion))));
    const f3 lightNormal = xdivs3(d3, distance);
    float normalFactor = normalFactorEx<150, FAST>(lightNormal, n);
    if (props.z >= 2.0f) {
        distanceFactor = xsub(1.0f, saturatef(xsub(distance, props.x)));
        normalFactor = 1.0f;
    } else if (props.z >= 1.0f) {
        distanceFactor = xmul(distanceFactor, distanceFactor);
    }
    return saturatef(xadd(xmul(normalFactor, distanceFactor), saturatef(xsub(props.x, distance))));
}

template <int FIELD>
ILB_DEV float computeAO(const DFGeometry& g, bool hasField, f3 p, f3 n, float aoRadius, float aoOpacity, bool visible) {  // AOCommon.fxh:1-20
    if ((aoRadius >= 0.5f) && hasField && visible) {
        const float distance = sampleFieldT<FIELD, false>(g, xadd3(p, mk3(0.0f, 0.0f, xmul(n.z, aoRadius))));
        const float clampedDistance = clampf(distance, 0.0f, aoRadius);
        float result = 1.0f - saturatef(clampedDistance / aoRadius);
        result *= result;
        result = 1.0f - result;
        return (1.0f - aoOpacity) + (result * aoOpacity);
    }
    return 1.0f;
}

// SphereLightPixelCore (SphereLightCore.fxh:58-158); returns false on discard
template <int FIELD, bool FAST>
ILB_DEV bool sphereCore(const DFGeometry& g, const DLight& L, float lightOcclusion, f3 p, f3 n, f3 center, float4 props,
                        float4 more, float& opacity, float& preTrace, Guard& bad) {
    const float distanceOpacity = sphereLightOpacity<FAST>(lightOcclusion, p, n, center, props, more.z, L.rcpRamp, bad);
    const bool visible = (distanceOpacity > 0.0f) && (p.x > -9999.0f);
    if (!visible) return false;
    const float aoRadius = xmul(more.x, fmaxf(0.0f, n.z));
    const float aoOpacity = computeAO<FIELD>(g, L.hasField != 0, p, n, aoRadius, more.w, visible);
    const float preTraceOpacity = distanceOpacity * aoOpacity;
    const bool traceShadows = (props.w != 0.0f) && (preTraceOpacity >= (0.75f / 255.0f));
    const float coneOpacity = coneTrace<FIELD, FAST>(g, L, center, props.x, props.y, 1.0f, xadd3(p, xscale3(n, 1.6f)), traceShadows, bad);
    opacity = preTraceOpacity * coneOpacity;
    preTrace = preTraceOpacity;  // SphereLightPixelCoreWithRamp hands it to the ramp lookup (coneOpacity = opacity / preTrace is not exact: the caller keeps it)
    return true;
}

// SphereLightPixelEpilogueWithRamp (SphereLightCore.fxh:99-119): RampTexture(preTraceOpacity, (angle + offset) * rate).rgb.
// Out of line: ramp-textured batches are rare and must not cost the sphere / directional pass registers.
__device__ __noinline__ float4 rampLookup(const RampTex* ramps, int id, float preTraceOpacity, float dx, float dy, float offset, float rate) {
    const RampTex t = ramps[id - 1];
    const float angle = atan2f(dy, dx);
    const float u = preTraceOpacity, v = (angle + offset) * rate;
    const float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f;
    const float x0f = floorf(x), y0f = floorf(y);
    const float fx = x - x0f, fy = y - y0f;
    const int x0 = (int)x0f, y0 = (int)y0f;
    const int xa = min(max(x0, 0), t.w - 1), xb = min(max(x0 + 1, 0), t.w - 1);          // U clamps
    const int ya = ((y0 % t.h) + t.h) % t.h, yb = (((y0 + 1) % t.h) + t.h) % t.h;        // V wraps
    const float4 t00 = __ldg(t.texels + (size_t)ya * t.w + xa), t10 = __ldg(t.texels + (size_t)ya * t.w + xb);
    const float4 t01 = __ldg(t.texels + (size_t)yb * t.w + xa), t11 = __ldg(t.texels + (size_t)yb * t.w + xb);
    const f4 top = lerp4(mk4(t00), mk4(t10), fx), bottom = lerp4(mk4(t01), mk4(t11), fx);
    return to_float4(lerp4(top, bottom, fy));
}

// DirectionalLightPixelCore (DirectionalLight.fx:52-93, useOpacityRamp = false)
template <int FIELD, bool FAST>
ILB_DEV bool directionalCore(const DFGeometry& g, const DLight& L, f3 p, f3 n, float4 dir, float4 props, float4 more,
                             float& opacity, Guard& bad) {
    // On a surface that faces straight up (floors, box tops) the normal factor of a directional light depends on the light
    // alone: the host evaluated it once (flattenLights, same operations, the oracle's own powf) into the record's spare word
    float lightOpacity;
    if (dir.w < 0.1f) lightOpacity = 1.0f;
    else if ((n.x == 0.0f) && (n.y == 0.0f) && (n.z == 1.0f)) lightOpacity = L.covX.z;
    else lightOpacity = normalFactorEx<350, FAST>(mk3(dir.x, dir.y, dir.z), n);
    const bool visible = (p.x > -9999.0f);
    const float aoRadius = xmul(more.x, fmaxf(0.0f, n.z));
    lightOpacity *= computeAO<FIELD>(g, L.hasField != 0, p, n, aoRadius, more.w, visible);
    const bool traceShadows = visible && (props.x != 0.0f) && (lightOpacity >= 1.0f / 256.0f) && (dir.w >= 0.1f);
    const f3 fakeLightCenter = xsub3(p, xscale3(mk3(dir.x, dir.y, dir.z), props.y));
    lightOpacity *= coneTrace<FIELD, FAST>(g, L, fakeLightCenter, props.z, more.y, props.w, xadd3(p, xscale3(n, 1.5f)), traceShadows, bad);
    if (!visible) return false;
    opacity = lightOpacity;
    return true;
}

// ---- line light (FBPBR.fxh:33-101, LineLightCore.fxh:17-120) -------------------------------------------------
// x-ops: the solid angle is a difference of four arc-cosines that nearly cancel (g0+g1+g2+g3 - 2*pi), so a few ulp in
// the normalised cross products change the illuminance by far more than 1e-4 relative -- keep it bit-identical.
template <bool FAST>
ILB_DEV float acosExact(float x, Guard& bad) {  // dm_acosf (include/ilb_detmath.h) with the sqrt of this build
    return dm_acos_finish(x, tsqrt<FAST>(xsub(1.0f, fabsf(x)), bad));
}
template <bool FAST>
ILB_DEV float rectangleSolidAngle(f3 v0, f3 v1, f3 v2, f3 v3, Guard& bad) {  // FBPBR.fxh:33-51, v_i = p_i - worldPos
    const f3 n0 = tnormalize3<FAST>(xcross3(v0, v1), bad), n1 = tnormalize3<FAST>(xcross3(v1, v2), bad);
    const f3 n2 = tnormalize3<FAST>(xcross3(v2, v3), bad), n3 = tnormalize3<FAST>(xcross3(v3, v0), bad);
    // deterministic acos: the four angles nearly cancel, so both sides must evaluate the same function
    const float g0 = acosExact<FAST>(xdot3(-n0, n1), bad), g1 = acosExact<FAST>(xdot3(-n1, n2), bad);
    const float g2 = acosExact<FAST>(xdot3(-n2, n3), bad), g3 = acosExact<FAST>(xdot3(-n3, n0), bad);
    return xsub(xadd(xadd(xadd(g0, g1), g2), g3), xmul(2.0f, ILB_PI));
}

template <bool FAST>
ILB_DEV float lineLightOpacity(f3 wp, f3 wn, f3 P0, f3 P1, float lightRadius, const DLine& D, f3& spherePosition, float& u, Guard& bad) {  // FBPBR.fxh:53-101
    const f3 lightLeft = xyz(mk4(D.left)), lightCenter = xyz(mk4(D.center)), ab = xyz(mk4(D.ab));
    // closestPointOnLineSegment3 DistanceFieldCommon.fxh:151-155 (exact: u places the three trace targets)
    u = saturatef(ldiv<FAST>(xdot3(xsub3(wp, P0), ab), D.ab.w, D.left.w));
    spherePosition = xadd3(P0, xscale3(ab, u));
    const f3 sphereUnormL = xsub3(spherePosition, wp);
    const float sqrSphereDistance = xdot3(sphereUnormL, sphereUnormL);
    const f3 forward = tnormalize3<FAST>(sphereUnormL, bad);  // == sphereL
    const f3 up = xcross3(lightLeft, forward);
    const f3 ru = xscale3(up, lightRadius);
    const f3 p0 = xadd3(P0, ru), p1 = xsub3(P0, ru);
    const f3 p2 = xsub3(P1, ru), p3 = xadd3(P1, ru);
    const f3 v0 = xsub3(p0, wp), v1 = xsub3(p1, wp), v2 = xsub3(p2, wp), v3 = xsub3(p3, wp);
    const float solidAngle = rectangleSolidAngle<FAST>(v0, v1, v2, v3, bad);
    float sum, forwardDotN;
    if (FAST && (wn.x == 0.0f) && (wn.y == 0.0f)) {
        // A surface that faces straight up or down (floors, the tops of raised boxes: most pixels of a top-down scene; the
        // branch is all but warp-uniform).  dot(normalize(v), n) = (0 * nx' + 0 * ny') + nz' * n.z, and a finite x times zero
        // plus a finite y times zero is a zero of either sign, which adds nothing to a non-zero product and leaves a zero
        // product a zero that saturate() maps to +0 either way: only the z component of each normalised vector is needed,
        // with the same bits (a non-finite component trips the range guard, and the fallback takes the general form below).
        auto nz = [&](f3 v) {
            const float d = xdot3(v, v);
            guardOperand(bad, d);
            return saturatef(xmul(xmul(v.z, grcp_core(gsqrt_core(d))), wn.z));
        };
        sum = xadd(xadd(xadd(xadd(nz(v0), nz(v1)), nz(v2)), nz(v3)), nz(xsub3(lightCenter, wp)));
        forwardDotN = saturatef(xmul(forward.z, wn.z));
    } else {
        sum = xadd(xadd(xadd(xadd(saturatef(xdot3(tnormalize3<FAST>(v0, bad), wn)), saturatef(xdot3(tnormalize3<FAST>(v1, bad), wn))),
                                  saturatef(xdot3(tnormalize3<FAST>(v2, bad), wn))),
                             saturatef(xdot3(tnormalize3<FAST>(v3, bad), wn))),
                        saturatef(xdot3(tnormalize3<FAST>(xsub3(lightCenter, wp), bad), wn)));
        forwardDotN = saturatef(xdot3(forward, wn));
    }
    float illuminance = xmul(xmul(solidAngle, 0.2f), sum);
    const float illuminanceSphere = xmul(xmul(ILB_PI, forwardDotN), xdiv(xmul(lightRadius, lightRadius), sqrSphereDistance));
    illuminance = xadd(illuminance, illuminanceSphere);
    return saturatef(illuminance);
}

template <int FIELD, bool INSIDE>
ILB_DEV void lineTraceMarch(const DFGeometry& g, const TraceConfig& cfg, Trace& a, Trace& b, Trace& c, float& stepsRemaining) {
    float liveness = 1.0f;
    if (!INSIDE) {
        a.tSafe = safeInsideLimit(g, a.origin, a.direction);
        b.tSafe = safeInsideLimit(g, b.origin, b.direction);
        c.tSafe = safeInsideLimit(g, c.origin, c.direction);
    }
    while (liveness > 0.0f) {
        const float stepLiveness = traceAdvanceEx<FIELD, INSIDE>(g, cfg, a) + traceAdvanceEx<FIELD, INSIDE>(g, cfg, b) + traceAdvanceEx<FIELD, INSIDE>(g, cfg, c);
        stepsRemaining -= 1.0f;
        liveness = stepsRemaining * stepLiveness;
    }
}

template <int FIELD, bool FAST>
ILB_DEV float lineConeTrace(const DFGeometry& g, const DLight& L, const DLine& D, f3 start, float u, float rampX, float rampY,
                            f3 shaded, bool enable, Guard& bad) {  // LineLightCore.fxh:17-68
    // a disabled trace returns 1 whatever its state (LineLightCore.fxh:62-67): most pixels of a line light's full-screen quad
    // lie in its far field (pre-trace opacity below 0.75 / 255) and skip the set-up of the three traces altogether
    if (!enable) return 1.0f;
    const TraceConfig cfg = makeTraceConfig(L, rampX, rampY, 1.0f);
    float stepsRemaining = cfg.stepLimit, visSum = 3.0f;
    if (L.hasField) {
        Trace a, b, c;
        const f3 delta = xyz(mk4(D.ab));
        const float offset = D.center.w;
        const f3 ta = xadd3(start, xscale3(delta, saturatef(xsub(u, offset)))), tb = xadd3(start, xscale3(delta, u));
        const f3 tc = xadd3(start, xscale3(delta, saturatef(xadd(u, offset))));
        const bool sa = traceInit<FAST>(a, shaded, ta, rampX, bad), sb = traceInit<FAST>(b, shaded, tb, rampX, bad);
        const bool sc = traceInit<FAST>(c, shaded, tc, rampX, bad);
        // t is clamped to len <= |target - shaded| for all three traces: samples stay on their segments
        const bool inside = sa && sb && sc && insideField(g, shaded) && insideField(g, ta) && insideField(g, tb) && insideField(g, tc);
#if ILB_NO_INSIDE_PATH
        lineTraceMarch<FIELD, false>(g, cfg, a, b, c, stepsRemaining);
#else
        if (inside) lineTraceMarch<FIELD, true>(g, cfg, a, b, c, stepsRemaining);
        else lineTraceMarch<FIELD, false>(g, cfg, a, b, c, stepsRemaining);
#endif
        visSum = a.vis + b.vis + c.vis;
    }
    return traceFinal(cfg, fminf(visSum / 3.0f, stepsRemaining / MAX_STEP_RAMP_WINDOW));
}

template <int FIELD, bool FAST>
ILB_DEV bool lineCore(const DFGeometry& g, const DLight& L, const DLine& D, f3 p, f3 n, f3 start, f3 end, float4 props, float4 more,
                      float& u, float& opacity, Guard& bad) {  // LineLightPixelCore :70-120
    f3 lightCenter;
    const float distanceOpacity = lineLightOpacity<FAST>(p, n, start, end, props.x, D, lightCenter, u, bad);
    const bool visible = (distanceOpacity > 0.0f) && (p.x > -9999.0f);
    if (!visible) return false;
    const float aoRadius = xmul(more.x, fmaxf(0.0f, n.z));
    const float aoOpacity = computeAO<FIELD>(g, L.hasField != 0, p, n, aoRadius, more.w, visible);
    const float preTraceOpacity = distanceOpacity * aoOpacity;
    const bool traceShadows = (props.w != 0.0f) && (preTraceOpacity >= (0.75f / 255.0f));
    const float coneOpacity = lineConeTrace<FIELD, FAST>(g, L, D, start, u, props.x, props.y, xadd3(p, xscale3(n, 1.5f)), traceShadows, bad);
    opacity = preTraceOpacity * coneOpacity;
    return true;
}

// ---- G-buffer decode (LightCommon.fxh:58-144, EnvironmentCommon.fxh:42-52) ---------------------------------
ILB_DEV float4 loadGBufferTexel(const LightingParams& P, int ix, int iy) {
    ix = min(max(ix, 0), P.gw - 1);
    iy = min(max(iy, 0), P.gh - 1);
    const size_t i = (size_t)iy * (size_t)P.gw + (size_t)ix;
    // streaming loads: a G-buffer texel is read once per pass and should not displace distance-field lines from L2
    if (P.gfmt == ILB_FORMAT_FLOAT4) return __ldcs((const float4*)P.gbuffer + i);
    const uint2 raw = __ldcs((const uint2*)P.gbuffer + i);
    const __half2 lo = *reinterpret_cast<const __half2*>(&raw.x), hi = *reinterpret_cast<const __half2*>(&raw.y);
    const float2 a = __half22float2(lo), b = __half22float2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
}

ILB_DEV Pixel decodePixel(const LightingParams& P, int px, int py) {
    Pixel r;
    r.enableShadows = true;
    r.fullbright = false;
    r.maskOk = true;
    float sx = (float)px, sy = (float)py;
    const float rsx = P.envZAndScale.z, rsy = P.envZAndScale.w;
    const float vsx = P.gbTexelAndMisc.z, vsy = P.gbTexelAndMisc.w;
    if (any2(P.gbTexelAndMisc.x, P.gbTexelAndMisc.y)) {
        float srcx = sx, srcy = sy;
        if (P.gbViewportRelative != 0.0f) {
            srcx = xadd(xdiv(srcx, vsx), P.vpx);
            srcy = xadd(xdiv(srcy, vsy), P.vpy);
        }
        const float u = xmul(xadd(srcx, 0.5f), P.gbTexelAndMisc.x), v = xmul(xadd(srcy, 0.5f), P.gbTexelAndMisc.y);
        const float4 s = loadGBufferTexel(P, (int)floorf(xmul(u, (float)P.gw)), (int)floorf(xmul(v, (float)P.gh)));
        if (P.stencil) {  // UpdateMaskFromGBuffer, GBufferMask.fx:26-44
            const float minW = -fabsf(P.envZAndScale.y) - 1.0f, maxW = -fabsf(P.envZAndScale.x) - 1.0f;
            r.maskOk = !((s.w >= 9999.0f) || (s.w < minW) || ((s.w < 0.0f) && (s.w > maxW)));
        }
        const float relativeY = s.z;
        float worldZ = s.w;
        if (worldZ < 0.0f) {
            worldZ = xadd(worldZ, 1.0f);
            worldZ = -worldZ;
            r.enableShadows = false;
        } else if (worldZ >= 9999.0f) {
            worldZ = 0.0f;
            r.enableShadows = false;
            r.fullbright = true;
        }
        worldZ = xsub(xmul(worldZ, 1024.0f), 1024.0f);
        sx = xdiv(sx, rsx);
        sy = xdiv(sy, rsy);
        r.camera = mk3(sx, sy, xadd(P.envZAndScale.y, 0.01f));
        r.pos = mk3(xadd(xdiv(sx, vsx), P.vpx), xadd(xdiv(xadd(sy, relativeY), vsy), P.vpy), worldZ);
        if (any2(s.x, s.y)) {
            const float ax = xsub(xmul(s.x, 2.0f), 1.0f), ay = xsub(xmul(s.y, 2.0f), 1.0f);
            float sn, cs;
            dm_sincosf(xmul(ax, ILB_PI), &sn, &cs);
            const float phx = xsqrt(xsub(1.0f, xmul(ay, ay)));
            r.normal = mk3(xmul(cs, phx), xmul(sn, phx), ay);
        } else {
            r.normal = mk3(0.0f);
        }
    } else {
        sx = xdiv(sx, rsx);
        sy = xdiv(sy, rsy);
        r.camera = mk3(sx, sy, xadd(P.envZAndScale.y, 0.01f));
        r.pos = mk3(xadd(xdiv(sx, vsx), P.vpx), xadd(xdiv(sy, vsy), P.vpy), P.envZAndScale.x);
        r.normal = mk3(0.0f, 0.0f, 1.0f);
    }
    return r;
}

// Rasterised coverage of the light's quad at this pixel centre (world space).
//   sphere: 12-vertex cross of FillSphereBuffer (LightingRenderer.cs:636-656) through SphereLightVertexShader
//           (SphereLightCore.fxh:13-56): covX = X(0), X(1/7), X(6/7), X(1); covY likewise (2.5D shift applied to w<0.5)
//   directional / line: one rectangle covX = (x0, x1), covY = (y0, y1) (DirectionalLight.fx:19-37, LineLightCore.fxh:122-173)
ILB_DEV bool coverage(const DLight& L, float wx, float wy) {
    if ((L.type & ILB_LIGHT_TYPE_MASK) == ILB_LIGHT_SPHERE) {
        const float4 X = L.covX, Y = L.covY;
        const bool a = (wx >= X.y) && (wx <= X.z) && (wy >= Y.x) && (wy <= Y.w);
        const bool b = (wx >= X.z) && (wx <= X.w) && (wy >= Y.y) && (wy <= Y.z);
        const bool c = (wx >= X.x) && (wx <= X.y) && (wy >= Y.y) && (wy <= Y.z);
        return a || b || c;
    }
    return (wx >= L.covX.x) && (wx <= L.covX.y) && (wy >= L.covY.x) && (wy <= L.covY.y);
}

ILB_DEV bool shadowFilterRejects(float filter, bool enableShadows) {  // checkShadowFilter LightCommon.fxh:146-152
    if (filter < 0.0f) return false;
    return (filter > 0.5f) != enableShadows;
}

ILB_DEV DLight loadLight(const DLight* lights, int i) {
    DLight L;
    const float4* src = reinterpret_cast<const float4*>(lights + i);
    float4* dst = reinterpret_cast<float4*>(&L);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(DLight) / 16); k++) dst[k] = __ldg(src + k);
    return L;
}
ILB_DEV DLine loadLine(const DLine* lines, int i) {
    DLine D;
    const float4* src = reinterpret_cast<const float4*>(lines + i);
    D.left = __ldg(src); D.center = __ldg(src + 1); D.ab = __ldg(src + 2);
    return D;
}

// One light at one pixel; returns false when the reference fragment would be discarded.
// TYPES: bit mask of ilb_light_type values this instantiation can meet (other branches are compiled out)
// CL: the light records of this frame are in the constant bank (c_lights / c_lines) and L refers into it
template <int FIELD, bool FAST, int TYPES, bool CL>
ILB_DEV bool shadeLight(const DFGeometry& df, float lightOcclusion, const DLight& L, const DLight* lights, const DLine* lines, const RampTex* ramps,
                        int lightIndex, const Pixel& px, f3& rgb, Guard& bad) {
    const float es = px.enableShadows ? 1.0f : 0.0f;
    if ((TYPES & ILB_LIGHT_SPHERE) && ((L.type & ILB_LIGHT_TYPE_MASK) == ILB_LIGHT_SPHERE || L.type == ILB_LIGHT_PARTICLE_BIT)) {  // SphereLightPixelShader SphereLight.fx:7-46, ParticleLightPixelShader ParticleLight.fx:84-118
        if (px.fullbright || shadowFilterRejects(L.evenMore.x, px.enableShadows)) return false;
        float4 props = L.props;
        props.w *= es;
        const f3 center = mk3(L.pos1.x, L.pos1.y, L.pos1.z);
        float opacity, preTrace;
        if (!sphereCore<FIELD, FAST>(df, L, lightOcclusion, px.pos, px.normal, center, props, L.more, opacity, preTrace, bad)) return false;
        const float4 color = L.color1, spec = L.color2;
        // (instantiations without the ramp bit in TYPES serve the frames that have no ramp-textured light: the branch and the
        // registers it keeps alive cost the sphere / directional pass 0.6 % when compiled in)
        if ((TYPES & ILB_LIGHT_RAMP_BIT) && (L.type & ILB_LIGHT_RAMP_BIT)) {  // uniform per light: SphereLightWithDistanceRampPixelShader (SphereLight.fx:48-87)
            // opacity3 = RampTexture(preTraceOpacity, angle).rgb * coneOpacity, coneOpacity = opacity / preTraceOpacity up to rounding;
            // a pixel that was not discarded has preTraceOpacity > 0 (distanceOpacity > 0 and the AO factor >= 1 - AO opacity).
            // The ramp's index, offset and rate are re-read from the light record (L1-resident) so that they do not occupy
            // registers across the cone trace of every ordinary sphere light.
            const float4 em = __ldg(&lights[lightIndex].evenMore);
            const float4 r = rampLookup(ramps, (int)em.y, preTrace, px.pos.x - center.x, px.pos.y - center.y, em.z, em.w);
            const float cone = (preTrace > 0.0f) ? opacity / preTrace : 0.0f;
            f3 o3 = mk3(r.x, r.y, r.z) * cone;
            rgb = mk3(color.x, color.y, color.z) * color.w * o3;
            if (any3(mk3(spec.x, spec.y, spec.z))) {
                const f3 lightDirection = px.pos - center;
                const f3 h = normalize3(normalize3(px.camera - px.pos) - lightDirection);
                rgb = rgb + (mk3(spec.x, spec.y, spec.z) * powf(saturatef(dot3(h, px.normal)), spec.w) * o3);
            }
            return true;
        }
        // colour products and the accumulation are individually rounded (the oracle's operations): every instantiation of the
        // kernel -- atlas / planes, constant bank or not, one pass or two -- then lands on the same bits
        rgb = xscale3(xscale3(mk3(color.x, color.y, color.z), color.w), opacity);
        if (any3(mk3(spec.x, spec.y, spec.z)) || L.type == ILB_LIGHT_PARTICLE_BIT) {  // CalcSphereLightSpecularity LightCommon.fxh:212-222
            const f3 lightDirection = px.pos - center;
            const f3 h = normalize3(normalize3(px.camera - px.pos) - lightDirection);
            const float specularity = powf(saturatef(dot3(h, px.normal)), spec.w);
            rgb = xadd3(rgb, xscale3(xscale3(mk3(spec.x, spec.y, spec.z), specularity), opacity));
        }
        return true;
    } else if ((TYPES & ILB_LIGHT_DIRECTIONAL) && L.type == ILB_LIGHT_DIRECTIONAL) {  // DirectionalLightPixelShader DirectionalLight.fx:95-127
        if (px.fullbright || shadowFilterRejects(L.evenMore.x, px.enableShadows)) return false;
        float4 props = L.props;
        props.x *= es;
        float opacity;
        if (!directionalCore<FIELD, FAST>(df, L, px.pos, px.normal, L.color2, props, L.more, opacity, bad)) return false;
        rgb = xscale3(xscale3(mk3(L.color1.x, L.color1.y, L.color1.z), L.color1.w), opacity);
        return true;
    } else if (TYPES & ILB_LIGHT_LINE) {  // LineLightPixelShader LineLight.fx:7-42
        if (px.fullbright) return false;
        float4 props = L.props;
        props.w *= es;
        float u, opacity;
        DLine Dg;
        if (!CL) Dg = loadLine(lines, lightIndex);
        const DLine& D = CL ? c_lines[lightIndex] : Dg;
        if (!lineCore<FIELD, FAST>(df, L, D, px.pos, px.normal, mk3(L.pos1.x, L.pos1.y, L.pos1.z), mk3(L.pos2.x, L.pos2.y, L.pos2.z), props,
                                   L.more, u, opacity, bad))
            return false;
        const f4 color = xlerp4(mk4(L.color1), mk4(L.color2), u);
        rgb = xscale3(xscale3(mk3(color.x, color.y, color.z), color.w), opacity);
        return true;
    }
    return false;
}

// The IEEE re-evaluation of one light-pixel whose fast evaluation tripped a range guard (an operand of a square root
// or reciprocal outside the fast window: zero-length vectors, parallel normals, denormal or huge values).  Out of
// line: it is never on the hot path and must not cost the hot path registers.  Returns rgb, w = 1 when lit.
template <int FIELD, int TYPES>
__device__ __noinline__ float4 shadeLightExact(const DFGeometry* df, float lightOcclusion, const DLight* lights, const DLine* lines,
                                               const RampTex* ramps, int lightIndex, float4 posShadows, float4 normalFullbright, float4 camera) {
    Pixel px;
    px.pos = mk3(posShadows.x, posShadows.y, posShadows.z);
    px.normal = mk3(normalFullbright.x, normalFullbright.y, normalFullbright.z);
    px.camera = mk3(camera.x, camera.y, camera.z);
    px.enableShadows = posShadows.w != 0.0f;
    px.fullbright = normalFullbright.w != 0.0f;
    px.maskOk = true;
    const DLight L = loadLight(lights, lightIndex);
    f3 rgb = mk3(0.0f);
    Guard bad = guardInit();
    const bool lit = shadeLight<FIELD, false, TYPES, false>(*df, lightOcclusion, L, lights, lines, ramps, lightIndex, px, rgb, bad);
#if ILB_BREAK_FALLBACK  // test hook: proves that a test reaches this path (tests/README: degenerate-geometry tests must fail with it)
    rgb.x += 1.0f;
#endif
    return make_float4(rgb.x, rgb.y, rgb.z, lit ? 1.0f : 0.0f);
}

// fast evaluation + fallback
template <int FIELD, int TYPES, bool CL>
ILB_DEV bool shadeLightGuarded(const DFGeometry& df, float lightOcclusion, const DLight& L, const DLight* lights, const DLine* lines,
                               const RampTex* ramps, int lightIndex, const Pixel& px, f3& rgb) {
#if ILB_NO_FAST_GUARD
    Guard bad = guardInit();
    return shadeLight<FIELD, false, TYPES, CL>(df, lightOcclusion, L, lights, lines, ramps, lightIndex, px, rgb, bad);
#else
    Guard bad = guardInit();
    bool lit = shadeLight<FIELD, true, TYPES, CL>(df, lightOcclusion, L, lights, lines, ramps, lightIndex, px, rgb, bad);
    if (guardTripped(bad)) {
        const float4 r = shadeLightExact<FIELD, TYPES>(&df, lightOcclusion, lights, lines, ramps, lightIndex,
                                                make_float4(px.pos.x, px.pos.y, px.pos.z, px.enableShadows ? 1.0f : 0.0f),
                                                make_float4(px.normal.x, px.normal.y, px.normal.z, px.fullbright ? 1.0f : 0.0f),
                                                make_float4(px.camera.x, px.camera.y, px.camera.z, 0.0f));
        rgb = mk3(r.x, r.y, r.z);
        lit = r.w != 0.0f;
    }
    return lit;
#endif
}

ILB_DEV void storeTexel(const LightingParams& P, size_t index, float r, float g, float b, float a) {
    if (P.out_format == ILB_FORMAT_FLOAT4) {
        const float4 v = make_float4(r, g, b, a);
        for (int o = 0; o < P.nouts; o++) reinterpret_cast<float4*>(P.outs[o])[index] = v;
    } else if (P.out_format == ILB_FORMAT_HALF4) {  // HalfVector4 lightmap (LightingRenderer.cs:477-479)
        const __half2 lo = __floats2half2_rn(r, g), hi = __floats2half2_rn(b, a);
        uint2 v;
        v.x = *reinterpret_cast<const uint32_t*>(&lo);
        v.y = *reinterpret_cast<const uint32_t*>(&hi);
        for (int o = 0; o < P.nouts; o++) reinterpret_cast<uint2*>(P.outs[o])[index] = v;
    } else {  // SurfaceFormat.Color
        const uint32_t R = (uint32_t)(saturatef(r) * 255.0f + 0.5f), G = (uint32_t)(saturatef(g) * 255.0f + 0.5f);
        const uint32_t B = (uint32_t)(saturatef(b) * 255.0f + 0.5f), A = (uint32_t)(saturatef(a) * 255.0f + 0.5f);
        const uint32_t v = R | (G << 8) | (B << 16) | (A << 24);
        for (int o = 0; o < P.nouts; o++) reinterpret_cast<uint32_t*>(P.outs[o])[index] = v;
    }
}

ILB_DEV float warpMin(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}
ILB_DEV float warpMax(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}

// Resident CTAs per SM (register budget = 65536 / (256 * n)): the trace is latency-bound (dependent IEEE ops and
// gathers through L1/L2), so more resident warps pay even at the price of a few spills.  Line lights (three interleaved
// traces) need the larger budget; sphere / directional lights run in their own instantiation with more warps.
#ifndef ILB_LIGHT_MINBLOCKS
#define ILB_LIGHT_MINBLOCKS 3
#endif
#ifndef ILB_LIGHT_MINBLOCKS_NOLINE
#define ILB_LIGHT_MINBLOCKS_NOLINE 5
#endif
// TYPES: the light types this pass shades (lights of other types are dropped by the tile culling).  A frame is one
// pass over all types, or -- when line lights and other lights are both present -- a line-light pass and a sphere +
// directional pass with their own register budgets (see lightingLaunchRows for how the two are scheduled).
struct TileSmem {
    float box[TILE_WARPS][6];
    int warpCount[TILE_WARPS];
    uint16_t list[TILE_THREADS];
    uint8_t mask[TILE_THREADS];  // per listed light: bit w set when the light can reach warp w's 8x4 pixel block
    int listCount;
    unsigned next, arrival;      // persistent CTAs: the tile taken from the queue, this pass's arrival rank at the tile
};

template <int FIELD, int TYPES, bool CL>
ILB_DEV void shadeTile(const LightingParams& P, unsigned tile, TileSmem& S) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // a warp covers an 8x4 pixel patch: neighbouring lanes trace neighbouring rays (coherent DF footprints)
    int tileX, tileY;
    if (ILB_SWIZZLE > 1) {
        const int band = tile / (ILB_SWIZZLE * P.tiles_x), inband = tile % (ILB_SWIZZLE * P.tiles_x);
        const int rowsInBand = min(ILB_SWIZZLE, P.tiles_y - band * ILB_SWIZZLE);
        tileX = inband / rowsInBand;
        tileY = band * ILB_SWIZZLE + inband % rowsInBand;
    } else {
        tileX = tile % P.tiles_x;
        tileY = tile / P.tiles_x;
    }
    const int px = tileX * TILE_W + (warp & 1) * 8 + (lane & 7);
    const int py = P.row_begin + tileY * TILE_H + (warp >> 1) * 4 + (lane >> 3);
    const bool valid = (px < P.width) && (py < P.row_end);

    Pixel pix;
    if (valid) pix = decodePixel(P, px, py);
    const bool shade = valid && pix.maskOk;

    // tile AABB of shaded world positions (warp shuffle + 8-entry shared reduction)
    const float BIG = 3.0e38f;
    float bx0 = shade ? pix.pos.x : BIG, bx1 = shade ? pix.pos.x : -BIG;
    float by0 = shade ? pix.pos.y : BIG, by1 = shade ? pix.pos.y : -BIG;
    float bz0 = shade ? pix.pos.z : BIG, bz1 = shade ? pix.pos.z : -BIG;
    bx0 = warpMin(bx0); bx1 = warpMax(bx1);
    by0 = warpMin(by0); by1 = warpMax(by1);
    bz0 = warpMin(bz0); bz1 = warpMax(bz1);
    if (lane == 0) {
        S.box[warp][0] = bx0; S.box[warp][1] = bx1; S.box[warp][2] = by0;
        S.box[warp][3] = by1; S.box[warp][4] = bz0; S.box[warp][5] = bz1;
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < TILE_WARPS; w++) {
        bx0 = fminf(bx0, S.box[w][0]); bx1 = fmaxf(bx1, S.box[w][1]);
        by0 = fminf(by0, S.box[w][2]); by1 = fmaxf(by1, S.box[w][3]);
        bz0 = fminf(bz0, S.box[w][4]); bz1 = fmaxf(bz1, S.box[w][5]);
    }
    const int tx0 = tileX * TILE_W, tx1 = tx0 + TILE_W - 1;
    const int ty0 = P.row_begin + tileY * TILE_H, ty1 = ty0 + TILE_H - 1;

    // pixel centre in world space for the coverage test (inverse of the light vertex shaders' transform)
    const float sxs = xmul(P.gbTexelAndMisc.z, P.envZAndScale.z), sys = xmul(P.gbTexelAndMisc.w, P.envZAndScale.w);
    const float wx = xadd(xdiv(xadd((float)px, 0.5f), sxs), P.vpx), wy = xadd(xdiv(xadd((float)py, 0.5f), sys), P.vpy);

    float accR = P.clear.x, accG = P.clear.y, accB = P.clear.z, accA = P.clear.w;
    const size_t scratchIndex = (size_t)(py - P.row_begin) * (size_t)P.width + (size_t)px;

    for (int base = 0; base < P.nlights; base += TILE_THREADS) {
        // ---- cull: thread t tests light base+t against the tile
        const int li = base + tid;
        bool keep = false;
        unsigned warpMask = 0xFFu;
        if (li < P.nlights) {
            const DLight* L = P.lights + li;
            const int4 r = __ldg(reinterpret_cast<const int4*>(&L->px0));
            const int type = __ldg(&L->type);
            keep = ((type & TYPES) != 0) && (r.x <= tx1) && (r.z >= tx0) && (r.y <= ty1) && (r.w >= ty0) && (bx0 <= bx1);
            if ((TYPES & ILB_LIGHT_SPHERE) && keep && ((type & ILB_LIGHT_TYPE_MASK) == ILB_LIGHT_SPHERE || type == ILB_LIGHT_PARTICLE_BIT)) {
                // sphere lights reach radius + rampLength (radius + 1 in RampMode None): reject the tile when the
                // closest point of its world AABB is farther (1 px of slack covers fp rounding)
                const float4 c = __ldg(&L->pos1), pr = __ldg(&L->props), mo = __ldg(&L->more);
                const float dx = fmaxf(fmaxf(bx0 - c.x, c.x - bx1), 0.0f);
                const float dy = fmaxf(fmaxf(by0 - c.y, c.y - by1), 0.0f) * fabsf(mo.z);
                const float dz = fmaxf(fmaxf(bz0 - c.z, c.z - bz1), 0.0f);
                const float reach = pr.x + fmaxf(pr.y, 1.0f) + 1.0f;
                keep = (dx * dx + dy * dy + dz * dz) <= reach * reach;
#if !ILB_NO_WARP_CULL
                if (keep) {  // second level: the same two tests against every warp's own 8x4 block (pixel rectangle, world AABB)
                    warpMask = 0u;
#pragma unroll
                    for (int w = 0; w < TILE_WARPS; w++) {
                        const int wx0 = tx0 + (w & 1) * 8, wy0 = ty0 + (w >> 1) * 4;
                        const float ex = fmaxf(fmaxf(S.box[w][0] - c.x, c.x - S.box[w][1]), 0.0f);
                        const float ey = fmaxf(fmaxf(S.box[w][2] - c.y, c.y - S.box[w][3]), 0.0f) * fabsf(mo.z);
                        const float ez = fmaxf(fmaxf(S.box[w][4] - c.z, c.z - S.box[w][5]), 0.0f);
                        const bool hit = (r.x <= wx0 + 7) && (r.z >= wx0) && (r.y <= wy0 + 3) && (r.w >= wy0) && (S.box[w][0] <= S.box[w][1]) &&
                                         ((ex * ex + ey * ey + ez * ez) <= reach * reach);
                        warpMask |= hit ? (1u << w) : 0u;
                    }
                    keep = warpMask != 0u;
                }
#endif
            }
        }
        // ---- ordered compaction (draw order is kept so accumulation order matches the reference)
        const unsigned ballot = __ballot_sync(0xFFFFFFFFu, keep);
        if (lane == 0) S.warpCount[warp] = __popc(ballot);
        __syncthreads();
        int offset = 0, total = 0;
#pragma unroll
        for (int w = 0; w < TILE_WARPS; w++) {
            const int c = S.warpCount[w];
            if (w < warp) offset += c;
            total += c;
        }
        if (keep) {
            const int slot = offset + __popc(ballot & ((1u << lane) - 1u));
            S.list[slot] = (uint16_t)tid;
            S.mask[slot] = (uint8_t)warpMask;
        }
        if (tid == 0) S.listCount = total;
        __syncthreads();

        // ---- shade
        const int n = S.listCount;
        for (int k = 0; k < n; k++) {
            if ((TYPES & ILB_LIGHT_SPHERE) && !((S.mask[k] >> warp) & 1u)) continue;  // warp-uniform: this block is out of the light's reach
            const int lightIndex = base + (int)S.list[k];
            DLight Lg;
            if (!CL) Lg = loadLight(P.lights, lightIndex);
            const DLight& L = CL ? c_lights[lightIndex] : Lg;
            if (shade && coverage(L, wx, wy)) {
                f3 rgb;
                if (shadeLightGuarded<FIELD, TYPES, CL>(P.df, P.envZToY.z, L, P.lights, P.lines, P.ramps, lightIndex, pix, rgb)) {
                    // BlendState.Additive with PS alpha 1: rgb += src.rgb, a += 1 (LightingRenderer.cs:206)
                    accR = xadd(accR, rgb.x); accG = xadd(accG, rgb.y); accB = xadd(accB, rgb.z); accA += 1.0f;
                }
            }
        }
        if (base + TILE_THREADS < P.nlights) __syncthreads();  // the list is rebuilt only when another round follows
    }

    const size_t outIndex = (size_t)(py - P.out_row_base) * (size_t)P.width + (size_t)px;
    if (P.tile_done) {
        // Concurrent passes: each pass leaves its fp32 sums of the tile in its own scratch buffer; the pass that arrives
        // second at the tile adds the two (line sums + other sums, a fixed operand order whoever arrives last) and stores
        // the lightmap texel.  Release: scratch stores -> fence -> arrival counter; acquire: counter -> fence -> L2 loads.
        if (valid) P.accum_out[scratchIndex] = make_float4(accR, accG, accB, accA);
        __threadfence();
        __syncthreads();
        if (tid == 0) S.arrival = atomicAdd(P.tile_done + tile, 1u);
        __syncthreads();
        if (S.arrival != 0u) {
            __threadfence();
            if (valid) {
                const float4 o = __ldcg(P.accum_in + scratchIndex);
                storeTexel(P, outIndex, accR + o.x, accG + o.y, accB + o.z, accA + o.w);
            }
        }
        return;
    }
    if (P.accum_in) {
        // second pass of a split frame: this pass summed its own lights from zero and now adds the first pass's sums (line sums +
        // these sums, the operand order of the concurrent mode as well).  Under programmatic dependent launch the CTAs of this
        // pass start while the first pass's last wave is still running and meet it here.
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (valid) {
            const float4 a = __ldcs(P.accum_in + scratchIndex);   // written once, read once: streaming
            accR = a.x + accR; accG = a.y + accG; accB = a.z + accB; accA = a.w + accA;
        }
    }
    if (!valid) return;
    if (P.accum_out) __stcs(P.accum_out + scratchIndex, make_float4(accR, accG, accB, accA));
    else storeTexel(P, outIndex, accR, accG, accB, accA);
}

// resident CTAs per SM for this pass's register budget (the budgets above are stated for 256-thread CTAs)
#if defined(ILB_LIGHT_CTAS_LINE) && defined(ILB_LIGHT_CTAS_NOLINE)   // dev knob: resident CTAs per SM stated directly (for other tile heights)
#define ILB_LIGHT_CTAS(TYPES) (((TYPES) & ILB_LIGHT_LINE) ? ILB_LIGHT_CTAS_LINE : ILB_LIGHT_CTAS_NOLINE)
#else
#define ILB_LIGHT_CTAS(TYPES) ((((TYPES) & ILB_LIGHT_LINE) ? ILB_LIGHT_MINBLOCKS : ILB_LIGHT_MINBLOCKS_NOLINE) * (256 / TILE_THREADS))
#endif

template <int FIELD, int TYPES, bool CL>
__global__ void __launch_bounds__(TILE_THREADS, ILB_LIGHT_CTAS(TYPES))
light_accumulate_kernel(const __grid_constant__ LightingParams P) {
    __shared__ TileSmem S;
    // first pass of a split frame: the second pass may be scheduled as soon as every CTA of this grid has started, i.e. into
    // the idle SM slots of this grid's last wave (it waits for this grid's results only at its very end, see shadeTile)
    if (P.accum_out && !P.tile_done) asm volatile("griddepcontrol.launch_dependents;");
    shadeTile<FIELD, TYPES, CL>(P, P.tile_order ? __ldg(P.tile_order + blockIdx.x) : blockIdx.x, S);
}

// Persistent form: a fixed number of resident CTAs per SM takes tiles from a queue (one atomic counter per pass), so that
// the line-light pass (issue-bound, 80 registers) and the sphere + directional pass (latency-bound on its gathers, 48
// registers) can be co-resident on every SM -- two CTAs of each fill the register file exactly -- and run side by side
// instead of back to back.  "Helper" grids of the same kernels, launched behind the main grids, become resident when one
// pass runs out of tiles and its CTAs exit, so the surviving pass gets its full occupancy back for the tail.
template <int FIELD, int TYPES>
__global__ void __launch_bounds__(TILE_THREADS, ILB_LIGHT_CTAS(TYPES))
light_accumulate_persistent_kernel(const __grid_constant__ LightingParams P) {
    __shared__ TileSmem S;
    const unsigned ntiles = (unsigned)P.tiles_x * (unsigned)P.tiles_y;
    for (;;) {
        __syncthreads();  // the previous tile's shared state is no longer read
        if (threadIdx.x == 0) S.next = atomicAdd(P.tile_counter, 1u);
        __syncthreads();
        const unsigned tile = S.next;
        if (tile >= ntiles) return;
        shadeTile<FIELD, TYPES, false>(P, tile, S);
    }
}

// ---- light probes (L11): SphereLightProbe.fx:19-44, DirectionalLight.fx:163-190, LineLightProbe.fx:23-48 ----------
struct ProbeParams {
    DFGeometry df;
    float lightOcclusion;
    const DLight* lights;
    const DLine* lines;
    int nlights;
    const float4* positions;
    const float4* normals;
    int nprobes;
    int out_format;
    void* out;
};

// One CTA per probe, one thread per light (rounds of PROBE_THREADS lights): the light-probe responses -- each with its own
// cone trace -- are evaluated in parallel and summed by one thread in draw order, so the sums are the ones a serial loop
// over the lights gives (the reference draws one additive pass per light into the probe target).
constexpr int PROBE_THREADS = 128;
template <int FIELD>
__global__ void __launch_bounds__(PROBE_THREADS) probe_accumulate_kernel(const __grid_constant__ ProbeParams P) {
    __shared__ float4 s_contribution[PROBE_THREADS];  // rgb, w = 1 when the light's fragment was not discarded
    const int i = blockIdx.x, tid = threadIdx.x;
    const float4 ps = __ldg(P.positions + i), ns = __ldg(P.normals + i);  // sampleLightProbeBuffer LightCommon.fxh:233-254
    float accR = 0.0f, accG = 0.0f, accB = 0.0f, accA = 0.0f;
    if (ps.w > 0.0f) {  // uniform over the CTA
        const f3 p = mk3(ps.x, ps.y, ps.z), n = mk3(ns.x, ns.y, ns.z);
        for (int base = 0; base < P.nlights; base += PROBE_THREADS) {
            const int k = base + tid;
            float4 contribution = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (k < P.nlights) {
                const DLight L = loadLight(P.lights, k);
                float4 props = L.props, more = L.more;
                more.x = 0.0f;
                more.w = 0.0f;
                float core;
                bool lit;
                Guard bad = guardInit();  // probes are few: plain IEEE x-ops (FAST = false), no guard
                if (L.type == ILB_LIGHT_DIRECTIONAL) {
                    props.x *= ns.w;
                    lit = directionalCore<FIELD, false>(P.df, L, p, n, L.color2, props, more, core, bad);
                } else {  // sphere, and line lights shaded as spheres at LightPosition1 (reference quirk, LineLightProbe.fx:4)
                    props.w *= ns.w;
                    float preTrace;
                    lit = sphereCore<FIELD, false>(P.df, L, P.lightOcclusion, p, n, mk3(L.pos1.x, L.pos1.y, L.pos1.z), props, more, core, preTrace, bad);
                }
                if (lit) {
                    const float opacity = ps.w * core;
                    const f3 rgb = mk3(L.color1.x, L.color1.y, L.color1.z) * L.color1.w * opacity;
                    contribution = make_float4(rgb.x, rgb.y, rgb.z, 1.0f);
                }
            }
            s_contribution[tid] = contribution;
            __syncthreads();
            if (tid == 0) {
                const int n_round = min(PROBE_THREADS, P.nlights - base);
                for (int j = 0; j < n_round; j++) {
                    const float4 c = s_contribution[j];
                    if (c.w != 0.0f) { accR += c.x; accG += c.y; accB += c.z; accA += 1.0f; }
                }
            }
            __syncthreads();
        }
    }
    if (tid != 0) return;
    if (P.out_format == ILB_FORMAT_FLOAT4) {
        reinterpret_cast<float4*>(P.out)[i] = make_float4(accR, accG, accB, accA);
    } else {
        const __half2 lo = __floats2half2_rn(accR, accG), hi = __floats2half2_rn(accB, accA);
        uint2 v;
        v.x = *reinterpret_cast<const uint32_t*>(&lo);
        v.y = *reinterpret_cast<const uint32_t*>(&hi);
        reinterpret_cast<uint2*>(P.out)[i] = v;
    }
}

// ---- host: flatten (batch, LightVertex) into DLight ---------------------------------------------------------
inline float4 h4(const ilb_float4& v) { return make_float4(v.x, v.y, v.z, v.w); }
inline float hlerp(float a, float b, float t) { return a + t * (b - a); }

int flattenLights(ilb_ctx* ctx, const ilb_df* df, const ilb_lighting_frame* f, const ilb_light_batch* batches, int batch_count,
                  const ilb_light_vertex* verts, int vertex_count, std::vector<DLight>& out, std::vector<DLine>& lines,
                  const ilb_df_uniforms** geometry, bool allowRamps = true) {
    *geometry = nullptr;
    // correctly rounded reciprocal of a uniform divisor for udiv(); 0 selects udiv's IEEE division
    auto rcp = [](float y) { const float a = std::fabs(y); return (a >= 1.0e-30f && a <= 1.0e30f) ? 1.0f / y : 0.0f; };
    const float invZToY = f->EnvironmentZToY.y, zToY = f->EnvironmentZToY.x;
    const float sxs = f->GBufferTexelSizeAndMisc.z * f->EnvironmentZAndScale.z;
    const float sys = f->GBufferTexelSizeAndMisc.w * f->EnvironmentZAndScale.w;
    if (!(sxs > 0.0f) || !(sys > 0.0f)) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "viewport scale * render scale must be > 0");
    for (int b = 0; b < batch_count; b++) {
        const ilb_light_batch& B = batches[b];
        if (B.light_type != ILB_LIGHT_SPHERE && B.light_type != ILB_LIGHT_DIRECTIONAL && B.light_type != ILB_LIGHT_LINE)
            return ilb_fail(ctx, ILB_ERR_UNSUPPORTED, "light type %d is outside the hot-path scope (sphere=1, directional=2, line=4)", B.light_type);
        if (B.first_vertex < 0 || B.vertex_count < 0 || B.first_vertex + B.vertex_count > vertex_count)
            return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "batch %d vertex range [%d,+%d) outside [0,%d)", b, B.first_vertex, B.vertex_count, vertex_count);
        if (B.ramp_texture != 0) {
            if (B.light_type != ILB_LIGHT_SPHERE)
                return ilb_fail(ctx, ILB_ERR_UNSUPPORTED, "batch %d: ramp textures are only supported on sphere lights (SphereLightWithDistanceRamp)", b);
            if (B.ramp_texture < 0 || (size_t)B.ramp_texture > ctx->ramps.size() || !ctx->ramps[(size_t)B.ramp_texture - 1].texels)
                return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "batch %d: unknown ramp texture %d", b, B.ramp_texture);
            if (!allowRamps) return ilb_fail(ctx, ILB_ERR_UNSUPPORTED, "batch %d: ramp textures on light probes are outside the hot-path scope", b);
        }
        const bool hasField = (df != nullptr) && (B.df.Extent.x > 0.0f);
        if (hasField) {
            if (*geometry) {
                const ilb_df_uniforms& g = **geometry;
                // one DistanceField per renderer in the reference: geometry members must agree across batches
                if (memcmp(&g.TextureSliceAndTexelSize, &B.df.TextureSliceAndTexelSize, sizeof(ilb_float4)) ||
                    memcmp(&g.TextureSliceCount, &B.df.TextureSliceCount, sizeof(ilb_float4)) ||
                    memcmp(&g.Extent, &B.df.Extent, sizeof(ilb_float4)) || g.ConeAndMisc.y != B.df.ConeAndMisc.y ||
                    g.ConeAndMisc.w != B.df.ConeAndMisc.w || g.StepAndMisc2.w != B.df.StepAndMisc2.w ||
                    g.Packed1.x != B.df.Packed1.x || g.Packed1.y != B.df.Packed1.y || g.Packed1.z != B.df.Packed1.z)
                    return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "batch %d: distance-field geometry differs from earlier batches", b);
            } else {
                *geometry = &B.df;
            }
        }
        for (int i = 0; i < B.vertex_count; i++) {
            const ilb_light_vertex& v = verts[B.first_vertex + i];
            DLight L;
            memset(&L, 0, sizeof(L));
            L.pos1 = h4(v.LightPosition1); L.pos2 = h4(v.LightPosition2);
            L.props = h4(v.LightProperties); L.more = h4(v.MoreLightProperties); L.evenMore = h4(v.EvenMoreLightProperties);
            L.color1 = h4(v.Color1); L.color2 = h4(v.Color2);
            L.quality = make_float4(B.df.ConeAndMisc.x, B.df.ConeAndMisc.z, B.df.StepAndMisc2.x, B.df.Packed1.w);
            L.longStep = B.df.StepAndMisc2.z;
            L.hasField = hasField ? 1 : 0;
            L.type = B.light_type | (B.ramp_texture ? ILB_LIGHT_RAMP_BIT : 0);
            L.rcpRamp = rcp(v.LightProperties.y);
            // EvenMoreLightProperties.y is always 0 on the host side (LightingRenderer.cs:1212-1214): the device record carries the
            // batch's ramp texture there
            L.evenMore.y = (float)B.ramp_texture;
            DLine D;
            memset(&D, 0, sizeof(D));
            if (B.light_type == ILB_LIGHT_LINE) {
                // the pixel-independent part of computeLineLightOpacity / lineConeTrace (FBPBR.fxh:53-60, LineLightCore.fxh:24-31),
                // same fp32 operations in the same order as the shaders (host code is built with -ffp-contract=off)
                const float P0[3] = {v.LightPosition1.x, v.LightPosition1.y, v.LightPosition1.z};
                const float P1[3] = {v.LightPosition2.x, v.LightPosition2.y, v.LightPosition2.z};
                const float ab[3] = {P1[0] - P0[0], P1[1] - P0[1], P1[2] - P0[2]};
                const float abab = (ab[0] * ab[0] + ab[1] * ab[1]) + ab[2] * ab[2];
                const float inv = (abab == 0.0f) ? 0.0f : 1.0f / std::sqrt(abab);   // normalize(): zero in, zero out
                const float deltaLength = std::sqrt(abab);
                const float ratio = (v.LightProperties.x + 1.0f) / deltaLength;
                const float sat = std::fmin(std::fmax(ratio, 0.0f), 1.0f);          // NaN -> 0 like saturatef
                D.left = make_float4(ab[0] * inv, ab[1] * inv, ab[2] * inv, rcp(abab));
                D.center = make_float4(P0[0] + 0.5f * ab[0], P0[1] + 0.5f * ab[1], P0[2] + 0.5f * ab[2], std::fmax(sat, 0.03f));
                D.ab = make_float4(ab[0], ab[1], ab[2], abab);
            }
            lines.push_back(D);
            float bx0, bx1, by0, by1;  // world-space bounds of the rasterised quad
            if (B.light_type == ILB_LIGHT_SPHERE) {  // SphereLightVertexShader SphereLightCore.fxh:13-56
                const float radius = v.LightProperties.x + v.LightProperties.y + 1;
                const float deltaY = (radius) - (radius / v.MoreLightProperties.z);
                const float rx = radius, ry = radius - (deltaY / 2.0f);
                const float tlx = v.LightPosition1.x - rx, brx = v.LightPosition1.x + rx;
                const float tly = v.LightPosition1.y - ry, bry = v.LightPosition1.y + ry;
                const float radiusOffset = radius * invZToY, zOffset = v.LightPosition1.z * zToY;
                const float cOne = 1.0f / 7.0f, mOne = 6.0f / 7.0f;
                auto Y = [&](float w) {
                    float y = hlerp(tly, bry, w);
                    if (w < 0.5f) { y -= radiusOffset; y -= zOffset; }
                    return y;
                };
                L.covX = make_float4(hlerp(tlx, brx, 0.0f), hlerp(tlx, brx, cOne), hlerp(tlx, brx, mOne), hlerp(tlx, brx, 1.0f));
                L.covY = make_float4(Y(0.0f), Y(cOne), Y(mOne), Y(1.0f));
                bx0 = std::min(L.covX.x, L.covX.w); bx1 = std::max(L.covX.x, L.covX.w);
                by0 = std::min(std::min(L.covY.x, L.covY.y), std::min(L.covY.z, L.covY.w));
                by1 = std::max(std::max(L.covY.x, L.covY.y), std::max(L.covY.z, L.covY.w));
            } else if (B.light_type == ILB_LIGHT_DIRECTIONAL) {  // DirectionalLight.fx:19-37
                bx0 = v.LightPosition1.x; bx1 = v.LightPosition2.x; by0 = v.LightPosition1.y; by1 = v.LightPosition2.y;
                // computeNormalFactorEx(direction, (0, 0, 1), 0.35, 0.35) (LightCommon.fxh:154-165): dot(-direction, n) = -direction.z
                // exactly for n = (0, 0, 1) and a finite direction; individually rounded like normalFactorEx<350>
                const float dz = -v.Color2.z, range = 0.35f;
                const float q = (dz + range) / range;
                const float flatFactor = std::pow(std::fmin(std::fmax(q, 0.0f), 1.0f), 0.85f);
                L.covX = make_float4(bx0, bx1, flatFactor, 0);
                L.covY = make_float4(by0, by1, 0, 0);
            } else {  // LineLightVertexShader LineLightCore.fxh:122-173 (bounds are +-9999 around the segment)
                const float radius = v.LightProperties.x + v.LightProperties.y + 1;
                bx0 = std::min(v.LightPosition1.x, v.LightPosition2.x) - 9999; bx1 = std::max(v.LightPosition1.x, v.LightPosition2.x) + 9999;
                by0 = std::min(v.LightPosition1.y, v.LightPosition2.y) - 9999; by1 = std::max(v.LightPosition1.y, v.LightPosition2.y) + 9999;
                by0 -= radius * invZToY;
                by0 -= v.LightPosition1.z * zToY;
                L.covX = make_float4(bx0, bx1, 0, 0);
                L.covY = make_float4(by0, by1, 0, 0);
            }
            // conservative pixel bounds: pixel centre (p + 0.5) / s + vp inside [b0, b1], one pixel of slack
            auto toPx = [](float w, float vp, float s, float slack) {
                double p = ((double)w - (double)vp) * (double)s - 0.5 + slack;
                if (!(p > -1.0e9)) p = -1.0e9;
                if (!(p < 1.0e9)) p = 1.0e9;
                return (int)std::floor(p);
            };
            L.px0 = toPx(bx0, f->ViewportPosition[0], sxs, -1.0f);
            L.px1 = toPx(bx1, f->ViewportPosition[0], sxs, 2.0f);
            L.py0 = toPx(by0, f->ViewportPosition[1], sys, -1.0f);
            L.py1 = toPx(by1, f->ViewportPosition[1], sys, 2.0f);
            out.push_back(L);
        }
    }
    return ILB_OK;
}

// device layout: DLine[n] (host lights only) followed by DLight[n + extra]; the `extra` records are appended on the device
// (particle lights)
std::mutex g_constBankMutex;
cudaEvent_t g_constBankLastUse = nullptr;   // behind the most recent launch (of any context) that reads c_lights / c_lines
unsigned long long g_constBankEpoch = 0;    // bumped by every upload into the bank: a context's cached list is there only while its epoch is current

// records that the work queued on `stream` so far reads the constant-bank light records
int constBankMarkUse(ilb_ctx* ctx, cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(g_constBankMutex);
    if (!g_constBankLastUse) ILB_CUDA(ctx, cudaEventCreateWithFlags(&g_constBankLastUse, cudaEventDisableTiming));
    ILB_CUDA(ctx, cudaEventRecord(g_constBankLastUse, stream));
    return ILB_OK;
}

int uploadLights(ilb_ctx* ctx, const std::vector<DLight>& lights, const std::vector<DLine>& lines, size_t extra, const DLight** d_lights,
                 const DLine** d_lines, bool toConstantBank = false) {
    const size_t n = lights.size();
    const size_t lineBytes = std::max<size_t>(n, 1) * sizeof(DLine), hostBytes = lineBytes + std::max<size_t>(n, 1) * sizeof(DLight);
    int rc = ilb_reserve(ctx, &ctx->d_lights, &ctx->d_lights_capacity, hostBytes + extra * sizeof(DLight), false);
    if (rc) return rc;
    *d_lines = reinterpret_cast<const DLine*>(ctx->d_lights);
    *d_lights = reinterpret_cast<const DLight*>(reinterpret_cast<const char*>(ctx->d_lights) + lineBytes);
    // Lights that did not change since the last upload (static lights, and the probe update that follows a frame with the same
    // list) are not copied again: the flattened records are compared with what the device copy -- and, when asked for, the
    // constant bank, which another context may have overwritten since (epoch) -- already hold.
    const size_t lb = n * sizeof(DLine), gb = n * sizeof(DLight);
    if (n && ctx->lights_cache_ptr == ctx->d_lights && ctx->lights_cache.size() == lb + gb &&
        memcmp(ctx->lights_cache.data(), lines.data(), lb) == 0 && memcmp(ctx->lights_cache.data() + lb, lights.data(), gb) == 0) {
        bool bankOk = !toConstantBank;
        if (toConstantBank) {
            std::lock_guard<std::mutex> lock(g_constBankMutex);
            bankOk = ctx->lights_cache_const_epoch != 0 && ctx->lights_cache_const_epoch == g_constBankEpoch;
        }
        if (bankOk) return ILB_OK;
    }
    // two pinned staging buffers alternate, each guarded by the event recorded behind its last copy: a frame never waits
    // for the stream unless the copy issued two frames ago is still pending
    const int slot = ctx->h_lights_next;
    ctx->h_lights_next ^= 1;
    if (!ctx->ev_lights[slot]) ILB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_lights[slot], cudaEventDisableTiming));
    else ILB_CUDA(ctx, cudaEventSynchronize(ctx->ev_lights[slot]));
    if (ctx->h_lights_capacity[slot] < hostBytes || !ctx->h_lights[slot]) {
        if (ctx->h_lights[slot]) cudaFreeHost(ctx->h_lights[slot]);
        ctx->h_lights[slot] = nullptr;
        ctx->h_lights_capacity[slot] = 0;
        ILB_CUDA(ctx, cudaMallocHost(&ctx->h_lights[slot], hostBytes * 2));
        ctx->h_lights_capacity[slot] = hostBytes * 2;
    }
    if (n) {
        memcpy(ctx->h_lights[slot], lines.data(), n * sizeof(DLine));
        memcpy(reinterpret_cast<char*>(ctx->h_lights[slot]) + lineBytes, lights.data(), n * sizeof(DLight));
        ILB_CUDA(ctx, cudaMemcpyAsync(ctx->d_lights, ctx->h_lights[slot], hostBytes, cudaMemcpyHostToDevice, ctx->stream));
        if (toConstantBank) {
            // The constant bank belongs to the module, i.e. to every context of the process on this device: the kernels of
            // another context's frame that still read it must have finished before it is overwritten (same-stream work is
            // ordered anyway).  One process-wide event, recorded behind every lighting launch that reads the bank.
            std::lock_guard<std::mutex> lock(g_constBankMutex);
            if (g_constBankLastUse) ILB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, g_constBankLastUse, 0));
            const char* staged = reinterpret_cast<const char*>(ctx->h_lights[slot]);
            ILB_CUDA(ctx, cudaMemcpyToSymbolAsync(c_lines, staged, n * sizeof(DLine), 0, cudaMemcpyHostToDevice, ctx->stream));
            ILB_CUDA(ctx, cudaMemcpyToSymbolAsync(c_lights, staged + lineBytes, n * sizeof(DLight), 0, cudaMemcpyHostToDevice, ctx->stream));
            ctx->lights_cache_const_epoch = ++g_constBankEpoch;
        } else {
            ctx->lights_cache_const_epoch = 0;
        }
        ILB_CUDA(ctx, cudaEventRecord(ctx->ev_lights[slot], ctx->stream));
        ctx->lights_cache.resize(lb + gb);
        memcpy(ctx->lights_cache.data(), lines.data(), lb);
        memcpy(ctx->lights_cache.data() + lb, lights.data(), gb);
        ctx->lights_cache_ptr = ctx->d_lights;
    }
    return ILB_OK;
}

// ---- particle lights ("next" row N4): ParticleLightSource (LightSource.cs:466-500) ---------------------------------------
// The reference draws the particle system's instanced quads with the ParticleLight material (LightingRenderer.cs:1126-1144):
// ParticleLightVertexShader (ParticleLight.fx:16-82) turns every live particle into a sphere light at its position with the
// template's properties and colour = unpremultiplied attribute colour * LightColor; the pixel shader is the sphere-light
// core without the shadow filter (:84-118).  Here the particle state never leaves HBM: three small kernels append one
// DLight per live, visible particle to the frame's light list in particle order (count per block, scan, ordered write),
// and the tile kernel shades them like sphere lights with a rectangular quad.  StippleFactor is taken as 1 (StippleReject
// lives in the un-vendored sq/Fracture DitherCommon.fxh).

struct PLightParams {
    const float4* P;      // PositionAndLife
    const float4* A;      // attributes (AttributeSampler)
    unsigned total;       // live_chunks * per_chunk
    unsigned* counts;     // per block of 256 particles
    unsigned* offsets;    // exclusive scan of counts; offsets[nblocks] = total lights
    unsigned nblocks;
    DLight* out;          // first particle light record
    float4 props, more, color, spec, quality;
    float longStep, rcpRamp;
    int hasField;
    float zToY, invZToY, sxs, sys, vpx, vpy;
};

ILB_DEV bool particleLightColor(const PLightParams& P, unsigned i, float4& position, float4& lightColor) {  // ParticleLight.fx:36-51,73-81
    position = __ldg(P.P + i);
    if (!(position.w > 0.0f)) return false;
    float4 c = __ldg(P.A + i);
    if (c.w > 0.0f) { c.x = xdiv(c.x, c.w); c.y = xdiv(c.y, c.w); c.z = xdiv(c.z, c.w); }  // unpremultiply
    lightColor = make_float4(xmul(c.x, P.color.x), xmul(c.y, P.color.y), xmul(c.z, P.color.z), xmul(c.w, P.color.w));
    return lightColor.w > 0.0f;
}

__global__ void __launch_bounds__(256) particle_light_count_kernel(const __grid_constant__ PLightParams P) {
    const unsigned i = blockIdx.x * 256u + threadIdx.x;
    float4 pos, col;
    const bool keep = (i < P.total) && particleLightColor(P, i, pos, col);
    const int n = __syncthreads_count(keep);
    if (threadIdx.x == 0) P.counts[blockIdx.x] = (unsigned)n;
}

__global__ void __launch_bounds__(1024) particle_light_scan_kernel(const __grid_constant__ PLightParams P) {  // one block
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < P.nblocks; base += 1024u) {
        const unsigned v = (base + tid < P.nblocks) ? P.counts[base + tid] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (unsigned)o) incl += t;
        }
        if (lane == 31u) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned w = s_warp[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xFFFFFFFFu, wi, o);
                if (lane >= (unsigned)o) wi += t;
            }
            s_warp[lane] = wi - w;  // exclusive warp offsets
        }
        __syncthreads();
        const unsigned carry = s_carry;
        if (base + tid < P.nblocks) P.offsets[base + tid] = carry + s_warp[warp] + incl - v;
        __syncthreads();
        if (tid == 1023u) s_carry = carry + s_warp[warp] + incl;
        __syncthreads();
    }
    if (tid == 0) P.offsets[P.nblocks] = s_carry;
}

__global__ void __launch_bounds__(256) particle_light_write_kernel(const __grid_constant__ PLightParams P) {
    __shared__ unsigned s_warp[8];
    const unsigned i = blockIdx.x * 256u + threadIdx.x, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    float4 pos, col;
    const bool keep = (i < P.total) && particleLightColor(P, i, pos, col);
    const unsigned ballot = __ballot_sync(0xFFFFFFFFu, keep);
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    if (!keep) return;
    unsigned rank = __popc(ballot & ((1u << lane) - 1u));
    for (unsigned w = 0; w < warp; w++) rank += s_warp[w];
    DLight L;
    L.pos1 = L.pos2 = make_float4(pos.x, pos.y, pos.z, 0.0f);
    L.props = P.props; L.more = P.more;
    L.evenMore = make_float4(-1.0f, 0.0f, 0.0f, 0.0f);  // no shadow filter in ParticleLightPixelShader
    L.color1 = col; L.color2 = P.spec;
    L.quality = P.quality; L.longStep = P.longStep; L.hasField = P.hasField; L.type = ILB_LIGHT_PARTICLE_BIT; L.rcpRamp = P.rcpRamp;
    // quad: lerp(tl, br, corner) with tl = center - radius, br = center + radius, tl.y -= radius * invZToY + z * zToY (:53-64)
    const float radius = xadd(xadd(P.props.x, P.props.y), 1.0f);
    const float x0 = xsub(pos.x, radius), x1 = xadd(pos.x, radius);
    float y0 = xsub(pos.y, radius);
    const float y1 = xadd(pos.y, radius);
    y0 = xsub(y0, xmul(radius, P.invZToY));
    y0 = xsub(y0, xmul(pos.z, P.zToY));
    L.covX = make_float4(x0, x1, 0.0f, 0.0f);
    L.covY = make_float4(y0, y1, 0.0f, 0.0f);
    // conservative pixel bounds (pixel centre (p + 0.5) / s + vp inside [b0, b1]), one pixel of slack plus float rounding
    const float BIG = 1.0e9f;
    L.px0 = (int)floorf(fminf(fmaxf((x0 - P.vpx) * P.sxs - 2.5f, -BIG), BIG));
    L.px1 = (int)floorf(fminf(fmaxf((x1 - P.vpx) * P.sxs + 2.5f, -BIG), BIG));
    L.py0 = (int)floorf(fminf(fmaxf((y0 - P.vpy) * P.sys - 2.5f, -BIG), BIG));
    L.py1 = (int)floorf(fminf(fmaxf((y1 - P.vpy) * P.sys + 2.5f, -BIG), BIG));
    float4* dst = reinterpret_cast<float4*>(P.out + P.offsets[blockIdx.x] + rank);
    const float4* src = reinterpret_cast<const float4*>(&L);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(DLight) / 16); k++) dst[k] = src[k];
}

}  // namespace

size_t ilb_format_bytes(int format) {
    return format == ILB_FORMAT_FLOAT4 ? 16 : (format == ILB_FORMAT_HALF4 ? 8 : 4);
}

namespace {

struct LightingPrepared {
    LightingParams P;   // everything but the row band and the outputs
    int nline = 0, nlights = 0;
    bool hasRamp = false;     // a sphere-light batch with a ramp texture: the instantiations with ILB_LIGHT_RAMP_BIT in TYPES
    bool constBank = false;   // the frame's light records are in c_lights / c_lines as well
    std::vector<int> sphereRects;   // px0, py0, px1, py1 of every sphere light of the host list: what tileOrderFor counts per tile
};

// Per-frame work that does not depend on the row band: validate, flatten + upload the light list, resolve the field.
int lightingPrepare(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* f, const ilb_light_batch* batches, int batch_count,
                    const ilb_light_vertex* vertices, int vertex_count, LightingPrepared* out) {
    if (!f || (batch_count > 0 && (!batches || !vertices))) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null argument");
    if (f->width <= 0 || f->height <= 0 || f->row_begin < 0 || f->row_end > f->height || f->row_begin > f->row_end)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad frame geometry %dx%d rows [%d,%d)", f->width, f->height, f->row_begin, f->row_end);
    if (f->lightmap_format != ILB_FORMAT_FLOAT4 && f->lightmap_format != ILB_FORMAT_HALF4 && f->lightmap_format != ILB_FORMAT_RGBA8)
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad lightmap format %d", f->lightmap_format);
    if (df && df->ctx != ctx) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "distance field belongs to another context");

    std::vector<DLight> lights;
    std::vector<DLine> lines;
    const ilb_df_uniforms* geometry = nullptr;
    auto rcp = [](float y) { const float a = std::fabs(y); return (a >= 1.0e-30f && a <= 1.0e30f) ? 1.0f / y : 0.0f; };
    int rc = flattenLights(ctx, df, f, batches, batch_count, vertices, vertex_count, lights, lines, &geometry);
    if (rc) return rc;
    if (lights.size() > 65535 * (size_t)TILE_THREADS) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "too many lights");

    LightingParams& P = out->P;
    memset(&P, 0, sizeof(P));
    // particle light sources: count the live, visible particles of every source (device), then append their records
    struct Pending { PLightParams pp; unsigned count; };
    std::vector<Pending> pending;
    size_t extra = 0;
    if (!ctx->particle_lights.empty()) {
        const float sxs = f->GBufferTexelSizeAndMisc.z * f->EnvironmentZAndScale.z, sys = f->GBufferTexelSizeAndMisc.w * f->EnvironmentZAndScale.w;
        size_t words = 0;
        for (const ilb_particle_light_source& src : ctx->particle_lights) {
            const ilb_psys* ps = src.system;
            const size_t total = (size_t)ps->live_chunks * ps->per_chunk;
            words += 2 * ((total + 255) / 256) + 1;
        }
        rc = ilb_reserve(ctx, &ctx->d_plight_scratch, &ctx->d_plight_scratch_capacity, std::max<size_t>(words, 1) * sizeof(unsigned), false);
        if (rc) return rc;
        unsigned* scratch = reinterpret_cast<unsigned*>(ctx->d_plight_scratch);
        for (const ilb_particle_light_source& src : ctx->particle_lights) {
            const ilb_psys* ps = src.system;
            const size_t total = (size_t)ps->live_chunks * ps->per_chunk;
            if (total == 0) continue;
            Pending pd;
            memset(&pd, 0, sizeof(pd));
            PLightParams& pp = pd.pp;
            pp.P = ps->buf[0]; pp.A = ps->buf[2];
            pp.total = (unsigned)total;
            pp.nblocks = (unsigned)((total + 255) / 256);
            pp.counts = scratch; pp.offsets = scratch + pp.nblocks;
            scratch += 2 * pp.nblocks + 1;
            pp.props = h4(src.LightProperties); pp.more = h4(src.MoreLightProperties);
            pp.color = h4(src.LightColor); pp.spec = h4(src.LightSpecularColor);
            const bool hasField = (df != nullptr) && (src.df.Extent.x > 0.0f);
            if (hasField) {
                if (!geometry) geometry = &src.df;
            }
            pp.quality = make_float4(src.df.ConeAndMisc.x, src.df.ConeAndMisc.z, src.df.StepAndMisc2.x, src.df.Packed1.w);
            pp.longStep = src.df.StepAndMisc2.z;
            pp.hasField = hasField ? 1 : 0;
            pp.rcpRamp = rcp(src.LightProperties.y);
            pp.zToY = f->EnvironmentZToY.x; pp.invZToY = f->EnvironmentZToY.y;
            pp.sxs = sxs; pp.sys = sys; pp.vpx = f->ViewportPosition[0]; pp.vpy = f->ViewportPosition[1];
            particle_light_count_kernel<<<pp.nblocks, 256, 0, ctx->stream>>>(pp);
            particle_light_scan_kernel<<<1, 1024, 0, ctx->stream>>>(pp);
            ctx->launches += 2;
            ILB_CUDA(ctx, cudaMemcpyAsync(&pd.count, pp.offsets + pp.nblocks, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
            pending.push_back(pd);
        }
        ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (const Pending& pd : pending) extra += pd.count;
        if (extra > ((size_t)1 << 20))
            return ilb_fail(ctx, ILB_ERR_UNSUPPORTED, "%zu particle lights in one frame (limit 1048576): the tile culling is brute force over the light list", extra);
    }
    out->constBank = ctx->opt[ILB_OPT_LIGHT_CONST_BANK] != 0 && !lights.empty() && lights.size() + extra <= (size_t)ILB_CONST_LIGHTS;
    rc = uploadLights(ctx, lights, lines, extra, &P.lights, &P.lines, out->constBank);
    if (rc) return rc;
    {
        size_t at = lights.size();
        for (Pending& pd : pending) {
            if (pd.count == 0) continue;
            pd.pp.out = const_cast<DLight*>(P.lights) + at;
            particle_light_write_kernel<<<pd.pp.nblocks, 256, 0, ctx->stream>>>(pd.pp);
            ctx->launches++;
            at += pd.count;
        }
        ILB_CUDA(ctx, cudaGetLastError());
        if (out->constBank && extra)   // the particle-light records were written on the device: append them to the bank from there
            ILB_CUDA(ctx, cudaMemcpyToSymbolAsync(c_lights, P.lights + lights.size(), extra * sizeof(DLight), lights.size() * sizeof(DLight),
                                                  cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (geometry) {
        if (!ilb_make_df_geometry(df, *geometry, &P.df)) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad distance-field uniforms");
        rc = ilb_planes_attach(ctx, df, *geometry, &P.df);
        if (rc) return rc;
    }
    if (!ctx->ramps.empty()) {  // the ramp table: a few 16-byte records, re-uploaded when a texture was created or destroyed
        const size_t bytes = sizeof(RampTex) * ctx->ramps.size();
        rc = ilb_reserve(ctx, &ctx->d_ramp_table, &ctx->d_ramp_table_capacity, bytes, false);
        if (rc) return rc;
        if (ctx->ramp_table_dirty) {
            std::vector<RampTex> table(ctx->ramps.size());
            for (size_t i = 0; i < table.size(); i++) table[i] = RampTex{ctx->ramps[i].texels, ctx->ramps[i].w, ctx->ramps[i].h, 0};
            ILB_CUDA(ctx, cudaMemcpyAsync(ctx->d_ramp_table, table.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
            ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `table` is a host vector
            ctx->ramp_table_dirty = false;
        }
        P.ramps = reinterpret_cast<const RampTex*>(ctx->d_ramp_table);
    }
    P.envZAndScale = h4(f->EnvironmentZAndScale);
    P.envZToY = h4(f->EnvironmentZToY);
    P.gbTexelAndMisc = h4(f->GBufferTexelSizeAndMisc);
    if (!ctx->gbuffer) P.gbTexelAndMisc.x = P.gbTexelAndMisc.y = 0.0f;
    P.clear = h4(f->ClearColor);
    P.gbViewportRelative = f->GBufferViewportRelative;
    P.vpx = f->ViewportPosition[0];
    P.vpy = f->ViewportPosition[1];
    P.gbuffer = ctx->gbuffer;
    P.gw = ctx->gb_w; P.gh = ctx->gb_h; P.gfmt = ctx->gb_fmt;
    P.nlights = (int)(lights.size() + extra);
    P.width = f->width; P.height = f->height;
    P.out_format = f->lightmap_format;
    P.stencil = f->stencil_culling;
    out->nlights = (int)(lights.size() + extra);
    out->nline = 0;
    for (const DLight& L : lights) {
        out->nline += ((L.type & ILB_LIGHT_TYPE_MASK) == ILB_LIGHT_LINE) ? 1 : 0;
        out->hasRamp = out->hasRamp || (L.type & ILB_LIGHT_RAMP_BIT) != 0;
        if ((L.type & ILB_LIGHT_TYPE_MASK) == ILB_LIGHT_SPHERE) {
            const int r[4] = {L.px0, L.py0, L.px1, L.py1};
            out->sphereRects.insert(out->sphereRects.end(), r, r + 4);
        }
    }
    return ILB_OK;
}

// ILB_OPT_LIGHT_TILE_ORDER: the tile indices of rows [row_begin, row_end) sorted by the number of sphere-light quads that
// cover the tile, heaviest first, equal tiles in their traversal order (a stable counting sort, so the 2-D locality of the
// traversal survives inside every class).  The per-pixel cost of the sphere + directional pass is a serial loop over the tile's
// lights, so a tile under a cluster of lights runs several times as long as an empty one: started last it would keep a few SMs
// busy while the rest of the device idles.  Counted on the host from the light list (a 2-D difference array over the tiles: a
// few microseconds per band), cached per (geometry, rectangles) in a ring of slots that own their staging and device memory.
// Purely a launch-order hint: a missing or stale order costs time, never results.  Returns null when there is nothing to sort.
const unsigned* tileOrderFor(ilb_ctx* ctx, const LightingPrepared& prep, int width, int row_begin, int row_end, int lane, cudaStream_t st) {
    const int tiles_x = (width + TILE_W - 1) / TILE_W, tiles_y = (row_end - row_begin + TILE_H - 1) / TILE_H;
    const size_t tiles = (size_t)tiles_x * (size_t)tiles_y;
    if (prep.sphereRects.empty() || tiles < 512) return nullptr;
    constexpr int SLOTS = (int)(sizeof(ctx->tile_orders) / sizeof(ctx->tile_orders[0]));
    std::vector<int> key;
    key.reserve(prep.sphereRects.size() + 3);
    key.push_back(width); key.push_back(row_begin); key.push_back(row_end);
    key.insert(key.end(), prep.sphereRects.begin(), prep.sphereRects.end());
    ilb_ctx::TileOrder* slot = nullptr;
    for (int i = 0; i < SLOTS; i++)
        if (ctx->tile_orders[i].d && ctx->tile_orders[i].key == key) { slot = &ctx->tile_orders[i]; break; }
    const bool hit = slot != nullptr;
    if (!hit) {
        slot = &ctx->tile_orders[0];
        for (int i = 1; i < SLOTS; i++)
            if (ctx->tile_orders[i].stamp < slot->stamp) slot = &ctx->tile_orders[i];
        for (int l = 0; l < 2; l++)   // kernels that still read the order this slot held
            if (slot->used[l] && cudaEventSynchronize(slot->used[l]) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
        if (slot->capacity < tiles) {
            if (slot->h) cudaFreeHost(slot->h);
            if (slot->d) cudaFree(slot->d);
            slot->h = slot->d = nullptr; slot->capacity = 0; slot->key.clear();
            const size_t want = tiles + tiles / 2;
            if (cudaMallocHost(reinterpret_cast<void**>(&slot->h), want * sizeof(unsigned)) != cudaSuccess ||
                cudaMalloc(reinterpret_cast<void**>(&slot->d), want * sizeof(unsigned)) != cudaSuccess) {
                (void)cudaGetLastError();
                if (slot->h) cudaFreeHost(slot->h);
                slot->h = slot->d = nullptr;
                return nullptr;
            }
            slot->capacity = want;
        }
        // sphere-light quads per tile: 2-D difference array, then prefix sums
        const int dw = tiles_x + 1;
        std::vector<int> cnt((size_t)(tiles_y + 1) * (size_t)dw, 0);
        for (size_t i = 0; i + 3 < prep.sphereRects.size(); i += 4) {
            const int x0 = std::max(prep.sphereRects[i], 0), y0 = std::max(prep.sphereRects[i + 1], row_begin);
            const int x1 = std::min(prep.sphereRects[i + 2], width - 1), y1 = std::min(prep.sphereRects[i + 3], row_end - 1);
            if (x0 > x1 || y0 > y1) continue;
            const int tx0 = x0 / TILE_W, tx1 = x1 / TILE_W, ty0 = (y0 - row_begin) / TILE_H, ty1 = (y1 - row_begin) / TILE_H;
            cnt[(size_t)ty0 * dw + tx0]++; cnt[(size_t)ty0 * dw + tx1 + 1]--;
            cnt[(size_t)(ty1 + 1) * dw + tx0]--; cnt[(size_t)(ty1 + 1) * dw + tx1 + 1]++;
        }
        int most = 0;
        for (int y = 0; y < tiles_y; y++)
            for (int x = 0; x < tiles_x; x++) {
                int v = cnt[(size_t)y * dw + x];
                if (x) v += cnt[(size_t)y * dw + x - 1];
                if (y) v += cnt[(size_t)(y - 1) * dw + x];
                if (x && y) v -= cnt[(size_t)(y - 1) * dw + x - 1];
                cnt[(size_t)y * dw + x] = v;
                most = std::max(most, v);
            }
        // stable counting sort of the kernel's linear tile indices (the traversal of shadeTile), heaviest class first
        auto tileCount = [&](unsigned tile) {
            int tileX, tileY;
            if (ILB_SWIZZLE > 1) {
                const int band = (int)(tile / (unsigned)(ILB_SWIZZLE * tiles_x)), inband = (int)(tile % (unsigned)(ILB_SWIZZLE * tiles_x));
                const int rowsInBand = std::min(ILB_SWIZZLE, tiles_y - band * ILB_SWIZZLE);
                tileX = inband / rowsInBand;
                tileY = band * ILB_SWIZZLE + inband % rowsInBand;
            } else {
                tileX = (int)(tile % (unsigned)tiles_x);
                tileY = (int)(tile / (unsigned)tiles_x);
            }
            return cnt[(size_t)tileY * dw + tileX];
        };
        std::vector<unsigned> start((size_t)most + 2, 0u);
        for (unsigned t = 0; t < (unsigned)tiles; t++) start[(size_t)(most - tileCount(t)) + 1]++;
        for (size_t c = 1; c < start.size(); c++) start[c] += start[c - 1];
        for (unsigned t = 0; t < (unsigned)tiles; t++) slot->h[start[(size_t)(most - tileCount(t))]++] = t;
        if (most == 0) { slot->key.clear(); return nullptr; }   // no quad touches the band: the traversal order is as good as any
        if (cudaMemcpyAsync(slot->d, slot->h, tiles * sizeof(unsigned), cudaMemcpyHostToDevice, st) != cudaSuccess) {
            (void)cudaGetLastError();
            slot->key.clear();
            return nullptr;
        }
        slot->key = std::move(key);
    }
    slot->stamp = ++ctx->tile_order_clock;
    return slot->d;
}

// records that the launches queued on lane `lane` so far read the order (see ilb_ctx::TileOrder)
void tileOrderMarkUse(ilb_ctx* ctx, const unsigned* order, int lane, cudaStream_t st) {
    for (ilb_ctx::TileOrder& t : ctx->tile_orders) {
        if (t.d != order) continue;
        cudaEvent_t& e = t.used[lane ? 1 : 0];
        if (!e && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { (void)cudaGetLastError(); e = nullptr; return; }
        if (cudaEventRecord(e, st) != cudaSuccess) (void)cudaGetLastError();
        return;
    }
}

// Shades rows [row_begin, row_end) of a prepared frame into d_outputs (band buffers whose row 0 is `out_row_base`).
// `lane` 1 runs the band on ctx->band_stream with its own scratch sums, so that the kernels of two neighbouring bands of a
// pipelined frame can be resident together (the tail of one band's last wave is filled by the next band's CTAs).
int lightingLaunchRows(ilb_ctx* ctx, const LightingPrepared& prep, int row_begin, int row_end, void* const* d_outputs, int output_count,
                       int out_row_base, int lane = 0) {
    const cudaStream_t st = lane ? ctx->band_stream : ctx->stream;
    void** accumBuffer = lane ? &ctx->d_accum2 : &ctx->d_accum;
    size_t* accumCapacity = lane ? &ctx->d_accum2_capacity : &ctx->d_accum_capacity;
    if (output_count < 1 || output_count > MAX_OUTPUTS) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "1..%d outputs", MAX_OUTPUTS);
    if (row_begin >= row_end) return ILB_OK;
    LightingParams P = prep.P;
    P.row_begin = row_begin; P.row_end = row_end;
    for (int o = 0; o < output_count; o++) {
        if (!d_outputs[o]) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null output %d", o);
        P.outs[o] = d_outputs[o];
    }
    P.nouts = output_count;
    P.out_row_base = out_row_base;
    P.tiles_x = (P.width + TILE_W - 1) / TILE_W;
    P.tiles_y = (row_end - row_begin + TILE_H - 1) / TILE_H;
    const unsigned tiles = (unsigned)P.tiles_x * (unsigned)P.tiles_y;
    constexpr int NOLINE = ILB_LIGHT_SPHERE | ILB_LIGHT_DIRECTIONAL | ILB_LIGHT_PARTICLE_BIT, ALL = NOLINE | ILB_LIGHT_LINE;
    bool split = prep.nline > 0 && prep.nline < prep.nlights;
    if (const char* e = getenv("ILB_SPLIT_PASSES")) split = split && e[0] != '0';
    const unsigned* order = nullptr;   // ILB_OPT_LIGHT_TILE_ORDER: heaviest tiles of the sphere + directional pass first
    int planesMask = 3;  // dev knob: bit 0 = line pass samples the planes, bit 1 = sphere / directional pass does
    if (const char* e = getenv("ILB_PLANES_MASK")) planesMask = atoi(e);
#define ILB_LIGHT_LAUNCH(TYPES)                                                                                       \
    do {                                                                                                              \
        if (prep.hasRamp && ((TYPES) & ILB_LIGHT_SPHERE)) {                                                           \
            if (P.df.planes && (planesMask & 2))                                                                      \
                light_accumulate_kernel<1, (TYPES) | ILB_LIGHT_RAMP_BIT, false><<<tiles, TILE_THREADS, 0, st>>>(P);   \
            else                                                                                                      \
                light_accumulate_kernel<0, (TYPES) | ILB_LIGHT_RAMP_BIT, false><<<tiles, TILE_THREADS, 0, st>>>(P);   \
        } else if (P.df.planes && (planesMask & (((TYPES) & ILB_LIGHT_LINE) ? 1 : 2))) {                              \
            if (prep.constBank) light_accumulate_kernel<1, TYPES, true><<<tiles, TILE_THREADS, 0, st>>>(P);           \
            else light_accumulate_kernel<1, TYPES, false><<<tiles, TILE_THREADS, 0, st>>>(P);                         \
        } else {                                                                                                      \
            light_accumulate_kernel<0, TYPES, false><<<tiles, TILE_THREADS, 0, st>>>(P);                              \
        }                                                                                                             \
        ctx->launches++;                                                                                              \
    } while (0)
    const bool concurrent = ctx->opt[ILB_OPT_LIGHT_CONCURRENT] != 0 &&  // every pass needs at least one grid, or its tiles are never shaded
                            ctx->opt[ILB_OPT_LIGHT_LINE_CTAS] + ctx->opt[ILB_OPT_LIGHT_LINE_HELPERS] > 0 &&
                            ctx->opt[ILB_OPT_LIGHT_OTHER_CTAS] + ctx->opt[ILB_OPT_LIGHT_OTHER_HELPERS] > 0;
    if (split && concurrent && lane == 0 && !prep.hasRamp) {
        // both passes at once: persistent grids sized to be co-resident, one tile queue per pass, the pass that reaches a
        // tile second adds the two fp32 partial sums and stores the texel (see light_accumulate_persistent_kernel)
        const size_t bytes = sizeof(float4) * (size_t)P.width * (size_t)(row_end - row_begin);
        int rc = ilb_reserve(ctx, &ctx->d_accum, &ctx->d_accum_capacity, bytes, false);
        if (rc) return rc;
        rc = ilb_reserve(ctx, &ctx->d_accum2, &ctx->d_accum2_capacity, bytes, false);
        if (rc) return rc;
        rc = ilb_reserve(ctx, &ctx->d_tilework, &ctx->d_tilework_capacity, sizeof(unsigned) * ((size_t)tiles + 2), false);
        if (rc) return rc;
        if (!ctx->light_aux[0]) {
            for (int i = 0; i < 3; i++) {
                ILB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->light_aux[i], cudaStreamNonBlocking));
                ILB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_light_join[i], cudaEventDisableTiming));
            }
            ILB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_light_fork, cudaEventDisableTiming));
        }
        unsigned* work = reinterpret_cast<unsigned*>(ctx->d_tilework);
        ILB_CUDA(ctx, cudaMemsetAsync(work, 0, sizeof(unsigned) * ((size_t)tiles + 2), ctx->stream));
        ILB_CUDA(ctx, cudaEventRecord(ctx->ev_light_fork, ctx->stream));
        for (int i = 0; i < 3; i++) ILB_CUDA(ctx, cudaStreamWaitEvent(ctx->light_aux[i], ctx->ev_light_fork, 0));
        LightingParams PL = P, PS = P;
        PL.accum_out = reinterpret_cast<float4*>(ctx->d_accum); PL.accum_in = reinterpret_cast<const float4*>(ctx->d_accum2);
        PS.accum_out = reinterpret_cast<float4*>(ctx->d_accum2); PS.accum_in = reinterpret_cast<const float4*>(ctx->d_accum);
        PL.tile_counter = work; PS.tile_counter = work + 1;
        PL.tile_done = PS.tile_done = work + 2;
        PS.clear = make_float4(0.0f, 0.0f, 0.0f, 0.0f);  // the clear colour enters through the line-light sums
        const unsigned sms = (unsigned)ctx->sm_count;
        auto grid = [&](int perSm) { return std::min<unsigned>(tiles, sms * (unsigned)std::max(perSm, 0)); };
#define ILB_LIGHT_PERSIST(PARAMS, TYPES, GRID, STREAM)                                                                 \
    do {                                                                                                              \
        if ((GRID) > 0) {                                                                                             \
            if (P.df.planes && (planesMask & (((TYPES) & ILB_LIGHT_LINE) ? 1 : 2)))                                   \
                light_accumulate_persistent_kernel<1, TYPES><<<(GRID), TILE_THREADS, 0, (STREAM)>>>(PARAMS);          \
            else                                                                                                      \
                light_accumulate_persistent_kernel<0, TYPES><<<(GRID), TILE_THREADS, 0, (STREAM)>>>(PARAMS);          \
            ctx->launches++;                                                                                          \
        }                                                                                                             \
    } while (0)
        ILB_LIGHT_PERSIST(PL, ILB_LIGHT_LINE, grid(ctx->opt[ILB_OPT_LIGHT_LINE_CTAS]), ctx->stream);
        ILB_LIGHT_PERSIST(PS, NOLINE, grid(ctx->opt[ILB_OPT_LIGHT_OTHER_CTAS]), ctx->light_aux[0]);
        ILB_LIGHT_PERSIST(PL, ILB_LIGHT_LINE, grid(ctx->opt[ILB_OPT_LIGHT_LINE_HELPERS]), ctx->light_aux[1]);
        ILB_LIGHT_PERSIST(PS, NOLINE, grid(ctx->opt[ILB_OPT_LIGHT_OTHER_HELPERS]), ctx->light_aux[2]);
#undef ILB_LIGHT_PERSIST
        for (int i = 0; i < 3; i++) {
            ILB_CUDA(ctx, cudaEventRecord(ctx->ev_light_join[i], ctx->light_aux[i]));
            ILB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_light_join[i], 0));
        }
    } else if (split) {
        const size_t bytes = sizeof(float4) * (size_t)P.width * (size_t)(row_end - row_begin);
        const int rc = ilb_reserve(ctx, accumBuffer, accumCapacity, bytes, false);
        if (rc) return rc;
        // (an upload of the order is queued here, ahead of the line pass: nothing may sit between the two launches of the
        // programmatic dependent pair)
        order = (ctx->opt[ILB_OPT_LIGHT_TILE_ORDER] != 0) ? tileOrderFor(ctx, prep, P.width, row_begin, row_end, lane, st) : nullptr;
        P.accum_out = reinterpret_cast<float4*>(*accumBuffer);
        ILB_LIGHT_LAUNCH(ILB_LIGHT_LINE);
        P.accum_in = P.accum_out;
        P.accum_out = nullptr;
        P.tile_order = order;
        P.clear = make_float4(0.0f, 0.0f, 0.0f, 0.0f);  // the clear colour entered through the line-light sums
        if (ctx->opt[ILB_OPT_LIGHT_PDL] && !prep.hasRamp) {
            // programmatic dependent launch: the second pass's CTAs fill the SM slots the first pass's last wave leaves idle
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(tiles); cfg.blockDim = dim3(TILE_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            if (P.df.planes && (planesMask & 2)) {
                if (prep.constBank) ILB_CUDA(ctx, cudaLaunchKernelEx(&cfg, light_accumulate_kernel<1, NOLINE, true>, P));
                else ILB_CUDA(ctx, cudaLaunchKernelEx(&cfg, light_accumulate_kernel<1, NOLINE, false>, P));
            } else {
                ILB_CUDA(ctx, cudaLaunchKernelEx(&cfg, light_accumulate_kernel<0, NOLINE, false>, P));
            }
            ctx->launches++;
        } else {
            ILB_LIGHT_LAUNCH(NOLINE);
        }
    } else if (prep.nline == 0) {
        P.tile_order = order = (ctx->opt[ILB_OPT_LIGHT_TILE_ORDER] != 0) ? tileOrderFor(ctx, prep, P.width, row_begin, row_end, lane, st) : nullptr;
        ILB_LIGHT_LAUNCH(NOLINE);
    } else {
        P.tile_order = order = (ctx->opt[ILB_OPT_LIGHT_TILE_ORDER] != 0) ? tileOrderFor(ctx, prep, P.width, row_begin, row_end, lane, st) : nullptr;
        ILB_LIGHT_LAUNCH(ALL);
    }
#undef ILB_LIGHT_LAUNCH
    ILB_CUDA(ctx, cudaGetLastError());
    if (order) tileOrderMarkUse(ctx, order, lane, st);
    if (prep.constBank) return constBankMarkUse(ctx, st);
    return ILB_OK;
}

}  // namespace

int ilb_lighting_launch(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* f, const ilb_light_batch* batches, int batch_count,
                        const ilb_light_vertex* vertices, int vertex_count, void* const* d_outputs, int output_count,
                        bool outputs_are_full_frames) {
    LightingPrepared prep;
    int rc = ilb_frames_drain(ctx);   // frames in flight share the scratch sums and the two compute lanes
    if (rc) return rc;
    rc = lightingPrepare(ctx, df, f, batches, batch_count, vertices, vertex_count, &prep);
    if (rc) return rc;
    const int rows = f->row_end - f->row_begin, outBase = outputs_are_full_frames ? 0 : f->row_begin;
    const bool splitFrame = prep.nline > 0 && prep.nline < prep.nlights;
    if (ctx->opt[ILB_OPT_LIGHT_SPLIT_BAND] != 0 && rows >= 4 * TILE_H && 2 * rows < f->height &&
        !(ctx->opt[ILB_OPT_LIGHT_CONCURRENT] != 0 && splitFrame)) {
        // A band of a sharded frame: its two halves run on two compute lanes (see lightingLaunchRows), the stream continues
        // behind both.  Whole frames are long enough for their tails not to matter and keep the single launch pair.
        const int mid = f->row_begin + ((rows / 2 + TILE_H - 1) / TILE_H) * TILE_H;
        if (!ctx->band_stream) {
            ILB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->band_stream, cudaStreamNonBlocking));
            ILB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_band_fork, cudaEventDisableTiming));
            ILB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_band_join, cudaEventDisableTiming));
        }
        if (splitFrame) {   // both scratch buffers at their final size before anything is in flight
            const size_t bytes = sizeof(float4) * (size_t)f->width * (size_t)std::max(mid - f->row_begin, f->row_end - mid);
            rc = ilb_reserve(ctx, &ctx->d_accum, &ctx->d_accum_capacity, bytes, false);
            if (rc) return rc;
            rc = ilb_reserve(ctx, &ctx->d_accum2, &ctx->d_accum2_capacity, bytes, false);
            if (rc) return rc;
        }
        ILB_CUDA(ctx, cudaEventRecord(ctx->ev_band_fork, ctx->stream));
        ILB_CUDA(ctx, cudaStreamWaitEvent(ctx->band_stream, ctx->ev_band_fork, 0));
        rc = lightingLaunchRows(ctx, prep, f->row_begin, mid, d_outputs, output_count, outBase, 0);
        if (rc) return rc;
        rc = lightingLaunchRows(ctx, prep, mid, f->row_end, d_outputs, output_count, outBase, 1);
        if (rc) return rc;
        ILB_CUDA(ctx, cudaEventRecord(ctx->ev_band_join, ctx->band_stream));
        ILB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_band_join, 0));
    } else {
        rc = lightingLaunchRows(ctx, prep, f->row_begin, f->row_end, d_outputs, output_count, outBase);
        if (rc) return rc;
    }
    // the kernels read the G-buffer texels of their own rows only (decodePixel); with a G-buffer of another size the mapping is
    // not one to one, so the launch counts as a reader of every row
    const bool oneToOne = ctx->gbuffer && ctx->gb_w == f->width && ctx->gb_h == f->height && f->GBufferViewportRelative == 0.0f;
    if (ctx->gbuffer) return ilb_gbuffer_note_user(ctx, oneToOne ? f->row_begin : 0, oneToOne ? f->row_end : ctx->gb_h);
    return ILB_OK;
}

// Host-to-host frame: G-buffer band up, shade, lightmap band down, software-pipelined over row bands on three streams
// so that the copies of neighbouring bands hide behind the kernels (the G-buffer decode reads only the pixel's own
// texel, so a band needs only its own rows).  Synchronous: returns when lightmap_out_host is complete.
int ilb_lighting_frame_from_host(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* f, const ilb_light_batch* batches, int batch_count,
                                 const ilb_light_vertex* vertices, int vertex_count, int gw, int gh, int gfmt, const void* gbuffer_host,
                                 void* lightmap_out_host, unsigned long long* out_ticket) {
    if (out_ticket) *out_ticket = 0;
    if (!f || !gbuffer_host || !lightmap_out_host) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null argument");
    if (gfmt != ILB_FORMAT_FLOAT4 && gfmt != ILB_FORMAT_HALF4) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "G-buffer format must be FLOAT4 or HALF4");
    if (gw != f->width || gh != f->height)  // one texel per pixel, so that a row band of the frame needs the same rows of the G-buffer
        return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "pipelined frames need a G-buffer of the frame's size (%dx%d), got %dx%d", f->width, f->height, gw, gh);
    if (f->GBufferViewportRelative != 0.0f) return ilb_fail(ctx, ILB_ERR_UNSUPPORTED, "pipelined frames need a screen-aligned G-buffer");
    const int rows = f->row_end - f->row_begin;
    if (rows < 0 || f->row_begin < 0 || f->row_end > f->height) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad frame geometry");
    // Bands of whole tile rows with heights 1 : 2 : 4 : 6 : 6 : 6 : 4 : 2 : 1 -- the first band is short so that the first kernel
    // starts after 1/32 of the upload, the last so that only 1/32 of the download is left when the last kernel ends, the middle
    // ones long so that few launches are paid (7.82 -> 7.69 ms per C4 frame against the 7-band split 1 : 2 : 3 : 4 : 3 : 2 : 1).
    // A rank's band of a frame sharded over 8 GPUs (270 rows, 4 080 tiles) still does best with the 9-band split, although a band is
    // then less than a wave of CTAs -- the two compute lanes interleave neighbouring bands, and what counts is how early the first
    // kernel starts and how little is left to download after the last (measured at 8 GPUs: 1 band 2.26 ms per host-to-host
    // frame, 2 bands 1.75, 4 bands 1.60, 7 bands 1.56, 9 bands 1.52).  Only small render targets take a coarser split.
    // ILB_BAND_SHARES selects the split (dev knob): 0 = 1:2:3:4:3:2:1, 1 = 1:2:4:6:6:6:4:2:1, 2 = 1:2:2:1, 3 = 1:1, 4 = one band
    static const int kShares[5][9] = {{1, 2, 3, 4, 3, 2, 1, 0, 0}, {1, 2, 4, 6, 6, 6, 4, 2, 1}, {1, 2, 2, 1, 0, 0, 0, 0, 0}, {1, 1, 0, 0, 0, 0, 0, 0, 0},
                                      {1, 0, 0, 0, 0, 0, 0, 0, 0}};
    static const int kCount[5] = {7, 9, 4, 2, 1}, kTotal[5] = {16, 32, 6, 2, 1};
    const long long tiles = (long long)((rows + TILE_H - 1) / TILE_H) * ((f->width + TILE_W - 1) / TILE_W);
    int which = tiles >= 2500 ? 1 : tiles >= 600 ? 2 : 3;
    if (const char* e = getenv("ILB_BAND_SHARES")) { const int v = atoi(e); if (v >= 0 && v <= 4) which = v; }
    const int* kShare = kShares[which];
    const int NB = kCount[which], total = kTotal[which];
    static_assert(ILB_PIPELINE_BANDS >= 9, "one event pair per band");
    int edge[10];
    edge[0] = f->row_begin;
    for (int b = 0, acc = 0; b < NB; b++) {
        acc += kShare[b];
        const int e = f->row_begin + (int)(((long long)rows * acc / total + TILE_H - 1) / TILE_H * TILE_H);
        edge[b + 1] = (b == NB - 1) ? f->row_end : std::min(std::max(e, edge[b]), f->row_end);
    }
    // Frames in flight (ilb_render_lighting_frame_async): this frame may be queued behind the one before it band by band when both
    // have the same geometry, buffers and band edges and nothing else has touched the G-buffer since; otherwise whatever is in
    // flight is waited for first, and the frame starts behind everything queued on the context's stream, as a lone frame does.
    const size_t gtexel0 = ilb_format_bytes(gfmt), ltexel0 = ilb_format_bytes(f->lightmap_format);
    bool chained = ctx->frame_ticket > ctx->frame_waited && ctx->pipe_generation == ctx->gb_generation && ctx->gbuffer_owned &&
                   ctx->gbuffer == ctx->pipe_gbuffer && ctx->d_lightmap == ctx->pipe_lightmap && ctx->pipe_w == f->width && ctx->pipe_h == f->height &&
                   ctx->pipe_gfmt == gfmt && ctx->pipe_lfmt == f->lightmap_format && ctx->pipe_nb == NB &&
                   ctx->gbuffer_capacity >= gtexel0 * (size_t)gw * (size_t)gh &&
                   ctx->d_lightmap_capacity >= std::max<size_t>(ltexel0 * (size_t)f->width * (size_t)std::max(rows, 1), 16);
    for (int b = 0; chained && b <= NB; b++) chained = ctx->pipe_edges[b] == edge[b];
    if (getenv("ILB_NO_FRAME_CHAINING")) chained = false;   // dev knob
    if (!chained) {
        const int rcd = ilb_frames_drain(ctx);
        if (rcd) return rcd;
    }
    ctx->pipe_generation = ~0ull;   // valid again only when this frame is queued completely
    const size_t gtexel = ilb_format_bytes(gfmt), ltexel = ilb_format_bytes(f->lightmap_format);
    const size_t gbytes = gtexel * (size_t)gw * (size_t)gh;
    if (!ctx->gbuffer_owned) { ctx->gbuffer = nullptr; ctx->gbuffer_capacity = 0; }
    int rc = ilb_reserve(ctx, &ctx->gbuffer, &ctx->gbuffer_capacity, gbytes, false);
    if (rc) return rc;
    ctx->gbuffer_owned = true;
    ctx->gb_w = gw; ctx->gb_h = gh; ctx->gb_fmt = gfmt;
    const size_t lbytes = ltexel * (size_t)f->width * (size_t)std::max(rows, 1);
    rc = ilb_reserve(ctx, &ctx->d_lightmap, &ctx->d_lightmap_capacity, std::max<size_t>(lbytes, 16), false);
    if (rc) return rc;
    if (!ctx->copy_in) {
        ILB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        ILB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
        for (int i = 0; i < ILB_PIPELINE_BANDS; i++) {
            ILB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
            ILB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
        }
    }
    if (!ctx->ev_frame[0]) {
        for (cudaEvent_t& e : ctx->ev_down) ILB_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (cudaEvent_t& e : ctx->ev_frame) ILB_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    LightingPrepared prep;
    rc = lightingPrepare(ctx, df, f, batches, batch_count, vertices, vertex_count, &prep);
    if (rc) return rc;
    if (rows == 0) return ILB_OK;
    // Two compute lanes: even bands run on the context's stream, odd bands on band_stream with their own scratch sums, so the
    // CTAs of band b + 1 fill the SM slots that the last wave of band b leaves idle (kernels of one stream run back to back,
    // and every band would otherwise pay the tails of both of its passes).  ILB_BAND_LANES=1 restores the single lane.
    int lanes = 2;
    if (const char* e = getenv("ILB_BAND_LANES")) lanes = atoi(e) >= 2 ? 2 : 1;
    const bool splitFrame = prep.nline > 0 && prep.nline < prep.nlights;
    if (ctx->opt[ILB_OPT_LIGHT_CONCURRENT] != 0 && splitFrame) lanes = 1;   // the co-resident pass mode owns both scratch buffers
    if (lanes == 2) {
        if (!ctx->band_stream) {
            ILB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->band_stream, cudaStreamNonBlocking));
            ILB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_band_fork, cudaEventDisableTiming));
            ILB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_band_join, cudaEventDisableTiming));
        }
        if (splitFrame) {   // both scratch buffers at their final size before anything is in flight
            int maxRows = 0;
            for (int b = 0; b < NB; b++) maxRows = std::max(maxRows, edge[b + 1] - edge[b]);
            const size_t bytes = sizeof(float4) * (size_t)f->width * (size_t)std::max(maxRows, 1);
            rc = ilb_reserve(ctx, &ctx->d_accum, &ctx->d_accum_capacity, bytes, false);
            if (rc) return rc;
            rc = ilb_reserve(ctx, &ctx->d_accum2, &ctx->d_accum2_capacity, bytes, false);
            if (rc) return rc;
        }
        // the second lane starts behind everything the frame's preparation queued on the context's stream (light records)
        ILB_CUDA(ctx, cudaEventRecord(ctx->ev_band_fork, ctx->stream));
        ILB_CUDA(ctx, cudaStreamWaitEvent(ctx->band_stream, ctx->ev_band_fork, 0));
    }
    const char* gsrc = reinterpret_cast<const char*>(gbuffer_host);
    char* gdst = reinterpret_cast<char*>(ctx->gbuffer);
    char* ldev = reinterpret_cast<char*>(ctx->d_lightmap);
    char* lhost = reinterpret_cast<char*>(lightmap_out_host);
    if (!chained) {   // everything already queued on the main stream (uploads, other renders) must be done before the G-buffer is overwritten
        ILB_CUDA(ctx, cudaEventRecord(ctx->ev_done[0], ctx->stream));
        ILB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, ctx->ev_done[0], 0));
    }
    for (int b = 0; b < NB; b++) {
        const int r0 = edge[b], r1 = edge[b + 1];
        if (r1 <= r0) continue;
        // chained: the rows of band b are free as soon as the kernels of band b of the frame before have run (the waits are queued
        // before this frame records ev_done[b] again, so they refer to that frame)
        if (chained) ILB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, ctx->ev_done[b], 0));
        const size_t goff = gtexel * (size_t)gw * (size_t)r0, gn = gtexel * (size_t)gw * (size_t)(r1 - r0);
        ILB_CUDA(ctx, cudaMemcpyAsync(gdst + goff, gsrc + goff, gn, cudaMemcpyHostToDevice, ctx->copy_in));
        ILB_CUDA(ctx, cudaEventRecord(ctx->ev_in[b], ctx->copy_in));
    }
    // every band's kernels are queued before the first download: with PAGEABLE caller buffers cudaMemcpyAsync blocks the host
    // until its copy is done, which must not hold back the launches of the bands behind it (pinned buffers never block)
    for (int b = 0; b < NB; b++) {
        const int r0 = edge[b], r1 = edge[b + 1];
        if (r1 <= r0) continue;
        const int lane = (lanes == 2) ? (b & 1) : 0;
        const cudaStream_t st = lane ? ctx->band_stream : ctx->stream;
        ILB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_in[b], 0));
        if (chained) ILB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_down[b], 0));   // the lightmap rows of band b have been downloaded
        void* outs[1] = {ctx->d_lightmap};
        rc = lightingLaunchRows(ctx, prep, r0, r1, outs, 1, f->row_begin, lane);
        if (rc) return rc;
        ILB_CUDA(ctx, cudaEventRecord(ctx->ev_done[b], st));
    }
    if (lanes == 2) {   // later work on the context's stream is ordered behind both lanes
        ILB_CUDA(ctx, cudaEventRecord(ctx->ev_band_join, ctx->band_stream));
        ILB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_band_join, 0));
    }
    for (int b = 0; b < NB; b++) {
        const int r0 = edge[b], r1 = edge[b + 1];
        if (r1 <= r0) continue;
        ILB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->ev_done[b], 0));
        const size_t loff = ltexel * (size_t)f->width * (size_t)(r0 - f->row_begin), ln = ltexel * (size_t)f->width * (size_t)(r1 - r0);
        ILB_CUDA(ctx, cudaMemcpyAsync(lhost + loff, ldev + loff, ln, cudaMemcpyDeviceToHost, ctx->copy_out));
        ILB_CUDA(ctx, cudaEventRecord(ctx->ev_down[b], ctx->copy_out));
    }
    const unsigned long long ticket = ++ctx->frame_ticket;
    ILB_CUDA(ctx, cudaEventRecord(ctx->ev_frame[ticket % 4], ctx->copy_out));
    ctx->pipe_generation = ctx->gb_generation;
    ctx->pipe_gbuffer = ctx->gbuffer; ctx->pipe_lightmap = ctx->d_lightmap;
    ctx->pipe_w = f->width; ctx->pipe_h = f->height; ctx->pipe_gfmt = gfmt; ctx->pipe_lfmt = f->lightmap_format; ctx->pipe_nb = NB;
    for (int b = 0; b <= NB; b++) ctx->pipe_edges[b] = edge[b];
    if (out_ticket) { *out_ticket = ticket; return ILB_OK; }   // asynchronous: ilb_lighting_frame_wait(ticket)
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_out));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->frame_waited = ticket;
    return ILB_OK;
}

int ilb_lighting_frame_wait(ilb_ctx* ctx, unsigned long long ticket) {
    if (ticket == 0 || ticket <= ctx->frame_waited) return ILB_OK;   // nothing was queued (an empty band), or known complete
    if (ticket > ctx->frame_ticket) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "unknown frame ticket %llu", ticket);
    // downloads complete in frame order, so a slot that a later frame has taken over still covers this one
    ILB_CUDA(ctx, cudaEventSynchronize(ctx->ev_frame[ticket % 4]));
    ctx->frame_waited = std::max(ctx->frame_waited, ticket);
    return ILB_OK;
}

int ilb_frames_drain(ilb_ctx* ctx) {
    if (ctx->frame_waited >= ctx->frame_ticket) return ILB_OK;
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_out));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->frame_waited = ctx->frame_ticket;
    return ILB_OK;
}


int ilb_probes_launch(ilb_ctx* ctx, ilb_df* df, const ilb_lighting_frame* f, const ilb_light_batch* batches, int batch_count,
                      const ilb_light_vertex* vertices, int vertex_count, const ilb_float4* positions, const ilb_float4* normals,
                      int probe_count, int output_format, void* probes_out_host, void* d_probes_out) {
    if (!f || !positions || !normals || (!probes_out_host && !d_probes_out) || probe_count < 0) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "null argument");
    if (output_format != ILB_FORMAT_FLOAT4 && output_format != ILB_FORMAT_HALF4) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "probe output must be FLOAT4 or HALF4");
    if (df && df->ctx != ctx) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "distance field belongs to another context");
    if (probe_count == 0) return ILB_OK;
    ilb_lighting_frame fr = *f;  // coverage is irrelevant for probes; guard the scale check
    if (!(fr.GBufferTexelSizeAndMisc.z > 0)) fr.GBufferTexelSizeAndMisc.z = 1;
    if (!(fr.GBufferTexelSizeAndMisc.w > 0)) fr.GBufferTexelSizeAndMisc.w = 1;
    std::vector<DLight> lights;
    std::vector<DLine> lines;
    const ilb_df_uniforms* geometry = nullptr;
    int rc = flattenLights(ctx, df, &fr, batches, batch_count, vertices, vertex_count, lights, lines, &geometry, false);
    if (rc) return rc;
    ProbeParams P;
    memset(&P, 0, sizeof(P));
    rc = uploadLights(ctx, lights, lines, 0, &P.lights, &P.lines);
    if (rc) return rc;
    const size_t in_bytes = sizeof(float4) * (size_t)probe_count, out_bytes = ilb_format_bytes(output_format) * (size_t)probe_count;
    rc = ilb_reserve(ctx, &ctx->d_probe_in, &ctx->d_probe_in_capacity, 2 * in_bytes + out_bytes, false);
    if (rc) return rc;
    char* base = reinterpret_cast<char*>(ctx->d_probe_in);
    ILB_CUDA(ctx, cudaMemcpyAsync(base, positions, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    ILB_CUDA(ctx, cudaMemcpyAsync(base + in_bytes, normals, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (geometry) {
        if (!ilb_make_df_geometry(df, *geometry, &P.df)) return ilb_fail(ctx, ILB_ERR_INVALID_ARGUMENT, "bad distance-field uniforms");
        rc = ilb_planes_attach(ctx, df, *geometry, &P.df);
        if (rc) return rc;
    }
    P.lightOcclusion = f->EnvironmentZToY.z;
    P.nlights = (int)lights.size();
    P.positions = reinterpret_cast<const float4*>(base);
    P.normals = reinterpret_cast<const float4*>(base + in_bytes);
    P.nprobes = probe_count;
    P.out_format = output_format;
    P.out = d_probes_out ? d_probes_out : base + 2 * in_bytes;
    if (P.df.planes) probe_accumulate_kernel<1><<<probe_count, PROBE_THREADS, 0, ctx->stream>>>(P);
    else probe_accumulate_kernel<0><<<probe_count, PROBE_THREADS, 0, ctx->stream>>>(P);
    ctx->launches++;
    ILB_CUDA(ctx, cudaGetLastError());
    if (d_probes_out) return ILB_OK;  // asynchronous: the caller reads the texels in stream order
    ILB_CUDA(ctx, cudaMemcpyAsync(probes_out_host, P.out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ILB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ILB_OK;
}
