// TEST INFRASTRUCTURE -- CPU oracle for the PARTICLE hot path (P1-P10 of SURVEY.md section 8).
//
// fp32 restatement of the reference particle shaders in the reference's own multi-pass form: spawner
// passes write in place, then one full-chunk pass per transform with ping-pong buffers (first pass's
// destination cleared), then the final Update pass with 4 outputs.  Only tests/, smoke() and bench.py's
// cpu_baseline / --impl reference legs may call this; the product never does.
// PARITY UNPINNED by reference outputs (no reference tests exist); pinned by the reference's own CPU
// mirror of Bezier.fxh (Bezier.cs ClampedBezier1/4.Evaluate) and closed-form cases (tests/test_oracle_kat.py).
//
// Citations are relative to /root/reference/Illuminant/.
#include <omp.h>

#include <cstring>
#include <memory>
#include <vector>

#include "../include/illuminant_b200.h"
#include "hlsl.hpp"
#include "oracle.h"

using namespace hlsl;

namespace {

inline float4 f4(const ilb_float4& v) { return float4(v.x, v.y, v.z, v.w); }
inline float4 ld4(const float* p, size_t i) { return float4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]); }
inline void st4(float* p, size_t i, float4 v) { p[4 * i] = v.x; p[4 * i + 1] = v.y; p[4 * i + 2] = v.z; p[4 * i + 3] = v.w; }

const float VelocityConstantScale = 1000;

// ---------------------------------------------------------------- Bezier.fxh
float tForScaledBezier(float4 rangeAndCount, float value, float& t) {  // Shaders/Bezier.fxh:21-63
    float minValue = rangeAndCount.x, invDivisor = rangeAndCount.y;
    uint32_t mode = (uint32_t)fabsf(rangeAndCount.w);
    bool repeating = mode > 255, bouncing = mode > 511;
    t = (value - minValue) * fabsf(invDivisor);
    if (bouncing) {
        t *= 2;
        if (invDivisor < 0) t = 2 - fmod(t, 2); else t = fmod(t, 2);
        if (t > 1) t = 1 - (t - 1);
    } else if (repeating) {
        if (invDivisor < 0) t = 1 - fmod(t, 1); else t = fmod(t, 1);
    } else {
        if (invDivisor < 0) t = 1 - saturate(t); else t = saturate(t);
    }
    switch (mode % 256) {
        default: break;
        case 1: t = dm_sinf(t * PI * 0.5f); break;
        case 2: t = t * t; break;
    }
    return rangeAndCount.z;
}

template <class T>
T evaluateBezierAtT(T a, T b, T c, T d, float count, float t) {  // Bezier.fxh:65-95, :141-171
    if (count <= 1.5f) return a;
    T ab = lerp(a, b, t);
    if (count <= 2.5f) return ab;
    if (count <= 3.5f) {
        if (t <= 0) return a;
        else if (t >= 1) return c;
        else return b;
    }
    T bc = lerp(b, c, t);
    T abbc = lerp(ab, bc, t);
    T cd = lerp(c, d, t);
    T bccd = lerp(bc, cd, t);
    return lerp(abbc, bccd, t);
}

float evaluateBezier1(const ilb_bezier1& b, float value) {  // :97-101
    float t;
    float count = tForScaledBezier(f4(b.RangeAndCount), value, t);
    return evaluateBezierAtT<float>(b.ABCD.x, b.ABCD.y, b.ABCD.z, b.ABCD.w, count, t);
}
float4 evaluateBezier4(const ilb_bezier4& b, float value) {  // :173-177
    float t;
    float count = tForScaledBezier(f4(b.RangeAndCount), value, t);
    return evaluateBezierAtT<float4>(f4(b.A), f4(b.B), f4(b.C), f4(b.D), count, t);
}

// ---------------------------------------------------------------- DistanceFunctionCommon.fxh (area weights)
float4 qmul(float4 q1, float4 q2) {  // :15-20
    return float4(q2.xyz() * q1.w + q1.xyz() * q2.w + cross(q1.xyz(), q2.xyz()), q1.w * q2.w - dot(q1.xyz(), q2.xyz()));
}
float3 rotateLocalPosition(float3 localPosition, float4 rotation) {  // :23-26
    float4 r_c = rotation * float4(-1, -1, -1, 1);
    return qmul(rotation, qmul(float4(localPosition, 0), r_c)).xyz();
}
float4 opElongate(float3 p, float3 h) {  // :43-46
    float3 q = abs(p) - h;
    return float4(sign(p) * max(q, float3(0.0f)), min(max(q.x, max(q.y, q.z)), 0.0f));
}
float evaluateBox(float3 worldPosition, float3 center, float3 size, float4 rotation) {  // :48-63
    float3 position = rotateLocalPosition(worldPosition - center, rotation);
    float3 d = abs(position) - size;
    return min(max(d.x, max(d.y, d.z)), 0.0f) + length(max(d, float3(0.0f)));
}
float evaluateSpheroid(float3 worldPosition, float3 center, float3 size, float4 rotation) {  // :65-75
    float3 position = rotateLocalPosition(worldPosition - center, rotation);
    float minSize = min(size.x, min(size.y, size.z));
    float3 elongation = size - minSize;
    float4 w = opElongate(position, elongation);
    return w.w + (length(w.xyz()) - minSize);
}
float evaluateEllipsoid(float3 worldPosition, float3 center, float3 size, float4 rotation) {  // :92-108
    float3 p = rotateLocalPosition(worldPosition - center, rotation), r = size;
    float k0 = length(p / r);
    float k1 = length(p / (r * r));
    return (k0 < 1.0f) ? (k0 - 1.0f) * min(min(r.x, r.y), r.z) : k0 * (k0 - 1.0f) / k1;
}
float sdCappedCylinder(float3 p, float h, float r) {  // :110-113
    float2 d = abs(float2(length(p.xy()), p.z)) - float2(r, h);
    return min(max(d.x, d.y), 0.0f) + length(max(d, float2(0.0f)));
}
float evaluateCylinder(float3 worldPosition, float3 center, float3 size, float4 rotation) {  // :115-121
    float3 position = rotateLocalPosition(worldPosition - center, rotation);
    return sdCappedCylinder(position, size.z, length(size.xy()));
}
float sdOctogonPrism(float3 p, float r, float h) {  // :139-152
    const float3 k = float3(-0.9238795325f, 0.3826834323f, 0.4142135623f);
    p = abs(p);
    float2 pxy = p.xy();
    pxy -= 2.0f * min(dot(float2(k.x, k.y), pxy), 0.0f) * float2(k.x, k.y);
    pxy -= 2.0f * min(dot(float2(-k.x, k.y), pxy), 0.0f) * float2(-k.x, k.y);
    pxy -= float2(clamp(pxy.x, -k.z * r, k.z * r), r);
    float2 d = float2(length(pxy) * sign(pxy.y), p.z - h);
    return min(max(d.x, d.y), 0.0f) + length(max(d, float2(0.0f)));
}
float evaluateOctagon(float3 worldPosition, float3 center, float3 size, float4 rotation) {  // :154-165
    float3 position = rotateLocalPosition(worldPosition - center, rotation);
    float minSize = min(size.x, size.y);
    float3 elongation = float3(size.xy() - minSize, 0);
    float4 w = opElongate(position, elongation);
    return w.w + sdOctogonPrism(w.xyz(), minSize, size.z);
}
}  // namespace

float orc_evaluate_by_type_id(int typeId, const float* wp, const float* c, const float* s, const float* rot) {  // :167-186
    float3 worldPosition(wp[0], wp[1], wp[2]), center(c[0], c[1], c[2]), size(s[0], s[1], s[2]);
    float4 rotation(rot[0], rot[1], rot[2], rot[3]);
    switch (typeId < 0 ? -typeId : typeId) {
        case 1: return evaluateEllipsoid(worldPosition, center, size, rotation);
        case 2: return evaluateBox(worldPosition, center, size, rotation);
        case 3: return evaluateCylinder(worldPosition, center, size, rotation);
        case 4: return evaluateSpheroid(worldPosition, center, size, rotation);
        case 5: return evaluateOctagon(worldPosition, center, size, rotation);
        default: return 0;
    }
}

namespace {

// computeWeight (FMA.fx:15-20, Noise.fx:21-26): the scalar AreaRotation is broadcast into the float4 quaternion
float computeWeight(const ilb_area& a, float3 worldPosition) {
    float rot[4] = {a.AreaRotation, a.AreaRotation, a.AreaRotation, a.AreaRotation};
    float wp[3] = {worldPosition.x, worldPosition.y, worldPosition.z};
    float distance = orc_evaluate_by_type_id(a.AreaType, wp, a.AreaCenter, a.AreaSize, rot);
    return (1 - saturate(distance / a.AreaFalloff)) * a.Strength;
}

bool checkCategoryFilter(float type, const float* typeMinMax) {  // ParticleCommon.fxh:198-200
    return (type >= typeMinMax[0]) && (type <= typeMinMax[1]);
}

// ---------------------------------------------------------------- system uniforms (ParticleCommon.fxh:29-92)
struct System {
    const ilb_psys_uniforms& u;
    explicit System(const ilb_psys_uniforms& u_) : u(u_) {}
    float getDeltaTimeSeconds() const { return u.GlobalSettings.x / VelocityConstantScale; }
    float getDeltaTime() const { return u.GlobalSettings.x; }
    float getFriction() const { return u.GlobalSettings.y; }
    float getMaximumVelocity() const { return u.GlobalSettings.z; }
    float getLifeDecayRate() const { return u.GlobalSettings.w; }
    float getEscapeVelocity() const { return u.CollisionSettings.x; }
    float getBounceVelocityMultiplier() const { return u.CollisionSettings.y; }
    float getCollisionDistance() const { return u.CollisionSettings.z; }
    float getCollisionLifePenalty() const { return u.CollisionSettings.w; }
    float getVelocityRotation() const { return u.AnimationRateAndRotationAndZToY.z; }
};

// ---------------------------------------------------------------- randomness (RandomCommon.fxh:17-34)
struct Randomness {
    const float* table;
    int w, h;
    // POINT sampled, WRAP/WRAP: texel = floor(uv * size) mod size
    float4 fetch(float2 uv) const {
        int ix = (int)floorf(uv.x * (float)w), iy = (int)floorf(uv.y * (float)h);
        ix %= w; if (ix < 0) ix += w;
        iy %= h; if (iy < 0) iy += h;
        return ld4(table, (size_t)iy * w + ix);
    }
    float4 randomCustom(float2 xy, float2 offset, float2 rate, float2 texel) const {
        float2 uv = ((xy * rate) + offset) * texel;
        return fetch(uv);
    }
};

// ---------------------------------------------------------------- spawner (SpawnerCommon.fxh, SpawnParticles.fx:10-30)
float3 generateRandomNormal3(float2 randomness) {  // :47-57
    float phi = randomness.x * PI * 2;
    float costheta = (randomness.y - 0.5f) * 2;
    float theta = dm_acosf(costheta);
    return float3(dm_sinf(theta) * dm_cosf(phi), dm_sinf(theta) * dm_sinf(phi), dm_cosf(theta));
}

float4 evaluateFormula(const ilb_spawn& s, float4 origin, float4 constant, float4 scale, float4 offset,
                       float4 randomness, float type) {  // :59-104
    float4 nonCircular = (randomness + offset) * scale;
    float4 type0 = constant + nonCircular;
    uint32_t itype = (uint32_t)fabsf(floorf(type));
    switch (itype) {
        default:
        case 0: return type0;
        case 3:
        case 1: {
            float3 axisMask(s.AxisMask[0], s.AxisMask[1], s.AxisMask[2]);
            float3 randomNormal = normalize(generateRandomNormal3(randomness.xy()) * axisMask);
            float3 circular = float3(randomNormal.x * randomness.z * scale.x, randomNormal.y * randomness.z * scale.y,
                                     randomNormal.z * randomness.z * scale.z);
            float3 result;
            if (itype == 3) {
                const float sqrt2 = 1.41421356237f;
                float3 edge = abs(offset.xyz());
                result = clamp(offset.xyz() * randomNormal * sqrt2, -edge, edge);
                result += constant.xyz() + circular;
            } else {
                circular += randomNormal * offset.xyz();
                result = constant.xyz() + circular;
            }
            return float4(result, type0.w);
        }
        case 2: {
            float3 distance = (constant - origin).xyz();
            float ldistance = length(distance);
            if (ldistance < 0.1f) return float4(0, 0, 0, constant.w);
            float3 direction = distance / ldistance;
            float3 randomSpeed = (randomness.x * scale.xyz() * direction);
            float3 fixedSpeed = (offset.xyz() * direction);
            return float4(randomSpeed + fixedSpeed, type0.w);
        }
    }
}

// evaluateRandomForIndex :106-117 ; random() = randomCustom(xy, RandomnessOffset, 1)
void evaluateRandomForIndex(const ilb_spawn& s, const Randomness& rng, float index, float4& random1, float4& random2, float4& random3) {
    float2 ro(s.RandomnessOffset[0], s.RandomnessOffset[1]), rt(s.RandomnessTexel[0], s.RandomnessTexel[1]);
    float2 randomOffset1 = float2(fmod(index, 8039), 0 + fmod(index, 57));
    float2 randomOffset2 = float2(fmod(index, 6180), 1 + fmod(index, 4031));
    float2 randomOffset3 = float2(fmod(index, 2025), 2 + fmod(index, 65531));
    random1 = rng.randomCustom(randomOffset1, ro, float2(1.0f), rt);
    random2 = rng.randomCustom(randomOffset2, ro, float2(1.0f), rt);
    random3 = rng.randomCustom(randomOffset3, ro, float2(1.0f), rt);
    if (s.AlignVelocityAndPosition != 0) { random2.x = random1.x; random2.y = random1.y; }
}

// Spawn_Stage1 :119-155.  Returns false when the texel is outside [first, last] (discard).
bool Spawn_Stage1(const ilb_spawn& s, const Randomness& rng, float2 xy, float4& random1, float4& random2, float4& random3,
                  int& index1, int& index2, float& positionIndexT) {
    float4 csi = f4(s.ChunkSizeAndIndices);
    float index = (xy.x) + (xy.y * csi.x);
    if ((index < csi.y) || (index > csi.z)) return false;

    evaluateRandomForIndex(s, rng, index, random1, random2, random3);

    float relativeIndex = (index - csi.y);
    if (s.PolygonRate > 0.05f) {
        float polyRate = s.PolygonRate;
        float positionIndexF = (relativeIndex / polyRate) + csi.w;
        float divisor = s.PositionConstantCount;
        float positionIndexI;
        positionIndexT = modff(positionIndexF, &positionIndexI);
        if (s.PolygonLoop != 0) {
            index1 = (int)fmod(positionIndexI, divisor);
            index2 = (int)fmod(positionIndexI + 1, divisor);
        } else {
            index1 = (int)fmod(positionIndexI, divisor);
            index2 = (int)min((float)(index1 + 1), divisor - 1);
        }
    } else {
        index1 = index2 = (int)fmod(relativeIndex + csi.w, s.PositionConstantCount);
        positionIndexT = 0;
    }
    return true;
}

// Spawn_Stage2 :157-190.  Returns false when the new particle is below the alpha threshold (discard).
bool Spawn_Stage2(const ilb_spawn& s, float4 positionConstant, float4 towardsNext, float4 random1, float4 random2, float4 random3,
                  float4& newPosition, float4& newVelocity, float4& newAttributes) {
    const ilb_float4* C = s.Configuration;
    float4 ft = f4(s.FormulaTypes);
    float4 tempPosition = evaluateFormula(s, float4(0.0f), positionConstant, f4(C[0]), f4(C[1]), random1, ft.x);
    newPosition = mul(float4(tempPosition.xyz(), 1), s.PositionMatrix);
    newPosition.w = tempPosition.w;

    float4 tempVelocity = evaluateFormula(s, tempPosition, f4(C[2]), f4(C[3]), f4(C[4]), random2, ft.y);
    newAttributes = evaluateFormula(s, float4(0.0f), f4(C[5]), f4(C[6]), f4(C[7]), random3, ft.z);

    float towardsDistance = length(towardsNext);
    if (towardsDistance > 0.0001f) {
        float towardsSpeed = evaluateFormula(s, float4(0.0f), float4(C[8].x), float4(C[8].y), float4(C[8].z), float4(random3.w), ft.w).x;
        tempVelocity += towardsSpeed * (towardsNext / towardsDistance);
    }
    newVelocity = mul(float4(tempVelocity.xyz(), 1), s.VelocityMatrix);
    newVelocity.w = tempVelocity.w;
    // (#if FNA zero-velocity hack :182-186 is compiled out on the XNA/D3D build this oracle restates)
    if (newAttributes.w < s.AttributeDiscardThreshold) return false;
    return true;
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// PS_Spawn SpawnParticles.fx:10-30
bool PS_Spawn(const ilb_spawn& s, const Randomness& rng, float2 xy, float4& newPosition, float4& newVelocity,
              float4& newAttributes) {
    int index1, index2;
    float positionIndexT;
    float4 random1, random2, random3;
    if (!Spawn_Stage1(s, rng, xy, random1, random2, random3, index1, index2, positionIndexT)) return false;
    index1 = clampi(index1, 0, 3);
    index2 = clampi(index2, 0, 3);
    float4 position1 = f4(s.InlinePositionConstants[index1]), position2 = f4(s.InlinePositionConstants[index2]);
    float4 positionConstant = lerp(position1, position2, positionIndexT);
    return Spawn_Stage2(s, positionConstant, position2 - position1, random1, random2, random3, newPosition, newVelocity, newAttributes);
}

// PS_SpawnFromPositionTexture SpawnParticles.fx:32-52.  The PositionBuffer (ParticleSpawner.cs:306-319) is a W x 1 Vector4
// texture point-sampled at u = index * (1 / W): texel `index` (CLAMP addressing).
bool PS_SpawnFromPositionTexture(const ilb_spawn& s, const float* positions, int positionCount, const Randomness& rng, float2 xy,
                                 float4& newPosition, float4& newVelocity, float4& newAttributes) {
    int index1, index2;
    float positionIndexT;
    float4 random1, random2, random3;
    if (!Spawn_Stage1(s, rng, xy, random1, random2, random3, index1, index2, positionIndexT)) return false;
    float4 position1 = ld4(positions, clampi(index1, 0, positionCount - 1)), position2 = ld4(positions, clampi(index2, 0, positionCount - 1));
    float4 positionConstant = lerp(position1, position2, positionIndexT);
    return Spawn_Stage2(s, positionConstant, position2 - position1, random1, random2, random3, newPosition, newVelocity, newAttributes);
}

// PS_SpawnFeedback SpawnParticles.fx:54-120.  srcP / srcV / srcRC: the source chunk's PositionAndLife, Velocity and RenderColor
// (ParticleTransform.cs:122-139: AttributeTexture = SourceChunk.RenderColor), srcSize^2 float4 each; point-sampled at
// sourceXy * (1 / srcSize): texel (floor(sourceX), sourceY), CLAMP addressing.
bool PS_SpawnFeedback(const ilb_spawn& s, const ilb_spawn_source& f, const float* srcP, const float* srcV, const float* srcRC, int srcSize,
                      const Randomness& rng, float2 xy, float4& newPosition, float4& newVelocity, float4& newAttributes) {
    float4 csi = f4(s.ChunkSizeAndIndices);
    float index = (xy.x) + (xy.y * csi.x);
    if ((index < csi.y) || (index > csi.z)) return false;

    float sourceIndex = ((index - csi.y) / f.InstanceMultiplier) + f.FeedbackSourceIndex;
    float sourceY, sourceX = modff(sourceIndex / (float)srcSize, &sourceY) * (float)srcSize;
    const int tx = clampi((int)floorf(sourceX), 0, srcSize - 1), ty = clampi((int)sourceY, 0, srcSize - 1);
    const size_t si = (size_t)ty * srcSize + tx;
    float4 sourcePosition = ld4(srcP, si), sourceVelocity = ld4(srcV, si), sourceAttributes = ld4(srcRC, si);
    if ((sourcePosition.w <= f.SourceLifeRange[0]) || (sourcePosition.w >= f.SourceLifeRange[1])) return false;

    float4 random1, random2, random3;
    evaluateRandomForIndex(s, rng, index, random1, random2, random3);

    const ilb_float4* C = s.Configuration;
    float4 ft = f4(s.FormulaTypes);
    float4 positionConstant = f4(s.InlinePositionConstants[0]);
    if (f.AlignPositionConstant != 0) {
        positionConstant.x += sourcePosition.x; positionConstant.y += sourcePosition.y; positionConstant.z += sourcePosition.z;
    }
    float4 tempPosition = evaluateFormula(s, float4(0.0f), positionConstant, f4(C[0]), f4(C[1]), random1, ft.x);

    float4 attributeConstant = f4(C[5]);
    if (f.MultiplyAttributeConstant != 0) attributeConstant *= sourceAttributes;

    newPosition = mul(float4(tempPosition.xyz(), 1), s.PositionMatrix);
    newPosition.w = tempPosition.w;
    if (f.MultiplyLife != 0) newPosition.w *= sourcePosition.w;

    float4 velocityConstant = f4(C[2]);
    float4 tempVelocity = evaluateFormula(s, tempPosition, velocityConstant, f4(C[3]), f4(C[4]), random2, ft.y);
    tempVelocity += sourceVelocity * f.SourceVelocityFactor;

    newVelocity = mul(float4(tempVelocity.xyz(), 1), s.VelocityMatrix);
    newVelocity.w = tempVelocity.w;

    newAttributes = evaluateFormula(s, tempPosition, attributeConstant, f4(C[6]), f4(C[7]), random3, ft.z);
    if (newAttributes.w < s.AttributeDiscardThreshold) return false;
    return true;
}

// ---- PatternSpawner.fx.  The pattern texture with its mip chain: SurfaceFormat.Color levels, level k+1 = rounding 2x2 box
// filter of level k (the reference's mips come from its content loader; un-pinned).
struct PatternTexture {
    std::vector<std::vector<uint8_t>> levels;
    std::vector<int> w, h;
    PatternTexture(const uint8_t* texels, int w0, int h0) {
        levels.emplace_back(texels, texels + (size_t)w0 * h0 * 4);
        w.push_back(w0); h.push_back(h0);
        while (w.back() > 1 || h.back() > 1) {
            const int pw = w.back(), ph = h.back(), nw = pw > 1 ? pw / 2 : 1, nh = ph > 1 ? ph / 2 : 1;
            const std::vector<uint8_t>& src = levels.back();
            std::vector<uint8_t> dst((size_t)nw * nh * 4);
            for (int y = 0; y < nh; y++)
                for (int x = 0; x < nw; x++) {
                    const int x0 = 2 * x < pw ? 2 * x : pw - 1, x1 = 2 * x + 1 < pw ? 2 * x + 1 : pw - 1;
                    const int y0 = 2 * y < ph ? 2 * y : ph - 1, y1 = 2 * y + 1 < ph ? 2 * y + 1 : ph - 1;
                    for (int c = 0; c < 4; c++) {
                        const unsigned sum = src[((size_t)y0 * pw + x0) * 4 + c] + src[((size_t)y0 * pw + x1) * 4 + c] +
                                             src[((size_t)y1 * pw + x0) * 4 + c] + src[((size_t)y1 * pw + x1) * 4 + c];
                        dst[((size_t)y * nw + x) * 4 + c] = (uint8_t)((sum + 2) >> 2);
                    }
                }
            levels.push_back(std::move(dst));
            w.push_back(nw); h.push_back(nh);
        }
    }
    float4 texel(int l, int x, int y) const {
        x = x < 0 ? 0 : (x >= w[l] ? w[l] - 1 : x);
        y = y < 0 ? 0 : (y >= h[l] ? h[l] - 1 : y);
        const uint8_t* t = &levels[l][((size_t)y * w[l] + x) * 4];
        return float4(t[0] / 255.0f, t[1] / 255.0f, t[2] / 255.0f, t[3] / 255.0f);
    }
    // tex2Dlod with PatternSampler (PatternSpawner.fx:11-19): LINEAR min/mag, POINT mip, CLAMP
    float4 sample(float2 uv, float lod) const {
        int l = (int)floorf(lod + 0.5f);
        l = l < 0 ? 0 : (l >= (int)levels.size() ? (int)levels.size() - 1 : l);
        const float fx = uv.x * (float)w[l] - 0.5f, fy = uv.y * (float)h[l] - 0.5f;
        const float x0 = floorf(fx), y0 = floorf(fy);
        const float tx = fx - x0, ty = fy - y0;
        const float4 top = lerp(texel(l, (int)x0, (int)y0), texel(l, (int)x0 + 1, (int)y0), tx);
        const float4 bottom = lerp(texel(l, (int)x0, (int)y0 + 1), texel(l, (int)x0 + 1, (int)y0 + 1), tx);
        return lerp(top, bottom, ty);
    }
};

// PS_SpawnPattern PatternSpawner.fx:21-96
bool PS_SpawnPattern(const ilb_spawn& s, const ilb_spawn_source& f, const PatternTexture& tex, const Randomness& rng, float2 xy,
                     float4& newPosition, float4& newVelocity, float4& newAttributes) {
    float4 csi = f4(s.ChunkSizeAndIndices);
    float index = floorf(xy.x) + (floorf(xy.y) * csi.x);
    if ((index < csi.y) || (index > csi.z)) return false;

    float relativeIndex = floorf(index - csi.y);
    float particlesPerRow = f.StepWidthAndSizeScale.y;
    float2 indexXy = float2(floorf(fmod(relativeIndex, particlesPerRow)), floorf(relativeIndex / particlesPerRow));
    indexXy.y += f.YOffsetsAndCoordScale.x;
    float2 texCoordXy = (indexXy * float2(f.StepWidthAndSizeScale.z, f.StepWidthAndSizeScale.w)) + float2(f.TexelOffsetAndMipBias.x, f.TexelOffsetAndMipBias.y);
    texCoordXy.y += f.YOffsetsAndCoordScale.y;
    float2 positionXy = indexXy * float2(f.YOffsetsAndCoordScale.z, f.YOffsetsAndCoordScale.w) + float2(f.CenteringOffset[0], f.CenteringOffset[1]);
    if ((texCoordXy.x > 1) || (texCoordXy.y > 1)) return false;

    float4 patternColor = tex.sample(texCoordXy, f.TexelOffsetAndMipBias.w);

    float4 random1, random2, random3;
    evaluateRandomForIndex(s, rng, index, random1, random2, random3);

    const ilb_float4* C = s.Configuration;
    float4 ft = f4(s.FormulaTypes);
    float4 positionConstant = f4(s.InlinePositionConstants[0]);
    float4 tempPosition = evaluateFormula(s, float4(0.0f), positionConstant, f4(C[0]), f4(C[1]), random1, ft.x);
    tempPosition.x += positionXy.x;
    tempPosition.y += positionXy.y;

    float4 attributeConstant = patternColor;
    if (f.MultiplyAttributeConstant != 0) attributeConstant *= f4(C[5]);
    else attributeConstant += f4(C[5]);

    newPosition = mul(float4(tempPosition.xyz(), 1), s.PositionMatrix);
    newPosition.w = tempPosition.w;

    float4 velocityConstant = f4(C[2]);
    float4 tempVelocity = evaluateFormula(s, tempPosition, velocityConstant, f4(C[3]), f4(C[4]), random2, ft.y);
    newVelocity = mul(float4(tempVelocity.xyz(), 1), s.VelocityMatrix);
    newVelocity.w = tempVelocity.w;

    newAttributes = evaluateFormula(s, tempPosition, attributeConstant, f4(C[6]), f4(C[7]), random3, ft.z);
    if (newAttributes.w < s.AttributeDiscardThreshold) return false;
    return true;
}

// ---------------------------------------------------------------- transforms
void PS_Gravity(const System& sys, const ilb_gravity& g, float4& newPosition, float4 oldVelocity, float4& newVelocity) {  // Gravity.fx:12-61
    if ((newPosition.w <= 0) || !checkCategoryFilter(oldVelocity.w, g.CategoryFilter)) {
        newVelocity = oldVelocity;
        return;
    }
    float3 acceleration(0.0f);
    for (int i = 0; i < g.AttractorCount; i++) {
        float3 apos = f4(g.AttractorPositions[i]).xyz();
        float3 ars = f4(g.AttractorRadiusesAndStrengths[i]).xyz();
        float3 toCenter = (apos - newPosition.xyz());
        float attraction = 0;
        if (ars.z >= 0.5f) {
            float distance = length(toCenter);
            attraction = 1 - saturate(distance / ars.x);
            if (ars.z >= 1.5f) attraction *= attraction;
            attraction = attraction * sys.getDeltaTime() / VelocityConstantScale;
        } else {
            float distanceSquared = dot(toCenter, toCenter) - ars.x;
            distanceSquared = max(distanceSquared, 0.001f);
            attraction = 1 / distanceSquared;
        }
        float3 newAccel = normalize(toCenter) * attraction * ars.y;
        acceleration += newAccel;
    }
    float maximumAcceleration = g.MaximumAcceleration * sys.getDeltaTime() / VelocityConstantScale;
    float currentLength = length(acceleration);
    if (currentLength > maximumAcceleration) acceleration = normalize(acceleration) * maximumAcceleration;
    // float4 + float3 truncates to float3 (Gravity.fx:59)
    newVelocity = float4(min(float3(sys.getMaximumVelocity()), oldVelocity.xyz() + acceleration), oldVelocity.w);
}

void PS_Noise(const System& sys, const ilb_noise& n, const Randomness& rng, float2 xy, float4 oldPosition,
              float4 oldVelocity, float4& newPosition, float4& newVelocity) {  // Noise.fx:28-72
    if (!checkCategoryFilter(oldVelocity.w, n.area.CategoryFilter)) {
        newPosition = oldPosition;
        newVelocity = oldVelocity;
        return;
    }
    float weight = computeWeight(n.area, oldPosition.xyz());
    float t = weight * sys.getDeltaTime() / n.TimeDivisor;

    float2 ro(n.RandomnessOffset[0], n.RandomnessOffset[1]), nro(n.NextRandomnessOffset[0], n.NextRandomnessOffset[1]);
    float2 rt(n.RandomnessTexel[0], n.RandomnessTexel[1]);
    float4 randomP1 = rng.randomCustom(xy, ro, rt, rt);
    float4 randomP2 = rng.randomCustom(xy, nro, rt, rt);
    float4 randomV1 = rng.randomCustom(xy + float2(2, 1), ro, rt, rt);
    float4 randomV2 = rng.randomCustom(xy + float2(2, 1), nro, rt, rt);
    float4 randomP = lerp(randomP1, randomP2, n.FrequencyLerp);
    float4 randomV = lerp(randomV1, randomV2, n.FrequencyLerp);

    float4 positionDelta = (randomP + f4(n.PositionOffset));
    positionDelta = sign(positionDelta) * max(abs(positionDelta), f4(n.PositionMinimum));
    positionDelta *= f4(n.PositionScale);
    float4 velocityDelta = (randomV + f4(n.VelocityOffset));
    velocityDelta = sign(velocityDelta) * max(abs(velocityDelta), f4(n.VelocityMinimum));
    velocityDelta *= f4(n.VelocityScale);

    newPosition = lerp(oldPosition, oldPosition + positionDelta, t);
    float3 nv;
    if (n.ReplaceOldVelocity != 0)
        nv = lerp(oldVelocity.xyz(), velocityDelta.xyz(), weight);
    else
        nv = lerp(oldVelocity.xyz(), oldVelocity.xyz() + velocityDelta.xyz(), t);
    nv += normalize(oldVelocity.xyz()) * velocityDelta.w;
    newVelocity = float4(nv, oldVelocity.w);
}

void PS_FMA(const System& sys, const ilb_fma& f, float4 oldPosition, float4 oldVelocity, float4& newPosition,
            float4& newVelocity) {  // FMA.fx:22-51
    if ((oldPosition.w <= 0) || !checkCategoryFilter(oldVelocity.w, f.area.CategoryFilter)) {
        newPosition = oldPosition;
        newVelocity = oldVelocity;
        return;
    }
    float weight = computeWeight(f.area, oldPosition.xyz());
    float t = weight * sys.getDeltaTime() / f.TimeDivisor;
    newPosition = lerp(oldPosition, (oldPosition * f4(f.PositionMultiply)) + f4(f.PositionAdd), t);
    newVelocity = lerp(oldVelocity, (oldVelocity * f4(f.VelocityMultiply)) + f4(f.VelocityAdd), t);
}

float4 mul3(float4 oldValue, const float* mat, float w) {  // ParticleCommon.fxh:187-196
    float4 temp = mul(float4(oldValue.xyz(), 1), mat);
    float3 divided;
    if (w != 0) divided = temp.xyz() / temp.w;
    else divided = temp.xyz();
    return float4(divided, oldValue.w);
}

void PS_MatrixMultiply(const System& sys, const ilb_matrix_multiply& m, float4 oldPosition, float4 oldVelocity,
                       float4& newPosition, float4& newVelocity) {  // MatrixMultiply.fx:14-52
    if ((oldPosition.w <= 0) || !checkCategoryFilter(oldVelocity.w, m.area.CategoryFilter)) {
        newPosition = oldPosition;
        newVelocity = oldVelocity;
        return;
    }
    float timeScale = (m.TimeDivisor >= 0) ? sys.getDeltaTime() / m.TimeDivisor : 1;
    float w = computeWeight(m.area, oldPosition.xyz()) * timeScale;
    newPosition = lerp(oldPosition, mul3(oldPosition, m.PositionMatrix, 1), w);
    newVelocity = lerp(oldVelocity, mul3(oldVelocity, m.VelocityMatrix, 0), w);
}

// ---------------------------------------------------------------- update (UpdateCommon.fxh)
float3 applyFrictionAndMaximum(const System& sys, float3 velocity) {  // :20-35
    float l = length(velocity);
    if (l <= 0.001f) return float3(0.0f);
    if (l > sys.getMaximumVelocity()) l = sys.getMaximumVelocity();
    float friction = l * sys.getFriction();
    l -= (friction * sys.getDeltaTimeSeconds());
    l = clamp(l, 0, sys.getMaximumVelocity());
    return normalize(velocity) * l;
}

// LifeRampSampler (UpdateCommon.fxh:6-13): POINT filter, AddressU CLAMP, AddressV WRAP.  The texture is set once per
// system with orc_set_life_ramp (float4 texels, row-major).
static const float* g_lifeRamp = nullptr;
static int g_lifeRampW = 0, g_lifeRampH = 0;
float4 readLifeRamp(float u, float v) {  // :36-38
    if (!g_lifeRamp) return float4(1.0f);  // Engine.DummyRampTexture (ParticleSystem.cs:921-924): white
    int ix = (int)floorf(u * (float)g_lifeRampW);
    ix = ix < 0 ? 0 : (ix > g_lifeRampW - 1 ? g_lifeRampW - 1 : ix);
    const float fy = floorf(v * (float)g_lifeRampH);
    int iy = (int)(fy - floorf(fy / (float)g_lifeRampH) * (float)g_lifeRampH);
    iy = iy < 0 ? 0 : (iy > g_lifeRampH - 1 ? g_lifeRampH - 1 : iy);
    const float* t = g_lifeRamp + 4 * ((size_t)iy * g_lifeRampW + ix);
    return float4(t[0], t[1], t[2], t[3]);
}

float getRotationForVelocity(float3 velocity) {  // :82-95
    float2 absvel = abs(velocity.xy());
    if ((absvel.x < 0.01f) && (absvel.y < 0.01f)) return 0;
    float result = atan2f(velocity.y, velocity.x);
    if (result < 0) result += 2 * PI;
    return result;
}

void computeRenderData(const System& sys, float2 vpos, float4 position, float4 velocity, float4 attributes,
                       float4& renderColor, float4& renderData) {  // :97-117
    if (position.w <= 0) {
        renderColor = renderData = float4(0.0f);
        return;
    }
    float index = vpos.x + (vpos.y * 256);  // hard-coded 256 in the reference (:107)
    float velocityLength = max(length(velocity.xyz()), 0.0001f);
    // getRampedColorForLifeValueAndIndex :67-80
    float4 ramped = evaluateBezier4(sys.u.ColorFromLife, position.w);
    ramped *= evaluateBezier4(sys.u.ColorFromVelocity, velocityLength);
    const ilb_float4 lrs = sys.u.LifeRampSettings;
    if (lrs.x != 0) {
        float u = (position.w - lrs.y) / lrs.z;
        if (lrs.x < 0) u = 1 - saturate(u);
        float v = index / lrs.w;
        ramped = lerp(ramped, readLifeRamp(u, v) * ramped, saturate(fabsf(lrs.x)));
    }
    renderColor = attributes * ramped;
    renderColor.w = saturate(renderColor.w);
    renderColor.x *= renderColor.w; renderColor.y *= renderColor.w; renderColor.z *= renderColor.w;
    float size = evaluateBezier1(sys.u.SizeFromLife, position.w);
    size *= evaluateBezier1(sys.u.SizeFromVelocity, velocityLength);
    renderData.x = size;
    renderData.y = (getRotationForVelocity(velocity.xyz()) * sys.getVelocityRotation()) +
                   ((position.w * sys.u.RotationFromLifeAndIndex[0]) + (index * sys.u.RotationFromLifeAndIndex[1]));
    renderData.z = velocityLength;
    renderData.w = velocity.w;
}

// returns false when discarded (dead): destination keeps its cleared zeros
bool PS_Update(const System& sys, float2 xy, float4 oldPosition, float4 oldVelocity, float4 attributes,
               float4& newPosition, float4& newVelocity, float4& renderColor, float4& renderData) {  // UpdateParticleSystem.fx:9-38
    if (oldPosition.w <= 0) return false;  // readStateOrDiscard ParticleCommon.fxh:162-181
    float3 velocity = applyFrictionAndMaximum(sys, oldVelocity.xyz());
    float3 scaledVelocity = velocity * sys.getDeltaTimeSeconds();
    float newLife = oldPosition.w - (sys.getLifeDecayRate() * sys.getDeltaTimeSeconds());
    if (newLife <= 0) {
        newPosition = float4(0.0f);
        newVelocity = float4(0.0f);
    } else {
        newPosition = float4(oldPosition.xyz() + scaledVelocity, newLife);
        newVelocity = float4(velocity, oldVelocity.w);
    }
    computeRenderData(sys, xy, newPosition, newVelocity, attributes, renderColor, renderData);
    return true;
}

}  // namespace

// The collision sampler is the lighting oracle's sampleDistanceFieldEx (same header in the reference).
extern "C" float orc_sample_distance_field(const uint16_t* df_tex, int tw, int th, const ilb_df_uniforms* u, float x, float y, float z);

namespace {

struct Field {
    const uint16_t* tex; int tw, th; const ilb_df_uniforms* u;
    float sample(float3 p) const { return orc_sample_distance_field(tex, tw, th, u, p.x, p.y, p.z); }
};

float3 estimateNormal4(const Field& df, float3 position) {  // VisualizeCommon.fxh:9-63
    const ilb_df_uniforms& u = *df.u;
    float3 texel(u.ConeAndMisc.w, u.StepAndMisc2.w, u.Extent.z / max(u.TextureSliceCount.w, 1));
    const float3 normalWeights[4] = {float3(1, -1, -1), float3(-1, -1, 1), float3(-1, 1, -1), float3(1, 1, 1)};
    float3 result(0.0f);
    for (int i = 0; i < 4; i++) {
        float3 weight = normalWeights[i];
        result += weight * df.sample(position + weight * texel);
    }
    return normalize(result);
}

bool PS_UpdateWithDistanceField(const System& sys, const Field& df, float2 xy, float4 oldPosition, float4 oldVelocity,
                                float4 attributes, float4& resultPosition, float4& newVelocity, float4& renderColor,
                                float4& renderData) {  // UpdateParticleSystemWithDistanceField.fx:29-147
    const int MAX_STEP_COUNT = 3;
    const float BOUNCE_DELAY = 3, NO_NORMAL_THRESHOLD = 0.33f;
    const float INITIAL_ESCAPE_SPEED = 0.33f, ESCAPE_SPEED_ACCELERATION = 1.1f;
    const float3 ESCAPE_MASK(1, 1, 0);

    resultPosition = newVelocity = renderColor = renderData = float4(0.0f);
    if (oldPosition.w <= 0) return false;  // readStateOrDiscard

    float newLife = oldPosition.w - (sys.getLifeDecayRate() * sys.getDeltaTimeSeconds());
    if (newLife <= 0) return true;  // zeros written (:45-51)

    float3 unitVector = normalize(oldVelocity.xyz());
    float3 velocity = applyFrictionAndMaximum(sys, oldVelocity.xyz());

    bool collided = false, escaping = false;
    float3 scaledVelocity = velocity * sys.getDeltaTimeSeconds();
    float3 previousPosition = oldPosition.xyz(), collisionPosition(0.0f), newPosition = previousPosition;

    float initialDistance = df.sample(oldPosition.xyz());
    bool wasColliding = initialDistance < sys.getCollisionDistance();
    float travelDistance = max(0, min(initialDistance, length(scaledVelocity)));
    int stepCount = MAX_STEP_COUNT;
    if (wasColliding) stepCount = 1;
    else if (travelDistance <= 0.001f) stepCount = 0;

    for (int i = 0; i < stepCount; i++) {
        float3 testPosition = oldPosition.xyz() + (travelDistance * unitVector);
        float stepDistance = df.sample(testPosition);
        if (stepDistance < sys.getCollisionDistance()) {
            collided = true;
            collisionPosition = testPosition;
        }
        escaping = stepDistance > initialDistance;
        if (collided && !escaping) {
            collisionPosition = testPosition;
            float offset = clamp(stepDistance + sys.getCollisionDistance(), 0.05f, 16);
            travelDistance = max(0, travelDistance - offset);
        } else
            stepCount = 0;
        if (travelDistance <= 0.001f) stepCount = 0;
    }

    if (collided) {
        bool bounce = oldVelocity.w <= 0;
        bool redirect = wasColliding && !escaping;
        float3 normal(0.0f);
        if (bounce || redirect) normal = estimateNormal4(df, collisionPosition);
        float escapeSpeed = min(sys.getMaximumVelocity(), sys.getEscapeVelocity());
        if (redirect) {
            normal *= ESCAPE_MASK;
            if (length(normal) < NO_NORMAL_THRESHOLD) {
                float a = (xy.x / 67) + (xy.y / 13);
                normal = float3(dm_sinf(a), dm_cosf(a), 0);  // sincos()
            }
            float3 escapeVector = normalize(normal);
            newVelocity = float4(escapeVector * escapeSpeed * INITIAL_ESCAPE_SPEED, BOUNCE_DELAY);
            float3 escapeDelta = newVelocity.xyz() * sys.getDeltaTimeSeconds();
            newPosition = oldPosition.xyz() + escapeDelta;
        } else if (bounce) {
            float3 bounceVector = -(2 * dot(normal, unitVector) * (normal - unitVector));
            if (length(bounceVector) < NO_NORMAL_THRESHOLD) bounceVector = -unitVector;
            else bounceVector = normalize(bounceVector);
            newPosition = collisionPosition;
            newVelocity = float4(bounceVector * (min(sys.getMaximumVelocity(), length(velocity) * sys.getBounceVelocityMultiplier())), BOUNCE_DELAY);
            newLife -= sys.getCollisionLifePenalty();
        } else {
            float currentSpeed = length(oldVelocity.xyz());
            float newSpeed = max(currentSpeed * ESCAPE_SPEED_ACCELERATION, escapeSpeed);
            float3 nv = unitVector * newSpeed;
            newVelocity = float4(nv, newVelocity.w);  // .w stays 0 from the initialiser (:36,:133)
            newPosition = oldPosition.xyz() + (travelDistance * unitVector);
        }
    } else {
        newVelocity = float4(velocity, max(oldVelocity.w - 1, 0));
        newPosition = oldPosition.xyz() + (travelDistance * unitVector);
    }

    if (newLife <= 0) {
        newPosition = float3(0.0f);
        newVelocity = float4(0.0f);
    }
    resultPosition = float4(newPosition, newLife);
    computeRenderData(sys, xy, resultPosition, newVelocity, attributes, renderColor, renderData);
    return true;
}

}  // namespace

extern "C" {

void orc_set_life_ramp(const float* texels, int w, int h) { g_lifeRamp = texels; g_lifeRampW = w; g_lifeRampH = h; }

float orc_bezier1(const ilb_bezier1* b, float value) { return evaluateBezier1(*b, value); }
void orc_bezier4(const ilb_bezier4* b, float value, float* out) {
    float4 r = evaluateBezier4(*b, value);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// One or more ParticleSystem.Update calls (Particles/ParticleSystem.cs:634-760) over chunks [0, live_chunks):
// spawners (in place, ParticleSystem.cs:725-741), then UpdateChunk (:791-856) per chunk.
// P, V, A, RC, RD: live_chunks * chunk_size^2 float4 each, chunk-major then row-major. In/out.
int orc_particles_step_sources(float* P, float* V, float* A, float* RC, float* RD, int chunk_size, int live_chunks,
                               const ilb_psys_uniforms* u, const ilb_spawn* spawns, const ilb_spawn_source* sources,
                               const orc_source_state* states, int nspawns, const ilb_op* ops, int nops,
                               const float* rng_table, int rw, int rh, const uint16_t* df_tex, int tw, int th, int steps,
                               int nthreads) {
    if (u->has_collision_field && !df_tex) return ILB_ERR_INVALID_OPERATION;
    if (nthreads > 0) omp_set_num_threads(nthreads);
    const System sys(*u);
    const Randomness rng{rng_table, rw, rh};
    const Field field{df_tex, tw, th, &u->CollisionField};
    const size_t per = (size_t)chunk_size * chunk_size, total = per * live_chunks;
    std::vector<float> P2(4 * total), V2(4 * total);
    float *pPrev = P, *vPrev = V, *pCurr = P2.data(), *vCurr = V2.data();

    for (int step = 0; step < steps; step++) {
        // the spawn list is one tick's spawns (ParticleSystem.Update runs its spawners once per update): first step only
        for (int si = 0; si < (step == 0 ? nspawns : 0); si++) {
            const ilb_spawn& s = spawns[si];
            if (s.chunk < 0 || s.chunk >= live_chunks) return ILB_ERR_INVALID_ARGUMENT;
            const int kind = sources ? sources[si].kind : ILB_SPAWN_INLINE;
            if (kind == ILB_SPAWN_INLINE && s.PositionConstantCount > 4) return ILB_ERR_UNSUPPORTED;
            if (kind == ILB_SPAWN_POSITION_TEXTURE && (!sources[si].positions || sources[si].position_count < 1)) return ILB_ERR_INVALID_ARGUMENT;
            if (kind == ILB_SPAWN_FEEDBACK && (!states || !states[si].P || states[si].chunk_size < 1)) return ILB_ERR_INVALID_ARGUMENT;
            if (kind == ILB_SPAWN_PATTERN && (!sources[si].pattern_texels || sources[si].pattern_width < 1 || sources[si].pattern_height < 1)) return ILB_ERR_INVALID_ARGUMENT;
            std::unique_ptr<PatternTexture> pattern;
            if (kind == ILB_SPAWN_PATTERN) pattern.reset(new PatternTexture(sources[si].pattern_texels, sources[si].pattern_width, sources[si].pattern_height));
            const size_t base = per * s.chunk;
#pragma omp parallel for schedule(static)
            for (long i = 0; i < (long)per; i++) {
                float2 xy((float)(i % chunk_size), (float)(i / chunk_size));
                float4 np, nv, na;
                bool spawned;
                if (kind == ILB_SPAWN_POSITION_TEXTURE)
                    spawned = PS_SpawnFromPositionTexture(s, &sources[si].positions[0].x, sources[si].position_count, rng, xy, np, nv, na);
                else if (kind == ILB_SPAWN_PATTERN)
                    spawned = PS_SpawnPattern(s, sources[si], *pattern, rng, xy, np, nv, na);
                else if (kind == ILB_SPAWN_FEEDBACK)
                    spawned = PS_SpawnFeedback(s, sources[si], states[si].P, states[si].V, states[si].RC, states[si].chunk_size, rng, xy, np, nv, na);
                else
                    spawned = PS_Spawn(s, rng, xy, np, nv, na);
                if (spawned) {
                    st4(pPrev, base + i, np);
                    st4(vPrev, base + i, nv);
                    st4(A, base + i, na);
                }
            }
        }
        // one full pass per transform, ping-pong (RotateBuffers, ParticleSystem.cs:602-616)
        for (int oi = 0; oi < nops; oi++) {
            const ilb_op& op = ops[oi];
#pragma omp parallel for schedule(static)
            for (long gi = 0; gi < (long)total; gi++) {
                long i = gi % (long)per;
                float2 xy((float)(i % chunk_size), (float)(i / chunk_size));
                float4 op_ = ld4(pPrev, gi), ov = ld4(vPrev, gi), np, nv;
                switch (op.kind) {
                    case ILB_OP_GRAVITY: np = op_; PS_Gravity(sys, op.u.gravity, np, ov, nv); break;
                    case ILB_OP_NOISE: PS_Noise(sys, op.u.noise, rng, xy, op_, ov, np, nv); break;
                    case ILB_OP_FMA: PS_FMA(sys, op.u.fma, op_, ov, np, nv); break;
                    case ILB_OP_MATRIX_MULTIPLY: PS_MatrixMultiply(sys, op.u.matrix, op_, ov, np, nv); break;
                    default: np = op_; nv = ov; break;
                }
                st4(pCurr, gi, np);
                st4(vCurr, gi, nv);
            }
            std::swap(pPrev, pCurr);
            std::swap(vPrev, vCurr);
        }
        // final update pass: destination cleared to 0 (shouldClear = true, ParticleSystem.cs:843,852), dead texels discard
#pragma omp parallel for schedule(static)
        for (long gi = 0; gi < (long)total; gi++) {
            long i = gi % (long)per;
            float2 xy((float)(i % chunk_size), (float)(i / chunk_size));
            float4 op_ = ld4(pPrev, gi), ov = ld4(vPrev, gi), at = ld4(A, gi);
            float4 np(0.0f), nv(0.0f), rc(0.0f), rd(0.0f);
            bool written = u->has_collision_field
                               ? PS_UpdateWithDistanceField(sys, field, xy, op_, ov, at, np, nv, rc, rd)
                               : PS_Update(sys, xy, op_, ov, at, np, nv, rc, rd);
            if (!written) np = nv = rc = rd = float4(0.0f);
            st4(pCurr, gi, np);
            st4(vCurr, gi, nv);
            if (u->write_render_outputs) {
                st4(RC, gi, rc);
                st4(RD, gi, rd);
            }
        }
        std::swap(pPrev, pCurr);
        std::swap(vPrev, vCurr);
    }
    if (pPrev != P) {
        memcpy(P, pPrev, sizeof(float) * 4 * total);
        memcpy(V, vPrev, sizeof(float) * 4 * total);
    }
    return 0;
}

int orc_particles_step(float* P, float* V, float* A, float* RC, float* RD, int chunk_size, int live_chunks,
                       const ilb_psys_uniforms* u, const ilb_spawn* spawns, int nspawns, const ilb_op* ops, int nops,
                       const float* rng_table, int rw, int rh, const uint16_t* df_tex, int tw, int th, int steps,
                       int nthreads) {
    return orc_particles_step_sources(P, V, A, RC, RD, chunk_size, live_chunks, u, spawns, nullptr, nullptr, nspawns, ops, nops, rng_table,
                                      rw, rh, df_tex, tw, th, steps, nthreads);
}


// ---------------------------------------------------------------- N2: ParticleSystem.Render (RasterizeParticleSystem.fx)
// The reference's own form: one quad per particle, in draw order, each covered pixel shaded and blended into the target at once.
// P, RD, RC: PositionAndLife / RenderData / RenderColor of the live chunks (total float4 each); texture: Color texels or NULL;
// target: width*height float4, read (clear == 0) and written.  Conventions as documented for ilb_particle_render.
int orc_particles_render(const float* P, const float* RD, const float* RC, long total, const ilb_particle_render* r,
                         const uint8_t* texture, float* target) {
    if (!P || !RD || !RC || !r || !target) return ILB_ERR_INVALID_ARGUMENT;
    if (r->StippleFactor != 1.0f || r->RenderingOptions.y >= 0.5f) return ILB_ERR_UNSUPPORTED;
    if (r->texture_filter != ILB_TEXTURE_NONE && !texture) return ILB_ERR_INVALID_ARGUMENT;
    const int W = r->width, H = r->height;
    if (r->clear)
        for (long i = 0; i < (long)W * H; i++) st4(target, i, f4(r->ClearColor));
    const float4 region = f4(r->BitmapTextureRegion), sfp = f4(r->SizeFactorAndPosition), scale = f4(r->Scale);
    const float4 texelAndSize = f4(r->TexelAndSize), anim = f4(r->AnimationRateAndRotationAndZToY), options = f4(r->RenderingOptions);
    const float4 globalColor = f4(r->GlobalColor);
    const float vpx = r->ViewportPosition[0], vpy = r->ViewportPosition[1], vsx = r->ViewportScale[0], vsy = r->ViewportScale[1];
    const int texW = r->texture_width, texH = r->texture_height;
    auto texel = [&](int x, int y) {
        x = x < 0 ? 0 : (x >= texW ? texW - 1 : x);
        y = y < 0 ? 0 : (y >= texH ? texH - 1 : y);
        const uint8_t* t = texture + ((size_t)y * texW + x) * 4;
        return float4(t[0] / 255.0f, t[1] / 255.0f, t[2] / 255.0f, t[3] / 255.0f);
    };
    auto sample = [&](float u, float v) {  // BitmapSampler (LINEAR) / BitmapPointSampler, CLAMP, one mip level
        if (r->texture_filter == ILB_TEXTURE_POINT) return texel((int)floorf(u * (float)texW), (int)floorf(v * (float)texH));
        const float fx = u * (float)texW - 0.5f, fy = v * (float)texH - 0.5f;
        const float x0 = floorf(fx), y0 = floorf(fy);
        const float tx = fx - x0, ty = fy - y0;
        const float4 top = lerp(texel((int)x0, (int)y0), texel((int)x0 + 1, (int)y0), tx);
        const float4 bottom = lerp(texel((int)x0, (int)y0 + 1), texel((int)x0 + 1, (int)y0 + 1), tx);
        return lerp(top, bottom, ty);
    };
    for (long i = 0; i < total; i++) {
        // ---- VS_PosVelAttr :62-150
        const float4 position = ld4(P, i), renderData = ld4(RD, i), color = ld4(RC, i);
        const float life = position.w;
        if (!(life > 0)) continue;
        const float angle = fmod(renderData.y, 2 * PI);
        const float zf = max(0.0f, 1.0f + (position.z * r->ZConfiguration.x));
        const float sx = ((renderData.x * texelAndSize.z) * sfp.x) * zf, sy = ((renderData.x * texelAndSize.w) * sfp.y) * zf;
        const float sn = dm_sinf(angle), cs = dm_cosf(angle);
        const float dispx = position.x * scale.x + sfp.z, dispy = (position.y - (position.z * anim.w)) * scale.y + sfp.w;
        const float cx = (dispx - vpx) * vsx, cy = (dispy - vpy) * vsy;
        const float kx = scale.x * vsx, ky = scale.y * vsy;
        const float ax = (cs * sx) * kx, ay = (sn * sx) * ky, bx = -((sn * sy) * kx), by = (cs * sy) * ky;
        const float det = ax * by - ay * bx;
        if (!(fabsf(det) > 0)) continue;
        const float m00 = by / det, m01 = -bx / det, m10 = -ay / det, m11 = ax / det;
        const float ex = fabsf(ax) + fabsf(bx), ey = fabsf(ay) + fabsf(by);
        const float fx0 = (cx - ex) - 1, fx1 = (cx + ex) + 1, fy0 = (cy - ey) - 1, fy1 = (cy + ey) + 1;
        if (!(fx0 <= fx1) || !(fy0 <= fy1) || std::isinf(fx0) || std::isinf(fx1) || std::isinf(fy0) || std::isinf(fy1)) continue;
        if (fx1 < 0 || fy1 < 0 || fx0 > (float)(W - 1) || fy0 > (float)(H - 1)) continue;
        const int x0 = (int)floorf(fmaxf(fx0, 0.0f)), x1 = (int)ceilf(fminf(fx1, (float)(W - 1)));
        const int y0 = (int)floorf(fmaxf(fy0, 0.0f)), y1 = (int)ceilf(fminf(fy1, (float)(H - 1)));
        float frameU = 0, frameV = 0;
        if (r->texture_filter != ILB_TEXTURE_NONE) {
            const float tsx = region.z - region.x, tsy = region.w - region.y;
            const float fcx = floorf(1.0f / tsx), fcy = floorf(1.0f / tsy);
            float fix = floorf(fabsf(anim.x) * life), fiy = floorf(fabsf(anim.y) * life);
            const float maxAngleX = (2 * PI) / fcx, maxAngleY = (2 * PI) / fcy;
            const float ffvx = floorf(angle / maxAngleX + 0.5f), ffvy = floorf(angle / maxAngleY + 0.5f);
            fiy += floorf(renderData.w);
            if (options.z != 0) fix += ffvx;
            if (options.w != 0) fiy += ffvy;
            fix = fmod(max(fix, 0.0f), fcx);
            fiy = clamp(fiy, 0.0f, fcy - 1);
            if (anim.x < 0) fix = (fcx - fix) - 1;
            if (anim.y < 0) fiy = (fcy - fiy) - 1;
            frameU = fix * tsx; frameV = fiy * tsy;
        }
        const float rounding = clamp(evaluateBezier1(r->RoundingPowerFromLife, life), 0.001f, 1.0f);
        for (int y = y0; y <= y1; y++)
            for (int x = x0; x <= x1; x++) {
                const float dx = ((float)x + 0.5f) - cx, dy = ((float)y + 0.5f) - cy;
                const float u = dx * m00 + dy * m01, v = dx * m10 + dy * m11;
                if (!(u >= -1 && u < 1 && v >= -1 && v < 1)) continue;
                // ---- PS_NoTexture / PS_Texture / PS_TexturePoint :190-254 ((1 / 512) is an integer division: 0)
                float4 result = color;
                if (r->texture_filter != ILB_TEXTURE_NONE) {
                    if (color.w > 0) {
                        const float ccx = u / 2 + 0.5f, ccy = v / 2 + 0.5f;
                        const float tu = lerp(region.x, region.z, ccx) + frameU, tv = lerp(region.y, region.w, ccy) + frameV;
                        result = result * sample(tu, tv);
                        result = result * globalColor;
                    }
                } else {
                    result = result * globalColor;
                }
                float circular = 1;
                if (options.x != 0) {  // computeCircularAlpha :152-163
                    const float distance = sqrtf(u * u + v * v);
                    const float power = max(rounding, 0.01f);
                    const float divisor = max(saturate(1 - power), 0.001f);
                    const float distanceFromEdge = saturate(distance - power) / divisor;
                    circular = saturate(1 - powf(distanceFromEdge, power));
                }
                result = result * circular;
                if (!(result.w > 0)) continue;
                const size_t pi = (size_t)y * W + x;
                const float4 dst = ld4(target, pi);
                float4 out;
                if (r->blend == ILB_BLEND_ALPHA) out = result + dst * (1 - result.w);
                else if (r->blend == ILB_BLEND_ADDITIVE) out = result * result.w + dst;
                else out = result;
                st4(target, pi, out);
            }
    }
    return 0;
}

}  // extern "C"
