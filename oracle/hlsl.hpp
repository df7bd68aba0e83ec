// TEST INFRASTRUCTURE -- CPU oracle for illuminant_b200 (see oracle/README.md).
// Minimal HLSL ps_3_0 vocabulary (float2/3/4 + intrinsics) so the restatements in
// oracle_lighting.cpp / oracle_particles.cpp can follow the reference shaders line by line.
// fp32 everywhere; compile with -ffp-contract=off (no FMA contraction, no fast-math).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

// deterministic sin / cos / acos shared bit-for-bit with the CUDA kernels (see the header for why)
#include "../include/ilb_detmath.h"

namespace hlsl {

static const float PI = 3.14159265358979323846f;

struct float2 {
    float x, y;
    float2() : x(0), y(0) {}
    float2(float s) : x(s), y(s) {}
    float2(float x_, float y_) : x(x_), y(y_) {}
};
struct float3 {
    float x, y, z;
    float3() : x(0), y(0), z(0) {}
    float3(float s) : x(s), y(s), z(s) {}
    float3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float3(float2 xy, float z_) : x(xy.x), y(xy.y), z(z_) {}
    float2 xy() const { return float2(x, y); }
};
struct float4 {
    float x, y, z, w;
    float4() : x(0), y(0), z(0), w(0) {}
    float4(float s) : x(s), y(s), z(s), w(s) {}
    float4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    float4(float3 v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    float3 xyz() const { return float3(x, y, z); }
    float2 xy() const { return float2(x, y); }
};

#define HLSL_OP2(T, op)                                                                         \
    static inline T operator op(T a, T b);                                                      \
    static inline T operator op(T a, float b) { return a op T(b); }                             \
    static inline T operator op(float a, T b) { return T(a) op b; }
HLSL_OP2(float2, +) HLSL_OP2(float2, -) HLSL_OP2(float2, *) HLSL_OP2(float2, /)
HLSL_OP2(float3, +) HLSL_OP2(float3, -) HLSL_OP2(float3, *) HLSL_OP2(float3, /)
HLSL_OP2(float4, +) HLSL_OP2(float4, -) HLSL_OP2(float4, *) HLSL_OP2(float4, /)
#undef HLSL_OP2
#define HLSL_DEF2(op)                                                                            \
    static inline float2 operator op(float2 a, float2 b) { return float2(a.x op b.x, a.y op b.y); } \
    static inline float3 operator op(float3 a, float3 b) { return float3(a.x op b.x, a.y op b.y, a.z op b.z); } \
    static inline float4 operator op(float4 a, float4 b) { return float4(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); }
HLSL_DEF2(+) HLSL_DEF2(-) HLSL_DEF2(*) HLSL_DEF2(/)
#undef HLSL_DEF2
static inline float2 operator-(float2 a) { return float2(-a.x, -a.y); }
static inline float3 operator-(float3 a) { return float3(-a.x, -a.y, -a.z); }
static inline float4 operator-(float4 a) { return float4(-a.x, -a.y, -a.z, -a.w); }
template <class T> static inline T& operator+=(T& a, T b) { a = a + b; return a; }
template <class T> static inline T& operator-=(T& a, T b) { a = a - b; return a; }
template <class T> static inline T& operator*=(T& a, T b) { a = a * b; return a; }
static inline float3& operator*=(float3& a, float b) { a = a * b; return a; }
static inline float4& operator*=(float4& a, float b) { a = a * b; return a; }

// scalar intrinsics
static inline float saturate(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }  // NaN -> 0 like HLSL
static inline float clamp(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
static inline float lerp(float a, float b, float t) { return a + t * (b - a); }
static inline float sign(float v) { return (v > 0.0f) ? 1.0f : ((v < 0.0f) ? -1.0f : 0.0f); }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline float abs(float a) { return fabsf(a); }
// HLSL `%` / fmod on floats: result has the sign of the dividend (truncated division)
static inline float fmod(float a, float b) { return fmodf(a, b); }

#define HLSL_MAP1(name, expr)                                                                    \
    static inline float2 name(float2 a) { return float2(expr(a.x), expr(a.y)); }                 \
    static inline float3 name(float3 a) { return float3(expr(a.x), expr(a.y), expr(a.z)); }      \
    static inline float4 name(float4 a) { return float4(expr(a.x), expr(a.y), expr(a.z), expr(a.w)); }
HLSL_MAP1(saturate, saturate) HLSL_MAP1(abs, fabsf) HLSL_MAP1(sign, sign) HLSL_MAP1(floor, floorf)
#undef HLSL_MAP1
#define HLSL_MAP2(name, expr)                                                                    \
    static inline float2 name(float2 a, float2 b) { return float2(expr(a.x, b.x), expr(a.y, b.y)); } \
    static inline float3 name(float3 a, float3 b) { return float3(expr(a.x, b.x), expr(a.y, b.y), expr(a.z, b.z)); } \
    static inline float4 name(float4 a, float4 b) { return float4(expr(a.x, b.x), expr(a.y, b.y), expr(a.z, b.z), expr(a.w, b.w)); }
HLSL_MAP2(min, fminf) HLSL_MAP2(max, fmaxf)
#undef HLSL_MAP2
static inline float2 clamp(float2 v, float2 lo, float2 hi) { return min(max(v, lo), hi); }
static inline float3 clamp(float3 v, float3 lo, float3 hi) { return min(max(v, lo), hi); }
static inline float2 lerp(float2 a, float2 b, float t) { return a + t * (b - a); }
static inline float3 lerp(float3 a, float3 b, float t) { return a + t * (b - a); }
static inline float4 lerp(float4 a, float4 b, float t) { return a + t * (b - a); }
static inline float3 lerp(float3 a, float3 b, float3 t) { return a + t * (b - a); }

static inline float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
static inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
static inline float length(float2 a) { return sqrtf(dot(a, a)); }
static inline float length(float3 a) { return sqrtf(dot(a, a)); }
static inline float length(float4 a) { return sqrtf(dot(a, a)); }
static inline float3 cross(float3 a, float3 b) {
    return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// ps_3_0 `nrm` is v * rsq(dot(v,v)); D3D9 defines rsq(0) = FLT_MAX, so normalize(0) = 0 (no NaN).
// Convention shared with the CUDA path: zero-length input -> zero vector, else v * (1 / sqrt(dot)) with the
// reciprocal square root formed as an IEEE division of an IEEE square root (reproducible on CPU and GPU).
static inline float3 normalize(float3 a) {
    float d = dot(a, a);
    if (d == 0.0f) return float3(0.0f);
    float r = 1.0f / sqrtf(d);
    return a * r;
}
static inline bool any(float2 a) { return (a.x != 0.0f) || (a.y != 0.0f); }
static inline bool any(float3 a) { return (a.x != 0.0f) || (a.y != 0.0f) || (a.z != 0.0f); }

// mul(row-vector float4, float4x4) with the matrix stored row-major (XNA Matrix M11..M44)
static inline float4 mul(float4 v, const float* m) {
    return float4(
        v.x * m[0] + v.y * m[4] + v.z * m[8] + v.w * m[12],
        v.x * m[1] + v.y * m[5] + v.z * m[9] + v.w * m[13],
        v.x * m[2] + v.y * m[6] + v.z * m[10] + v.w * m[14],
        v.x * m[3] + v.y * m[7] + v.z * m[11] + v.w * m[15]);
}

}  // namespace hlsl
