// Bezier.fxh (ClampedBezier1 / ClampedBezier4 evaluation) shared by the particle update (particles.cu) and the particle
// rasteriser (raster.cu).  Include inside the translation unit's anonymous namespace, after ilb_device.cuh.
#pragma once

// ---- Bezier.fxh ------------------------------------------------------------------------------------------
ILB_DEV float tForScaledBezier(const ilb_float4& rangeAndCount, float value, float& t) {  // :21-63
    const float minValue = rangeAndCount.x, invDivisor = rangeAndCount.y;
    const uint32_t mode = (uint32_t)fabsf(rangeAndCount.w);
    const bool repeating = mode > 255, bouncing = mode > 511;
    t = (value - minValue) * fabsf(invDivisor);
    if (bouncing) {
        t *= 2.0f;
        if (invDivisor < 0.0f) t = 2.0f - fmodf(t, 2.0f); else t = fmodf(t, 2.0f);
        if (t > 1.0f) t = 1.0f - (t - 1.0f);
    } else if (repeating) {
        if (invDivisor < 0.0f) t = 1.0f - fmodf(t, 1.0f); else t = fmodf(t, 1.0f);
    } else {
        if (invDivisor < 0.0f) t = 1.0f - saturatef(t); else t = saturatef(t);
    }
    switch (mode % 256) {
        default: break;
        case 1: t = dm_sinf(xmul(xmul(t, ILB_PI), 0.5f)); break;
        case 2: t = t * t; break;
    }
    return rangeAndCount.z;
}
ILB_DEV float bezierScalar(float a, float b, float c, float d, float count, float t) {  // :65-95
    if (count <= 1.5f) return a;
    const float ab = lerpf(a, b, t);
    if (count <= 2.5f) return ab;
    if (count <= 3.5f) return (t <= 0.0f) ? a : ((t >= 1.0f) ? c : b);
    const float bc = lerpf(b, c, t), abbc = lerpf(ab, bc, t), cd = lerpf(c, d, t), bccd = lerpf(bc, cd, t);
    return lerpf(abbc, bccd, t);
}
ILB_DEV float evaluateBezier1(const ilb_bezier1& b, float value) {  // :97-101
    if (b.RangeAndCount.z <= 1.5f) return b.ABCD.x;  // uniform: a one-point curve (the default) never looks at t
    float t;
    const float count = tForScaledBezier(b.RangeAndCount, value, t);
    return bezierScalar(b.ABCD.x, b.ABCD.y, b.ABCD.z, b.ABCD.w, count, t);
}
ILB_DEV f4 evaluateBezier4(const ilb_bezier4& b, float value) {  // :141-177
    if (b.RangeAndCount.z <= 1.5f) return mk4(b.A);  // uniform: a one-point curve (the default) never looks at t
    float t;
    const float count = tForScaledBezier(b.RangeAndCount, value, t);
    return mk4(bezierScalar(b.A.x, b.B.x, b.C.x, b.D.x, count, t), bezierScalar(b.A.y, b.B.y, b.C.y, b.D.y, count, t),
               bezierScalar(b.A.z, b.B.z, b.C.z, b.D.z, count, t), bezierScalar(b.A.w, b.B.w, b.C.w, b.D.w, count, t));
}

