/*
 * ilb_detmath.h -- deterministic fp32 sin / cos / acos shared by the CUDA kernels and the CPU oracle.
 *
 * The reference's transcendental intrinsics are whatever the D3D driver lowers `sincos` / `acos` to (ps_3_0 `sincos`
 * is itself a short polynomial macro), so no particular bit pattern is "the reference's".  What parity needs is that
 * both sides of a comparison evaluate the SAME function: results of these calls feed discontinuities (spawned particle
 * state feeds collision tests; decoded normals offset the cone-trace origin; the line-light solid angle is a
 * near-cancelling sum of four arc-cosines), where 1 ulp between two libm implementations becomes an O(1) difference.
 *
 * Algorithms: Cephes single-precision sinf / cosf (Moshier): absolute error below 1e-7 for |x| <= 8192 (measured
 * 7.7e-8 against float64 libm); acos from Abramowitz & Stegun 4.4.46: absolute error below 5e-7 on [-1, 1] (measured
 * 4.3e-7, i.e. 1.8 ulp of pi).  tests/test_detmath.py asserts these bounds and, on the GPU, that the device build
 * returns the same bits as the host build.  Arguments must satisfy |x| < 1.6e9 (the octant index is an int).  Every
 * operation is an individually rounded IEEE fp32 add / multiply / sqrt -- or, in the arc-cosine polynomial, an explicit
 * fused multiply-add -- through the DM_* macros, so the CPU build (-ffp-contract=off: nothing is fused behind the
 * source's back) and the GPU build (__fadd_rn / __fmul_rn / __fsqrt_rn / __fmaf_rn) agree bit for bit.
 *
 * Usage: define DM_FN (function qualifiers) and optionally DM_ADD / DM_MUL / DM_SQRT before including.
 */
#ifndef ILB_DETMATH_H
#define ILB_DETMATH_H

#ifndef DM_FN
#define DM_FN static inline
#endif
#ifndef DM_ADD
#define DM_ADD(a, b) ((a) + (b))
#define DM_MUL(a, b) ((a) * (b))
#define DM_SQRT(a) sqrtf(a)
#endif
#define DM_SUB(a, b) DM_ADD((a), -(b))
/* the fused multiply-add of IEEE 754-2008 (one rounding): fmaf() on the host, __fmaf_rn on the device -- the same bits */
#ifndef DM_FMA
#define DM_FMA(a, b, c) __builtin_fmaf((a), (b), (c)) /* gcc / clang: no header, so no ::abs / ::fmod leak into the includer */
#endif

#define DM_FOPI 1.27323954473516f /* 4/pi */
#define DM_DP1 0.78515625f
#define DM_DP2 2.4187564849853515625e-4f
#define DM_DP3 3.77489497744594108e-8f
#define DM_PIF 3.141592653589793238f
#define DM_PIO2F 1.5707963267948966192f

/* polynomial kernels on the reduced argument |x| <= pi/4, z = x*x */
DM_FN float dm_sin_kernel(float x, float z) {
    float y = DM_ADD(DM_MUL(-1.9515295891E-4f, z), 8.3321608736E-3f);
    y = DM_ADD(DM_MUL(y, z), -1.6666654611E-1f);
    y = DM_MUL(DM_MUL(y, z), x);
    return DM_ADD(y, x);
}
DM_FN float dm_cos_kernel(float z) {
    float y = DM_ADD(DM_MUL(2.443315711809948E-005f, z), -1.388731625493765E-003f);
    y = DM_ADD(DM_MUL(y, z), 4.166664568298827E-002f);
    y = DM_MUL(DM_MUL(y, z), z);
    y = DM_SUB(y, DM_MUL(0.5f, z));
    return DM_ADD(y, 1.0f);
}

/* octant reduction shared by sin and cos: returns the octant j in [0,7], the reduced argument in *r */
DM_FN int dm_reduce(float ax, float* r) {
    int j = (int)DM_MUL(DM_FOPI, ax); /* truncation; ax >= 0 */
    float y = (float)j;
    if (j & 1) {
        j += 1;
        y = DM_ADD(y, 1.0f);
    }
    *r = DM_SUB(DM_SUB(DM_SUB(ax, DM_MUL(y, DM_DP1)), DM_MUL(y, DM_DP2)), DM_MUL(y, DM_DP3));
    return j & 7;
}

DM_FN float dm_sinf(float xx) {
    int neg = xx < 0.0f;
    float x;
    int j = dm_reduce(neg ? -xx : xx, &x);
    if (j > 3) {
        neg = !neg;
        j -= 4;
    }
    const float z = DM_MUL(x, x);
    const float y = (j == 1 || j == 2) ? dm_cos_kernel(z) : dm_sin_kernel(x, z);
    return neg ? -y : y;
}

DM_FN float dm_cosf(float xx) {
    float x;
    int j = dm_reduce(xx < 0.0f ? -xx : xx, &x);
    int neg = 0;
    if (j > 3) {
        j -= 4;
        neg = !neg;
    }
    if (j > 1) neg = !neg;
    const float z = DM_MUL(x, x);
    const float y = (j == 1 || j == 2) ? dm_sin_kernel(x, z) : dm_cos_kernel(z);
    return neg ? -y : y;
}

DM_FN void dm_sincosf(float x, float* s, float* c) {
    *s = dm_sinf(x);
    *c = dm_cosf(x);
}

/* acos: Abramowitz & Stegun 4.4.46, acos(x) = sqrt(1 - x) * P7(x) on [0, 1] (|error| <= 2e-8 before rounding),
 * reflected for x < 0.  Branch-free apart from the final select; |x| > 1 gives sqrt(negative) = NaN like acos(). */
DM_FN float dm_acos_poly(float x) { /* P7(|x|), Horner's rule in fused multiply-adds: 7 operations, 7 roundings */
    float p = DM_FMA(-0.0012624911f, x, 0.0066700901f);
    p = DM_FMA(p, x, -0.0170881256f);
    p = DM_FMA(p, x, 0.0308918810f);
    p = DM_FMA(p, x, -0.0501743046f);
    p = DM_FMA(p, x, 0.0889789874f);
    p = DM_FMA(p, x, -0.2145988016f);
    return DM_FMA(p, x, 1.5707963050f);
}
/* s = sqrt(1 - |xx|), supplied by the caller (the kernels have a cheaper correctly rounded sqrt for in-range operands) */
DM_FN float dm_acos_finish(float xx, float s) {
    const float x = xx < 0.0f ? -xx : xx;
    const float r = DM_MUL(s, dm_acos_poly(x));
    return xx < 0.0f ? DM_SUB(DM_PIF, r) : r;
}
DM_FN float dm_acosf(float xx) {
    const float x = xx < 0.0f ? -xx : xx;
    return dm_acos_finish(xx, DM_SQRT(DM_SUB(1.0f, x)));
}

#endif /* ILB_DETMATH_H */
