// Analytic distance functions (Shaders/DistanceFunctionCommon.fxh), used by the particle area weights
// (FMA.fx:15-20, Noise.fx:21-26) and by the distance-field generator (DistanceFunction.fx).
#pragma once
#include "ilb_device.cuh"

ILB_DEV f4 qmul(f4 q1, f4 q2) {  // :15-20
    const f3 a = xyz(q1), b = xyz(q2);
    return mk4(b * q1.w + a * q2.w + cross3(a, b), q1.w * q2.w - dot3(a, b));
}
ILB_DEV f3 rotateLocalPosition(f3 p, f4 rotation) {  // :23-26
    const f4 r_c = rotation * mk4(-1.0f, -1.0f, -1.0f, 1.0f);
    return xyz(qmul(rotation, qmul(mk4(p, 0.0f), r_c)));
}
ILB_DEV f4 opElongate(f3 p, f3 h) {  // :43-46
    const f3 q = abs3(p) - h;
    return mk4(sign3(p) * max3(q, mk3(0.0f)), fminf(fmaxf(q.x, fmaxf(q.y, q.z)), 0.0f));
}
ILB_DEV float evaluateBox(f3 position, f3 size) {  // :48-63
    const f3 d = abs3(position) - size;
    return fminf(fmaxf(d.x, fmaxf(d.y, d.z)), 0.0f) + length3(max3(d, mk3(0.0f)));
}
ILB_DEV float evaluateSpheroid(f3 position, f3 size) {  // :65-75
    const float minSize = fminf(size.x, fminf(size.y, size.z));
    const f4 w = opElongate(position, size - minSize);
    return w.w + (length3(xyz(w)) - minSize);
}
ILB_DEV float evaluateEllipsoid(f3 p, f3 r) {  // sdEllipsoid_improvedV2 :92-108
    const float k0 = length3(p / r);
    const float k1 = length3(p / (r * r));
    return (k0 < 1.0f) ? (k0 - 1.0f) * fminf(fminf(r.x, r.y), r.z) : k0 * (k0 - 1.0f) / k1;
}
ILB_DEV float sdCappedCylinder(f3 p, float h, float r) {  // :110-113
    const float dx = fabsf(length2(mk2(p.x, p.y))) - r, dy = fabsf(p.z) - h;
    return fminf(fmaxf(dx, dy), 0.0f) + length2(mk2(fmaxf(dx, 0.0f), fmaxf(dy, 0.0f)));
}
ILB_DEV float evaluateCylinder(f3 position, f3 size) {  // :115-121
    return sdCappedCylinder(position, size.z, length2(mk2(size.x, size.y)));
}
ILB_DEV float sdOctogonPrism(f3 p, float r, float h) {  // :139-152
    const float kx = -0.9238795325f, ky = 0.3826834323f, kz = 0.4142135623f;
    p = abs3(p);
    f2 q = mk2(p.x, p.y);
    q = q - 2.0f * fminf(dot2(mk2(kx, ky), q), 0.0f) * mk2(kx, ky);
    q = q - 2.0f * fminf(dot2(mk2(-kx, ky), q), 0.0f) * mk2(-kx, ky);
    q = q - mk2(clampf(q.x, -kz * r, kz * r), r);
    const float dx = length2(q) * signf(q.y), dy = p.z - h;
    return fminf(fmaxf(dx, dy), 0.0f) + length2(mk2(fmaxf(dx, 0.0f), fmaxf(dy, 0.0f)));
}
ILB_DEV float evaluateOctagon(f3 position, f3 size) {  // :154-165
    const float minSize = fminf(size.x, size.y);
    const f4 w = opElongate(position, mk3(size.x - minSize, size.y - minSize, 0.0f));
    return w.w + sdOctogonPrism(xyz(w), minSize, size.z);
}
ILB_DEV float evaluateByTypeId(int typeId, f3 worldPosition, f3 center, f3 size, f4 rotation) {  // :167-186
    const int t = typeId < 0 ? -typeId : typeId;
    if (t < 1 || t > 5) return 0.0f;
    const f3 position = rotateLocalPosition(worldPosition - center, rotation);
    switch (t) {
        case 1: return evaluateEllipsoid(position, size);
        case 2: return evaluateBox(position, size);
        case 3: return evaluateCylinder(position, size);
        case 4: return evaluateSpheroid(position, size);
        default: return evaluateOctagon(position, size);
    }
}
