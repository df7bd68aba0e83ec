// Analytic distance functions (Shaders/DistanceFunctionCommon.fxh), used by the particle area weights
// (FMA.fx:15-20, Noise.fx:21-26) and by the distance-field generator (DistanceFunction.fx).
// x-ops (see ilb_device.cuh): an area weight scales a particle's position / velocity update, and particle state feeds
// the collision thresholds, so these stay bit-identical to the oracle.
#pragma once
#include "ilb_device.cuh"

ILB_DEV f4 qmul(f4 q1, f4 q2) {  // :15-20
    const f3 a = xyz(q1), b = xyz(q2);
    return mk4(xadd3(xadd3(xscale3(b, q1.w), xscale3(a, q2.w)), xcross3(a, b)), xsub(xmul(q1.w, q2.w), xdot3(a, b)));
}
ILB_DEV f3 rotateLocalPosition(f3 p, f4 rotation) {  // :23-26
    const f4 r_c = mk4(-rotation.x, -rotation.y, -rotation.z, rotation.w);
    return xyz(qmul(rotation, qmul(mk4(p, 0.0f), r_c)));
}
ILB_DEV f4 opElongate(f3 p, f3 h) {  // :43-46
    const f3 q = xsub3(abs3(p), h);
    return mk4(xmul3(sign3(p), max3(q, mk3(0.0f))), fminf(fmaxf(q.x, fmaxf(q.y, q.z)), 0.0f));
}
ILB_DEV float evaluateBox(f3 position, f3 size) {  // :48-63
    const f3 d = xsub3(abs3(position), size);
    return xadd(fminf(fmaxf(d.x, fmaxf(d.y, d.z)), 0.0f), xlength3z(max3(d, mk3(0.0f))));
}
ILB_DEV float evaluateSpheroid(f3 position, f3 size) {  // :65-75
    const float minSize = fminf(size.x, fminf(size.y, size.z));
    const f4 w = opElongate(position, mk3(xsub(size.x, minSize), xsub(size.y, minSize), xsub(size.z, minSize)));
    return xadd(w.w, xsub(xlength3z(xyz(w)), minSize));
}
ILB_DEV float evaluateEllipsoid(f3 p, f3 r) {  // sdEllipsoid_improvedV2 :92-108
    const float k0 = xlength3z(xdiv3z(p, r));
    const float k1 = xlength3z(xdiv3z(p, xmul3(r, r)));
    return (k0 < 1.0f) ? xmul(xsub(k0, 1.0f), fminf(fminf(r.x, r.y), r.z)) : xdivz(xmul(k0, xsub(k0, 1.0f)), k1);
}
ILB_DEV float sdCappedCylinder(f3 p, float h, float r) {  // :110-113
    const float dx = xsub(fabsf(xlength2z(mk2(p.x, p.y))), r), dy = xsub(fabsf(p.z), h);
    return xadd(fminf(fmaxf(dx, dy), 0.0f), xlength2z(mk2(fmaxf(dx, 0.0f), fmaxf(dy, 0.0f))));
}
ILB_DEV float evaluateCylinder(f3 position, f3 size) {  // :115-121
    return sdCappedCylinder(position, size.z, xlength2z(mk2(size.x, size.y)));
}
ILB_DEV float sdOctogonPrism(f3 p, float r, float h) {  // :139-152
    const float kx = -0.9238795325f, ky = 0.3826834323f, kz = 0.4142135623f;
    p = abs3(p);
    f2 q = mk2(p.x, p.y);
    float s = xmul(2.0f, fminf(xdot2(mk2(kx, ky), q), 0.0f));
    q = mk2(xsub(q.x, xmul(s, kx)), xsub(q.y, xmul(s, ky)));
    s = xmul(2.0f, fminf(xdot2(mk2(-kx, ky), q), 0.0f));
    q = mk2(xsub(q.x, xmul(s, -kx)), xsub(q.y, xmul(s, ky)));
    q = mk2(xsub(q.x, clampf(q.x, xmul(-kz, r), xmul(kz, r))), xsub(q.y, r));
    const float dx = xmul(xlength2z(q), signf(q.y)), dy = xsub(p.z, h);
    return xadd(fminf(fmaxf(dx, dy), 0.0f), xlength2z(mk2(fmaxf(dx, 0.0f), fmaxf(dy, 0.0f))));
}
ILB_DEV float evaluateOctagon(f3 position, f3 size) {  // :154-165
    const float minSize = fminf(size.x, size.y);
    const f4 w = opElongate(position, mk3(xsub(size.x, minSize), xsub(size.y, minSize), 0.0f));
    return xadd(w.w, sdOctogonPrism(xyz(w), minSize, size.z));
}
ILB_DEV float evaluateByTypeId(int typeId, f3 worldPosition, f3 center, f3 size, f4 rotation) {  // :167-186
    const int t = typeId < 0 ? -typeId : typeId;
    if (t < 1 || t > 5) return 0.0f;
    const f3 position = rotateLocalPosition(xsub3(worldPosition, center), rotation);
    switch (t) {
        case 1: return evaluateEllipsoid(position, size);
        case 2: return evaluateBox(position, size);
        case 3: return evaluateCylinder(position, size);
        case 4: return evaluateSpheroid(position, size);
        default: return evaluateOctagon(position, size);
    }
}
