// 3dgrut_b200/csrc/subtile_cull.cuh -- the arithmetic of the sub-tile culling of the render kernels (gut_render.cu, section comment there):
// the block frame, a ray's projective coordinates in it, and the conservative "can any ray of the block be accepted" decision.  In a
// header so that the CPU suite can compile the very same code with g++ and fuzz the NECESSARY-condition claim against the exact test
// (tests/host_emul/cull_host.cpp, tests/test_cull_host_emul.py).
#pragma once
#include "gut_common.cuh"

namespace gutb200 {

struct WarpFrame {
    float e1x, e1y, e1z, e2x, e2y, e2z, e3x, e3y, e3z;   // frame around the direction of the block's first live pixel
    float ulo, uhi, vlo, vhi, umax;                      // the block's rectangle in that frame
    bool on;
};

// frame (e1, e2, e3) around a direction: e3 = the normalised direction, e1 = the camera axis least aligned with it made orthogonal, e2 = e3 x e1
__device__ __forceinline__ bool frame_axes(const float* s2w, float dx, float dy, float dz, WarpFrame& wf) {
    const float n2 = dx * dx + dy * dy + dz * dz;
    if (!(n2 > 1e-20f) || !(n2 < 1e20f)) return false;
    const float in = rsqrtf(n2);
    dx *= in; dy *= in; dz *= in;
    const float* m = s2w;  // complement: the camera axis least aligned with the direction
    const float cxa = fabsf(m[0] * dx + m[1] * dy + m[2] * dz), cya = fabsf(m[3] * dx + m[4] * dy + m[5] * dz);
    const float ax = cxa <= cya ? m[0] : m[3], ay = cxa <= cya ? m[1] : m[4], az = cxa <= cya ? m[2] : m[5];
    const float k = ax * dx + ay * dy + az * dz;
    float e1x = ax - k * dx, e1y = ay - k * dy, e1z = az - k * dz;
    const float l1 = e1x * e1x + e1y * e1y + e1z * e1z;
    if (!(l1 > 1e-6f)) return false;
    const float i1 = rsqrtf(l1);
    e1x *= i1; e1y *= i1; e1z *= i1;
    wf.e1x = e1x; wf.e1y = e1y; wf.e1z = e1z;
    wf.e2x = dy * e1z - dz * e1y; wf.e2y = dz * e1x - dx * e1z; wf.e2z = dx * e1y - dy * e1x;
    wf.e3x = dx; wf.e3y = dy; wf.e3z = dz;
    return true;
}

// projective coordinates (u, v) of a ray direction in the frame; false when the ray is more than 60 degrees off e3
__device__ __forceinline__ bool ray_uv(const WarpFrame& wf, float rdx, float rdy, float rdz, float& u, float& v) {
    const float w = rdx * wf.e3x + rdy * wf.e3y + rdz * wf.e3z;
    const float r2 = rdx * rdx + rdy * rdy + rdz * rdz;
    const bool fine = (w > 0.f) && (w * w > 0.25f * r2);
    const float iw = fine ? 1.0f / w : 0.f;
    u = (rdx * wf.e1x + rdy * wf.e1y + rdz * wf.e1z) * iw;
    v = (rdx * wf.e2x + rdy * wf.e2y + rdz * wf.e2z) * iw;
    return fine;
}

// Can any ray of the warp's block be accepted by the particle with canonical transform rows (m0, m1, m2), canonical ray origin g
// and density dns?  Conservative (see the section comment).
template <int DEG>
__device__ __forceinline__ bool block_candidate(const FrameConfig& cfg, const WarpFrame& wf, float m0x, float m0y, float m0z, float m1x,
                                                float m1y, float m1z, float m2x, float m2y, float m2z, float gx, float gy, float gz,
                                                float dns) {
    const float tau = fmaxf(cfg.min_kernel_density, dns > 0.f ? cfg.min_alpha / dns : 2.f);
    if (!(tau < 1.f)) return false;  // response <= 1: never accepted
    const float ln = -logf(tau);
    float r2 = DEG == 4 ? sqrtf(18.f * ln) : 2.f * ln;   // exp(-gray^2/18) > tau  |  exp(-gray/2) > tau
    // relative + absolute inflation: the absolute part matters for densities just above min_alpha (r2 -> 0), where the approximate
    // exp / division of the exact test (-use_fast_math) moves the accept boundary by ~1e-6 in r2
    r2 = r2 * 1.002f + 2e-4f;
    const float a1x = m0x * wf.e1x + m0y * wf.e1y + m0z * wf.e1z, a1y = m1x * wf.e1x + m1y * wf.e1y + m1z * wf.e1z,
                a1z = m2x * wf.e1x + m2y * wf.e1y + m2z * wf.e1z;
    const float a2x = m0x * wf.e2x + m0y * wf.e2y + m0z * wf.e2z, a2y = m1x * wf.e2x + m1y * wf.e2y + m1z * wf.e2z,
                a2z = m2x * wf.e2x + m2y * wf.e2y + m2z * wf.e2z;
    const float a3x = m0x * wf.e3x + m0y * wf.e3y + m0z * wf.e3z, a3y = m1x * wf.e3x + m1y * wf.e3y + m1z * wf.e3z,
                a3z = m2x * wf.e3x + m2y * wf.e3y + m2z * wf.e3z;
    const float c1x = a1y * gz - a1z * gy, c1y = a1z * gx - a1x * gz, c1z = a1x * gy - a1y * gx;
    const float c2x = a2y * gz - a2z * gy, c2y = a2z * gx - a2x * gz, c2z = a2x * gy - a2y * gx;
    const float c3x = a3y * gz - a3z * gy, c3y = a3z * gx - a3x * gz, c3z = a3x * gy - a3y * gx;
    const float a33 = a3x * a3x + a3y * a3y + a3z * a3z, c33 = c3x * c3x + c3y * c3y + c3z * c3z;
    const float A00 = (c1x * c1x + c1y * c1y + c1z * c1z) - r2 * (a1x * a1x + a1y * a1y + a1z * a1z);
    const float A01 = (c1x * c2x + c1y * c2y + c1z * c2z) - r2 * (a1x * a2x + a1y * a2y + a1z * a2z);
    const float A11 = (c2x * c2x + c2y * c2y + c2z * c2z) - r2 * (a2x * a2x + a2y * a2y + a2z * a2z);
    const float B0 = (c1x * c3x + c1y * c3y + c1z * c3z) - r2 * (a1x * a3x + a1y * a3y + a1z * a3z);
    const float B1 = (c2x * c3x + c2y * c3y + c2z * c3z) - r2 * (a2x * a3x + a2y * a3y + a2z * a3z);
    const float K = c33 - r2 * a33;
    // f is convex iff A > 0; otherwise {f < 0} is unbounded (particle around / behind the origin): keep the entry
    if (!(A00 > 0.f) || !(A11 > 0.f)) return true;
    // bound of the fp32 error of f over the block and of the exact test's own rounding near the boundary
    const float U = wf.umax;
    const float g2 = gx * gx + gy * gy + gz * gz;
    const float err = 1e-5f * ((A00 + 2.f * fabsf(A01) + A11) * U * U + 2.f * (fabsf(B0) + fabsf(B1)) * U + c33 + r2 * a33) +
                      4e-6f * sqrtf(fmaxf(c33, r2 * a33) * a33 * g2);
    // Only the sign of f matters: bring the coefficients to O(1) before forming products of two of them.  For needle-like Gaussians
    // (scales ~1e-5 a few units away) the raw coefficients reach 1e21 and det = A00 A11 - A01^2 overflowed, which made the
    // centre-inside test fail and culled real hits (found by the host fuzz, tests/test_cull_host_emul.py).
    const float sc = 1.0f / fmaxf(A00, A11);
    const float a00 = A00 * sc, a01 = A01 * sc, a11 = A11 * sc, b0 = B0 * sc, b1 = B1 * sc, kp = (K - err) * sc;
    const float det = a00 * a11 - a01 * a01;
    if (!(det > 0.f)) return true;  // not convex, or not a number
    // minimum of f(u,v) = a00 u^2 + 2 a01 u v + a11 v^2 + 2 b0 u + 2 b1 v + kp over the rectangle: the unconstrained minimiser
    // if it lies inside, else the smallest of the four edge minima (1-D convex quadratics, clamped)
    const float u0 = wf.ulo, u1 = wf.uhi, v0 = wf.vlo, v1 = wf.vhi;
    const float ia00 = 1.0f / a00, ia11 = 1.0f / a11;
    const float qa = fmaf(a01, u0, b1), qb = fmaf(a01, u1, b1), pa = fmaf(a01, v0, b0), pb = fmaf(a01, v1, b0);
    const float va = fminf(fmaxf(-qa * ia11, v0), v1), vb = fminf(fmaxf(-qb * ia11, v0), v1);
    const float ua = fminf(fmaxf(-pa * ia00, u0), u1), ub = fminf(fmaxf(-pb * ia00, u0), u1);
    const float fa = fmaf(va, fmaf(a11, va, 2.f * qa), fmaf(u0, fmaf(a00, u0, 2.f * b0), kp));
    const float fb = fmaf(vb, fmaf(a11, vb, 2.f * qb), fmaf(u1, fmaf(a00, u1, 2.f * b0), kp));
    const float fc = fmaf(ua, fmaf(a00, ua, 2.f * pa), fmaf(v0, fmaf(a11, v0, 2.f * b1), kp));
    const float fd = fmaf(ub, fmaf(a00, ub, 2.f * pb), fmaf(v1, fmaf(a11, v1, 2.f * b1), kp));
    const float edge_min = fminf(fminf(fa, fb), fminf(fc, fd));
    // centre inside the rectangle: the interior minimum is below every edge value; keep the entry unless even f(centre) > 0,
    // which the edge values cannot tell -- so test the centre explicitly
    const float cu = (a01 * b1 - a11 * b0), cv = (a01 * b0 - a00 * b1);  // times det
    const bool inside = (cu >= u0 * det) && (cu <= u1 * det) && (cv >= v0 * det) && (cv <= v1 * det);
    const bool finite = (edge_min == edge_min) && (cu == cu) && (cv == cv);  // NaN anywhere: keep the entry
    return !finite || inside || (edge_min < 0.f);
}

}  // namespace gutb200
