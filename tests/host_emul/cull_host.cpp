// tests/host_emul/cull_host.cpp -- TEST INFRASTRUCTURE: compiles the sub-tile culling arithmetic of the render kernels
// (3dgrut_b200/csrc/subtile_cull.cuh, the header gut_render.cu is built from) and the exact per-ray accept test (hit_math.cuh) with g++
// and fuzzes the claim the culling rests on: block_candidate() == false  =>  no ray of the block is accepted by the exact test.
// One case = a random 8x4 block of pinhole-like rays with a common origin + a random Gaussian placed near the bundle.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <random>
#include <cstdio>
#include <cstdlib>
using std::max;
using std::min;
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
#define __expf(x) expf(x)
static inline float4 __ldg(const float4* p) { return *p; }
#define __CUDACC__ 1
#include "../../3dgrut_b200/csrc/subtile_cull.cuh"
#include "../../3dgrut_b200/csrc/hit_math.cuh"

using namespace gutb200;

extern "C" {

// returns the number of violations (must be 0); stats[0] = cases with at least one accepted ray, stats[1] = cases culled,
// stats[2] = cases kept although no ray is accepted (lost opportunity), worst[12] = particle record of the first violation
int64_t cull_fuzz(uint64_t seed, int64_t cases, int degree, float min_density, float min_alpha, float max_alpha, float slack, int64_t* stats,
                  float* worst) {
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    std::normal_distribution<float> N(0.f, 1.f);
    FrameConfig cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.kernel_degree = degree; cfg.min_kernel_density = min_density; cfg.min_alpha = min_alpha; cfg.max_alpha = max_alpha;
    const float s2w[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
    int64_t violations = 0;
    stats[0] = stats[1] = stats[2] = 0;
    for (int64_t it = 0; it < cases; ++it) {
        // camera-like bundle: origin, central direction within 50 degrees of +z, pixel pitch 1e-4 .. 3e-2 rad
        const float ox = N(rng), oy = N(rng), oz = N(rng);
        const float ang = 0.87f * U(rng), az = 6.2831853f * U(rng);
        const float cdx = sinf(ang) * cosf(az), cdy = sinf(ang) * sinf(az), cdz = cosf(ang);
        const float pitch = expf(logf(1e-4f) + U(rng) * (logf(3e-2f) - logf(1e-4f)));
        // image-plane axes
        float ax = 1.f - cdx * cdx, ay = -cdx * cdy, az_ = -cdx * cdz;
        const float al = sqrtf(ax * ax + ay * ay + az_ * az_) + 1e-20f;
        ax /= al; ay /= al; az_ /= al;
        const float bx = cdy * az_ - cdz * ay, by = cdz * ax - cdx * az_, bz = cdx * ay - cdy * ax;
        float rdx[32], rdy[32], rdz[32];
        const float offx = (U(rng) - 0.5f) * 40.f, offy = (U(rng) - 0.5f) * 40.f;  // the block is somewhere in a 40-pixel neighbourhood
        for (int l = 0; l < 32; ++l) {
            const float pu = (offx + (l & 7)) * pitch, pv = (offy + (l >> 3)) * pitch;
            float dx = cdx + pu * ax + pv * bx, dy = cdy + pu * ay + pv * by, dz = cdz + pu * az_ + pv * bz;
            const float il = (U(rng) < 0.5f) ? 1.0f / sqrtf(dx * dx + dy * dy + dz * dz) : 1.0f;  // normalised or not
            rdx[l] = dx * il; rdy[l] = dy * il; rdz[l] = dz * il;
        }
        // particle near the bundle: depth 0.05 .. 50, lateral offset up to a few extents, scales 1e-4 .. 3 with anisotropy
        const float depth = expf(logf(0.05f) + U(rng) * (logf(50.f) - logf(0.05f)));
        const float base = expf(logf(1e-4f) + U(rng) * (logf(3.f) - logf(1e-4f)));
        float sc[3];
        for (int k = 0; k < 3; ++k) sc[k] = base * expf(2.3f * (U(rng) - 0.5f) * (U(rng) < 0.3f ? 3.f : 1.f));
        const float smax = fmaxf(sc[0], fmaxf(sc[1], sc[2]));
        const float lat = (4.f * smax + 20.f * pitch * depth) * (U(rng) < 0.5f ? U(rng) : 3.f * U(rng));
        const float la = 6.2831853f * U(rng);
        const int lref = static_cast<int>(U(rng) * 32) & 31;
        const float nl = 1.0f / sqrtf(rdx[lref] * rdx[lref] + rdy[lref] * rdy[lref] + rdz[lref] * rdz[lref]);
        float particle[12];
        particle[0] = ox + depth * rdx[lref] * nl + lat * (cosf(la) * ax + sinf(la) * bx);
        particle[1] = oy + depth * rdy[lref] * nl + lat * (cosf(la) * ay + sinf(la) * by);
        particle[2] = oz + depth * rdz[lref] * nl + lat * (cosf(la) * az_ + sinf(la) * bz);
        particle[3] = U(rng) < 0.2f ? 0.004f + 0.01f * U(rng) : U(rng);   // density, some near the alpha threshold
        float q[4] = {N(rng), N(rng), N(rng), N(rng)};
        const float ql = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) + 1e-20f;
        for (int k = 0; k < 4; ++k) particle[4 + k] = q[k] / ql;
        particle[8] = sc[0]; particle[9] = sc[1]; particle[10] = sc[2]; particle[11] = 0.f;
        const ParticleFrame f = load_frame(particle, 0);
        // exact test per ray
        bool any = false;
        for (int l = 0; l < 32; ++l) {
            // `slack` < 1 lowers both accept thresholds of the exact test a little: the kernels evaluate it with approximate exp / rsqrt /
            // division (-use_fast_math), so rays this close to the thresholds may be accepted on the GPU -- the culling must keep them too
            const CanonicalHit h = degree == 4 ? canonical_hit<4>(f, ox, oy, oz, rdx[l], rdy[l], rdz[l], min_density * slack, min_alpha * slack, max_alpha)
                                               : canonical_hit<2>(f, ox, oy, oz, rdx[l], rdy[l], rdz[l], min_density * slack, min_alpha * slack, max_alpha);
            any = any || h.accept;
        }
        // the block's frame and rectangle exactly as make_warp_frame builds them (first ray = frame axis)
        WarpFrame wf;
        wf.on = false;
        if (!frame_axes(s2w, rdx[0], rdy[0], rdz[0], wf)) continue;
        float ulo = 3.0e38f, uhi = -3.0e38f, vlo = 3.0e38f, vhi = -3.0e38f;
        bool fine = true;
        for (int l = 0; l < 32; ++l) {
            float u, v;
            fine = fine && ray_uv(wf, rdx[l], rdy[l], rdz[l], u, v);
            ulo = fminf(ulo, u); uhi = fmaxf(uhi, u); vlo = fminf(vlo, v); vhi = fmaxf(vhi, v);
        }
        if (!fine) continue;  // the kernels do not cull such blocks
        wf.ulo = ulo; wf.uhi = uhi; wf.vlo = vlo; wf.vhi = vhi;
        wf.umax = fmaxf(fmaxf(fabsf(ulo), fabsf(uhi)), fmaxf(fabsf(vlo), fabsf(vhi)));
        wf.on = true;
        // M = S^-1 R^T rows and g = M (o - mu), as the staging code of the kernels forms them
        const float m0x = f.isx * f.r0x, m0y = f.isx * f.r0y, m0z = f.isx * f.r0z;
        const float m1x = f.isy * f.r1x, m1y = f.isy * f.r1y, m1z = f.isy * f.r1z;
        const float m2x = f.isz * f.r2x, m2y = f.isz * f.r2y, m2z = f.isz * f.r2z;
        const float vx = ox - f.px, vy = oy - f.py, vz = oz - f.pz;
        const float gx = m0x * vx + m0y * vy + m0z * vz, gy = m1x * vx + m1y * vy + m1z * vz, gz = m2x * vx + m2y * vy + m2z * vz;
        const bool cand = degree == 4 ? block_candidate<4>(cfg, wf, m0x, m0y, m0z, m1x, m1y, m1z, m2x, m2y, m2z, gx, gy, gz, f.dns)
                                      : block_candidate<2>(cfg, wf, m0x, m0y, m0z, m1x, m1y, m1z, m2x, m2y, m2z, gx, gy, gz, f.dns);
        if (any) stats[0]++;
        if (!cand) stats[1]++;
        if (cand && !any) stats[2]++;
        if (any && !cand) {
            if (violations == 0) {
                memcpy(worst, particle, sizeof(particle));
                if (getenv("CULL_FUZZ_DEBUG")) {
                    fprintf(stderr, "violation: origin %g %g %g pitch %g depth %g scales %g %g %g dns %g |g| %g rect u[%g,%g] v[%g,%g]\n", ox, oy, oz, pitch, depth,
                            sc[0], sc[1], sc[2], particle[3], sqrtf(gx * gx + gy * gy + gz * gz), ulo, uhi, vlo, vhi);
                    for (int l = 0; l < 32; ++l) {
                        const CanonicalHit h = canonical_hit<2>(f, ox, oy, oz, rdx[l], rdy[l], rdz[l], min_density, min_alpha, max_alpha);
                        double gd[3] = {gx, gy, gz}, dd[3];
                        dd[0] = (double)m0x * rdx[l] + (double)m0y * rdy[l] + (double)m0z * rdz[l];
                        dd[1] = (double)m1x * rdx[l] + (double)m1y * rdy[l] + (double)m1z * rdz[l];
                        dd[2] = (double)m2x * rdx[l] + (double)m2y * rdy[l] + (double)m2z * rdz[l];
                        const double nl2 = dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2];
                        const double cx = dd[1] * gd[2] - dd[2] * gd[1], cy = dd[2] * gd[0] - dd[0] * gd[2], cz = dd[0] * gd[1] - dd[1] * gd[0];
                        if (h.accept) fprintf(stderr, "  lane %d accepted: gray(fp32) %g gray(fp64 on the same inputs) %g\n", l, h.gray, (cx * cx + cy * cy + cz * cz) / nl2);
                    }
                }
            }
            violations++;
        }
    }
    return violations;
}

}  // extern "C"
