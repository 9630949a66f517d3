// 3dgrut_b200/csrc/kbuffer_walk.cuh -- device-side core of the sorted (k-buffer) 3DGUT variant: ray set-up, the K-slot buffer and the walk
// over a tile's depth-sorted list that yields the hits in compositing order.  Kept in a header so that the same code is compiled into the
// kernels (gut_render_kbuffer.cu) and into the host emulation the CPU tests run (tests/host_emul/kbuffer_host.cpp).
// Reference semantics: renderers/gutKBufferRenderer.cuh:62-112 (buffer), :274-352 (walk); kernels/cuda/common/rayPayload.cuh:76-108 (ray).
#pragma once
#include "gut_common.cuh"
#include "hit_math.cuh"

namespace gutb200 {

constexpr int kMaxK = 16;

struct KRay {
    float ox, oy, oz, dx, dy, dz, tmin, tmax;
    bool alive;
};

// initializeRay (kernels/cuda/common/rayPayload.cuh:76-108) with the +-1e6 scene box of splatRaster.cpp:240 (same as gut_render.cu)
__device__ __forceinline__ KRay make_kray(const FrameCamera& cam, const float* __restrict__ rays_o, const float* __restrict__ rays_d, int64_t pix) {
    KRay r;
    const float rox = rays_o[pix * 3 + 0], roy = rays_o[pix * 3 + 1], roz = rays_o[pix * 3 + 2];
    const float rdx = rays_d[pix * 3 + 0], rdy = rays_d[pix * 3 + 1], rdz = rays_d[pix * 3 + 2];
    const float* m = cam.s2w;
    r.ox = m[0] * rox + m[3] * roy + m[6] * roz + m[9];
    r.oy = m[1] * rox + m[4] * roy + m[7] * roz + m[10];
    r.oz = m[2] * rox + m[5] * roy + m[8] * roz + m[11];
    r.dx = m[0] * rdx + m[3] * rdy + m[6] * rdz;
    r.dy = m[1] * rdx + m[4] * rdy + m[7] * rdz;
    r.dz = m[2] * rdx + m[5] * rdy + m[8] * rdz;
    const float lo = -1e06f, hi = 1e06f;
    float tmin = (lo - r.ox) / r.dx, tmax = (hi - r.ox) / r.dx, t;
    if (tmin > tmax) { t = tmin; tmin = tmax; tmax = t; }
    float tymin = (lo - r.oy) / r.dy, tymax = (hi - r.oy) / r.dy;
    if (tymin > tymax) { t = tymin; tymin = tymax; tymax = t; }
    bool miss = (tmin > tymax) || (tymin > tmax);
    tmin = fmaxf(tmin, tymin);
    tmax = fminf(tmax, tymax);
    float tzmin = (lo - r.oz) / r.dz, tzmax = (hi - r.oz) / r.dz;
    if (tzmin > tzmax) { t = tzmin; tzmin = tzmax; tzmax = t; }
    miss = miss || (tmin > tzmax) || (tzmin > tmax);
    tmin = fmaxf(tmin, tzmin);
    tmax = fminf(tmax, tzmax);
    r.tmin = miss ? 3.4028235e+38f : fmaxf(tmin, 0.0f);
    r.tmax = miss ? 3.4028235e+38f : tmax;
    r.alive = r.tmax > r.tmin;
    return r;
}

struct KBuffer {
    float t[kMaxK];
    float alpha[kMaxK];
    uint32_t idx[kMaxK];
    int num;
};

// insert (:78-92): slots [0, K) ascending in t with the invalid (-1) entries in front; a full buffer loses its closest entry first
__device__ __forceinline__ void kb_insert(KBuffer& kb, int K, float t, float alpha, uint32_t idx) {
    if (kb.num == K) kb.t[0] = -1.0f; else kb.num++;
    for (int i = K - 1; i >= 0; --i) {
        if (t > kb.t[i]) {
            const float tt = kb.t[i], ta = kb.alpha[i];
            const uint32_t ti = kb.idx[i];
            kb.t[i] = t; kb.alpha[i] = alpha; kb.idx[i] = idx;
            t = tt; alpha = ta; idx = ti;
        }
    }
}

// Walks the tile's list for one ray and calls process(t, alpha, idx) for every hit in compositing order; process returns false when
// the ray is finished.  Shared by the forward and the backward kernel (both need the same sequence).
template <int DEG, typename Process>
__device__ __forceinline__ void walk_kbuffer(const FrameConfig& cfg, int K, const KRay& ray, bool alive, uint32_t begin, uint32_t end,
                                             const float* __restrict__ particles, const uint32_t* __restrict__ sorted_values, Process process) {
    KBuffer kb;
    kb.num = 0;
    for (int i = 0; i < kMaxK; ++i) { kb.t[i] = -1.0f; kb.alpha[i] = 0.f; kb.idx[i] = kInvalid; }
    for (uint32_t base = begin; base < end; base += 32) {
        if (__syncthreads_and(!alive)) break;
        const uint32_t stop = min(end, base + 32);
        for (uint32_t k = base; alive && k < stop; ++k) {
            const uint32_t idx = sorted_values[k];
            const ParticleFrame f = load_frame(particles, idx);
            const CanonicalHit h = canonical_hit<DEG>(f, ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz, cfg.min_kernel_density, cfg.min_alpha,
                                                      cfg.max_alpha);
            if (!h.accept) continue;
            const float t = hit_distance(f, h);
            if (!((t > ray.tmin) && (t < ray.tmax))) continue;
            if (kb.num == K) alive = process(kb.t[0], kb.alpha[0], kb.idx[0]);   // closestHit (:101-103)
            kb_insert(kb, K, t, h.alpha, idx);
        }
    }
    for (int i = 0; alive && i < kb.num; ++i) alive = process(kb.t[K - kb.num + i], kb.alpha[K - kb.num + i], kb.idx[K - kb.num + i]);
}

// ---- what happens to a hit when it leaves the buffer ------------------------------------------------------------------------------

struct KForward {  // densityIntegrateHit + featureIntegrateFwd (gutKBufferRenderer.cuh:199-217)
    float T = 1.f, cr = 0.f, cg = 0.f, cb = 0.f, dist = 0.f;
    uint32_t hits = 0;
};

__device__ __forceinline__ bool kb_forward_hit(const FrameConfig& cfg, const float* __restrict__ rgb, KForward& a, float t, float alpha,
                                               uint32_t idx) {
    const float w = alpha * a.T;
    a.dist += t * w;
    a.T *= (1.f - alpha);
    if (w > 0.f) {
        a.cr += fmaxf(rgb[idx * 3 + 0], 0.f) * w;
        a.cg += fmaxf(rgb[idx * 3 + 1], 0.f) * w;
        a.cb += fmaxf(rgb[idx * 3 + 2], 0.f) * w;
        a.hits++;
    }
    return !(a.T < cfg.min_transmittance);
}

struct KBackward {  // initializeBackwardRay (kernels/cuda/common/rayPayloadBackward.cuh:31-73) + the running state of the replay
    float Cix = 0.f, Ciy = 0.f, Ciz = 0.f, Cgx = 0.f, Cgy = 0.f, Cgz = 0.f, Tint = 1.f, Tgrad = 0.f, Dint = 0.f, Dgrad = 0.f;
    float T = 1.f, Cx = 0.f, Cy = 0.f, Cz = 0.f, D = 0.f;
};

// adjoint of one processed hit; sink(idx, g[11] = d(pos3, density, quat4, scale3), rg[3] = d(radiance)) scatters the gradients
template <int DEG, typename Sink>
__device__ __forceinline__ bool kb_backward_hit(const FrameConfig& cfg, const KRay& ray, const float* __restrict__ particles,
                                                const float* __restrict__ rgb, KBackward& b, uint32_t idx, Sink sink) {
    const ParticleFrame f = load_frame(particles, idx);
    const CanonicalHit h = canonical_hit<DEG>(f, ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz, cfg.min_kernel_density, cfg.min_alpha, cfg.max_alpha);
    float g[11], rg[3];
    hit_adjoint<DEG>(f, h, ray.dx, ray.dy, ray.dz, fmaxf(rgb[idx * 3 + 0], 0.f), fmaxf(rgb[idx * 3 + 1], 0.f), fmaxf(rgb[idx * 3 + 2], 0.f),
                     cfg.min_transmittance, b.Tint, b.Tgrad, b.Cix, b.Ciy, b.Ciz, b.Cgx, b.Cgy, b.Cgz, b.Dint, b.Dgrad, b.T, b.Cx, b.Cy, b.Cz, b.D, g, rg);
    sink(idx, g, rg);
    return !(b.T < cfg.min_transmittance);
}

}  // namespace gutb200
