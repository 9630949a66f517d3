// 3dgrut_b200/csrc/gut_render_kbuffer.cu -- sorted 3DGUT: per-tile compositing with a per-ray k-buffer (GAUSSIAN_K_BUFFER_SIZE = K > 0)
// and its adjoint.  EXPERIMENTAL: written at the end of round 1 after the GPU budget was spent -- compiled, never run on hardware; gated
// behind GUTB200_EXPERIMENTAL_KBUFFER=1 (gut_api.cu check_args) until tests/test_kbuffer_gpu.py has passed on a B200.
// First version, written for parity, not for speed (the default configuration is K = 0, gut_render.cu).
//
// Reference semantics restated (not copied) from threedgut_tracer/include/3dgut/kernels/cuda/renderers/gutKBufferRenderer.cuh:
//   :62-112   HitParticleKBufferT: K hits sorted by hit distance; when full, the closest one is composited before the new hit enters
//   :150-225  processHitParticle (forward branch: densityIntegrateHit + featureIntegrateFwd; backward branch: the hit's adjoint)
//   :274-352  evalKBuffer: walk the tile's depth-sorted list, then composite what is left in the buffer in order
// The reference obtains the backward of this variant from Slang autodiff; here the hand-derived per-hit adjoint of the K = 0 path
// (hit_math.cuh: hit_adjoint, restated from models/gaussianParticles.cuh:484-751) is applied in the buffer's processing order, which is
// the exact gradient of the same forward (the CPU restatement used by the tests checks this against torch autograd).
// One CTA per 16x16 tile, one thread per pixel; particle records are read through the read-only cache (every thread of the CTA reads the
// same record); gradients are scattered with four 16-byte vector reductions per processed hit into the [N,16] accumulator G8 consumes.
#include "gut_common.cuh"
#include "hit_math.cuh"

namespace gutb200 {

namespace {

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

template <int DEG>
__global__ void __launch_bounds__(kTilePixels) render_forward_kbuffer_kernel(FrameCamera cam, FrameConfig cfg, int K,
                                                                             const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                                             const float* __restrict__ particles, const float* __restrict__ rgb,
                                                                             const uint32_t* __restrict__ sorted_values,
                                                                             const uint32_t* __restrict__ ranges, float* __restrict__ out_rgba,
                                                                             float* __restrict__ out_dist, float* __restrict__ out_hits) {
    const int tile = blockIdx.x;
    const int tx = tile % cam.grid_x, ty = tile / cam.grid_x;
    const int px = tx * kTile + (threadIdx.x & 15), py = ty * kTile + (threadIdx.x >> 4);
    const bool inside = (px < cam.width) && (py < cam.height);
    const int64_t pix = static_cast<int64_t>(py) * cam.width + px;
    KRay ray;
    ray.alive = false;
    if (inside) ray = make_kray(cam, rays_o, rays_d, pix);
    const bool valid = inside && ray.alive;
    float T = 1.f, cr = 0.f, cg = 0.f, cb = 0.f, dist = 0.f;
    uint32_t hits = 0;
    walk_kbuffer<DEG>(cfg, K, ray, valid, ranges[tile * 2], ranges[tile * 2 + 1], particles, sorted_values,
                      [&](float t, float alpha, uint32_t idx) -> bool {
                          const float w = alpha * T;   // densityIntegrateHit + featureIntegrateFwd (:199-217)
                          dist += t * w;
                          T *= (1.f - alpha);
                          if (w > 0.f) {
                              cr += fmaxf(rgb[idx * 3 + 0], 0.f) * w;
                              cg += fmaxf(rgb[idx * 3 + 1], 0.f) * w;
                              cb += fmaxf(rgb[idx * 3 + 2], 0.f) * w;
                              hits++;
                          }
                          return !(T < cfg.min_transmittance);
                      });
    if (valid) {
        reinterpret_cast<float4*>(out_rgba)[pix] = make_float4(cr, cg, cb, 1.0f - T);
        out_dist[pix] = dist;
        out_hits[pix] = static_cast<float>(hits);
    } else if (inside) {
        reinterpret_cast<float4*>(out_rgba)[pix] = make_float4(0.f, 0.f, 0.f, 0.f);
        out_dist[pix] = 1e06f;
        out_hits[pix] = 0.f;
    }
}

template <int DEG>
__global__ void __launch_bounds__(kTilePixels) render_backward_kbuffer_kernel(FrameCamera cam, FrameConfig cfg, int K,
                                                                              const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                                              const float* __restrict__ particles, const float* __restrict__ rgb,
                                                                              const uint32_t* __restrict__ sorted_values,
                                                                              const uint32_t* __restrict__ ranges, const float* __restrict__ out_rgba,
                                                                              const float* __restrict__ d_rgba, const float* __restrict__ out_dist,
                                                                              const float* __restrict__ d_dist, float* __restrict__ grad_acc) {
    const int tile = blockIdx.x;
    const int tx = tile % cam.grid_x, ty = tile / cam.grid_x;
    const int px = tx * kTile + (threadIdx.x & 15), py = ty * kTile + (threadIdx.x >> 4);
    const bool inside = (px < cam.width) && (py < cam.height);
    const int64_t pix = static_cast<int64_t>(py) * cam.width + px;
    KRay ray;
    ray.alive = false;
    if (inside) ray = make_kray(cam, rays_o, rays_d, pix);
    const bool valid = inside && ray.alive;
    // initializeBackwardRay (kernels/cuda/common/rayPayloadBackward.cuh:31-73)
    float Cix = 0.f, Ciy = 0.f, Ciz = 0.f, Cgx = 0.f, Cgy = 0.f, Cgz = 0.f, Tint = 1.f, Tgrad = 0.f, Dint = 0.f, Dgrad = 0.f;
    if (valid) {
        const float4 o = reinterpret_cast<const float4*>(out_rgba)[pix];
        const float4 g = reinterpret_cast<const float4*>(d_rgba)[pix];
        Cix = o.x; Ciy = o.y; Ciz = o.z;
        Cgx = g.x; Cgy = g.y; Cgz = g.z;
        Tint = 1.f - o.w;
        Tgrad = -1.f * g.w;
        Dint = out_dist[pix];
        Dgrad = d_dist[pix];
    }
    float T = 1.f, Cx = 0.f, Cy = 0.f, Cz = 0.f, D = 0.f;
    walk_kbuffer<DEG>(cfg, K, ray, valid, ranges[tile * 2], ranges[tile * 2 + 1], particles, sorted_values,
                      [&](float /*t*/, float /*alpha*/, uint32_t idx) -> bool {
                          const ParticleFrame f = load_frame(particles, idx);
                          const CanonicalHit h = canonical_hit<DEG>(f, ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz, cfg.min_kernel_density,
                                                                    cfg.min_alpha, cfg.max_alpha);
                          float g[11], rg[3];
                          hit_adjoint<DEG>(f, h, ray.dx, ray.dy, ray.dz, fmaxf(rgb[idx * 3 + 0], 0.f), fmaxf(rgb[idx * 3 + 1], 0.f),
                                           fmaxf(rgb[idx * 3 + 2], 0.f), cfg.min_transmittance, Tint, Tgrad, Cix, Ciy, Ciz, Cgx, Cgy, Cgz, Dint, Dgrad,
                                           T, Cx, Cy, Cz, D, g, rg);
                          float4* row = reinterpret_cast<float4*>(grad_acc + static_cast<size_t>(idx) * kGradRow);
                          atomicAdd(row + 0, make_float4(g[0], g[1], g[2], g[3]));
                          atomicAdd(row + 1, make_float4(g[4], g[5], g[6], g[7]));
                          atomicAdd(row + 2, make_float4(g[8], g[9], g[10], 0.f));
                          atomicAdd(row + 3, make_float4(rg[0], rg[1], rg[2], 0.f));
                          return !(T < cfg.min_transmittance);
                      });
}

}  // namespace

void launch_render_forward_kbuffer(cudaStream_t s, const FrameCamera& cam, const FrameConfig& cfg, int K, const float* rays_o, const float* rays_d,
                                   const float* particles, const float* rgb, const uint32_t* sorted_values, const uint32_t* ranges,
                                   float* out_rgba, float* out_dist, float* out_hits) {
    const unsigned grid = cam.grid_x * cam.grid_y;
    if (cfg.kernel_degree == 4)
        render_forward_kbuffer_kernel<4><<<grid, kTilePixels, 0, s>>>(cam, cfg, K, rays_o, rays_d, particles, rgb, sorted_values, ranges, out_rgba, out_dist, out_hits);
    else
        render_forward_kbuffer_kernel<2><<<grid, kTilePixels, 0, s>>>(cam, cfg, K, rays_o, rays_d, particles, rgb, sorted_values, ranges, out_rgba, out_dist, out_hits);
}

void launch_render_backward_kbuffer(cudaStream_t s, const FrameCamera& cam, const FrameConfig& cfg, int K, const float* rays_o, const float* rays_d,
                                    const float* particles, const float* rgb, const uint32_t* sorted_values, const uint32_t* ranges,
                                    const float* out_rgba, const float* d_rgba, const float* out_dist, const float* d_dist, float* grad_acc) {
    const unsigned grid = cam.grid_x * cam.grid_y;
    if (cfg.kernel_degree == 4)
        render_backward_kbuffer_kernel<4><<<grid, kTilePixels, 0, s>>>(cam, cfg, K, rays_o, rays_d, particles, rgb, sorted_values, ranges, out_rgba, d_rgba, out_dist, d_dist, grad_acc);
    else
        render_backward_kbuffer_kernel<2><<<grid, kTilePixels, 0, s>>>(cam, cfg, K, rays_o, rays_d, particles, rgb, sorted_values, ranges, out_rgba, d_rgba, out_dist, d_dist, grad_acc);
}

}  // namespace gutb200
