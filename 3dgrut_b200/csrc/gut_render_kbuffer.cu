// 3dgrut_b200/csrc/gut_render_kbuffer.cu -- sorted 3DGUT: per-tile compositing with a per-ray k-buffer (GAUSSIAN_K_BUFFER_SIZE = K > 0)
// and its adjoint.  Parity vs the CPU restatement verified on a B200 in round 2 (tests/test_kbuffer_gpu.py: image 2e-7 mean, gradients rel-L2 3e-6).
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
// same record); gradients are scattered with four 16-byte vector reductions per processed hit into the [N,20] accumulator (kGradRow floats per row, the first 15 used here) G8 consumes.
#include "kbuffer_walk.cuh"

namespace gutb200 {

namespace {

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
    KForward acc;
    walk_kbuffer<DEG>(cfg, K, ray, valid, ranges[tile * 2], ranges[tile * 2 + 1], particles, sorted_values,
                      [&](float t, float alpha, uint32_t idx) -> bool { return kb_forward_hit(cfg, rgb, acc, t, alpha, idx); });
    const float T = acc.T, cr = acc.cr, cg = acc.cg, cb = acc.cb, dist = acc.dist;
    const uint32_t hits = acc.hits;
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
    KBackward st;  // initializeBackwardRay (kernels/cuda/common/rayPayloadBackward.cuh:31-73)
    if (valid) {
        const float4 o = reinterpret_cast<const float4*>(out_rgba)[pix];
        const float4 g = reinterpret_cast<const float4*>(d_rgba)[pix];
        st.Cix = o.x; st.Ciy = o.y; st.Ciz = o.z;
        st.Cgx = g.x; st.Cgy = g.y; st.Cgz = g.z;
        st.Tint = 1.f - o.w;
        st.Tgrad = -1.f * g.w;
        st.Dint = out_dist[pix];
        st.Dgrad = d_dist[pix];
    }
    auto scatter = [&](uint32_t idx, const float* g, const float* rg) {  // four 16-byte vector reductions into the accumulator row
        float4* row = reinterpret_cast<float4*>(grad_acc + static_cast<size_t>(idx) * kGradRow);
        atomicAdd(row + 0, make_float4(g[0], g[1], g[2], g[3]));
        atomicAdd(row + 1, make_float4(g[4], g[5], g[6], g[7]));
        atomicAdd(row + 2, make_float4(g[8], g[9], g[10], 0.f));
        atomicAdd(row + 3, make_float4(rg[0], rg[1], rg[2], 0.f));
    };
    walk_kbuffer<DEG>(cfg, K, ray, valid, ranges[tile * 2], ranges[tile * 2 + 1], particles, sorted_values,
                      [&](float /*t*/, float /*alpha*/, uint32_t idx) -> bool { return kb_backward_hit<DEG>(cfg, ray, particles, rgb, st, idx, scatter); });
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
