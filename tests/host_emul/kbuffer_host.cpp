// tests/host_emul/kbuffer_host.cpp -- TEST INFRASTRUCTURE: compiles the DEVICE-side core of the sorted 3DGUT kernels
// (3dgrut_b200/csrc/kbuffer_walk.cuh + hit_math.cuh, the very headers gut_render_kbuffer.cu is built from) with g++ and runs it one pixel
// at a time on the host, so that the CPU suite can check the k-buffer walk, the per-hit compositing and the per-hit adjoint against the
// oracle before the kernels ever reach a GPU.  What it cannot cover: the thread / tile indexing, the launch configuration and the
// vector atomics of the kernels themselves.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
using std::max;
using std::min;
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
#define __expf(x) expf(x)  // glibc declares (but does not export) __expf: map the device intrinsic onto expf
static inline float4 __ldg(const float4* p) { return *p; }
static inline int __syncthreads_and(int p) { return p; }  // a one-thread block
#define __CUDACC__ 1
#include "../../3dgrut_b200/csrc/kbuffer_walk.cuh"

using namespace gutb200;

extern "C" {

// cam_s2w: 12 floats (4 columns x 3, sensor -> world at mid exposure); lists as the CUDA pipeline produces them
void kbuffer_host_forward(int degree, float min_density, float min_alpha, float max_alpha, float min_t, int K, int width, int height,
                          const float* cam_s2w, const float* rays_o, const float* rays_d, const float* particles, const float* rgb,
                          const uint32_t* sorted_values, const uint32_t* ranges, float* out_rgba, float* out_dist, float* out_hits) {
    FrameCamera cam;
    memset(&cam, 0, sizeof(cam));
    cam.width = width; cam.height = height;
    cam.grid_x = (width + kTile - 1) / kTile; cam.grid_y = (height + kTile - 1) / kTile;
    memcpy(cam.s2w, cam_s2w, sizeof(cam.s2w));
    FrameConfig cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.kernel_degree = degree; cfg.min_kernel_density = min_density; cfg.min_alpha = min_alpha; cfg.max_alpha = max_alpha; cfg.min_transmittance = min_t;
    for (int py = 0; py < height; ++py)
        for (int px = 0; px < width; ++px) {
            const int64_t pix = static_cast<int64_t>(py) * width + px;
            const int tile = (py / kTile) * cam.grid_x + px / kTile;
            const KRay ray = make_kray(cam, rays_o, rays_d, pix);
            KForward acc;
            auto proc = [&](float t, float alpha, uint32_t idx) -> bool { return kb_forward_hit(cfg, rgb, acc, t, alpha, idx); };
            if (degree == 4) walk_kbuffer<4>(cfg, K, ray, ray.alive, ranges[tile * 2], ranges[tile * 2 + 1], particles, sorted_values, proc);
            else walk_kbuffer<2>(cfg, K, ray, ray.alive, ranges[tile * 2], ranges[tile * 2 + 1], particles, sorted_values, proc);
            if (ray.alive) {
                out_rgba[pix * 4] = acc.cr; out_rgba[pix * 4 + 1] = acc.cg; out_rgba[pix * 4 + 2] = acc.cb; out_rgba[pix * 4 + 3] = 1.0f - acc.T;
                out_dist[pix] = acc.dist; out_hits[pix] = static_cast<float>(acc.hits);
            } else {
                out_rgba[pix * 4] = out_rgba[pix * 4 + 1] = out_rgba[pix * 4 + 2] = out_rgba[pix * 4 + 3] = 0.f;
                out_dist[pix] = 1e06f; out_hits[pix] = 0.f;
            }
        }
}

// grad_acc: [N,16] rows as the CUDA accumulator (0..2 pos, 3 density, 4..7 quat, 8..10 scale, 12..14 radiance), zeroed by the caller
void kbuffer_host_backward(int degree, float min_density, float min_alpha, float max_alpha, float min_t, int K, int width, int height,
                           const float* cam_s2w, const float* rays_o, const float* rays_d, const float* particles, const float* rgb,
                           const uint32_t* sorted_values, const uint32_t* ranges, const float* out_rgba, const float* d_rgba,
                           const float* out_dist, const float* d_dist, double* grad_acc) {
    FrameCamera cam;
    memset(&cam, 0, sizeof(cam));
    cam.width = width; cam.height = height;
    cam.grid_x = (width + kTile - 1) / kTile; cam.grid_y = (height + kTile - 1) / kTile;
    memcpy(cam.s2w, cam_s2w, sizeof(cam.s2w));
    FrameConfig cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.kernel_degree = degree; cfg.min_kernel_density = min_density; cfg.min_alpha = min_alpha; cfg.max_alpha = max_alpha; cfg.min_transmittance = min_t;
    for (int py = 0; py < height; ++py)
        for (int px = 0; px < width; ++px) {
            const int64_t pix = static_cast<int64_t>(py) * width + px;
            const int tile = (py / kTile) * cam.grid_x + px / kTile;
            const KRay ray = make_kray(cam, rays_o, rays_d, pix);
            if (!ray.alive) continue;
            KBackward st;
            st.Cix = out_rgba[pix * 4]; st.Ciy = out_rgba[pix * 4 + 1]; st.Ciz = out_rgba[pix * 4 + 2];
            st.Cgx = d_rgba[pix * 4]; st.Cgy = d_rgba[pix * 4 + 1]; st.Cgz = d_rgba[pix * 4 + 2];
            st.Tint = 1.f - out_rgba[pix * 4 + 3];
            st.Tgrad = -1.f * d_rgba[pix * 4 + 3];
            st.Dint = out_dist[pix];
            st.Dgrad = d_dist[pix];
            auto scatter = [&](uint32_t idx, const float* g, const float* rg) {
                double* row = grad_acc + static_cast<size_t>(idx) * 16;
                for (int q = 0; q < 11; ++q) row[q] += g[q];
                for (int q = 0; q < 3; ++q) row[12 + q] += rg[q];
            };
            auto proc = [&](float, float, uint32_t idx) -> bool {
                return degree == 4 ? kb_backward_hit<4>(cfg, ray, particles, rgb, st, idx, scatter) : kb_backward_hit<2>(cfg, ray, particles, rgb, st, idx, scatter);
            };
            if (degree == 4) walk_kbuffer<4>(cfg, K, ray, true, ranges[tile * 2], ranges[tile * 2 + 1], particles, sorted_values, proc);
            else walk_kbuffer<2>(cfg, K, ray, true, ranges[tile * 2], ranges[tile * 2 + 1], particles, sorted_values, proc);
        }
}

}  // extern "C"
