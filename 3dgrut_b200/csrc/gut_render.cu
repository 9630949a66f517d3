// 3dgrut_b200/csrc/gut_render.cu -- per-tile compositing (G6), its adjoint (G7) and the per-particle
// spherical-harmonics adjoint (G8) of the 3DGUT path.
//
// Reference semantics restated (not copied) from
//   G6  threedgut_tracer/include/3dgut/kernels/cuda/renderers/gutKBufferRenderer.cuh:274-352 (k-buffer = 0)
//       + kernels/slang/models/gaussianParticles.slang:96-274 (canonical ray, response, integrate)
//   G7  gutKBufferRenderer.cuh:642-716 + kernels/cuda/models/gaussianParticles.cuh:484-751 (hand-written adjoint)
//   G8  kernels/cuda/renderers/gutProjector.cuh:390-430 + slang/common/sphericalHarmonics.slang:21-64
//
// Design (sub-tile culling: see the section comment below): one CTA per 16x16 tile (the tile id is part of the sort key, so the tile
// shape is fixed by parity), 256 threads = 256 pixels, sorted particle lists consumed in batches staged in
// shared memory as render-ready records (scale folded into the rotation rows once per staged particle
// instead of once per pixel test).  Both kernels are FP32/SFU-issue bound, not HBM bound.
// G7 reduces the 14 per-particle gradient floats across the warp with a 16-value transposing butterfly
// (16 SHFL instead of 70) and lands them with ONE coalesced 64-byte vector RED per (warp, particle) into a
// [N,16] accumulator that G8 consumes and re-zeroes.
#include "gut_common.cuh"
#include "hit_math.cuh"
#include "subtile_cull.cuh"
#include "tma.cuh"

namespace gutb200 {

namespace {

constexpr int kBatch = 256;
constexpr unsigned kFull = 0xFFFFFFFFu;

struct Ray {
    float ox, oy, oz, dx, dy, dz, tmin, tmax;
    bool alive;
};

// initializeRay (kernels/cuda/common/rayPayload.cuh:76-108) with the +-1e6 scene box of splatRaster.cpp:240
__device__ __forceinline__ Ray make_ray(const FrameCamera& cam, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                        int64_t pix) {
    Ray r;
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

template <int DEG>
__device__ __forceinline__ float kernel_response(float gray) {
    // generalized Gaussian exp(-4.5/3^DEG * |x|^DEG) on the squared canonical distance (gaussianParticles.cuh:267-308)
    if (DEG == 4) return __expf(-0.0555555555556f * gray * gray);
    return __expf(-0.5f * gray);
}

template <int DEG>
__device__ __forceinline__ float kernel_response_grad(float gray, float gres, float gres_grad) {
    if (DEG == 4) return (-0.0555555555556f * 2.0f) * gray * gres * gres_grad;  // gaussianParticles.cuh:239-243
    return -0.5f * gres * gres_grad;                                             // :259-263
}

// ----------------------------------------------------------------------------------------------------------
// Sub-tile culling (ours; the reference runs the full test on all 256 pixels of a tile for every entry of the tile's list).
//
// For rays with a common origin o the accept test of a (pixel, particle) pair,
//     |normalize(M d) x g|^2 < r^2,   M = S^-1 R^T, g = M (o - mu),  r^2 = r^2(min response, min alpha / density),
// is a quadratic inequality in the projective coordinates (u, v) of the ray in a local frame (e1, e2, e3), d ~ e3 + u e1 + v e2:
//     f(u,v) = |c3 + u c1 + v c2|^2 - r^2 |a3 + u a1 + v a2|^2 < 0,   a_i = M e_i,  c_i = a_i x g
// (same cross-product formulation as the exact test, hence the same conditioning).  Each warp owns an 8x4 pixel block; before it
// walks a chunk of 32 list entries, LANE k decides for ENTRY k whether {f < 0} can meet the block's (u, v) rectangle at all --
// the minimum of the convex quadratic over the rectangle, with r^2 inflated and a bound of the fp32 evaluation error subtracted,
// so the decision is a NECESSARY condition of the exact test.  The warp then runs the exact test only on the entries whose
// ballot bit is set (48 % of the warp iterations on C2).  Pairs dropped could never be accepted: outputs are bit-identical with
// the switch on and off (tests/test_gut_parity_gpu.py::test_subtile_culling_is_bit_identical).

__device__ __forceinline__ WarpFrame make_warp_frame(const FrameCamera& cam, const Ray& ray, bool alive, bool enabled, int lane) {
    WarpFrame wf;
    wf.on = false;
    const unsigned live = __ballot_sync(kFull, alive);
    if (!enabled || live == 0u) return wf;
    const int src = __ffs(live) - 1;
    float dx = __shfl_sync(kFull, ray.dx, src), dy = __shfl_sync(kFull, ray.dy, src), dz = __shfl_sync(kFull, ray.dz, src);
    if (!frame_axes(cam.s2w, dx, dy, dz, wf)) return wf;
    // this lane's ray in the frame; every live ray must point within 60 degrees of e3
    float u = 0.f, v = 0.f;
    const bool fine = !alive || ray_uv(wf, ray.dx, ray.dy, ray.dz, u, v);
    float ulo = alive ? u : 3.0e38f, uhi = alive ? u : -3.0e38f, vlo = alive ? v : 3.0e38f, vhi = alive ? v : -3.0e38f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ulo = fminf(ulo, __shfl_xor_sync(kFull, ulo, o));
        uhi = fmaxf(uhi, __shfl_xor_sync(kFull, uhi, o));
        vlo = fminf(vlo, __shfl_xor_sync(kFull, vlo, o));
        vhi = fmaxf(vhi, __shfl_xor_sync(kFull, vhi, o));
    }
    // half a pixel of slack is not needed: the rectangle is the hull of the rays themselves
    wf.ulo = ulo; wf.uhi = uhi; wf.vlo = vlo; wf.vhi = vhi;
    wf.umax = fmaxf(fmaxf(fabsf(ulo), fabsf(uhi)), fmaxf(fabsf(vlo), fabsf(vhi)));
    wf.on = __all_sync(kFull, fine);
    return wf;
}

// ----------------------------------------------------------------------------------------------------------
// G6 forward
// staged record, 5 x float4: rows of M = diag(1/s) R^T with the particle position in .w, then (s, density), (rgb)

struct FwdSmem {
    float4 m0[kBatch], m1[kBatch], m2[kBatch], sd[kBatch], col[kBatch];
};

// pixel of thread `tid` in tile (tx,ty): a warp covers an 8x4 pixel block (not a 16x2 strip) -- hits are spatially
// coherent, so a compact footprint keeps more lanes on the same side of the accept branch
__device__ __forceinline__ void tile_pixel(int tile, int grid_x, int tid, int& px, int& py) {
    const int tx = tile % grid_x, ty = tile / grid_x;
    px = tx * kTile + ((tid >> 5) & 1) * 8 + (tid & 7);
    py = ty * kTile + (tid >> 6) * 4 + ((tid >> 3) & 3);
}

// world-space origin of the tile's first pixel; when every ray of the tile starts there (always the case for the
// camera rays the projection stage assumes) the canonical origin S^-1 R^T (o - mu) is computed once per staged
// particle instead of once per (pixel, particle)
__device__ __forceinline__ bool tile_common_origin(const FrameCamera& cam, const float* __restrict__ rays_o, int tile, bool inside,
                                                   int64_t pix, float& ox, float& oy, float& oz) {
    const int tx = tile % cam.grid_x, ty = tile / cam.grid_x;
    const int64_t pix0 = static_cast<int64_t>(ty * kTile) * cam.width + tx * kTile;
    const float ax = rays_o[pix0 * 3 + 0], ay = rays_o[pix0 * 3 + 1], az = rays_o[pix0 * 3 + 2];
    bool same = true;
    if (inside) same = (rays_o[pix * 3 + 0] == ax) && (rays_o[pix * 3 + 1] == ay) && (rays_o[pix * 3 + 2] == az);
    const float* m = cam.s2w;
    ox = m[0] * ax + m[3] * ay + m[6] * az + m[9];
    oy = m[1] * ax + m[4] * ay + m[7] * az + m[10];
    oz = m[2] * ax + m[5] * ay + m[8] * az + m[11];
    return __syncthreads_and(same);
}

// exact test + compositing of one (pixel, staged entry j) pair
template <int DEG, bool UNIFORM>
__device__ __forceinline__ void forward_pair(const FrameConfig& cfg, const FwdSmem& sm, int j, const Ray& ray, bool& alive, float& T, float& cr,
                                             float& cg, float& cb, float& dist, uint32_t& hits) {
    const float4 m0 = sm.m0[j], m1 = sm.m1[j], m2 = sm.m2[j];
    float gox, goy, goz;
    if (UNIFORM) {
        gox = m0.w; goy = m1.w; goz = m2.w;
    } else {
        const float vx = ray.ox - m0.w, vy = ray.oy - m1.w, vz = ray.oz - m2.w;
        gox = m0.x * vx + m0.y * vy + m0.z * vz;
        goy = m1.x * vx + m1.y * vy + m1.z * vz;
        goz = m2.x * vx + m2.y * vy + m2.z * vz;
    }
    const float ax = m0.x * ray.dx + m0.y * ray.dy + m0.z * ray.dz;
    const float ay = m1.x * ray.dx + m1.y * ray.dy + m1.z * ray.dz;
    const float az = m2.x * ray.dx + m2.y * ray.dy + m2.z * ray.dz;
    const float l = ax * ax + ay * ay + az * az;
    const float il = l > 0.f ? rsqrtf(l) : 1.f;
    const float gdx = ax * il, gdy = ay * il, gdz = az * il;
    const float ccx = gdy * goz - gdz * goy, ccy = gdz * gox - gdx * goz, ccz = gdx * goy - gdy * gox;
    const float gray = ccx * ccx + ccy * ccy + ccz * ccz;
    const float gres = kernel_response<DEG>(gray);
    const float4 sd = sm.sd[j];
    const float alpha = fminf(cfg.max_alpha, gres * sd.w);
    if ((gres > cfg.min_kernel_density) && (alpha > cfg.min_alpha)) {
        const float pd = -(gdx * gox + gdy * goy + gdz * goz);
        const float hx = sd.x * gdx * pd, hy = sd.y * gdy * pd, hz = sd.z * gdz * pd;
        const float t = sqrtf(hx * hx + hy * hy + hz * hz);
        if ((t > ray.tmin) && (t < ray.tmax)) {
            const float w = alpha * T;
            dist += t * w;
            T *= (1.f - alpha);
            if (w > 0.f) {
                const float4 c = sm.col[j];
                cr += c.x * w;
                cg += c.y * w;
                cb += c.z * w;
                hits++;
            }
            if (T < cfg.min_transmittance) alive = false;
        }
    }
}

template <int DEG, bool UNIFORM>
__device__ __forceinline__ void forward_tile(const FrameConfig& cfg, FwdSmem& sm, const WarpFrame& wf, const Ray& ray, float o0x, float o0y,
                                             float o0z, int tid, uint32_t begin, uint32_t end, const float* __restrict__ particles,
                                             const float* __restrict__ rgb, const uint32_t* __restrict__ sorted_values, bool& alive,
                                             float& T, float& cr, float& cg, float& cb, float& dist, uint32_t& hits) {
    const int lane = tid & 31;
    for (uint32_t base = begin; base < end; base += kBatch) {
        if (__syncthreads_and(!alive)) break;
        const uint32_t k = base + tid;
        if (k < end) {
            const uint32_t idx = sorted_values[k];
            const float4* p4 = reinterpret_cast<const float4*>(particles) + static_cast<size_t>(idx) * 3;
            const float4 a = __ldg(p4), q = __ldg(p4 + 1), s = __ldg(p4 + 2);
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
            const float rx = r * x, ry = r * y, rz = r * z;
            const float isx = 1.0f / s.x, isy = 1.0f / s.y, isz = 1.0f / s.z;
            float4 m0 = make_float4(isx * (1.f - 2.f * (yy + zz)), isx * (2.f * (xy + rz)), isx * (2.f * (xz - ry)), a.x);
            float4 m1 = make_float4(isy * (2.f * (xy - rz)), isy * (1.f - 2.f * (xx + zz)), isy * (2.f * (yz + rx)), a.y);
            float4 m2 = make_float4(isz * (2.f * (xz + ry)), isz * (2.f * (yz - rx)), isz * (1.f - 2.f * (xx + yy)), a.z);
            if (UNIFORM) {  // .w carries the canonical origin instead of the particle position
                const float vx = o0x - a.x, vy = o0y - a.y, vz = o0z - a.z;
                m0.w = m0.x * vx + m0.y * vy + m0.z * vz;
                m1.w = m1.x * vx + m1.y * vy + m1.z * vz;
                m2.w = m2.x * vx + m2.y * vy + m2.z * vz;
            }
            sm.m0[tid] = m0;
            sm.m1[tid] = m1;
            sm.m2[tid] = m2;
            sm.sd[tid] = make_float4(s.x, s.y, s.z, a.w);
            sm.col[tid] = make_float4(fmaxf(rgb[idx * 3 + 0], 0.f), fmaxf(rgb[idx * 3 + 1], 0.f), fmaxf(rgb[idx * 3 + 2], 0.f), 0.f);
        }
        __syncthreads();
        const int count = min(kBatch, static_cast<int>(end - base));
        if (UNIFORM) {
            // chunks of 32 entries: lane k screens entry k against the warp's pixel block, the warp walks the survivors
            for (int c = 0; c < count; c += 32) {
                if (!__any_sync(kFull, alive)) break;
                const int e = c + lane;
                bool cand = e < count;
                if (wf.on && cand) {
                    const float4 m0 = sm.m0[e], m1 = sm.m1[e], m2 = sm.m2[e];
                    cand = block_candidate<DEG>(cfg, wf, m0.x, m0.y, m0.z, m1.x, m1.y, m1.z, m2.x, m2.y, m2.z, m0.w, m1.w, m2.w, sm.sd[e].w);
                }
                unsigned todo = __ballot_sync(kFull, cand);
                while (todo) {
                    const int j = c + __ffs(todo) - 1;
                    todo &= todo - 1;
                    if (alive) forward_pair<DEG, true>(cfg, sm, j, ray, alive, T, cr, cg, cb, dist, hits);
                }
            }
        } else {
            for (int j = 0; alive && j < count; ++j) forward_pair<DEG, false>(cfg, sm, j, ray, alive, T, cr, cg, cb, dist, hits);
        }
    }
}

template <int DEG>
__global__ void __launch_bounds__(kTilePixels) render_forward_kernel(FrameCamera cam, FrameConfig cfg,
                                                                     const float* __restrict__ rays_o,
                                                                     const float* __restrict__ rays_d,
                                                                     const float* __restrict__ particles,
                                                                     const float* __restrict__ rgb,
                                                                     const uint32_t* __restrict__ sorted_values,
                                                                     const uint32_t* __restrict__ ranges,
                                                                     const uint32_t* __restrict__ tile_order, float* __restrict__ out_rgba,
                                                                     float* __restrict__ out_dist, float* __restrict__ out_hits) {
    __shared__ FwdSmem sm;
    const int tile = tile_order[blockIdx.x];  // heaviest tiles first (tile_order_kernel): shortens the tail of the grid
    const int tid = threadIdx.x;
    int px, py;
    tile_pixel(tile, cam.grid_x, tid, px, py);
    const bool inside = (px < cam.width) && (py < cam.height);
    const int64_t pix = static_cast<int64_t>(py) * cam.width + px;

    Ray ray;
    ray.alive = false;
    if (inside) ray = make_ray(cam, rays_o, rays_d, pix);
    const bool valid = inside && ray.alive;
    float o0x, o0y, o0z;
    const bool uniform = tile_common_origin(cam, rays_o, tile, inside, pix, o0x, o0y, o0z);
    const WarpFrame wf = make_warp_frame(cam, ray, valid, uniform && (cfg.subtile_culling & 2), tid & 31);

    float T = 1.f, cr = 0.f, cg = 0.f, cb = 0.f, dist = 0.f;
    uint32_t hits = 0;
    bool alive = valid;
    const uint32_t begin = ranges[tile * 2], end = ranges[tile * 2 + 1];
    if (uniform)
        forward_tile<DEG, true>(cfg, sm, wf, ray, o0x, o0y, o0z, tid, begin, end, particles, rgb, sorted_values, alive, T, cr, cg, cb, dist, hits);
    else
        forward_tile<DEG, false>(cfg, sm, wf, ray, o0x, o0y, o0z, tid, begin, end, particles, rgb, sorted_values, alive, T, cr, cg, cb, dist, hits);

    if (valid) {  // finalizeRay (rayPayload.cuh:160-193); invalid rays keep the initial buffer values
        reinterpret_cast<float4*>(out_rgba)[pix] = make_float4(cr, cg, cb, 1.0f - T);
        out_dist[pix] = dist;
        out_hits[pix] = static_cast<float>(hits);
    } else if (inside) {
        reinterpret_cast<float4*>(out_rgba)[pix] = make_float4(0.f, 0.f, 0.f, 0.f);
        out_dist[pix] = 1e06f;  // torch::ones(...)*1e6 (splatRaster.cpp:213)
        out_hits[pix] = 0.f;
    }
}

// Order tiles by decreasing list length (bucketed by log2): one small single-CTA kernel per frame.
__global__ void __launch_bounds__(1024) tile_order_kernel(int num_tiles, const uint32_t* __restrict__ ranges, uint32_t* __restrict__ order) {
    __shared__ uint32_t hist[34];
    if (threadIdx.x < 34) hist[threadIdx.x] = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < num_tiles; t += blockDim.x) {
        const uint32_t c = ranges[t * 2 + 1] - ranges[t * 2];
        atomicAdd(&hist[__clz(c) + 1], 1u);  // __clz(0) = 32 -> last bucket; long lists -> small bucket index
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int b = 1; b < 34; ++b) {
            const uint32_t h = hist[b];
            hist[b] = run;
            run += h;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < num_tiles; t += blockDim.x) {
        const uint32_t c = ranges[t * 2 + 1] - ranges[t * 2];
        order[atomicAdd(&hist[__clz(c) + 1], 1u)] = static_cast<uint32_t>(t);
    }
}

// ----------------------------------------------------------------------------------------------------------
// G7 backward
// staged record, 7 x float4:
//   r0 = rot row0, pos.x   r1 = rot row1, pos.y   r2 = rot row2, pos.z   (rows of quaternionWXYZToMatrix = columns of R)
//   sc = scale.xyz, density     is = 1/scale.xyz, _     qt = quat wxyz     cl = clamped rgb, particle index bits

struct BwdSmem {
    float4 r0[kBatch], r1[kBatch], r2[kBatch], sc[kBatch], is[kBatch], qt[kBatch], cl[kBatch];
    float4 go[kBatch];  // canonical origin of the tile's common ray origin (UNIFORM path)
};

// sum 16 per-lane values over the warp; lane L returns the total of component (L >> 1). 16 SHFL in all.
__device__ __forceinline__ float warp_transpose_reduce16(float (&v)[16], int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool up = lane & 16;
        const float send = up ? v[i] : v[i + 8];
        const float keep = up ? v[i + 8] : v[i];
        v[i] = keep + __shfl_xor_sync(kFull, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool up = lane & 8;
        const float send = up ? v[i] : v[i + 4];
        const float keep = up ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(kFull, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const bool up = lane & 4;
        const float send = up ? v[i] : v[i + 2];
        const float keep = up ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(kFull, send, 4);
    }
    {
        const bool up = lane & 2;
        const float send = up ? v[0] : v[1];
        const float keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(kFull, send, 2);
    }
    v[0] += __shfl_xor_sync(kFull, v[0], 1);
    return v[0];
}

template <int DEG, bool UNIFORM>
__device__ __forceinline__ void backward_tile(const FrameConfig& cfg, BwdSmem& sm, const WarpFrame& wf, const Ray& ray, float o0x, float o0y,
                                              float o0z, int tid, int lane, uint32_t begin, uint32_t end, const float* __restrict__ particles,
                                              const float* __restrict__ rgb, const uint32_t* __restrict__ sorted_values, bool alive,
                                              float Cix, float Ciy, float Ciz, float Cgx, float Cgy, float Cgz, float Tint, float Tgrad,
                                              float Dint, float Dgrad, float* __restrict__ grad_acc) {
    float T = 1.f, Cx = 0.f, Cy = 0.f, Cz = 0.f, D = 0.f;
    for (uint32_t base = begin; base < end; base += kBatch) {
        if (__syncthreads_and(!alive)) break;
        const uint32_t k = base + tid;
        if (k < end) {
            const uint32_t idx = sorted_values[k];
            const float4* p4 = reinterpret_cast<const float4*>(particles) + static_cast<size_t>(idx) * 3;
            const float4 a = __ldg(p4), q = __ldg(p4 + 1), s = __ldg(p4 + 2);
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
            const float rx = r * x, ry = r * y, rz = r * z;
            sm.r0[tid] = make_float4(1.f - 2.f * (yy + zz), 2.f * (xy + rz), 2.f * (xz - ry), a.x);
            sm.r1[tid] = make_float4(2.f * (xy - rz), 1.f - 2.f * (xx + zz), 2.f * (yz + rx), a.y);
            sm.r2[tid] = make_float4(2.f * (xz + ry), 2.f * (yz - rx), 1.f - 2.f * (xx + yy), a.z);
            sm.sc[tid] = make_float4(s.x, s.y, s.z, a.w);
            sm.is[tid] = make_float4(1.0f / s.x, 1.0f / s.y, 1.0f / s.z, 0.f);
            sm.qt[tid] = q;
            sm.cl[tid] = make_float4(fmaxf(rgb[idx * 3 + 0], 0.f), fmaxf(rgb[idx * 3 + 1], 0.f), fmaxf(rgb[idx * 3 + 2], 0.f),
                                     __uint_as_float(idx));
            if (UNIFORM) {
                const float vx = o0x - a.x, vy = o0y - a.y, vz = o0z - a.z;
                const float4 t0 = sm.r0[tid], t1 = sm.r1[tid], t2 = sm.r2[tid];
                sm.go[tid] = make_float4((t0.x * vx + t0.y * vy + t0.z * vz) / s.x, (t1.x * vx + t1.y * vy + t1.z * vz) / s.y,
                                         (t2.x * vx + t2.y * vy + t2.z * vz) / s.z, 0.f);
            }
        }
        __syncthreads();
        const int count = min(kBatch, static_cast<int>(end - base));
        // chunks of 32 entries: lane k screens entry k against the warp's pixel block (UNIFORM tiles), the warp walks the survivors
        for (int c = 0; c < count; c += 32) {
            if (__all_sync(kFull, !alive)) break;
            const int e = c + lane;
            bool cand = e < count;
            if (UNIFORM && wf.on && cand) {
                const float4 r0 = sm.r0[e], r1 = sm.r1[e], r2 = sm.r2[e], is = sm.is[e], g0 = sm.go[e];
                cand = block_candidate<DEG>(cfg, wf, is.x * r0.x, is.x * r0.y, is.x * r0.z, is.y * r1.x, is.y * r1.y, is.y * r1.z, is.z * r2.x,
                                            is.z * r2.y, is.z * r2.z, g0.x, g0.y, g0.z, sm.sc[e].w);
            }
            unsigned todo = __ballot_sync(kFull, cand);
            while (todo) {
            const int j = c + __ffs(todo) - 1;
            todo &= todo - 1;
            float g[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) g[i] = 0.f;
            bool hit = false;
            if (alive) {
                const float4 r0 = sm.r0[j], r1 = sm.r1[j], r2 = sm.r2[j], sc = sm.sc[j], is = sm.is[j];
                // canonical ray (processHitBwd, gaussianParticles.cuh:520-532)
                const float pcx = ray.ox - r0.w, pcy = ray.oy - r1.w, pcz = ray.oz - r2.w;              // gposc
                float gox, goy, goz;                                                                      // gro
                if (UNIFORM) {
                    const float4 g0 = sm.go[j];
                    gox = g0.x; goy = g0.y; goz = g0.z;
                } else {
                    gox = is.x * (r0.x * pcx + r0.y * pcy + r0.z * pcz);
                    goy = is.y * (r1.x * pcx + r1.y * pcy + r1.z * pcz);
                    goz = is.z * (r2.x * pcx + r2.y * pcy + r2.z * pcz);
                }
                const float drx = r0.x * ray.dx + r0.y * ray.dy + r0.z * ray.dz;                          // rayDirR
                const float dry = r1.x * ray.dx + r1.y * ray.dy + r1.z * ray.dz;
                const float drz = r2.x * ray.dx + r2.y * ray.dy + r2.z * ray.dz;
                const float ux = is.x * drx, uy = is.y * dry, uz = is.z * drz;                            // grdu
                const float l = ux * ux + uy * uy + uz * uz;
                const float il = l > 0.f ? rsqrtf(l) : 1.f;
                const float gdx = ux * il, gdy = uy * il, gdz = uz * il;                                  // grd
                const float ccx = gdy * goz - gdz * goy, ccy = gdz * gox - gdx * goz, ccz = gdx * goy - gdy * gox;  // gcrod
                const float gray = ccx * ccx + ccy * ccy + ccz * ccz;
                const float gres = kernel_response<DEG>(gray);
                const float dns = sc.w;
                const float alpha = fminf(cfg.max_alpha, gres * dns);
                if ((gres > cfg.min_kernel_density) && (alpha > cfg.min_alpha)) {
                    hit = true;
                    const float4 cl = sm.cl[j];
                    const float pd = -(gdx * gox + gdy * goy + gdz * goz);
                    const float ddx = gdx * pd, ddy = gdy * pd, ddz = gdz * pd;                            // grdd
                    const float hx = sc.x * ddx, hy = sc.y * ddy, hz = sc.z * ddz;                         // grds
                    const float gsq = hx * hx + hy * hy + hz * hz;
                    const float gdist = sqrtf(gsq);
                    const float weight = alpha * T;
                    const float nextT = (1.f - alpha) * T;
                    const bool last = nextT <= cfg.min_transmittance;
                    const float inv_next = last ? 0.f : 1.0f / nextT;

                    // depth branch (:545-580)
                    D += weight * gdist;
                    const float resD = fmaxf((Dint - D) * inv_next, 0.f);
                    const float a_hit = (gdist - resD) * T * Dgrad;
                    const float hs = gsq > 0.f ? (weight / gdist) * Dgrad : 0.f;
                    const float hgx = hx * hs, hgy = hy * hs, hgz = hz * hs;                               // grdsRayHitGrd
                    const float sd = hgx * sc.x * gdx + hgy * sc.y * gdy + hgz * sc.z * gdz;               // grdScaledDot
                    // opacity branch (:586-587)
                    const float resT = alpha < 0.999999f ? Tint / (1.f - alpha) : T;
                    const float a_dns = resT * -Tgrad;
                    // radiance branch (:602-612)
                    g[12] = Cgx * weight; g[13] = Cgy * weight; g[14] = Cgz * weight;
                    Cx += weight * cl.x; Cy += weight * cl.y; Cz += weight * cl.z;
                    const float rcx = fmaxf((Cix - Cx) * inv_next, 0.f);
                    const float rcy = fmaxf((Ciy - Cy) * inv_next, 0.f);
                    const float rcz = fmaxf((Ciz - Cz) * inv_next, 0.f);
                    const float common = a_hit + a_dns + T * (cl.x - rcx) * Cgx + T * (cl.y - rcy) * Cgy + T * (cl.z - rcz) * Cgz;
                    g[3] = gres * common;                                                                 // d density (:624-627)
                    const float gray_g = kernel_response_grad<DEG>(gray, gres, dns * common);             // (:639-648)
                    // gray = |grd x gro|^2  (:684-702)
                    const float kx = 2.f * ccx * gray_g, ky = 2.f * ccy * gray_g, kz = 2.f * ccz * gray_g;  // gcrodGrd
                    const float gd_gx = kz * goy - ky * goz + (sc.x * hgx * pd - gox * sd);                 // grdGrd + grdRayHitGrd
                    const float gd_gy = kx * goz - kz * gox + (sc.y * hgy * pd - goy * sd);
                    const float gd_gz = ky * gox - kx * goy + (sc.z * hgz * pd - goz * sd);
                    const float go_gx = ky * gdz - kz * gdy - gdx * sd;                                     // groGrd + groRayHitGrd
                    const float go_gy = kz * gdx - kx * gdz - gdy * sd;
                    const float go_gz = kx * gdy - ky * gdx - gdz * sd;
                    // gro = (1/s) gposcr  (:705-713)
                    const float prg_x = is.x * go_gx, prg_y = is.y * go_gy, prg_z = is.z * go_gz;          // gposcrGrd
                    float sgx = ddx * hgx - gox * prg_x;               // gsclRayHitGrd + gsclGrdGro (gposcr/s^2 = gro/s)
                    float sgy = ddy * hgy - goy * prg_y;
                    float sgz = ddz * hgz - goz * prg_z;
                    // gposcr = R^T gposc  (:715-726)
                    g[0] = -(prg_x * r0.x + prg_y * r1.x + prg_z * r2.x);
                    g[1] = -(prg_x * r0.y + prg_y * r1.y + prg_z * r2.y);
                    g[2] = -(prg_x * r0.z + prg_y * r1.z + prg_z * r2.z);
                    // grd = normalize(grdu)  (:729-731, safe_normalize_bw mathUtils.cuh:410-420)
                    const float il3 = il * il * il;
                    const float du = gd_gx * ux + gd_gy * uy + gd_gz * uz;
                    const float ug_x = l > 0.f ? il * gd_gx - il3 * ux * du : 0.f;                          // grduGrd
                    const float ug_y = l > 0.f ? il * gd_gy - il3 * uy * du : 0.f;
                    const float ug_z = l > 0.f ? il * gd_gz - il3 * uz * du : 0.f;
                    // grdu = (1/s) rayDirR  (:733-738)
                    const float rdg_x = is.x * ug_x, rdg_y = is.y * ug_y, rdg_z = is.z * ug_z;             // rayDirRGrd
                    sgx -= ux * rdg_x;                                  // rayDirR/s^2 = grdu/s
                    sgy -= uy * rdg_y;
                    sgz -= uz * rdg_z;
                    g[8] = sgx; g[9] = sgy; g[10] = sgz;
                    // rotation rows m_i receive dM_i = prg_i * gposc + rdg_i * d   (matmul_bw_quat twice, :719-747)
                    const float m00 = prg_x * pcx + rdg_x * ray.dx, m01 = prg_x * pcy + rdg_x * ray.dy, m02 = prg_x * pcz + rdg_x * ray.dz;
                    const float m10 = prg_y * pcx + rdg_y * ray.dx, m11 = prg_y * pcy + rdg_y * ray.dy, m12 = prg_y * pcz + rdg_y * ray.dz;
                    const float m20 = prg_z * pcx + rdg_z * ray.dx, m21 = prg_z * pcy + rdg_z * ray.dy, m22 = prg_z * pcz + rdg_z * ray.dz;
                    const float4 q = sm.qt[j];
                    const float qr = q.x, qx = q.y, qy = q.z, qz = q.w;
                    g[4] = 2.f * (qz * (m01 - m10) + qy * (m20 - m02) + qx * (m12 - m21));
                    g[5] = 2.f * (qy * (m01 + m10) + qz * (m02 + m20) + qr * (m12 - m21)) - 4.f * qx * (m11 + m22);
                    g[6] = 2.f * (qx * (m01 + m10) + qr * (m20 - m02) + qz * (m12 + m21)) - 4.f * qy * (m00 + m22);
                    g[7] = 2.f * (qr * (m01 - m10) + qx * (m02 + m20) + qy * (m12 + m21)) - 4.f * qz * (m00 + m11);
                    T = nextT;
                    if (T < cfg.min_transmittance) alive = false;
                }
            }
            if (__any_sync(kFull, hit)) {
                const float total = warp_transpose_reduce16(g, lane);
                if ((lane & 1) == 0) {
                    const uint32_t idx = __float_as_uint(sm.cl[j].w);
                    atomicAdd(grad_acc + static_cast<size_t>(idx) * kGradRow + (lane >> 1), total);
                }
                if (__all_sync(kFull, !alive)) break;
            }
            }
        }
    }
}

template <int DEG>
__global__ void __launch_bounds__(kTilePixels, 3) render_backward_kernel(FrameCamera cam, FrameConfig cfg,
                                                                      const float* __restrict__ rays_o,
                                                                      const float* __restrict__ rays_d,
                                                                      const float* __restrict__ particles,
                                                                      const float* __restrict__ rgb,
                                                                      const uint32_t* __restrict__ sorted_values,
                                                                      const uint32_t* __restrict__ ranges,
                                                                      const uint32_t* __restrict__ tile_order,
                                                                      const float* __restrict__ out_rgba, const float* __restrict__ d_rgba,
                                                                      const float* __restrict__ out_dist, const float* __restrict__ d_dist,
                                                                      float* __restrict__ grad_acc) {
    __shared__ BwdSmem sm;
    const int tile = tile_order[blockIdx.x];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    int px, py;
    tile_pixel(tile, cam.grid_x, tid, px, py);
    const bool inside = (px < cam.width) && (py < cam.height);
    const int64_t pix = static_cast<int64_t>(py) * cam.width + px;

    Ray ray;
    ray.alive = false;
    if (inside) ray = make_ray(cam, rays_o, rays_d, pix);
    bool alive = inside && ray.alive;

    // initializeBackwardRay (kernels/cuda/common/rayPayloadBackward.cuh:31-73)
    float Cix = 0.f, Ciy = 0.f, Ciz = 0.f, Cgx = 0.f, Cgy = 0.f, Cgz = 0.f, Tint = 1.f, Tgrad = 0.f, Dint = 0.f, Dgrad = 0.f;
    if (alive) {
        const float4 o = reinterpret_cast<const float4*>(out_rgba)[pix];
        const float4 g = reinterpret_cast<const float4*>(d_rgba)[pix];
        Cix = o.x; Ciy = o.y; Ciz = o.z;
        Cgx = g.x; Cgy = g.y; Cgz = g.z;
        Tint = 1.f - o.w;
        Tgrad = -1.f * g.w;
        Dint = out_dist[pix];
        Dgrad = d_dist[pix];
    }

    float o0x, o0y, o0z;
    const bool uniform = tile_common_origin(cam, rays_o, tile, inside, pix, o0x, o0y, o0z);
    // the warp-uniform frame lives in shared memory: 15 fewer live registers in the adjoint loop
    __shared__ WarpFrame wfs[kTilePixels / 32];
    WarpFrame& wf = wfs[tid >> 5];
    {
        const WarpFrame tmp = make_warp_frame(cam, ray, alive, uniform && (cfg.subtile_culling & 1), lane);
        if (lane == 0) wf = tmp;
        __syncwarp();
    }
    const uint32_t begin = ranges[tile * 2], end = ranges[tile * 2 + 1];
    if (uniform)
        backward_tile<DEG, true>(cfg, sm, wf, ray, o0x, o0y, o0z, tid, lane, begin, end, particles, rgb, sorted_values, alive, Cix, Ciy, Ciz,
                                 Cgx, Cgy, Cgz, Tint, Tgrad, Dint, Dgrad, grad_acc);
    else
        backward_tile<DEG, false>(cfg, sm, wf, ray, o0x, o0y, o0z, tid, lane, begin, end, particles, rgb, sorted_values, alive, Cix, Ciy, Ciz,
                                  Cgx, Cgy, Cgz, Tint, Tgrad, Dint, Dgrad, grad_acc);
}

// ----------------------------------------------------------------------------------------------------------
// G8 per-particle SH adjoint + emission of the final gradient rows; re-zeroes the accumulator for the next frame.

constexpr float kC0 = 0.28209479177387814f;
constexpr float kC1 = 0.4886025119029199f;
__constant__ float kC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                             0.5462742152960396f};
__constant__ float kC3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                             -0.4570457994644658f, 1.445305721320277f,  -0.5900435899266435f};

// Block of 128 particles, all global traffic through TMA bulk copies (cp.async.bulk + mbarrier): one copy brings the block's
// accumulator rows, one 192-byte copy per VISIBLE particle brings its SH coefficients, and three bulk stores write the
// d_sph rows, the d_particles rows and the re-zeroed accumulator rows (every output row is written, zeros for invisible
// particles, so the caller needs no memset).  Threads only touch shared memory, with 128-bit accesses.
constexpr int kPbThreads = 128;

// COMPACT (view-parallel training): instead of the [N,48] SH gradient row the kernel emits the masked radiance gradient (3 floats) the
// row is the outer product of -- d_sph[j][c] = basis_j(direction) * g[c] -- so ranks exchange 16 instead of 192 bytes per particle and
// rebuild the summed rows with sph_from_views_kernel.
template <bool COMPACT>
__global__ void __launch_bounds__(kPbThreads) project_backward_kernel(FrameCamera cam, int64_t n, const float* __restrict__ particles,
                                                                      const float* __restrict__ sph, int deg, const float* __restrict__ rgb,
                                                                      const uint32_t* __restrict__ tiles_count, float* __restrict__ grad_acc,
                                                                      float* __restrict__ d_particles, float* __restrict__ d_sph) {
    __shared__ __align__(128) float4 s_acc[kPbThreads * 4];   // in: accumulator rows, out: zeros
    __shared__ __align__(128) float4 s_sh[kPbThreads * 12];   // in: SH coefficients, out: d_sph rows
    __shared__ __align__(128) float4 s_dp[kPbThreads * 3];    // out: d_particles rows
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x;
    const int64_t base = static_cast<int64_t>(blockIdx.x) * kPbThreads;
    const int cnt = static_cast<int>(min(static_cast<int64_t>(kPbThreads), n - base));
    if (tid == 0) {
        mbar_init(&s_bar, kPbThreads);
        fence_proxy_async();
    }
    __syncthreads();
    const int64_t i = base + tid;
    const bool in_range = tid < cnt;
    const bool vis = in_range && (tiles_count[i] != 0u);
    const bool want_sh = vis && (deg > 0);
    mbar_expect_tx(&s_bar, (tid == 0 ? static_cast<uint32_t>(cnt) * 64u : 0u) + (want_sh ? 192u : 0u));
    if (tid == 0) tma_bulk_g2s(s_acc, grad_acc + base * kGradRow, static_cast<uint32_t>(cnt) * 64u, &s_bar);
    if (want_sh) tma_bulk_g2s(s_sh + tid * 12, sph + i * 48, 192u, &s_bar);
    // overlap the remaining scalar loads with the bulk copies
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
    if (vis) {
        p = __ldg(reinterpret_cast<const float4*>(particles + i * 12));
        c0 = rgb[i * 3 + 0]; c1 = rgb[i * 3 + 1]; c2 = rgb[i * 3 + 2];
    }
    mbar_wait(&s_bar, 0);

    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in_range) {
        const float4 a0 = s_acc[tid * 4 + 0], a1 = s_acc[tid * 4 + 1], a2 = s_acc[tid * 4 + 2], a3 = s_acc[tid * 4 + 3];
        s_acc[tid * 4 + 0] = zero; s_acc[tid * 4 + 1] = zero; s_acc[tid * 4 + 2] = zero; s_acc[tid * 4 + 3] = zero;
        float dpx = a0.x, dpy = a0.y, dpz = a0.z;
        float4* row = s_sh + tid * 12;
        if (!vis) {
#pragma unroll
            for (int k = 0; k < 12; ++k) row[k] = zero;
        } else {
            // incident direction = normalize(position - sensor position) (gutProjector.cuh:418)
            const float vx = p.x - cam.cam_pos[0], vy = p.y - cam.cam_pos[1], vz = p.z - cam.cam_pos[2];
            const float len = sqrtf(vx * vx + vy * vy + vz * vz);
            const float inv_len = len > 0.f ? 1.0f / len : 0.f;
            const float x = len > 0.f ? vx * inv_len : 1.f, y = vy * inv_len, z = vz * inv_len;
            // clamp mask of max(f + 0.5, 0) (sphericalHarmonics.slang:63); rgb holds the unclamped f + 0.5
            const float mgr = c0 > 0.f ? a3.x : 0.f, mgg = c1 > 0.f ? a3.y : 0.f, mgb = c2 > 0.f ? a3.z : 0.f;
            float bs[16];
            sh_basis16(deg, x, y, z, bs);
            if (deg > 0 && len > 0.f) {
                // s[j] = sum_c coeff[j][c] * masked_grad[c]; then d(rgb)/d(direction) . grad, then through normalize
                // (gaussianParticles.slang:545-558)
                float cf[48];
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                    const float4 v = row[k];
                    cf[k * 4] = v.x; cf[k * 4 + 1] = v.y; cf[k * 4 + 2] = v.z; cf[k * 4 + 3] = v.w;
                }
                float sc[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) sc[k] = cf[k * 3] * mgr + cf[k * 3 + 1] * mgg + cf[k * 3 + 2] * mgb;
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                float gx = -kC1 * sc[3], gy = -kC1 * sc[1], gz = kC1 * sc[2];
                if (deg > 1) {
                    gx += kC2[0] * y * sc[4] + kC2[2] * (-2.f * x) * sc[6] + kC2[3] * z * sc[7] + kC2[4] * (2.f * x) * sc[8];
                    gy += kC2[0] * x * sc[4] + kC2[1] * z * sc[5] + kC2[2] * (-2.f * y) * sc[6] + kC2[4] * (-2.f * y) * sc[8];
                    gz += kC2[1] * y * sc[5] + kC2[2] * (4.f * z) * sc[6] + kC2[3] * x * sc[7];
                    if (deg > 2) {
                        gx += kC3[0] * (6.f * xy) * sc[9] + kC3[1] * yz * sc[10] + kC3[2] * (-2.f * xy) * sc[11] + kC3[3] * (-6.f * xz) * sc[12] +
                              kC3[4] * (4.f * zz - 3.f * xx - yy) * sc[13] + kC3[5] * (2.f * xz) * sc[14] + kC3[6] * (3.f * xx - 3.f * yy) * sc[15];
                        gy += kC3[0] * (3.f * xx - 3.f * yy) * sc[9] + kC3[1] * xz * sc[10] + kC3[2] * (4.f * zz - xx - 3.f * yy) * sc[11] +
                              kC3[3] * (-6.f * yz) * sc[12] + kC3[4] * (-2.f * xy) * sc[13] + kC3[5] * (-2.f * yz) * sc[14] +
                              kC3[6] * (-6.f * xy) * sc[15];
                        gz += kC3[1] * xy * sc[10] + kC3[2] * (8.f * yz) * sc[11] + kC3[3] * (6.f * zz - 3.f * xx - 3.f * yy) * sc[12] +
                              kC3[4] * (8.f * xz) * sc[13] + kC3[5] * (xx - yy) * sc[14];
                    }
                }
                const float dd = x * gx + y * gy + z * gz;
                dpx += (gx - x * dd) * inv_len;
                dpy += (gy - y * dd) * inv_len;
                dpz += (gz - z * dd) * inv_len;
            }
            if (COMPACT) {
                row[0] = make_float4(mgr, mgg, mgb, 0.f);
            } else {
                float o[48];  // d SH = basis x masked gradient
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    o[k * 3 + 0] = bs[k] * mgr;
                    o[k * 3 + 1] = bs[k] * mgg;
                    o[k * 3 + 2] = bs[k] * mgb;
                }
#pragma unroll
                for (int k = 0; k < 12; ++k) row[k] = make_float4(o[k * 4], o[k * 4 + 1], o[k * 4 + 2], o[k * 4 + 3]);
            }
        }
        s_dp[tid * 3 + 0] = make_float4(dpx, dpy, dpz, a0.w);
        s_dp[tid * 3 + 1] = a1;
        s_dp[tid * 3 + 2] = make_float4(a2.x, a2.y, a2.z, 0.f);
    }
    if (COMPACT) {  // pack the 16-byte radiance gradients of the block contiguously (front of s_sh) for one bulk store
        const float4 mine = in_range ? s_sh[tid * 12] : zero;
        __syncthreads();
        s_sh[tid] = mine;
    }
    fence_proxy_async();  // our shared-memory writes must be visible to the TMA engine
    __syncthreads();
    if (tid == 0) {
        if (COMPACT)
            tma_bulk_s2g(d_sph + base * 4, s_sh, static_cast<uint32_t>(cnt) * 16u);
        else
            tma_bulk_s2g(d_sph + base * 48, s_sh, static_cast<uint32_t>(cnt) * 192u);
        tma_bulk_s2g(d_particles + base * 12, s_dp, static_cast<uint32_t>(cnt) * 48u);
        tma_bulk_s2g(grad_acc + base * kGradRow, s_acc, static_cast<uint32_t>(cnt) * 64u);
        tma_commit_group();
        tma_wait_group_read0();  // shared memory must stay valid until the engine has read it
    }
}

// d_sph[p] = sum over views v of basis(direction of particle p seen from view v) x radiance gradient of view v (the rows the
// non-compact G8 of each view would have written, summed in view order -- the same order on every rank).
struct ViewPositions {
    float pos[64][3];
};

__global__ void __launch_bounds__(128) sph_from_views_kernel(int64_t n, const float* __restrict__ particles, int deg, int views, ViewPositions vp,
                                                              const float4* __restrict__ d_radiance_all, float4* __restrict__ d_sph) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = __ldg(reinterpret_cast<const float4*>(particles + i * 12));
    float o[48];
#pragma unroll
    for (int k = 0; k < 48; ++k) o[k] = 0.f;
    for (int v = 0; v < views; ++v) {
        const float4 g = __ldg(d_radiance_all + static_cast<int64_t>(v) * n + i);
        if (g.x == 0.f && g.y == 0.f && g.z == 0.f) continue;  // invisible in that view (or clamped): a zero row
        const float vx = p.x - vp.pos[v][0], vy = p.y - vp.pos[v][1], vz = p.z - vp.pos[v][2];
        const float len = sqrtf(vx * vx + vy * vy + vz * vz);
        const float inv_len = len > 0.f ? 1.0f / len : 0.f;
        const float x = len > 0.f ? vx * inv_len : 1.f, y = vy * inv_len, z = vz * inv_len;
        float bs[16];
        sh_basis16(deg, x, y, z, bs);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            o[k * 3 + 0] += bs[k] * g.x;
            o[k * 3 + 1] += bs[k] * g.y;
            o[k * 3 + 2] += bs[k] * g.z;
        }
    }
    float4* row = d_sph + i * 12;
#pragma unroll
    for (int k = 0; k < 12; ++k) row[k] = make_float4(o[k * 4], o[k * 4 + 1], o[k * 4 + 2], o[k * 4 + 3]);
}

}  // namespace

void launch_tile_order(cudaStream_t s, const FrameCamera& cam, const uint32_t* ranges, uint32_t* tile_order) {
    tile_order_kernel<<<1, 1024, 0, s>>>(cam.grid_x * cam.grid_y, ranges, tile_order);
}

void launch_render_forward(cudaStream_t s, const FrameCamera& cam, const FrameConfig& cfg, const float* rays_o,
                           const float* rays_d, const float* particles, const float* rgb, const uint32_t* sorted_values,
                           const uint32_t* ranges, const uint32_t* tile_order, float* out_rgba, float* out_dist, float* out_hits) {
    const unsigned grid = cam.grid_x * cam.grid_y;
    if (cfg.kernel_degree == 4)
        render_forward_kernel<4><<<grid, kTilePixels, 0, s>>>(cam, cfg, rays_o, rays_d, particles, rgb, sorted_values, ranges, tile_order, out_rgba, out_dist, out_hits);
    else
        render_forward_kernel<2><<<grid, kTilePixels, 0, s>>>(cam, cfg, rays_o, rays_d, particles, rgb, sorted_values, ranges, tile_order, out_rgba, out_dist, out_hits);
}

void launch_render_backward(cudaStream_t s, const FrameCamera& cam, const FrameConfig& cfg, const float* rays_o,
                            const float* rays_d, const float* particles, const float* rgb, const uint32_t* sorted_values,
                            const uint32_t* ranges, const uint32_t* tile_order, const float* out_rgba, const float* d_rgba,
                            const float* out_dist, const float* d_dist, float* grad_acc) {
    const unsigned grid = cam.grid_x * cam.grid_y;
    if (cfg.kernel_degree == 4)
        render_backward_kernel<4><<<grid, kTilePixels, 0, s>>>(cam, cfg, rays_o, rays_d, particles, rgb, sorted_values, ranges, tile_order, out_rgba, d_rgba, out_dist, d_dist, grad_acc);
    else
        render_backward_kernel<2><<<grid, kTilePixels, 0, s>>>(cam, cfg, rays_o, rays_d, particles, rgb, sorted_values, ranges, tile_order, out_rgba, d_rgba, out_dist, d_dist, grad_acc);
}

void launch_project_backward(cudaStream_t s, const FrameCamera& cam, int64_t n, const float* particles, const float* sph,
                             int sph_degree, const float* rgb, const uint32_t* tiles_count, float* grad_acc, float* d_particles,
                             float* d_sph, bool compact) {
    if (n <= 0) return;
    const unsigned blocks = static_cast<unsigned>((n + kPbThreads - 1) / kPbThreads);
    if (compact)
        project_backward_kernel<true><<<blocks, kPbThreads, 0, s>>>(cam, n, particles, sph, sph_degree, rgb, tiles_count, grad_acc, d_particles, d_sph);
    else
        project_backward_kernel<false><<<blocks, kPbThreads, 0, s>>>(cam, n, particles, sph, sph_degree, rgb, tiles_count, grad_acc, d_particles, d_sph);
}

void launch_sph_from_views(cudaStream_t s, int64_t n, const float* particles, int sph_degree, int views, const float* view_positions /*host [views,3]*/,
                           const float* d_radiance_all, float* d_sph) {
    if (n <= 0) return;
    ViewPositions vp;
    for (int v = 0; v < views; ++v)
        for (int k = 0; k < 3; ++k) vp.pos[v][k] = view_positions[v * 3 + k];
    const unsigned blocks = static_cast<unsigned>((n + 127) / 128);
    sph_from_views_kernel<<<blocks, 128, 0, s>>>(n, particles, sph_degree, views, vp, reinterpret_cast<const float4*>(d_radiance_all),
                                                 reinterpret_cast<float4*>(d_sph));
}

}  // namespace gutb200
