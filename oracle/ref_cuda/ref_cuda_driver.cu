// oracle/ref_cuda/ref_cuda_driver.cu -- TEST / BASELINE INFRASTRUCTURE ONLY (never linked into the product).
//
// extern "C" driver (OUR code) around the REFERENCE's own 3DGUT renderer, compiled for sm_100a by nvcc from the sources
// where they lie under /root/reference (oracle/Makefile target `refcuda`; nothing is copied into this repo):
//     threedgut_tracer/src/gutRenderer.cu   -- GUTRenderer::renderForward / renderBackward: projectOnTiles, CUB inclusive scan,
//                                              host read of the intersection count, expandTileProjections, the 44-bit CUB
//                                              SortPairs<uint64,uint32>, computeSortedTileRangeIndices, render, renderBackward,
//                                              projectBackward, grow-only scratch (included below as a translation unit)
//     threedgut_tracer/src/cudaBuffer.cpp   -- its buffer class (compiled separately)
// Only the slangc output `threedgutSlang.cuh` is replaced by our hand translation (oracle/ref_cuda/threedgutSlang.cuh).
// This driver plays the role of src/splatRaster.cpp:184-382 (the torch/pybind layer): it fills RenderParameters and
// GUTRenderer::Parameters from raw device pointers, zero-fills the outputs the way `torch::zeros` / `ones*1e6` do there
// (:213-217,:297-299), and hooks the reference's own device-launch callback (utils/logger.h:58-62) to time its stages with
// CUDA events.  Uses: (1) parity pin of oracle/ and of the product against the reference's kernels on a GPU,
// (2) the same-box GPU denominator bench.py prints next to our numbers.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#define private public  // read-only access to GUTRenderer::m_forwardContext for the parity tests (debug copies below)
#include "/root/reference/threedgut_tracer/src/gutRenderer.cu"
#undef private

namespace {

struct StageEvents {
    cudaEvent_t a = nullptr, b = nullptr;
    bool pending = false;
    double ms = 0;
    int calls = 0;
};

struct RefCtx {
    std::map<std::string, StageEvents> stages;
    bool timing = false;
    threedgut::Logger* logger = nullptr;
    threedgut::GUTRenderer* renderer = nullptr;
    threedgut::GUTRenderer::Parameters params;
    std::string error;
    uint32_t last_isect = 0;
};

void log_cb(uint8_t level, const char* msg, void* data) {
    RefCtx* c = static_cast<RefCtx*>(data);
    if (level <= threedgut::LoggerParameters::Error) {
        c->error = msg;
        fprintf(stderr, "[ref_cuda] %s\n", msg);
    }
}

void launch_cb(bool start, const char* tag, int /*deviceIndex*/, uint64_t queue, void* data) {
    RefCtx* c = static_cast<RefCtx*>(data);
    if (!c->timing) return;
    StageEvents& s = c->stages[tag];
    if (!s.a) {
        cudaEventCreate(&s.a);
        cudaEventCreate(&s.b);
    }
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(queue);
    if (start) {
        cudaEventRecord(s.a, stream);
    } else {
        cudaEventRecord(s.b, stream);
        s.pending = true;
    }
}

void drain(RefCtx* c) {
    for (auto& kv : c->stages) {
        StageEvents& s = kv.second;
        float ms = 0.f;
        if (s.pending && cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
            s.ms += ms;
            s.calls++;
        }
        s.pending = false;
    }
}

__global__ void fill_kernel(float* p, float v, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

threedgut::RenderParameters make_params(uint32_t frame, int width, int height, const float* focal, const float* pp, const float* pose0,
                                        const float* pose1) {
    threedgut::RenderParameters rp;
    rp.id               = frame;
    rp.resolution       = tcnn::ivec2{width, height};
    rp.hitTransmittance = 0.f;
    threedgut::TSensorModel m;
    m.shutterType = threedgut::TSensorModel::GlobalShutter;
    m.modelType   = threedgut::TSensorModel::OpenCVPinholeModel;
    memset(&m.ocvPinholeParams, 0, sizeof(m.ocvPinholeParams));
    m.ocvPinholeParams.principalPoint = tcnn::vec2{pp[0], pp[1]};
    m.ocvPinholeParams.focalLength    = tcnn::vec2{focal[0], focal[1]};
    rp.sensorModel = m;
    threedgut::TSensorState st;
    st.startTimestamp = 0;
    st.endTimestamp   = 1;
    for (int i = 0; i < 7; ++i) {
        st.startPose[i] = pose0[i];
        st.endPose[i]   = pose1[i];
    }
    rp.sensorState = st;
    rp.objectAABB  = threedgut::BoundingBox{tcnn::vec3{-1e06f, -1e06f, -1e06f}, tcnn::vec3{1e06f, 1e06f, 1e06f}};  // splatRaster.cpp:240
    return rp;
}

}  // namespace

extern "C" {

void* refcuda_create(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return nullptr;
    RefCtx* c = new RefCtx();
    threedgut::LoggerParameters lp;
    lp.maximumLevel             = threedgut::LoggerParameters::Error;
    lp.callback                 = log_cb;
    lp.callbackData             = c;
    lp.deviceLauncCallback      = launch_cb;
    lp.deviceLaunchCallbackData = c;
    c->logger   = new threedgut::Logger(lp);
    c->renderer = new threedgut::GUTRenderer(nlohmann::json::object(), *c->logger);
    // splatRaster.cpp:163-176
    c->params.valuesBuffer.resize(sizeof(c->params.values), 0, *c->logger);
    c->params.parametersBuffer.resize(sizeof(c->params.parameters), 0, *c->logger);
    c->params.gradientsBuffer.resize(sizeof(c->params.gradients), 0, *c->logger);
    c->params.parameters.dptrValuesBuffer = c->params.valuesBuffer.data();
    c->params.m_dptrParametersBuffer      = (uint64_t*)c->params.parametersBuffer.data();
    c->params.m_dptrGradientsBuffer       = (uint64_t*)c->params.gradientsBuffer.data();
    return c;
}

void refcuda_destroy(void* h) {
    RefCtx* c = static_cast<RefCtx*>(h);
    if (!c) return;
    cudaDeviceSynchronize();
    delete c->renderer;
    delete c->logger;
    for (auto& kv : c->stages) {
        if (kv.second.a) cudaEventDestroy(kv.second.a);
        if (kv.second.b) cudaEventDestroy(kv.second.b);
    }
    delete c;
}

const char* refcuda_last_error(void* h) { return h ? static_cast<RefCtx*>(h)->error.c_str() : "null"; }

void refcuda_set_timing(void* h, int on) {
    RefCtx* c = static_cast<RefCtx*>(h);
    drain(c);
    c->timing = on != 0;
}

// SplatRaster::trace (splatRaster.cpp:184-262); all pointers are device pointers; visibility is the [N] int buffer the
// reference exposes as a float tensor
int refcuda_forward(void* h, void* stream, uint32_t frame, int sph_degree, int64_t n, const float* particles, const float* sph, int width,
                    int height, const float* focal, const float* pp, const float* pose0, const float* pose1, const float* rays_o,
                    const float* rays_d, float* out_rgba, float* out_dist, float* out_hits, float* visibility) {
    RefCtx* c = static_cast<RefCtx*>(h);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t px = static_cast<int64_t>(width) * height;
    // torch::zeros / ones * 1e6 of splatRaster.cpp:213-217
    cudaMemsetAsync(out_rgba, 0, px * 16, s);
    fill_kernel<<<static_cast<unsigned>((px + 255) / 256), 256, 0, s>>>(out_dist, 1e06f, px);
    cudaMemsetAsync(out_hits, 0, px * 4, s);
    cudaMemsetAsync(visibility, 0, n * 4, s);
    c->params.values.numParticles               = static_cast<uint32_t>(n);
    c->params.values.radianceSphDegree          = sph_degree;
    c->params.parameters.dptrDensityParameters  = const_cast<float*>(particles);
    c->params.parameters.dptrRadianceParameters = const_cast<float*>(sph);
    c->params.valuesBuffer.setFromHost(&c->params.values, sizeof(c->params.values), reinterpret_cast<uint64_t>(s), *c->logger);
    c->params.parametersBuffer.setFromHost(&c->params.parameters, sizeof(c->params.parameters), reinterpret_cast<uint64_t>(s), *c->logger);
    const threedgut::RenderParameters rp = make_params(frame, width, height, focal, pp, pose0, pose1);
    int dev = 0;
    cudaGetDevice(&dev);
    const threedgut::Status st =
        c->renderer->renderForward(rp, reinterpret_cast<const tcnn::vec3*>(rays_o), reinterpret_cast<const tcnn::vec3*>(rays_d), out_hits,
                                   out_dist, out_rgba, reinterpret_cast<int*>(visibility), c->params, dev, s);
    return (cudaGetLastError() == cudaSuccess && st == threedgut::ErrorCode::None) ? 0 : 1;
}

// SplatRaster::traceBwd (splatRaster.cpp:264-350)
int refcuda_backward(void* h, void* stream, uint32_t frame, int sph_degree, int64_t n, const float* particles, const float* sph, int width,
                     int height, const float* focal, const float* pp, const float* pose0, const float* pose1, const float* rays_o,
                     const float* rays_d, const float* out_rgba, const float* d_rgba, const float* out_dist, const float* d_dist,
                     float* d_particles, float* d_sph) {
    RefCtx* c = static_cast<RefCtx*>(h);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(d_particles, 0, n * 48, s);  // torch::zeros of splatRaster.cpp:297-298
    cudaMemsetAsync(d_sph, 0, n * 192, s);
    c->params.values.numParticles               = static_cast<uint32_t>(n);
    c->params.values.radianceSphDegree          = sph_degree;
    c->params.parameters.dptrDensityParameters  = const_cast<float*>(particles);
    c->params.parameters.dptrRadianceParameters = const_cast<float*>(sph);
    c->params.gradients.dptrDensityGradients    = d_particles;
    c->params.gradients.dptrRadianceGradients   = d_sph;
    c->params.valuesBuffer.setFromHost(&c->params.values, sizeof(c->params.values), reinterpret_cast<uint64_t>(s), *c->logger);
    c->params.parametersBuffer.setFromHost(&c->params.parameters, sizeof(c->params.parameters), reinterpret_cast<uint64_t>(s), *c->logger);
    c->params.gradientsBuffer.setFromHost(&c->params.gradients, sizeof(c->params.gradients), reinterpret_cast<uint64_t>(s), *c->logger);
    const threedgut::RenderParameters rp = make_params(frame, width, height, focal, pp, pose0, pose1);
    int dev = 0;
    cudaGetDevice(&dev);
    const threedgut::Status st = c->renderer->renderBackward(
        rp, reinterpret_cast<const tcnn::vec3*>(rays_o), reinterpret_cast<const tcnn::vec3*>(rays_d), out_dist, d_dist, out_rgba, d_rgba, nullptr,
        nullptr, c->params, dev, s);
    return (cudaGetLastError() == cudaSuccess && st == threedgut::ErrorCode::None) ? 0 : 1;
}

// mean ms per call of the reference's own launch scopes since the last call; order:
// render::project, render::prepare-expand, render::expand, render::sort, render::render, render (whole forward),
// render-backward::render, render-backward::project, render-backward (whole backward)
int refcuda_stage_times(void* h, float* ms9) {
    RefCtx* c = static_cast<RefCtx*>(h);
    drain(c);
    static const char* tags[9] = {"render::project", "render::prepare-expand", "render::expand", "render::sort", "render::render", "render",
                                  "render-backward::render", "render-backward::project", "render-backward"};
    for (int i = 0; i < 9; ++i) {
        auto it = c->stages.find(tags[i]);
        ms9[i]  = (it != c->stages.end() && it->second.calls) ? static_cast<float>(it->second.ms / it->second.calls) : 0.f;
        if (it != c->stages.end()) {
            it->second.ms    = 0;
            it->second.calls = 0;
        }
    }
    return 0;
}

// host-side camera maths of the reference for one pose (debug): view matrix columns [12], sensor position [3], and the RenderParameters
// fields as the kernels receive them [resolution 2, focal 2, principal 2, model type, shutter type]
void refcuda_debug_camera(int width, int height, const float* focal, const float* pp, const float* pose0, const float* pose1, float* out20) {
    const threedgut::RenderParameters rp = make_params(0, width, height, focal, pp, pose0, pose1);
    const threedgut::TSensorPose pose = threedgut::interpolatedSensorPose(rp.sensorState.startPose, rp.sensorState.endPose, 0.5f);
    const tcnn::mat4x3 view = threedgut::sensorPoseToMat(pose);
    const threedgut::TSensorPose inv = threedgut::sensorPoseInverse(pose);
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 3; ++r) out20[c * 3 + r] = view[c][r];
    for (int i = 0; i < 3; ++i) out20[12 + i] = inv[i];
    out20[15] = static_cast<float>(rp.resolution.x);
    out20[16] = static_cast<float>(rp.resolution.y);
    out20[17] = rp.sensorModel.ocvPinholeParams.focalLength.x;
    out20[18] = rp.sensorModel.ocvPinholeParams.principalPoint.y;
    out20[19] = static_cast<float>(static_cast<int>(rp.sensorModel.modelType) * 10 + static_cast<int>(rp.sensorModel.shutterType));
}

// debug copies of the forward context for the parity tests: 0 tiles count [N] u32, 1 sorted keys [I] u64, 2 sorted values [I] u32,
// 3 tile ranges [T,2] u32, 4 depth [N] f32, 5 precomputed features [N,3] f32.  Returns the byte size (dst may be null to query).
int64_t refcuda_debug_copy(void* h, int what, void* dst, int64_t n, int64_t tiles) {
    RefCtx* c = static_cast<RefCtx*>(h);
    auto* f   = c->renderer->m_forwardContext.get();
    if (!f) return -1;
    cudaDeviceSynchronize();
    uint32_t total = 0;
    if (n > 0) cudaMemcpy(&total, static_cast<const uint32_t*>(f->particlesTilesOffset.data()) + (n - 1), 4, cudaMemcpyDeviceToHost);
    const void* src = nullptr;
    int64_t bytes   = 0;
    switch (what) {
        case 0: src = f->particlesTilesCount.data(); bytes = n * 4; break;
        case 1: src = f->sortedTileDepthKeys.data(); bytes = static_cast<int64_t>(total) * 8; break;
        case 2: src = f->sortedTileParticleIdx.data(); bytes = static_cast<int64_t>(total) * 4; break;
        case 3: src = f->sortedTileRangeIndices.data(); bytes = tiles * 8; break;
        case 4: src = f->particlesGlobalDepth.data(); bytes = n * 4; break;
        case 5: src = f->particlesPrecomputedFeatures.data(); bytes = n * 12; break;
        default: return -1;
    }
    if (dst && bytes > 0 && cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return bytes;
}

}  // extern "C"
