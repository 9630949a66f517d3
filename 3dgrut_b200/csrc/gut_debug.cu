// 3dgrut_b200/csrc/gut_debug.cu -- measurement helpers behind the debug entry points of include/gut_b200.h (never on the render path).
#include "gut_common.cuh"

namespace gutb200 {

namespace {

// FP32 FMA throughput micro-benchmark ("builder to measure with an FMA micro-benchmark"): 8 independent
// dependency chains per thread, 2 flops per FMA.  The denominator of bench.py's roofline_fp32 block.
__global__ void __launch_bounds__(256) fma_peak_kernel(int iters, float* __restrict__ sink) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float b = 0.999f, c = 1e-4f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    const float r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456f) sink[0] = r;  // never true: keeps the chains alive
}

}  // namespace

// flops of one launch = blocks * 256 threads * iters * 16 * 8 FMAs * 2
void launch_fma_peak(cudaStream_t s, int blocks, int iters, float* sink) { fma_peak_kernel<<<blocks, 256, 0, s>>>(iters, sink); }

}  // namespace gutb200
