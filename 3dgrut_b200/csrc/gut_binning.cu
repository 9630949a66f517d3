// 3dgrut_b200/csrc/gut_binning.cu -- G2 / G4 / G5 of the 3DGUT forward without a global sort (ours; no library calls).
//
// Reference (threedgut_tracer/src/gutRenderer.cu:303-373): inclusive scan of the per-particle tile counts (CUB), host read of the
// total, expand into 64-bit (tile << 32 | depth) keys, ONE stable 44-bit CUB radix sort of all I keys, tile ranges from the sorted
// stream.  The order it defines inside a tile is (depth bits, particle index) -- a stable sort of keys emitted in particle order.
//
// Here:
//   project        counts every particle's tiles AND bumps a per-tile histogram           (gut_project.cu, atomics on T counters)
//   tile_scan      one CTA: exclusive scan of the histogram -> tile ranges (= G5, no pass over the keys), hit-word slices,
//                  heaviest-first tile order, total I and the capacity check on the device
//   expand_place   every (particle, tile) pair is dropped into its tile's slice at an atomically claimed slot as the 64-bit key
//                  (depth bits << 32 | particle index)                                      (gut_project.cu)
//   tile_sort      one CTA per tile sorts its slice by that key -- a bitonic network in shared memory, all comparators ascending
//                  ("flip" + "disperse" steps), so the slots beyond the list length act as +infinity without being stored; lists
//                  longer than the shared-memory chunk run the outer steps of the same network through global memory
// The keys are unique inside a tile (a particle enters a tile once), so the result is exactly the reference's order whatever order
// the atomics claimed the slots in; only sorted artefacts are observable and they stay bit-identical (tests/test_gut_parity_gpu.py).
// Work: I x 8 B written once, sorted in place on chip; no N-sized depth sort, no scan over N, no multi-pass radix sort over I.
#include "gut_common.cuh"

namespace gutb200 {

namespace {

constexpr unsigned kFullMask = 0xFFFFFFFFu;

// ----------------------------------------------------------------------------------------------------------
// tile_scan: single CTA.  Every tile owns kTileSubs sub-counters (a particle bumps sub-counter `particle & (kTileSubs - 1)`): the atomics
// of a hot tile -- thousands of increments of one address serialise in the L2 -- spread over kTileSubs addresses, and the slots of a
// tile's slice are claimed per sub-bucket the same way.  counts[T][kTileSubs] -> ranges[T][2] ((0, 0) for an empty tile, as the
// reference's zero-filled range buffer reads), sub_base[T][kTileSubs] (first slot of each sub-bucket), chunk_base[T], order[T]
// (decreasing list length, bucketed by log2), fill[T][kTileSubs] = 0, totals[0] = I, totals[1] = 1 if I exceeds the capacity of the key
// buffers (empty ranges are published then and the host re-launches the dependent kernels after growing the buffers).
__global__ void __launch_bounds__(1024) tile_scan_kernel(int num_tiles, const uint32_t* __restrict__ counts, uint32_t capacity,
                                                         uint32_t* __restrict__ ranges, uint32_t* __restrict__ sub_base,
                                                         uint32_t* __restrict__ chunk_base, uint32_t* __restrict__ order,
                                                         uint32_t* __restrict__ fill, uint32_t* __restrict__ totals) {
    __shared__ uint32_t hist[34];
    __shared__ uint32_t warp_a[32], warp_b[32];
    __shared__ uint32_t s_overflow;
    if (threadIdx.x < 34) hist[threadIdx.x] = 0;
    __syncthreads();
    // every thread owns a contiguous strip of tiles
    const int strip = (num_tiles + static_cast<int>(blockDim.x) - 1) / static_cast<int>(blockDim.x);
    const int t0 = min(static_cast<int>(threadIdx.x) * strip, num_tiles), t1 = min(t0 + strip, num_tiles);
    auto tile_total = [&](int t) {
        const uint4* c4 = reinterpret_cast<const uint4*>(counts + static_cast<size_t>(t) * kTileSubs);
        uint32_t c = 0;
#pragma unroll
        for (int k = 0; k < kTileSubs / 4; ++k) {
            const uint4 v = c4[k];
            c += v.x + v.y + v.z + v.w;
        }
        return c;
    };
    uint32_t sum_n = 0, sum_c = 0;
    for (int t = t0; t < t1; ++t) {
        const uint32_t c = tile_total(t);
        sum_n += c;
        sum_c += (c + 31u) >> 5;
        atomicAdd(&hist[__clz(c) + 1], 1u);  // __clz(0) = 32 -> last bucket; long lists -> small bucket index
    }
    uint32_t inc_n = sum_n, inc_c = sum_c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(kFullMask, inc_n, o), b = __shfl_up_sync(kFullMask, inc_c, o);
        if (lane >= o) {
            inc_n += a;
            inc_c += b;
        }
    }
    if (lane == 31) {
        warp_a[warp] = inc_n;
        warp_b[warp] = inc_c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int b = 1; b < 34; ++b) {
            const uint32_t h = hist[b];
            hist[b] = run;
            run += h;
        }
        uint32_t ra = 0, rb = 0;
        for (int w = 0; w < 32; ++w) {
            const uint32_t a = warp_a[w], b = warp_b[w];
            warp_a[w] = ra;
            warp_b[w] = rb;
            ra += a;
            rb += b;
        }
        totals[0] = ra;
        totals[1] = ra > capacity ? 1u : 0u;
        s_overflow = ra > capacity ? 1u : 0u;
    }
    __syncthreads();
    const bool overflow = s_overflow != 0u;  // the lists do not fit the key buffer: publish empty ranges, the host grows and re-queues
    uint32_t run_n = warp_a[warp] + inc_n - sum_n, run_c = warp_b[warp] + inc_c - sum_c;
    for (int t = t0; t < t1; ++t) {
        uint32_t sub = run_n;
        for (int k = 0; k < kTileSubs; ++k) {
            const size_t at = static_cast<size_t>(t) * kTileSubs + k;
            const uint32_t c = counts[at];
            sub_base[at] = sub;
            fill[at] = overflow ? 0xC0000000u : 0u;   // a huge fill level makes every claim fall outside its sub-bucket
            sub += c;
        }
        const uint32_t c = sub - run_n;
        const bool empty = overflow || (c == 0u);
        ranges[t * 2] = empty ? 0u : run_n;
        ranges[t * 2 + 1] = empty ? 0u : run_n + c;
        chunk_base[t] = overflow ? 0u : run_c;
        run_n += c;
        run_c += (c + 31u) >> 5;
        order[atomicAdd(&hist[__clz(c) + 1], 1u)] = static_cast<uint32_t>(t);
    }
}

// ----------------------------------------------------------------------------------------------------------
// tile_sort

__device__ __forceinline__ void compare_exchange(unsigned long long* s, uint32_t lo, uint32_t hi) {
    const unsigned long long a = s[lo], b = s[hi];
    if (a > b) {
        s[lo] = b;
        s[hi] = a;
    }
}

// all steps of the ascending bitonic network with block sizes k = k_first .. k_last on s[0 .. len) (len <= chunk, local indices);
// `flip_first`: whether the first k starts with its flip step (false = only the disperse steps j < k_first / 2 ... of an outer merge)
// (block sizes and strides are powers of two: all index arithmetic is shifts and masks)
template <int THREADS>
__device__ __forceinline__ void bitonic_local(unsigned long long* s, uint32_t len, uint32_t span, uint32_t k_first, uint32_t k_last, bool inner_only) {
    for (uint32_t k = k_first; k <= k_last; k <<= 1) {
        const uint32_t lk = 31u - __clz(k);   // log2 k
        if (!inner_only) {
            const uint32_t half_mask = (k >> 1) - 1u;
            for (uint32_t i = threadIdx.x; i < span / 2; i += THREADS) {  // flip: i-th element of a block's lower half <-> its mirror
                const uint32_t blk = i >> (lk - 1u), off = i & half_mask;
                const uint32_t lo = (blk << lk) + off, hi = (blk << lk) + (k - 1u - off);
                if (hi < len) compare_exchange(s, lo, hi);
            }
            __syncthreads();
        }
        for (uint32_t j = inner_only ? (k >> 1) : (k >> 2); j >= 1; j >>= 1) {  // disperse
            const uint32_t lj = 31u - __clz(j);
            for (uint32_t i = threadIdx.x; i < span / 2; i += THREADS) {
                const uint32_t lo = ((i >> lj) << (lj + 1u)) + (i & (j - 1u)), hi = lo + j;
                if (hi < len) compare_exchange(s, lo, hi);
            }
            __syncthreads();
        }
    }
}

// One CTA per tile.  CHUNK = keys held in shared memory at a time (a power of two).  Tiles whose list length is outside
// (min_len, max_len] leave at once: two launches (small / large chunk) cover all tiles with the occupancy each class wants.
template <int CHUNK, int THREADS>
__global__ void __launch_bounds__(THREADS) tile_sort_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ ranges,
                                                            const uint32_t* __restrict__ totals, uint32_t min_len, uint32_t max_len,
                                                            unsigned long long* __restrict__ keys, uint32_t* __restrict__ sorted_values) {
    extern __shared__ __align__(16) unsigned long long s_keys[];
    if (totals[1] != 0u) return;  // capacity exceeded: the host grows the buffers and launches again
    const uint32_t tile = order[blockIdx.x];
    const uint32_t begin = ranges[tile * 2], n = ranges[tile * 2 + 1] - begin;
    if (n <= min_len || n > max_len) return;
    unsigned long long* g = keys + begin;
    uint32_t n2 = 1;
    while (n2 < n) n2 <<= 1;
    if (n2 <= CHUNK) {
        for (uint32_t i = threadIdx.x; i < n; i += THREADS) s_keys[i] = g[i];
        __syncthreads();
        bitonic_local<THREADS>(s_keys, n, n2, 2, n2, false);
        for (uint32_t i = threadIdx.x; i < n; i += THREADS) sorted_values[begin + i] = static_cast<uint32_t>(s_keys[i]);
        return;
    }
    // long list: sort CHUNK-sized pieces on chip, then merge with the outer steps of the same network through global memory
    for (uint32_t c0 = 0; c0 < n; c0 += CHUNK) {
        const uint32_t len = min(static_cast<uint32_t>(CHUNK), n - c0);
        for (uint32_t i = threadIdx.x; i < len; i += THREADS) s_keys[i] = g[c0 + i];
        __syncthreads();
        bitonic_local<THREADS>(s_keys, len, CHUNK, 2, CHUNK, false);
        for (uint32_t i = threadIdx.x; i < len; i += THREADS) g[c0 + i] = s_keys[i];
        __syncthreads();
    }
    for (uint32_t k = 2u * CHUNK; k <= n2; k <<= 1) {
        const uint32_t lk = 31u - __clz(k), half_mask = (k >> 1) - 1u;
        for (uint32_t i = threadIdx.x; i < n2 / 2; i += THREADS) {  // flip over global memory
            const uint32_t blk = i >> (lk - 1u), off = i & half_mask;
            const uint32_t lo = (blk << lk) + off, hi = (blk << lk) + (k - 1u - off);
            if (hi < n) compare_exchange(g, lo, hi);
        }
        __syncthreads();
        for (uint32_t j = k >> 2; j >= CHUNK; j >>= 1) {  // disperse steps wider than a chunk
            const uint32_t lj = 31u - __clz(j);
            for (uint32_t i = threadIdx.x; i < n2 / 2; i += THREADS) {
                const uint32_t lo = ((i >> lj) << (lj + 1u)) + (i & (j - 1u)), hi = lo + j;
                if (hi < n) compare_exchange(g, lo, hi);
            }
            __syncthreads();
        }
        for (uint32_t c0 = 0; c0 < n; c0 += CHUNK) {  // the remaining steps (j < CHUNK) stay inside a chunk
            const uint32_t len = min(static_cast<uint32_t>(CHUNK), n - c0);
            for (uint32_t i = threadIdx.x; i < len; i += THREADS) s_keys[i] = g[c0 + i];
            __syncthreads();
            bitonic_local<THREADS>(s_keys, len, CHUNK, CHUNK, CHUNK, true);
            for (uint32_t i = threadIdx.x; i < len; i += THREADS) g[c0 + i] = s_keys[i];
            __syncthreads();
        }
    }
    for (uint32_t i = threadIdx.x; i < n; i += THREADS) sorted_values[begin + i] = static_cast<uint32_t>(g[i]);
}

// test-only: the reference's sorted 64-bit keys (tile << 32 | depth bits), rebuilt per tile from the sorted values
__global__ void __launch_bounds__(256) synth_tile_keys_kernel(const uint32_t* __restrict__ ranges, const uint32_t* __restrict__ vals,
                                                              const float* __restrict__ depth, uint64_t* __restrict__ out) {
    const uint32_t tile = blockIdx.x;
    const uint32_t begin = ranges[tile * 2], end = ranges[tile * 2 + 1];
    for (uint32_t k = begin + threadIdx.x; k < end; k += blockDim.x)
        out[k] = (static_cast<uint64_t>(tile) << 32) | __float_as_uint(depth[vals[k]]);
}

constexpr int kSmallChunk = 2048, kLargeChunk = 8192, kSortThreads = 256;

}  // namespace

void launch_tile_scan(cudaStream_t s, int num_tiles, const uint32_t* counts, uint32_t capacity, uint32_t* ranges, uint32_t* sub_base,
                      uint32_t* chunk_base, uint32_t* order, uint32_t* fill, uint32_t* totals) {
    tile_scan_kernel<<<1, 1024, 0, s>>>(num_tiles, counts, capacity, ranges, sub_base, chunk_base, order, fill, totals);
}

cudaError_t launch_tile_sort(cudaStream_t s, int num_tiles, const uint32_t* order, const uint32_t* ranges, const uint32_t* totals,
                             unsigned long long* keys, uint32_t* sorted_values) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tile_sort_kernel<kLargeChunk, kSortThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kLargeChunk * 8);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (num_tiles <= 0) return cudaSuccess;
    // heaviest tiles come first in `order`: the long lists start before the bulk of the short ones
    tile_sort_kernel<kLargeChunk, kSortThreads><<<num_tiles, kSortThreads, kLargeChunk * 8, s>>>(order, ranges, totals, kSmallChunk, 0xFFFFFFFFu, keys, sorted_values);
    tile_sort_kernel<kSmallChunk, kSortThreads><<<num_tiles, kSortThreads, kSmallChunk * 8, s>>>(order, ranges, totals, 0u, kSmallChunk, keys, sorted_values);
    return cudaGetLastError();
}

void launch_synth_tile_keys(cudaStream_t s, int num_tiles, const uint32_t* ranges, const uint32_t* vals, const float* depth, uint64_t* out) {
    if (num_tiles <= 0) return;
    synth_tile_keys_kernel<<<num_tiles, 256, 0, s>>>(ranges, vals, depth, out);
}

}  // namespace gutb200
