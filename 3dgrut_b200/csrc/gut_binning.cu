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
//   tile_sort      one CTA per tile sorts its slice by that key: a stable LSD radix sort over the depth bits (8 bits a pass, passes
//                  with a single digit skipped, slice ping-ponged through the L2) + a fix-up of equal-depth runs by particle index
//                  (a first version used a bitonic network: 125 M compare-exchanges at C2 cost 0.13 ms, 2.0 ms at C3 -- replaced)
// The keys are unique inside a tile (a particle enters a tile once), so the result is exactly the reference's order whatever order
// the atomics claimed the slots in; only sorted artefacts are observable and they stay bit-identical (tests/test_gut_parity_gpu.py).
// Work: I x 8 B written once, sorted in place on chip; no N-sized depth sort, no scan over N, no multi-pass radix sort over I.
#include <cstdlib>

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
    static_assert(kTileSubs == 16, "a half-warp scans one tile's sub-counters");
    __shared__ uint32_t hist[34];
    __shared__ uint32_t warp_a[32], warp_b[32];
    __shared__ uint32_t s_overflow;
    if (threadIdx.x < 34) hist[threadIdx.x] = 0;
    __syncthreads();
    // pass 1: tile totals.  Every thread owns a contiguous strip of tiles; the 16 counters of a tile are four 16-byte loads
    const int strip = (num_tiles + static_cast<int>(blockDim.x) - 1) / static_cast<int>(blockDim.x);
    const int t0 = min(static_cast<int>(threadIdx.x) * strip, num_tiles), t1 = min(t0 + strip, num_tiles);
    uint32_t sum_n = 0, sum_c = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    // The length buckets are few and most tiles fall into two or three of them, so a warp's lanes that land in the same bucket go through
    // ONE shared-memory atomic (match.any + population count) instead of serialising on its address; the loops are warp-uniform for that.
    // (Measured at C2: 0.027 ms with or without -- the kernel is bound by the latency of its two dependent passes, not by these atomics.)
    for (int k = 0; k < strip; ++k) {
        const int t = t0 + k;
        const bool live = t < t1;
        uint32_t c = 0;
        if (live) {
            const uint4* c4 = reinterpret_cast<const uint4*>(counts + static_cast<size_t>(t) * kTileSubs);
            const uint4 v0 = c4[0], v1 = c4[1], v2 = c4[2], v3 = c4[3];
            c = (v0.x + v0.y + v0.z + v0.w) + (v1.x + v1.y + v1.z + v1.w) + (v2.x + v2.y + v2.z + v2.w) + (v3.x + v3.y + v3.z + v3.w);
            sum_n += c;
            sum_c += (c + 31u) >> 5;
        }
        const int bucket = live ? __clz(c) + 1 : 64 + lane;  // __clz(0) = 32 -> last bucket; long lists -> small bucket index
        const unsigned peers = __match_any_sync(kFullMask, bucket);
        if (live && (peers & lt_mask) == 0u) atomicAdd(&hist[bucket], static_cast<uint32_t>(__popc(peers)));
    }
    uint32_t inc_n = sum_n, inc_c = sum_c;
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
    // pass 2: per-tile outputs.  The thread that owns a tile has its 16 sub-counters at hand (four 16-byte loads, L1 hits after pass 1):
    // it lays out the sub-buckets itself and writes them with 16-byte stores (a separate coalesced pass over the counters was 2 us).
    uint32_t run_n = warp_a[warp] + inc_n - sum_n, run_c = warp_b[warp] + inc_c - sum_c;
    const uint32_t fill0 = overflow ? 0xC0000000u : 0u;   // a huge fill level makes every claim fall outside its sub-bucket
    const uint4 fill4 = make_uint4(fill0, fill0, fill0, fill0);
    for (int k = 0; k < strip; ++k) {
        const int t = t0 + k;
        const bool live = t < t1;
        uint32_t c = 0;
        if (live) {
            const uint4* c4 = reinterpret_cast<const uint4*>(counts + static_cast<size_t>(t) * kTileSubs);
            const uint4 v0 = c4[0], v1 = c4[1], v2 = c4[2], v3 = c4[3];
            uint32_t r = run_n;
            uint4 b0, b1, b2, b3;
            b0.x = r; r += v0.x; b0.y = r; r += v0.y; b0.z = r; r += v0.z; b0.w = r; r += v0.w;
            b1.x = r; r += v1.x; b1.y = r; r += v1.y; b1.z = r; r += v1.z; b1.w = r; r += v1.w;
            b2.x = r; r += v2.x; b2.y = r; r += v2.y; b2.z = r; r += v2.z; b2.w = r; r += v2.w;
            b3.x = r; r += v3.x; b3.y = r; r += v3.y; b3.z = r; r += v3.z; b3.w = r; r += v3.w;
            c = r - run_n;
            uint4* sb = reinterpret_cast<uint4*>(sub_base + static_cast<size_t>(t) * kTileSubs);
            sb[0] = b0; sb[1] = b1; sb[2] = b2; sb[3] = b3;
            uint4* fl = reinterpret_cast<uint4*>(fill + static_cast<size_t>(t) * kTileSubs);
            fl[0] = fill4; fl[1] = fill4; fl[2] = fill4; fl[3] = fill4;
            // empty tile (or nothing fits): (0, 0) like the reference's zero-filled range buffer
            reinterpret_cast<uint2*>(ranges)[t] = (overflow || c == 0u) ? make_uint2(0u, 0u) : make_uint2(run_n, r);
            chunk_base[t] = overflow ? 0u : run_c;
            run_n = r;
            run_c += (c + 31u) >> 5;
        }
        const int bucket = live ? __clz(c) + 1 : 64 + lane;
        const unsigned peers = __match_any_sync(kFullMask, bucket);
        const int leader = __ffs(peers) - 1;
        uint32_t slot = 0;
        if (live && lane == leader) slot = atomicAdd(&hist[bucket], static_cast<uint32_t>(__popc(peers)));
        slot = __shfl_sync(kFullMask, slot, leader);
        if (live) order[slot + __popc(peers & lt_mask)] = static_cast<uint32_t>(t);
    }
}

// ----------------------------------------------------------------------------------------------------------
// tile_sort: one CTA per tile sorts the tile's slice of 64-bit keys (depth bits << 32 | particle) -- a least-significant-digit radix
// sort over the 32 depth bits, 8 bits a pass, ping-ponging between the key buffer and its twin (both stay in the L2: a slice is a few
// tens of KB), shared memory holding only the digit counters.  Each of the CTA's warps owns a contiguous segment of the slice and
// keeps its elements in order (32 at a time: `match.any` groups equal digits, the group's first lane claims their slots), so a pass is
// stable; passes in which every key carries the same digit (typically the exponent byte) are skipped.  The depth order the passes
// produce is completed to (depth, particle) order by a fix-up of runs of EQUAL depth bits -- duplicates of a position, e.g. freshly
// cloned Gaussians; the slots were claimed in arbitrary order, so the sort cannot rely on stability for them.

constexpr int kSortThreads = 512, kSortWarps = kSortThreads / 32;   // 16 warps per tile: the long lists (6-30 k keys at C3) are latency-bound
constexpr int kSortUnroll = 4;                                      // keys in flight per lane in the count / scatter loops

template <bool BALLOT>
__global__ void __launch_bounds__(kSortThreads) tile_sort_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ ranges,
                                                                 const uint32_t* __restrict__ totals, unsigned long long* keys,
                                                                 unsigned long long* keys_alt, uint32_t* __restrict__ sorted_values) {
    __shared__ uint32_t s_cnt[kSortWarps][256];   // per-warp digit counts -> running slot of (warp, digit)
    __shared__ uint32_t s_tot[8];
    __shared__ int s_flag;
    if (totals[1] != 0u) return;  // capacity exceeded: the host grows the buffers and launches again
    const uint32_t tile = order[blockIdx.x];
    const uint32_t begin = ranges[tile * 2], n = ranges[tile * 2 + 1] - begin;
    if (n == 0u) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned long long* src = keys + begin;
    unsigned long long* dst = keys_alt + begin;
    // Short lists use fewer warps (>= 128 keys each): the per-pass fixed cost -- zeroing, summing and scanning one counter row per warp --
    // scales with the warps that take part, and most C2 tiles hold a few hundred keys.
    const int aw = static_cast<int>(min(static_cast<uint32_t>(kSortWarps), max(1u, (n + 127u) / 128u)));
    // warp w < aw owns elements [w * seg, min(n, (w + 1) * seg)), seg a multiple of 32
    const uint32_t seg = ((n + aw - 1) / aw + 31u) & ~31u;
    const uint32_t w0 = min(n, warp * seg), w1 = min(n, w0 + seg);
    for (int shift = 32; shift < 64; shift += 8) {
        for (int i = threadIdx.x; i < aw * 256; i += kSortThreads) (&s_cnt[0][0])[i] = 0u;
        if (threadIdx.x == 0) s_flag = 0;
        __syncthreads();
        for (uint32_t i0 = w0; i0 < w1; i0 += 32 * kSortUnroll) {  // digit counts of this warp's segment, kSortUnroll loads in flight
            unsigned long long k[kSortUnroll];
#pragma unroll
            for (int u = 0; u < kSortUnroll; ++u) {
                const uint32_t i = i0 + u * 32 + lane;
                k[u] = i < w1 ? src[i] : ~0ull;
            }
#pragma unroll
            for (int u = 0; u < kSortUnroll; ++u)
                if (i0 + u * 32 + lane < w1) atomicAdd(&s_cnt[warp][static_cast<uint32_t>(k[u] >> shift) & 255u], 1u);
        }
        __syncthreads();
        // threads 0..255: digit d's total, an exclusive scan over the digits, then the first slot of (warp, d)
        const int d = threadIdx.x;
        uint32_t tot = 0, incl = 0;
        if (d < 256) {
            for (int w = 0; w < aw; ++w) tot += s_cnt[w][d];
            if (tot == n) s_flag = 1;   // every key has this digit: the pass would be the identity
            incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(kFullMask, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_tot[warp] = incl;   // totals of the 8 groups of 32 digits
        }
        __syncthreads();
        const bool skip = s_flag != 0;   // uniform across the CTA
        if (!skip && d < 256) {
            uint32_t run = incl - tot;
            for (int w = 0; w < warp; ++w) run += s_tot[w];
            for (int w = 0; w < aw; ++w) {
                const uint32_t c = s_cnt[w][d];
                s_cnt[w][d] = run;
                run += c;
            }
        }
        __syncthreads();
        if (skip) continue;
        for (uint32_t i0 = w0; i0 < w1; i0 += 32 * kSortUnroll) {
            unsigned long long k[kSortUnroll];
#pragma unroll
            for (int u = 0; u < kSortUnroll; ++u) {
                const uint32_t i = i0 + u * 32 + lane;
                k[u] = i < w1 ? src[i] : 0ull;
            }
#pragma unroll
            for (int u = 0; u < kSortUnroll; ++u) {   // 32 keys at a time, in order: the pass is stable
                if (i0 + u * 32 >= w1) break;         // warp-uniform
                const bool have = i0 + u * 32 + lane < w1;
                const uint32_t dg = have ? (static_cast<uint32_t>(k[u] >> shift) & 255u) : 256u + lane;  // idle lanes: unique pseudo-digits
                unsigned peers;
                if (BALLOT) {  // lanes with the same 8-bit digit, from one ballot per bit (the multi-split CUB's ranking uses)
                    peers = __ballot_sync(kFullMask, have);
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        const bool bit = (dg >> b) & 1u;
                        const unsigned bal = __ballot_sync(kFullMask, bit);
                        peers &= bit ? bal : ~bal;
                    }
                    if (!have) peers = 1u << lane;
                } else {
                    peers = __match_any_sync(kFullMask, dg);
                }
                const int leader = __ffs(peers) - 1;
                uint32_t slot = 0;
                if (have && lane == leader) slot = atomicAdd(&s_cnt[warp][dg], static_cast<uint32_t>(__popc(peers)));
                slot = __shfl_sync(kFullMask, slot, leader);
                if (have) dst[slot + __popc(peers & lt_mask)] = k[u];
            }
        }
        __syncthreads();
        unsigned long long* t = src;
        src = dst;
        dst = t;
    }
    // `src` holds the slice in depth order.  Runs of EQUAL depth bits are ordered by particle index by the thread whose element starts
    // the run (runs are disjoint, a run is two or three keys long; a serial walk by one thread cost milliseconds here: every second C3
    // tile has such a pair)
    for (uint32_t i = threadIdx.x; i + 1 < n; i += kSortThreads) {
        const unsigned long long a = src[i];
        if ((a >> 32) != (src[i + 1] >> 32)) continue;
        if (i > 0 && (src[i - 1] >> 32) == (a >> 32)) continue;  // inside a run: its first element's thread handles it
        uint32_t e = i + 2;
        while (e < n && (src[e] >> 32) == (a >> 32)) ++e;
        for (uint32_t p = i + 1; p < e; ++p) {  // insertion sort of src[i .. e)
            const unsigned long long k = src[p];
            uint32_t q = p;
            while (q > i && src[q - 1] > k) {
                src[q] = src[q - 1];
                --q;
            }
            src[q] = k;
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += kSortThreads) sorted_values[begin + i] = static_cast<uint32_t>(src[i]);
}

// test-only: the reference's sorted 64-bit keys (tile << 32 | depth bits), rebuilt per tile from the sorted values
__global__ void __launch_bounds__(256) synth_tile_keys_kernel(const uint32_t* __restrict__ ranges, const uint32_t* __restrict__ vals,
                                                              const float* __restrict__ depth, uint64_t* __restrict__ out) {
    const uint32_t tile = blockIdx.x;
    const uint32_t begin = ranges[tile * 2], end = ranges[tile * 2 + 1];
    for (uint32_t k = begin + threadIdx.x; k < end; k += blockDim.x)
        out[k] = (static_cast<uint64_t>(tile) << 32) | __float_as_uint(depth[vals[k]]);
}

}  // namespace

void launch_tile_scan(cudaStream_t s, int num_tiles, const uint32_t* counts, uint32_t capacity, uint32_t* ranges, uint32_t* sub_base,
                      uint32_t* chunk_base, uint32_t* order, uint32_t* fill, uint32_t* totals) {
    tile_scan_kernel<<<1, 1024, 0, s>>>(num_tiles, counts, capacity, ranges, sub_base, chunk_base, order, fill, totals);
}

// heaviest tiles come first in `order`: the long lists start before the bulk of the short ones
cudaError_t launch_tile_sort(cudaStream_t s, int num_tiles, const uint32_t* order, const uint32_t* ranges, const uint32_t* totals,
                             unsigned long long* keys, unsigned long long* keys_alt, uint32_t* sorted_values) {
    if (num_tiles <= 0) return cudaSuccess;
    static const bool use_match = [] { const char* e = std::getenv("GUTB200_SORT_MATCH"); return e && std::atoi(e) != 0; }();  // A/B switch
    if (use_match)
        tile_sort_kernel<false><<<num_tiles, kSortThreads, 0, s>>>(order, ranges, totals, keys, keys_alt, sorted_values);
    else
        tile_sort_kernel<true><<<num_tiles, kSortThreads, 0, s>>>(order, ranges, totals, keys, keys_alt, sorted_values);
    return cudaGetLastError();
}

void launch_synth_tile_keys(cudaStream_t s, int num_tiles, const uint32_t* ranges, const uint32_t* vals, const float* depth, uint64_t* out) {
    if (num_tiles <= 0) return;
    synth_tile_keys_kernel<<<num_tiles, 256, 0, s>>>(ranges, vals, depth, out);
}

}  // namespace gutb200
