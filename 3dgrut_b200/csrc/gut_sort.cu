// 3dgrut_b200/csrc/gut_sort.cu -- G2 prefix sum and G4 (tile,depth) key sort.
//
// Round-1 status: the scan and the radix-sort passes call CUB (header-only CCCL shipped with the CUDA toolkit), the same
// library the reference calls (threedgut_tracer/src/gutRenderer.cu:303,356-365); they are LIBRARY code, not counted as
// ours.  What is ours is the decomposition: depth-sort the N particles once, then sort only the
// tile bits of the I intersections, instead of one 44-bit sort of I 64-bit keys.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "gut_common.cuh"

namespace gutb200 {

size_t scan_temp_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, bytes, static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr),
                                  static_cast<int>(n));
    return bytes;
}

void run_inclusive_scan(cudaStream_t s, void* temp, size_t temp_bytes, const uint32_t* in, uint32_t* out, int64_t n) {
    cub::DeviceScan::InclusiveSum(temp, temp_bytes, in, out, static_cast<int>(n), s);
}

// stable LSD radix sort of (32-bit key, 32-bit payload) pairs on bits [0, end_bit): depth keys and tile keys of the 3DGUT
// binning, Morton codes of the 3DGRT LBVH build
size_t sort32_temp_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr),
                                    static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), n, 0, 32);
    return bytes;
}

void run_sort32_pairs(cudaStream_t s, void* temp, size_t temp_bytes, const uint32_t* kin, uint32_t* kout, const uint32_t* vin,
                      uint32_t* vout, int64_t n, int end_bit) {
    cub::DeviceRadixSort::SortPairs(temp, temp_bytes, kin, kout, vin, vout, n, 0, end_bit, s);
}

}  // namespace gutb200
