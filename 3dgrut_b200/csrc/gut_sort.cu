// 3dgrut_b200/csrc/gut_sort.cu -- the one library sort left: Morton codes of the 3DGRT LBVH build (grt.cu).
//
// CUB (header-only CCCL shipped with the CUDA toolkit) is LIBRARY code, not counted as ours.  The 3DGUT path no longer calls it: its
// binning is gut_binning.cu (per-tile histogram -> one-CTA scan -> atomic placement -> per-tile radix sort).
#include <cub/device/device_radix_sort.cuh>

#include "gut_common.cuh"

namespace gutb200 {

// stable LSD radix sort of (32-bit key, 32-bit payload) pairs on bits [0, end_bit)
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
