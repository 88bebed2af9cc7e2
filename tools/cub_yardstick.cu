// cub_yardstick.cu — the external bar for the instance sort: cub::DeviceRadixSort::SortPairs(begin_bit, end_bit) on the
// SAME (tile << 32 | depth bits) keys and 1-based Gaussian ids, i.e. what a CUDA port of the reference's
// sortperm! + 2 x _permute! (rasterizer.jl:357-372) would call.  Measurement infrastructure only: built into its own
// tools/libcub_yardstick.so by tools/cub_yardstick.py, never linked into libgsrast.so.
#include <cub/device/device_radix_sort.cuh>
#include <cstdint>

extern "C" {

// returns the temp-storage bytes CUB asks for
__attribute__((visibility("default"))) size_t cub_sort_pairs_temp_bytes(int64_t m, int begin_bit, int end_bit) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const uint32_t *)nullptr,
                                    (uint32_t *)nullptr, (int)m, begin_bit, end_bit, 0);
    return bytes;
}

__attribute__((visibility("default"))) int cub_sort_pairs(void *temp, size_t temp_bytes, const uint64_t *keys_in, uint64_t *keys_out,
                                                          const uint32_t *vals_in, uint32_t *vals_out, int64_t m, int begin_bit,
                                                          int end_bit, void *stream) {
    return (int)cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, (int)m, begin_bit, end_bit,
                                                static_cast<cudaStream_t>(stream));
}
}
