// Internal interface of radix_sort.cu / scan.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace mwe {

size_t sort_workspace_bytes(int64_t N);
// Stable sort on the low key_bits of the keys.  The sorted pairs end up either in (keys, vals) or
// in the ping-pong buffers inside the workspace; *keys_sorted / *vals_sorted say which.
int sort_pairs(uint64_t* keys, uint32_t* vals, int64_t N, int key_bits, void* ws, size_t ws_bytes,
               cudaStream_t stream, uint64_t** keys_sorted, uint32_t** vals_sorted);

// Exclusive prefix sum of n int32 values (out may alias in). total_out (device, nullable) gets the sum.
size_t scan_workspace_bytes(int64_t n);
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int64_t* total_out, void* ws, size_t ws_bytes,
                       cudaStream_t stream);

}  // namespace mwe
