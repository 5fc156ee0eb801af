// Warp-level transposes for thread-per-row tcgen05 epilogues.
//
// After tcgen05.ld every lane owns one ROW of the accumulator tile, so a naive store writes 16 bytes per lane to 32
// different rows: one store instruction touches 32 cache lines and the LSU, not HBM, bounds the epilogue (measured:
// 2.6x off the HBM roofline on the K = 64 ResNet layers).  These helpers bounce a 32-row x 64-byte block through a
// per-warp shared-memory scratch so that a global access instruction covers 8 rows x 64 contiguous bytes instead
// (4x fewer line transactions), in both directions.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>

namespace ehb {

constexpr int EPI_PITCH = 80;                       // bytes per scratch row: 64 + 16 keeps the 128-bit phases conflict-free
constexpr int EPI_SCRATCH_BYTES = 32 * EPI_PITCH;   // per warp

// Each lane holds the 64 bytes (32 fp16) of ITS row in `mine`; the block goes to rows [0, 32) of `gbase` (row pitch
// `ld` halves), rows >= rows_valid are skipped.
__device__ __forceinline__ void warp_store_rows_64B(uint8_t* scratch, const uint4 (&mine)[4], __half* gbase, size_t ld,
                                                    long long rows_valid, int lane) {
  uint4* srow = reinterpret_cast<uint4*>(scratch + lane * EPI_PITCH);
#pragma unroll
  for (int j = 0; j < 4; ++j) srow[j] = mine[j];
  __syncwarp();
  const int piece = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2);
    const uint4 val = *reinterpret_cast<const uint4*>(scratch + r * EPI_PITCH + piece * 16);
    if (r < rows_valid) *reinterpret_cast<uint4*>(gbase + static_cast<size_t>(r) * ld + piece * 8) = val;
  }
  __syncwarp();
}

// The reverse: rows [0, 32) x 64 bytes of `gbase` -> every lane gets its own row in `mine` (zeros for invalid rows).
__device__ __forceinline__ void warp_load_rows_64B(uint8_t* scratch, uint4 (&mine)[4], const __half* gbase, size_t ld,
                                                   long long rows_valid, int lane) {
  const int piece = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2);
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r < rows_valid) val = __ldg(reinterpret_cast<const uint4*>(gbase + static_cast<size_t>(r) * ld + piece * 8));
    *reinterpret_cast<uint4*>(scratch + r * EPI_PITCH + piece * 16) = val;
  }
  __syncwarp();
  const uint4* srow = reinterpret_cast<const uint4*>(scratch + lane * EPI_PITCH);
#pragma unroll
  for (int j = 0; j < 4; ++j) mine[j] = srow[j];
  __syncwarp();
}

// Split version of the load for software pipelining: `issue` starts the coalesced global loads of a hi block and a lo
// block (8 loads in flight, results in registers), `finish` transposes them through the scratch into per-row data.
__device__ __forceinline__ void warp_issue_rows_64B(uint4 (&raw)[4], const __half* gbase, size_t ld, long long rows_valid,
                                                    int lane) {
  const int piece = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2);
    raw[i] = r < rows_valid ? __ldg(reinterpret_cast<const uint4*>(gbase + static_cast<size_t>(r) * ld + piece * 8))
                            : make_uint4(0, 0, 0, 0);
  }
}
__device__ __forceinline__ void warp_finish_rows_64B(uint8_t* scratch, const uint4 (&raw)[4], uint4 (&mine)[4], int lane) {
  const int piece = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2);
    *reinterpret_cast<uint4*>(scratch + r * EPI_PITCH + piece * 16) = raw[i];
  }
  __syncwarp();
  const uint4* srow = reinterpret_cast<const uint4*>(scratch + lane * EPI_PITCH);
#pragma unroll
  for (int j = 0; j < 4; ++j) mine[j] = srow[j];
  __syncwarp();
}

}  // namespace ehb
