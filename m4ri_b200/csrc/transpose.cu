// transpose.cu — bit-matrix transpose on the device (reference: mzd_transpose,
// m4ri/mzd.c:1104-1139, a recursive 64x64 word-swap scheme tuned for CPU caches).
//
// HBM-bound by nature (every bit read once, written once), so the design goal is full 128-byte
// sectors on BOTH sides: a CTA owns a 512 x 512-bit tile, loads it with coalesced 64-byte row segments
// into shared memory (row stride 17 words: column accesses are conflict-free), transposes it in place
// as 32 x 32 blocks — each block is transposed inside a warp with five shuffle/mask stages, block
// (i, j) swaps with block (j, i) — and writes the 512 transposed rows back as 64-byte segments.  34 KB
// of shared memory per CTA lets several CTAs per SM overlap their load, shuffle and store phases.  Rows/columns outside the source read as zero, so the destination keeps the "bits beyond
// ncols are zero" invariant of device matrices.
#include "dev.h"

namespace m4b {
namespace {

constexpr int kT = 512;               // tile edge in bits
constexpr int kW = kT / 32;           // 16 words per tile row
constexpr int kStride = kW + 1;       // padded (odd) row stride in shared memory: column accesses hit 32 banks
constexpr int kThreads = 256;         // 34 KB of shared memory per CTA -> several CTAs per SM overlap their phases

// lane i holds row i of a 32 x 32 bit block (column j in bit j); returns column lane as a row
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
  uint32_t m = 0x0000FFFFu;
#pragma unroll
  for (int j = 16; j; j >>= 1) {
    uint32_t const y = __shfl_xor_sync(0xffffffffu, x, j);
    if (lane & j) x = (x & (m << j)) | ((y >> j) & m);   // keep own columns with bit j set
    else          x = (x & m) | ((y & m) << j);          // keep own columns with bit j clear
    m ^= m << (j >> 1);                                  // 0x0000FFFF -> 0x00FF00FF -> 0x0F0F0F0F -> ...
  }
  return x;
}

__global__ void __launch_bounds__(kThreads) transpose_kernel(uint32_t const *__restrict__ src, long long spitch32,
                                                             int srows, int sw32, uint32_t *__restrict__ dst,
                                                             long long dpitch32, int drows, int dw32) {
  __shared__ uint32_t tile[kT * kStride];
  __shared__ uint8_t pair_bi[kW * (kW + 1) / 2], pair_bj[kW * (kW + 1) / 2];
  constexpr int kPairs = kW * (kW + 1) / 2;
  if (threadIdx.x < kPairs) {            // block pair p -> (bi, bj), bi <= bj, once per CTA
    int bi = 0, rem = threadIdx.x;
    while (rem >= kW - bi) { rem -= kW - bi; ++bi; }
    pair_bi[threadIdx.x] = (uint8_t)bi;
    pair_bj[threadIdx.x] = (uint8_t)(bi + rem);
  }
  int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kThreads / 32;
  int const half = lane >> 4, w = lane & 15;      // a warp moves two 64-byte row segments per instruction
  int const r0 = blockIdx.y * kT;      // first source row of the tile
  int const c0w = blockIdx.x * kW;     // first source word of the tile

  constexpr int kRowsPerWarp = kT / (kThreads / 32);   // 64 rows, 32 two-row loads per warp
#pragma unroll 8
  for (int i = 0; i < kRowsPerWarp / 2; ++i) {
    int const r = warp * kRowsPerWarp + 2 * i + half;
    uint32_t v = 0;
    if (r0 + r < srows && c0w + w < sw32) v = src[(long long)(r0 + r) * spitch32 + c0w + w];
    tile[r * kStride + w] = v;
  }
  __syncthreads();

  // 32 x 32 blocks; the pairs (bi, bj), bi <= bj, are independent of each other
  for (int p = warp; p < kPairs; p += nwarps) {
    int const bi = pair_bi[p], bj = pair_bj[p];
    uint32_t const a = transpose32(tile[(32 * bi + lane) * kStride + bj], lane);
    if (bi == bj) {
      tile[(32 * bi + lane) * kStride + bi] = a;
    } else {
      uint32_t const b = transpose32(tile[(32 * bj + lane) * kStride + bi], lane);
      tile[(32 * bj + lane) * kStride + bi] = a;
      tile[(32 * bi + lane) * kStride + bj] = b;
    }
  }
  __syncthreads();

  int const dr0 = blockIdx.x * kT;     // first destination row = first source column
  int const dc0w = blockIdx.y * kW;    // first destination word
#pragma unroll 8
  for (int i = 0; i < kRowsPerWarp / 2; ++i) {
    int const r = warp * kRowsPerWarp + 2 * i + half;
    if (dr0 + r < drows && dc0w + w < dw32) dst[(long long)(dr0 + r) * dpitch32 + dc0w + w] = tile[r * kStride + w];
  }
}

}  // namespace

// dst (n x m) = src (m x n)^T on device views; dst must not alias src.
void launch_transpose(DView dst, DView src, cudaStream_t s) {
  if (src.nrows <= 0 || src.ncols <= 0) return;
  dim3 const grid((src.ncols + kT - 1) / kT, (src.nrows + kT - 1) / kT);
  transpose_kernel<<<grid, kThreads, 0, s>>>(reinterpret_cast<uint32_t const *>(src.data), src.pitch * 2, src.nrows,
                                                ((src.ncols + 127) / 128) * 4, reinterpret_cast<uint32_t *>(dst.data),
                                                dst.pitch * 2, dst.nrows, ((dst.ncols + 127) / 128) * 4);
  M4B_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}

}  // namespace m4b
