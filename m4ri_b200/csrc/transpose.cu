// transpose.cu — bit-matrix transpose on the device (reference: mzd_transpose,
// m4ri/mzd.c:1104-1139, a recursive 64x64 word-swap scheme tuned for CPU caches).
//
// HBM-bound by nature (every bit read once, written once), so the design goal is full 128-byte
// transactions on BOTH sides: a CTA owns a 1024 x 1024-bit tile, loads it with coalesced 128-byte row
// reads into shared memory (row stride 33 words: column accesses are conflict-free), transposes it in
// place as 32 x 32 blocks — each block is transposed inside a warp with five shuffle/mask stages, block
// (i, j) swaps with block (j, i) — and writes the 1024 transposed rows back with coalesced 128-byte
// stores.  Rows/columns outside the source read as zero, so the destination keeps the "bits beyond
// ncols are zero" invariant of device matrices.
#include "dev.h"

namespace m4b {
namespace {

constexpr int kT = 1024;              // tile edge in bits
constexpr int kW = kT / 32;           // 32 words per tile row
constexpr int kStride = kW + 1;       // padded row stride in shared memory
constexpr int kThreads = 512;

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
  extern __shared__ uint32_t tile[];   // [kT][kStride]
  int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kThreads / 32;
  int const r0 = blockIdx.y * kT;      // first source row of the tile
  int const c0w = blockIdx.x * kW;     // first source word of the tile

  for (int r = warp; r < kT; r += nwarps) {
    uint32_t v = 0;
    if (r0 + r < srows && c0w + lane < sw32) v = src[(long long)(r0 + r) * spitch32 + c0w + lane];
    tile[r * kStride + lane] = v;
  }
  __syncthreads();

  // 32 x 32 blocks; the pairs (bi, bj), bi <= bj, are independent of each other
  for (int p = warp; p < kW * (kW + 1) / 2; p += nwarps) {
    int bi = 0, rem = p;
    while (rem >= kW - bi) { rem -= kW - bi; ++bi; }
    int const bj = bi + rem;
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
  for (int r = warp; r < kT; r += nwarps)
    if (dr0 + r < drows && dc0w + lane < dw32) dst[(long long)(dr0 + r) * dpitch32 + dc0w + lane] = tile[r * kStride + lane];
}

}  // namespace

// dst (n x m) = src (m x n)^T on device views; dst must not alias src.
void launch_transpose(DView dst, DView src, cudaStream_t s) {
  if (src.nrows <= 0 || src.ncols <= 0) return;
  static bool configured[64] = {};
  int dev = 0;
  M4B_CUDA(cudaGetDevice(&dev));
  size_t const smem = (size_t)kT * kStride * sizeof(uint32_t);
  if (!configured[dev & 63]) {
    M4B_CUDA(cudaFuncSetAttribute(transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dev & 63] = true;
  }
  dim3 const grid((src.ncols + kT - 1) / kT, (src.nrows + kT - 1) / kT);
  transpose_kernel<<<grid, kThreads, smem, s>>>(reinterpret_cast<uint32_t const *>(src.data), src.pitch * 2, src.nrows,
                                                ((src.ncols + 127) / 128) * 4, reinterpret_cast<uint32_t *>(dst.data),
                                                dst.pitch * 2, dst.nrows, ((dst.ncols + 127) / 128) * 4);
  M4B_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}

}  // namespace m4b
