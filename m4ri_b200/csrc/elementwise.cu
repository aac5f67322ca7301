// elementwise.cu — HBM-bound helpers around the leaf: the device forms of _mzd_add
// (m4ri/mzd.c:1471-1583), mzd_copy (mzd.c:1363-1382) and mzd_set_ui(.,0) (mzd.c:1294-1301) on
// 128-bit aligned views, plus the excess-bit mask applied after uploads.
//
// All views handed to these kernels are 16-byte aligned and 128-bit padded (dev.h), so every
// access is a coalesced 128-bit load/store; grid = a multiple of the SM count, grid-stride loop.
#include "dev.h"

namespace m4b {

unsigned long long g_kernel_launches = 0;

namespace {

struct V128 {
  uint4  *p;
  int64_t pitch;   // in uint4 units
};

// streaming load (no L1 allocation).  Not .nc: C may alias A or B (in-place adds), each element
// is read and then written by the same thread only.
__device__ __forceinline__ uint4 ldg_stream(uint4 const *p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// mode 0: C = A ^ B   mode 1: C = A   mode 2: C = 0
template <int MODE>
__global__ void __launch_bounds__(256) ew_kernel(V128 C, V128 A, V128 B, int rows, int w128) {
  int64_t const total = (int64_t)rows * w128;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t const r = i / w128;
    int const     c = (int)(i - r * w128);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (MODE <= 1) v = ldg_stream(A.p + r * A.pitch + c);
    if (MODE == 0) {
      uint4 const b = ldg_stream(B.p + r * B.pitch + c);
      v.x ^= b.x; v.y ^= b.y; v.z ^= b.z; v.w ^= b.w;
    }
    C.p[r * C.pitch + c] = v;
  }
}

// Clear bits [ncols, round_up(ncols,128)) of every row (one thread per row).
__global__ void mask_excess_kernel(word *data, int64_t pitch, int rows, int ncols) {
  int const r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  int const w = ncols / 64, bit = ncols % 64;
  word *row = data + (int64_t)r * pitch;
  if (bit) row[w] &= (~(word)0) >> (64 - bit);
  int const first_clear = bit ? w + 1 : w;
  int const end = ((ncols + 127) / 128) * 2;
  for (int j = first_clear; j < end; ++j) row[j] = 0;
}

V128 v128(DView const &V) { return V128{reinterpret_cast<uint4 *>(V.data), V.pitch / 2}; }

template <int MODE>
void launch_ew(DView C, DView A, DView B, cudaStream_t s) {
  if (C.nrows <= 0 || C.ncols <= 0) return;
  int const w128 = (C.ncols + 127) / 128;
  int64_t const total = (int64_t)C.nrows * w128;
  int64_t blocks = (total + 255) / 256;
  int64_t const cap = (int64_t)m4rm_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  ew_kernel<MODE><<<(unsigned)blocks, 256, 0, s>>>(v128(C), v128(A), v128(B), C.nrows, w128);
  M4B_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}

}  // namespace

void launch_xor(DView C, DView A, DView B, cudaStream_t s) { launch_ew<0>(C, A, B, s); }
void launch_copy(DView C, DView A, cudaStream_t s) { launch_ew<1>(C, A, A, s); }
void launch_zero(DView C, cudaStream_t s) { launch_ew<2>(C, C, C, s); }

void launch_mask_excess(DView C, cudaStream_t s) {
  if (C.nrows <= 0 || C.ncols % 128 == 0) return;
  mask_excess_kernel<<<(C.nrows + 127) / 128, 128, 0, s>>>(C.data, C.pitch, C.nrows, C.ncols);
  M4B_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}

}  // namespace m4b
