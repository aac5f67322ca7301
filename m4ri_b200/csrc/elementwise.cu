// elementwise.cu — HBM-bound helpers around the leaf: the device forms of _mzd_add
// (m4ri/mzd.c:1471-1583), mzd_copy (mzd.c:1363-1382) and mzd_set_ui(.,0) (mzd.c:1294-1301) on
// 128-bit aligned views, plus the excess-bit mask applied after uploads.
//
// All views handed to these kernels are 16-byte aligned and 128-bit padded (dev.h), so every
// access is a coalesced 128-bit load/store; grid = a multiple of the SM count, grid-stride loop.
#include <string.h>

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

// ---- fused Winograd pre/post additions (strassen.cu: winograd_node) ------------------------------
// One launch replaces the 4 + 4 operand additions, resp. the 7 result additions, of a Strassen-Winograd
// node: every 128-bit element of the quadrants is read once and written once.
struct VSet {
  V128 v[8];
};

__device__ __forceinline__ uint4 x4(uint4 a, uint4 const &b) { a.x ^= b.x; a.y ^= b.y; a.z ^= b.z; a.w ^= b.w; return a; }

// MODE 0: in = A11 A12 A21 A22 -> out = S1 S2 S3 S4      (S1 = A21+A22, S2 = S1+A11, S3 = A11+A21, S4 = A12+S2)
// MODE 1: in = B11 B12 B21 B22 -> out = T1 T2 T3 T4      (T1 = B12+B11, T2 = B22+T1, T3 = B22+B12, T4 = T2+B21)
// MODE 2: in = P1..P7 -> out = C11 C12 C21 C22 (overwrite)   MODE 3: same, accumulated onto C
constexpr int kMaxNodes = 7;     // nodes of identical shape handled by one launch (blockIdx.y)
struct VBatch {
  VSet in[kMaxNodes], out[kMaxNodes];
};

template <int MODE>
__global__ void __launch_bounds__(256) winograd_ew_kernel(const __grid_constant__ VBatch batch, int rows, int w128) {
  VSet const &in = batch.in[blockIdx.y], &out = batch.out[blockIdx.y];
  int64_t const total = (int64_t)rows * w128;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t const r = i / w128;
    int const     c = (int)(i - r * w128);
    auto ld = [&](int k) { return ldg_stream(in.v[k].p + r * in.v[k].pitch + c); };
    auto st = [&](int k, uint4 const &v) { out.v[k].p[r * out.v[k].pitch + c] = v; };
    if (MODE == 0) {
      uint4 const a11 = ld(0), a12 = ld(1), a21 = ld(2), a22 = ld(3);
      uint4 const s1 = x4(a21, a22), s2 = x4(s1, a11);
      st(0, s1); st(1, s2); st(2, x4(a11, a21)); st(3, x4(a12, s2));
    } else if (MODE == 1) {
      uint4 const b11 = ld(0), b12 = ld(1), b21 = ld(2), b22 = ld(3);
      uint4 const t1 = x4(b12, b11), t2 = x4(b22, t1);
      st(0, t1); st(1, t2); st(2, x4(b22, b12)); st(3, x4(t2, b21));
    } else {
      uint4 const p1 = ld(0), p2 = ld(1), p3 = ld(2), p4 = ld(3), p5 = ld(4), p6 = ld(5), p7 = ld(6);
      uint4 const u2 = x4(p1, p6), u3 = x4(u2, p7);
      uint4 c11 = x4(p1, p2), c12 = x4(x4(u2, p5), p3), c21 = x4(u3, p4), c22 = x4(u3, p5);
      if (MODE == 3) {
        c11 = x4(c11, out.v[0].p[r * out.v[0].pitch + c]);
        c12 = x4(c12, out.v[1].p[r * out.v[1].pitch + c]);
        c21 = x4(c21, out.v[2].p[r * out.v[2].pitch + c]);
        c22 = x4(c22, out.v[3].p[r * out.v[3].pitch + c]);
      }
      st(0, c11); st(1, c12); st(2, c21); st(3, c22);
    }
  }
}

// `nodes` Winograd nodes of identical shape in one launch: in[node * nin + k], out[node * nout + k]
template <int MODE>
void launch_winograd_ew(DView const *in, int nin, DView const *out, int nout, cudaStream_t s, int nodes = 1) {
  int const rows = out[0].nrows, w128 = (out[0].ncols + 127) / 128;
  if (rows <= 0 || w128 <= 0) return;
  if (nodes > kMaxNodes) die("m4ri_b200: %d Winograd nodes in one launch exceed %d\n", nodes, kMaxNodes);
  VBatch b;
  memset(&b, 0, sizeof b);
  for (int nd = 0; nd < nodes; ++nd) {
    for (int k = 0; k < nin; ++k) b.in[nd].v[k] = v128(in[nd * nin + k]);
    for (int k = 0; k < nout; ++k) b.out[nd].v[k] = v128(out[nd * nout + k]);
  }
  int64_t const total = (int64_t)rows * w128;
  int64_t blocks = (total + 255) / 256;
  int64_t const cap = ((int64_t)m4rm_num_sms() * 8 + nodes - 1) / nodes;
  if (blocks > cap) blocks = cap;
  winograd_ew_kernel<MODE><<<dim3((unsigned)blocks, (unsigned)nodes), 256, 0, s>>>(b, rows, w128);
  M4B_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}

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

void launch_winograd_pre_a(DView const a[4], DView const s_out[4], cudaStream_t s) { launch_winograd_ew<0>(a, 4, s_out, 4, s); }
void launch_winograd_pre_b(DView const b[4], DView const t_out[4], cudaStream_t s) { launch_winograd_ew<1>(b, 4, t_out, 4, s); }
void launch_winograd_post(DView const p[7], DView const c[4], bool accumulate, cudaStream_t s) {
  if (accumulate) launch_winograd_ew<3>(p, 7, c, 4, s);
  else            launch_winograd_ew<2>(p, 7, c, 4, s);
}

// the same for `nodes` (<= 7) nodes of identical shape: a[node * 4 + q] etc.
void launch_winograd_pre_a_batch(int nodes, DView const *a, DView const *s_out, cudaStream_t s) { launch_winograd_ew<0>(a, 4, s_out, 4, s, nodes); }
void launch_winograd_pre_b_batch(int nodes, DView const *b, DView const *t_out, cudaStream_t s) { launch_winograd_ew<1>(b, 4, t_out, 4, s, nodes); }
void launch_winograd_post_batch(int nodes, DView const *p, DView const *c, cudaStream_t s) { launch_winograd_ew<2>(p, 7, c, 4, s, nodes); }

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
