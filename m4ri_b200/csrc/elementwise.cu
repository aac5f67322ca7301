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

// ---- two Winograd levels in one pass (strassen.cu: winograd_node2) ------------------------------------------------
// A (or B) is seen as 4 x 4 sub-blocks a[q1][q2] = quadrant q2 of quadrant q1.  One launch writes the four level-1 sums
// (as their 16 sub-blocks: they are operands of the level-2 products through quadrant views) AND the 28 level-2 sums
// of the seven level-1 operands, each an XOR of a subset of the 16 inputs given by the level masks below; the result
// side reads the 49 products once and writes the 16 sub-blocks of C.  Against running the two levels as separate
// passes this saves re-reading the level-1 operands (28 + 28 sub-blocks) and writing + re-reading the seven
// intermediate results (56 sub-blocks) per node: 224 MB of 692 MB at 4096^3 leaves.
struct V16 {
  V128 v[16];
};
struct V44 {
  V128 v[44];
};
struct V49 {
  V128 v[49];
};

// level masks over quadrant indices (bit q = quadrant q; order 11, 12, 21, 22)
__host__ __device__ constexpr unsigned sum_mask(int side, int s) {        // S1..S4 / T1..T4 of winograd_ew_kernel<0/1>
  return side == 0 ? (s == 0 ? 0xCu : s == 1 ? 0xDu : s == 2 ? 0x5u : 0xFu) : (s == 0 ? 0x3u : s == 1 ? 0xBu : s == 2 ? 0xAu : 0xFu);
}
__host__ __device__ constexpr unsigned operand_mask(int side, int i) {    // X1[i] / Y1[i] of winograd_node in terms of quadrants
  return side == 0 ? (i == 0 ? 0x1u : i == 1 ? 0x2u : i == 2 ? sum_mask(0, 3) : i == 3 ? 0x8u : sum_mask(0, i - 4))
                   : (i == 0 ? 0x1u : i == 1 ? 0x4u : i == 2 ? 0x8u : i == 3 ? sum_mask(1, 3) : sum_mask(1, i - 4));
}
__host__ __device__ constexpr unsigned result_mask(int q) {               // C quadrant q over P1..P7 (bit i = P_{i+1})
  return q == 0 ? 0x03u : q == 1 ? 0x35u : q == 2 ? 0x69u : 0x71u;
}

// out[4*s + q2] = level-1 sum s, sub-block q2;  out[16 + 4*i + t] = level-2 sum t of level-1 operand i
template <int SIDE>
__global__ void __launch_bounds__(256) winograd_pre2_kernel(const __grid_constant__ V16 in, const __grid_constant__ V44 out, int rows,
                                                            int w128) {
  int64_t const total = (int64_t)rows * w128;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t const r = e / w128;
    int const     c = (int)(e - r * w128);
    uint4 a[4][4];
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1)
#pragma unroll
      for (int q2 = 0; q2 < 4; ++q2) a[q1][q2] = ldg_stream(in.v[4 * q1 + q2].p + r * in.v[4 * q1 + q2].pitch + c);
    auto st = [&](int k, uint4 const &v) { out.v[k].p[r * out.v[k].pitch + c] = v; };
    // level-1 sums, sub-block by sub-block
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
      for (int q2 = 0; q2 < 4; ++q2) {
        uint4 v = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int q1 = 0; q1 < 4; ++q1)
          if ((sum_mask(SIDE, s) >> q1) & 1u) v = x4(v, a[q1][q2]);
        st(4 * s + q2, v);
      }
    // row sums R[q1][t] = XOR over q2 in sum_mask(t) of a[q1][q2], then the level-2 sums of every level-1 operand
    uint4 R[4][4];
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1)
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        uint4 v = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int q2 = 0; q2 < 4; ++q2)
          if ((sum_mask(SIDE, t) >> q2) & 1u) v = x4(v, a[q1][q2]);
        R[q1][t] = v;
      }
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        uint4 v = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int q1 = 0; q1 < 4; ++q1)
          if ((operand_mask(SIDE, i) >> q1) & 1u) v = x4(v, R[q1][t]);
        st(16 + 4 * i + t, v);
      }
  }
}

// in[7*i + j] = product j of inner node i;  out[4*Q1 + Q2] = sub-block Q2 of quadrant Q1 of C
template <bool ACC>
__global__ void __launch_bounds__(256) winograd_post2_kernel(const __grid_constant__ V49 in, const __grid_constant__ V16 out, int rows,
                                                             int w128) {
  int64_t const total = (int64_t)rows * w128;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t const r = e / w128;
    int const     c = (int)(e - r * w128);
    uint4 acc[4][4];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q >> 2][q & 3] = ACC ? out.v[q].p[r * out.v[q].pitch + c] : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      uint4 p[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) p[j] = ldg_stream(in.v[7 * i + j].p + r * in.v[7 * i + j].pitch + c);
#pragma unroll
      for (int Q2 = 0; Q2 < 4; ++Q2) {                 // quadrant Q2 of the inner node's result
        uint4 v = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < 7; ++j)
          if ((result_mask(Q2) >> j) & 1u) v = x4(v, p[j]);
#pragma unroll
        for (int Q1 = 0; Q1 < 4; ++Q1)
          if ((result_mask(Q1) >> i) & 1u) acc[Q1][Q2] = x4(acc[Q1][Q2], v);
      }
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) out.v[q].p[r * out.v[q].pitch + c] = acc[q >> 2][q & 3];
  }
}

unsigned ew_blocks(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  int64_t const cap = (int64_t)m4rm_num_sms() * 8;
  return (unsigned)(blocks > cap ? cap : blocks);
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

// two-level forms: sub[4 * q1 + q2] = quadrant q2 of quadrant q1 of the operand; sums[0..16) = the level-1 sums by
// sub-block, sums[16 + 4 * i + t] = level-2 sum t of level-1 operand i; prods[7 * i + j]; csub[4 * Q1 + Q2]
void launch_winograd_pre2(int side, DView const *sub, DView const *sums, cudaStream_t s) {
  int const rows = sub[0].nrows, w128 = (sub[0].ncols + 127) / 128;
  if (rows <= 0 || w128 <= 0) return;
  V16 vi;
  V44 vo;
  for (int k = 0; k < 16; ++k) vi.v[k] = v128(sub[k]);
  for (int k = 0; k < 44; ++k) vo.v[k] = v128(sums[k]);
  unsigned const blocks = ew_blocks((int64_t)rows * w128);
  if (side == 0) winograd_pre2_kernel<0><<<blocks, 256, 0, s>>>(vi, vo, rows, w128);
  else           winograd_pre2_kernel<1><<<blocks, 256, 0, s>>>(vi, vo, rows, w128);
  M4B_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}

void launch_winograd_post2(DView const *prods, DView const *csub, bool accumulate, cudaStream_t s) {
  int const rows = csub[0].nrows, w128 = (csub[0].ncols + 127) / 128;
  if (rows <= 0 || w128 <= 0) return;
  V49 vi;
  V16 vo;
  for (int k = 0; k < 49; ++k) vi.v[k] = v128(prods[k]);
  for (int k = 0; k < 16; ++k) vo.v[k] = v128(csub[k]);
  unsigned const blocks = ew_blocks((int64_t)rows * w128);
  if (accumulate) winograd_post2_kernel<true><<<blocks, 256, 0, s>>>(vi, vo, rows, w128);
  else            winograd_post2_kernel<false><<<blocks, 256, 0, s>>>(vi, vo, rows, w128);
  M4B_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}

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
