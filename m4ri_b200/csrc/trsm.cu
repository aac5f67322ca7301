// trsm.cu — triangular solves with matrices, left variants, device resident (the first
// caller of the multiplication path that is widened into).
//
// Device counterpart of _mzd_trsm_lower_left / _mzd_trsm_upper_left (m4ri/triangular.c:406-455,
// 467-516): L X = B resp. U X = B over GF(2), X overwrites B, unit diagonal implied, only the strict
// triangle of the triangular operand is read.  Same recursion as the reference —
//     lower:  X0 = L00^-1 B0 ;  B1 ^= L10 X0 ;  X1 = L11^-1 B1
//     upper:  X1 = U11^-1 B1 ;  B0 ^= U01 X1 ;  X0 = U00^-1 B0
// — but the operands stay in HBM for the whole solve, the update is the Strassen/M4RM product of this
// library (accumulate form) on sub-views, and the recursion splits on 128-column boundaries down to a
// 128-row base case.  The reference's base cases (64-row substitution and the "russian" table variant,
// triangular_russian.c:50-168) are replaced by products as well: all 128 x 128 diagonal blocks are
// inverted up front in one small launch (they do not depend on B), and a base step is
// X_block = inv(T_block) * B_block on the M4RM leaf, which is parallel over all columns of B.
#include "dev.h"
#include "workspace.h"

namespace m4b {
namespace {

constexpr int kBaseRows = 128;

// Inverse of every 128 x 128 diagonal block of T in ONE launch (one warp per block; the blocks are
// independent of B and of each other).  Row i of the inverse is e_i plus the XOR of the already
// inverted rows k selected by the strict triangle of T's row i; the 32 lanes split the k range and
// combine with shuffles.  inv: m rows x 128 bits (pitch 2 words), block b in rows [128 b, 128 b + 128).
template <bool UPPER>
__global__ void __launch_bounds__(32) tri_inv128_kernel(uint32_t const *__restrict__ t, long long tpitch32,
                                                        uint32_t *__restrict__ inv, int m) {
  __shared__ uint32_t rows[kBaseRows][4];
  int const blk = blockIdx.x, lane = threadIdx.x;
  int const r0 = blk * kBaseRows;
  int const r = m - r0 < kBaseRows ? m - r0 : kBaseRows;       // rows in this block
  for (int step = 0; step < r; ++step) {
    int const i = UPPER ? r - 1 - step : step;
    uint32_t const *trow = t + (long long)(r0 + i) * tpitch32 + blk * 4;
    uint32_t acc[4] = {0, 0, 0, 0};
    for (int k = lane; k < r; k += 32) {
      bool const in_triangle = UPPER ? k > i : k < i;
      if (in_triangle && ((trow[k >> 5] >> (k & 31)) & 1u)) {
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] ^= rows[k][c];
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int off = 16; off; off >>= 1) acc[c] ^= __shfl_xor_sync(0xffffffffu, acc[c], off);
    }
    if (lane < 4) rows[i][lane] = acc[lane] ^ ((i >> 5) == lane ? 1u << (i & 31) : 0u);
    __syncwarp();
  }
  for (int i = lane; i < r * 4; i += 32) inv[(long long)(r0 + (i >> 2)) * 4 + (i & 3)] = rows[i >> 2][i & 3];
}

// C ^= A*B on views with the deepest Strassen recursion the views' alignment allows
void addmul_views(DView C, DView A, DView B, int cutoff, Workspace &ws, cudaStream_t s) {
  int levels = strassen_levels(A.nrows, A.ncols, B.ncols, cutoff);
  while (levels > 0 && (A.nrows % (1 << levels) || A.ncols % (128 << levels) || B.ncols % (128 << levels))) --levels;
  strassen_mul(C, A, B, levels, false, ws, s);
}

}  // namespace

// t = order of the triangular matrix, (m, n) = shape of B (left variants: t == m, right: t == n)
size_t trsm_workspace_bytes(int t, int m, int n, int cutoff) {
  // every update product is bounded by max-dims; its Strassen temporaries bound all the smaller ones
  int const a = (t > m ? t : m), b = (t > n ? t : n);
  int const ap = (a + 127) / 128 * 128, bp = (b + 127) / 128 * 128;
  return strassen_workspace_bytes(ap, bp, bp, strassen_levels(a, b, b, cutoff)) + Workspace::bytes_for(t, kBaseRows) +
         Workspace::bytes_for(m, kBaseRows) + Workspace::bytes_for(kBaseRows, n);
}

namespace {

// recursion below the top: Tinv holds the inverted diagonal blocks of the WHOLE matrix, row0 = first
// row of this sub-problem inside it
void trsm_rec(DView T, DView B, DView Tinv, int row0, bool upper, int cutoff, Workspace &ws, cudaStream_t s) {
  int const m = B.nrows, n = B.ncols;
  if (m <= kBaseRows) {
    // X = inv(T_block) * B_block: one 128-row product of the M4RM leaf (full-chip parallel over the
    // columns of B), IN PLACE: with a single K slab every C tile has one owner CTA, which has staged its
    // whole B column strip before it stores
    launch_m4rm_overwrite(B, Tinv.sub(row0, 0, row0 + m, m), B, s);
    return;
  }
  int const m1 = ((m + 127) / 128 / 2) * 128;          // multiple of 128, 0 < m1 < m
  DView const T00 = T.sub(0, 0, m1, m1), T11 = T.sub(m1, m1, m, m);
  DView const B0 = B.sub(0, 0, m1, n), B1 = B.sub(m1, 0, m, n);
  if (!upper) {
    trsm_rec(T00, B0, Tinv, row0, false, cutoff, ws, s);
    addmul_views(B1, T.sub(m1, 0, m, m1), B0, cutoff, ws, s);
    trsm_rec(T11, B1, Tinv, row0 + m1, false, cutoff, ws, s);
  } else {
    trsm_rec(T11, B1, Tinv, row0 + m1, true, cutoff, ws, s);
    addmul_views(B0, T.sub(0, m1, m1, m), B1, cutoff, ws, s);
    trsm_rec(T00, B0, Tinv, row0, true, cutoff, ws, s);
  }
}

// right variants, X T = B (m4ri/triangular.c:40-148, 300-392): the recursion runs over column blocks
//     upper:  X0 = B0 T00^-1 ;  B1 ^= X0 T01 ;  X1 = B1 T11^-1
//     lower:  X1 = B1 T11^-1 ;  B0 ^= X1 T10 ;  X0 = B0 T00^-1
// col0 = first column of this sub-problem inside the whole B (= first row of its blocks in Tinv)
void trsm_right_rec(DView T, DView B, DView Tinv, int col0, bool upper, int cutoff, Workspace &ws, cudaStream_t s) {
  int const m = B.nrows, n = B.ncols;
  if (n <= kBaseRows) {
    launch_m4rm_overwrite(B, B, Tinv.sub(col0, 0, col0 + n, n), s);   // X = B_block * inv(T_block), in place
    return;
  }
  int const n1 = ((n + 127) / 128 / 2) * 128;
  DView const T00 = T.sub(0, 0, n1, n1), T11 = T.sub(n1, n1, n, n);
  DView const B0 = B.sub(0, 0, m, n1), B1 = B.sub(0, n1, m, n);
  if (upper) {
    trsm_right_rec(T00, B0, Tinv, col0, true, cutoff, ws, s);
    addmul_views(B1, B0, T.sub(0, n1, n1, n), cutoff, ws, s);
    trsm_right_rec(T11, B1, Tinv, col0 + n1, true, cutoff, ws, s);
  } else {
    trsm_right_rec(T11, B1, Tinv, col0 + n1, false, cutoff, ws, s);
    addmul_views(B0, B1, T.sub(n1, 0, n, n1), cutoff, ws, s);
    trsm_right_rec(T00, B0, Tinv, col0, false, cutoff, ws, s);
  }
}

DView invert_diagonal_blocks(DView T, bool upper, Workspace &ws, cudaStream_t s) {
  int const m = T.nrows;
  DView Tinv = ws.alloc(m, kBaseRows);
  unsigned const blocks = (m + kBaseRows - 1) / kBaseRows;
  auto const *t = reinterpret_cast<uint32_t const *>(T.data);
  auto *inv = reinterpret_cast<uint32_t *>(Tinv.data);
  if (upper) tri_inv128_kernel<true><<<blocks, 32, 0, s>>>(t, T.pitch * 2, inv, m);
  else       tri_inv128_kernel<false><<<blocks, 32, 0, s>>>(t, T.pitch * 2, inv, m);
  M4B_CUDA(cudaGetLastError());
  ++g_kernel_launches;
  return Tinv;
}

}  // namespace

void trsm_right(DView T, DView B, bool upper, int cutoff, Workspace &ws, cudaStream_t s) {
  if (B.nrows <= 0 || B.ncols <= 0) return;
  size_t const mark = ws.mark();
  DView Tinv = invert_diagonal_blocks(T, upper, ws, s);
  trsm_right_rec(T, B, Tinv, 0, upper, cutoff, ws, s);
  ws.release(mark);
}

void trsm_left(DView T, DView B, bool upper, int cutoff, Workspace &ws, cudaStream_t s) {
  int const m = B.nrows, n = B.ncols;
  if (m <= 0 || n <= 0) return;
  size_t const mark = ws.mark();
  DView Tinv = invert_diagonal_blocks(T, upper, ws, s);
  trsm_rec(T, B, Tinv, 0, upper, cutoff, ws, s);
  ws.release(mark);
}

}  // namespace m4b
