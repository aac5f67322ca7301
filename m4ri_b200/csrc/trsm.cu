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
// triangular_russian.c:50-168) are replaced by one kernel: columns of B are independent, so each thread
// owns one 32-bit column word of all <= 128 rows (staged in shared memory, one bank per thread) and
// substitutes along the set bits of the triangular block.
#include "dev.h"
#include "workspace.h"

namespace m4b {
namespace {

constexpr int kBaseRows = 128;
constexpr int kBaseThreads = 64;

// T: rows [0, m) x cols [0, m) of the diagonal block (bit j of row i at t[i*tpitch32 + j/32]);
// B: m rows of nw32 32-bit words.
template <bool UPPER>
__global__ void __launch_bounds__(kBaseThreads) trsm_base_kernel(uint32_t const *__restrict__ t, long long tpitch32,
                                                                 uint32_t *__restrict__ b, long long bpitch32, int m,
                                                                 int nw32) {
  __shared__ uint32_t tri[kBaseRows][4];
  __shared__ uint32_t x[kBaseRows][kBaseThreads];
  int const tid = threadIdx.x;
  int const w = blockIdx.x * kBaseThreads + tid;
  for (int i = tid; i < kBaseRows * 4; i += kBaseThreads) {
    int const r = i >> 2, c = i & 3;
    tri[r][c] = (r < m && c * 32 < m) ? t[r * tpitch32 + c] : 0u;
  }
  if (w < nw32)
    for (int i = 0; i < m; ++i) x[i][tid] = b[i * bpitch32 + w];
  __syncthreads();
  if (w >= nw32) return;
  if (!UPPER) {
    for (int i = 1; i < m; ++i) {            // X_i = B_i + sum_{k < i, L[i][k]} X_k
      uint32_t acc = x[i][tid];
      for (int c = 0; c * 32 < i; ++c) {
        uint32_t bits = tri[i][c];
        if (i - c * 32 < 32) bits &= (1u << (i - c * 32)) - 1u;     // strictly below the diagonal
        while (bits) {
          int const k = c * 32 + __ffs(bits) - 1;
          bits &= bits - 1;
          acc ^= x[k][tid];
        }
      }
      x[i][tid] = acc;
    }
  } else {
    for (int i = m - 2; i >= 0; --i) {       // X_i = B_i + sum_{k > i, U[i][k]} X_k
      uint32_t acc = x[i][tid];
      for (int c = (i + 1) >> 5; c * 32 < m; ++c) {
        uint32_t bits = tri[i][c];
        if (c * 32 <= i) bits &= ~((2u << (i - c * 32)) - 1u);      // strictly above the diagonal
        if (m - c * 32 < 32) bits &= (1u << (m - c * 32)) - 1u;     // columns >= m do not exist
        while (bits) {
          int const k = c * 32 + __ffs(bits) - 1;
          bits &= bits - 1;
          acc ^= x[k][tid];
        }
      }
      x[i][tid] = acc;
    }
  }
  for (int i = 0; i < m; ++i) b[i * bpitch32 + w] = x[i][tid];
}

void launch_base(DView T, DView B, bool upper, cudaStream_t s) {
  int const nw32 = ((B.ncols + 127) / 128) * 4;
  unsigned const grid = (nw32 + kBaseThreads - 1) / kBaseThreads;
  auto const *t = reinterpret_cast<uint32_t const *>(T.data);
  auto *b = reinterpret_cast<uint32_t *>(B.data);
  if (upper)
    trsm_base_kernel<true><<<grid, kBaseThreads, 0, s>>>(t, T.pitch * 2, b, B.pitch * 2, B.nrows, nw32);
  else
    trsm_base_kernel<false><<<grid, kBaseThreads, 0, s>>>(t, T.pitch * 2, b, B.pitch * 2, B.nrows, nw32);
  M4B_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}

// C ^= A*B on views with the deepest Strassen recursion the views' alignment allows
void addmul_views(DView C, DView A, DView B, int cutoff, Workspace &ws, cudaStream_t s) {
  int levels = strassen_levels(A.nrows, A.ncols, B.ncols, cutoff);
  while (levels > 0 && (A.nrows % (1 << levels) || A.ncols % (128 << levels) || B.ncols % (128 << levels))) --levels;
  strassen_mul(C, A, B, levels, false, ws, s);
}

}  // namespace

size_t trsm_workspace_bytes(int m, int n, int cutoff) {
  // every update product is at most m x m x n; its Strassen temporaries bound all the smaller ones
  int const mp = (m + 127) / 128 * 128, np = (n + 127) / 128 * 128;
  return strassen_workspace_bytes(mp, mp, np, strassen_levels(m, m, n, cutoff));
}

void trsm_left(DView T, DView B, bool upper, int cutoff, Workspace &ws, cudaStream_t s) {
  int const m = B.nrows, n = B.ncols;
  if (m <= 0 || n <= 0) return;
  if (m <= kBaseRows) {
    launch_base(T, B, upper, s);
    return;
  }
  int const m1 = ((m + 127) / 128 / 2) * 128;          // multiple of 128, 0 < m1 < m
  DView const T00 = T.sub(0, 0, m1, m1), T11 = T.sub(m1, m1, m, m);
  DView const B0 = B.sub(0, 0, m1, n), B1 = B.sub(m1, 0, m, n);
  if (!upper) {
    trsm_left(T00, B0, false, cutoff, ws, s);
    addmul_views(B1, T.sub(m1, 0, m, m1), B0, cutoff, ws, s);
    trsm_left(T11, B1, false, cutoff, ws, s);
  } else {
    trsm_left(T11, B1, true, cutoff, ws, s);
    addmul_views(B0, T.sub(0, m1, m1, m), B1, cutoff, ws, s);
    trsm_left(T00, B0, true, cutoff, ws, s);
  }
}

}  // namespace m4b
