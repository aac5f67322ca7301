// dev.h — internal declarations shared by the CUDA kernels and the host scheduler.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../include/m4ri_b200.h"

namespace m4b {

// m4ri_die convention (m4ri/misc.c:36-42): message to stderr, then abort().
[[noreturn]] void die(char const *fmt, ...);

#define M4B_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      ::m4b::die("m4ri_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, \
                 __LINE__, cudaGetErrorString(e_));                                          \
  } while (0)

// A view of a device-resident bit-packed matrix.  Invariants: data 16-byte aligned,
// pitch (64-bit words) even, all bits in columns [ncols, 64*pitch) of every row that
// belong to THIS allocation are zero (sub-views of Strassen only ever split on
// 128-bit boundaries, so a view's "padding" is never someone else's data unless
// ncols % 128 == 0).
struct DView {
  word   *data;
  int64_t pitch;
  int     nrows;
  int     ncols;
  DView sub(int r0, int c0, int r1, int c1) const {  // c0 % 128 == 0
    return DView{data + (int64_t)r0 * pitch + c0 / 64, pitch, r1 - r0, c1 - c0};
  }
};

extern unsigned long long g_kernel_launches;  // counted by every launcher in this library

// ---- leaf: C ^= A*B (M4RM, stream-K persistent kernel) -------------------------------
// C must already hold the addend (zeros for a plain product).
void launch_m4rm(DView C, DView A, DView B, cudaStream_t stream);
// products of identical shape in ONE persistent launch: 7 (the last Strassen level) with either leaf, up to 49
// (the last TWO levels) where the tall-tile leaf suits — m4rm_batch_limit() says which
void launch_m4rm_batch(int count, DView const *C, DView const *A, DView const *B, cudaStream_t stream);
// the same as C = A*B: C need not be initialised (the tall-tile leaf stores the tiles a CTA owns alone and zero-fills
// only the products that hold stream-K tail tiles; the 1024-row leaf zero-fills everything first)
void launch_m4rm_batch_clear(int count, DView const *C, DView const *A, DView const *B, cudaStream_t stream);
int  m4rm_batch_limit(int m, int l, int n);
// C = A*B for l <= 128 with plain stores; C may alias A or B
void launch_m4rm_overwrite(DView C, DView A, DView B, cudaStream_t stream);
int  m4rm_num_sms();
// tall-tile leaf (m4rm_leaf2.cu): 4096 x 256-bit C tiles; worth it only when (nearly) all 4096 rows are real
bool leaf2_suits(int m, int l, int n);
void launch_m4rm_leaf2(int count, DView const *C, DView const *A, DView const *B, cudaStream_t stream, bool overwrite = false);
// 0 = automatic (leaf2 where it suits), 1 = always the 1024-row leaf, 2 = leaf2 whenever m >= 1
extern int g_leaf_variant;
extern int g_last_leaf;
void leaf_profile_begin();
unsigned long long leaf_profile_end(double *ms, double *bitops);
void leaf_profile_reset();   // drop the event pool (its events belong to one device)

// ---- element-wise helpers on views (all 128-bit vectorised) --------------------------
void launch_xor(DView C, DView A, DView B, cudaStream_t stream);        // C = A ^ B
void launch_zero(DView C, cudaStream_t stream);                          // C = 0
void launch_copy(DView C, DView A, cudaStream_t stream);                 // C = A
void launch_mask_excess(DView C, cudaStream_t stream);                   // clear bits >= ncols of last word(s)
void launch_transpose(DView dst, DView src, cudaStream_t stream);        // dst = src^T (transpose.cu)
// fused additions of one Strassen-Winograd node (quadrant order 11, 12, 21, 22)
void launch_winograd_pre_a(DView const a[4], DView const s_out[4], cudaStream_t stream);
void launch_winograd_pre_b(DView const b[4], DView const t_out[4], cudaStream_t stream);
void launch_winograd_post(DView const p[7], DView const c[4], bool accumulate, cudaStream_t stream);
// the same for up to 7 nodes of identical shape in one launch (arrays indexed [node * 4 + q] / [node * 7 + i])
void launch_winograd_pre_a_batch(int nodes, DView const *a, DView const *s_out, cudaStream_t stream);
void launch_winograd_pre_b_batch(int nodes, DView const *b, DView const *t_out, cudaStream_t stream);
void launch_winograd_post_batch(int nodes, DView const *p, DView const *c, cudaStream_t stream);
// two Winograd levels in one pass (elementwise.cu): side 0 = A, 1 = B; sub[4*q1+q2] = quadrant q2 of quadrant q1;
// sums[0..16) = level-1 sums by sub-block, sums[16+4*i+t] = level-2 sum t of level-1 operand i; prods[7*i+j]; csub[4*Q1+Q2]
void launch_winograd_pre2(int side, DView const *sub, DView const *sums, cudaStream_t stream);
void launch_winograd_post2(DView const *prods, DView const *csub, bool accumulate, cudaStream_t stream);

// ---- host <-> device transfers (capi.cu) ----------------------------------------------
class Stager;   // staging.h: pinned-ring transfers for pageable host memory (optional)
void upload(DView dst, mzd_t const *src, cudaStream_t s, Stager *st = nullptr);   // excess bits cleared on device
void download(mzd_t *dst, DView src, cudaStream_t s, std::vector<word> &tmp, Stager *st = nullptr);   // only valid bits of dst change
void zero_async(DView v, cudaStream_t s);

// ---- multi-GPU row-block product (multi.cu) --------------------------------------------
// GPUs base_device .. base_device + num_devices - 1; C cut into pr x pc blocks (multi_grid)
void multi_product(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff, bool clear, int num_devices, int base_device,
                   char *path_out, size_t path_len);
void multi_grid(int num_devices, int ncols, int *pr, int *pc);
void multi_release();

// ---- host scheduler ------------------------------------------------------------------
struct Workspace;  // bump allocator over one cached device slab
int  strassen_levels(int m, int k, int n, int cutoff);

// Optional callbacks of the TOP level of the schedule, used by the host path to overlap PCIe
// transfers with compute: need_*(q) is called before the first kernel that reads quadrant q
// (0 = 11, 1 = 12, 2 = 21, 3 = 22) of A / B / C-as-addend is enqueued, done_c(q) after the last
// kernel that writes quadrant q of C.
struct TopHooks {
  virtual void need_a(int) {}
  virtual void need_b(int) {}
  virtual void need_c(int) {}
  virtual void done_c(int) {}
  virtual ~TopHooks() {}
};
void strassen_mul(DView C, DView A, DView B, int levels, bool clear, Workspace &ws, cudaStream_t s,
                  TopHooks *hooks = nullptr);
// top level (levels >= 1) on explicit quadrants 11, 12, 21, 22, each possibly its own allocation
void strassen_mul_quads(DView const c[4], DView const a[4], DView const b[4], int levels, bool clear, Workspace &ws,
                        cudaStream_t s, TopHooks &hooks);

// ---- triangular solve, left variants (trsm.cu) -----------------------------------------------
// T: m x m (strict triangle read, unit diagonal implied), B: m x n, X overwrites B.
void   trsm_left(DView T, DView B, bool upper, int cutoff, Workspace &ws, cudaStream_t s);    // T X = B
void   trsm_right(DView T, DView B, bool upper, int cutoff, Workspace &ws, cudaStream_t s);   // X T = B
size_t trsm_workspace_bytes(int t, int m, int n, int cutoff);

// ---- reduced row echelon form (echelon.cu) ----------------------------------------------------
int    echelonize_device(DView A, Workspace &ws, cudaStream_t s);     // in place, returns the rank, synchronises s
size_t echelon_workspace_bytes(int m, int n, int64_t pitch_words = 0);   // pitch_words: A's pitch if above the minimal one

// ---- experimental tensor-core leaf (tc_leaf.cu): C (^)= A * Bt^T with tcgen05.mma kind::mxf4 on expanded operands ----
void launch_tc_leaf_simple(DView C, DView A, DView Bt, bool accumulate, cudaStream_t s);
void launch_tc_leaf2(DView C, DView A, DView B, cudaStream_t s);        // pipelined B-stationary form, C = A * B
bool tc_leaf_suits(int m, int l, int n);
void tc_scratch_release();
void launch_tc_batch(int count, DView const *C, DView const *A, DView const *B, cudaStream_t s, bool accumulate);   // <= 49 products

// ---- PLE decomposition (ple.cu) ----------------------------------------------------------------
// in place on a device-resident matrix; P (nrows ints) and Q (ncols ints) on the host; returns the rank, synchronises s
int    ple_device(DView A, int *P, int *Q, int cutoff, Workspace &ws, cudaStream_t s);
size_t ple_workspace_bytes(int m, int n, int cutoff);

}  // namespace m4b
