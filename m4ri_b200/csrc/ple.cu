// ple.cu — PLE decomposition with the matrix resident in HBM (the L4 callers of the
// multiplication path; reference: _mzd_ple, m4ri/ple.c:62-178, base case _mzd_ple_naive :222-272,
// _mzd_compress_l, m4ri/mzp.c:294-397, mzd_apply_p_left, mzp.c:65-73).
//
// A = P L E in place: E (row echelon form) above, L (unit lower triangular, compressed into the first `rank`
// columns) below the pivots, P as LAPACK-style row transpositions, Q[i] = column of the i-th pivot.  The result
// of the reference — matrix, P, rank, Q[0..rank) — does not depend on its algorithm variant
// (tests/test_ple_reference_canonical.py) because every variant honours ONE pivot rule: leftmost column with a 1
// at or below the current row, FIRST such row in the current (already swapped) order.  This file reproduces that
// rule, so it is bit-exact against the reference:
//
//   * recursion over column halves as in the reference (first half, apply P to the second half, TRSM with L00,
//     Schur update A11 ^= A10 * A01, second half, apply P2 to A10, compress L) — the TRSM and the update are this
//     library's device-resident kernels (trsm.cu, the M4RM leaf), nothing leaves the GPU between the upload and the
//     download of A; split points are multiples of 128 columns (16-byte aligned device views) instead of 64, which
//     cannot change a bit of the result;
//   * base case = a strip of at most 128 columns and ALL remaining rows: one CTA walks the columns, elects the
//     first row with the bit (block-wide minimum over the current order), swaps it up and eliminates below —
//     every row's 128-bit strip word is updated by its own thread;
//   * the host only composes permutations (it needs the ranks to cut the windows anyway).
#include <cooperative_groups.h>
#include <string.h>
#include <time.h>

#include <unordered_map>
#include <vector>

#include "dev.h"
#include "workspace.h"

namespace m4b {
namespace {

constexpr int kStripCols   = 128;
constexpr int kStripThreads = 1024;

struct StripResult {
  int rank;
  int P[kStripCols];   // P[i] = row (relative to the window) swapped with row i
  int Q[kStripCols];   // Q[i] = column (relative to the window) of pivot i
};

struct U128 {
  unsigned long long lo, hi;
};

__device__ __forceinline__ U128 ld128(word const *p) {
  ulonglong2 const v = *reinterpret_cast<ulonglong2 const *>(p);
  return U128{v.x, v.y};
}
__device__ __forceinline__ void st128(word *p, U128 v) { *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2(v.lo, v.hi); }
__device__ __forceinline__ bool bit128(U128 v, int j) { return ((j < 64 ? v.lo >> j : v.hi >> (j - 64)) & 1ull) != 0; }
// bits j+1 .. 127
__device__ __forceinline__ U128 above(int j) {
  U128 m;
  if (j < 63)       { m.lo = ~0ull << (j + 1); m.hi = ~0ull; }
  else if (j == 63) { m.lo = 0;                m.hi = ~0ull; }
  else if (j < 127) { m.lo = 0;                m.hi = ~0ull << (j - 63); }
  else              { m.lo = 0;                m.hi = 0; }
  return m;
}
__device__ __forceinline__ U128 swap_bits(U128 v, int a, int b) {   // exchange bits a and b
  bool const x = bit128(v, a), y = bit128(v, b);
  if (x != y) {
    if (a < 64) v.lo ^= 1ull << a; else v.hi ^= 1ull << (a - 64);
    if (b < 64) v.lo ^= 1ull << b; else v.hi ^= 1ull << (b - 64);
  }
  return v;
}

// PLE of the strip: rows [0, nr) x columns [0, nc) of the window whose first word is `base` (16-byte aligned),
// rows `pitch` words apart.  One CTA.  _mzd_ple_naive (m4ri/ple.c:222-272) on the window, including the final
// compression of L (ple.c:260-266).
__global__ void __launch_bounds__(kStripThreads) ple_strip_kernel(word *base, long long pitch, int nr, int nc, StripResult *out) {
  __shared__ int s_min[kStripThreads / 32];
  __shared__ int s_pivot;
  __shared__ int s_Q[kStripCols];
  __shared__ unsigned long long s_pw[2];
  int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int rpos = 0;
  for (int j = 0; j < nc && rpos < nr; ++j) {
    // ---- first row >= rpos (current order) with bit j: rounds of 1024 consecutive rows ----
    int found = -1;
    for (int r0 = rpos; r0 < nr; r0 += kStripThreads) {
      int const i = r0 + tid;
      bool const hit = i < nr && bit128(ld128(base + (long long)i * pitch), j);
      unsigned const ballot = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) s_min[warp] = ballot ? r0 + warp * 32 + (__ffs(ballot) - 1) : 0x7fffffff;
      __syncthreads();
      if (tid < 32) {
        int v = s_min[tid];                         // kStripThreads / 32 == 32 warps
#pragma unroll
        for (int off = 16; off; off >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, off));
        if (tid == 0) s_pivot = v;
      }
      __syncthreads();
      found = s_pivot;
      if (found != 0x7fffffff) break;               // uniform: every thread reads the same value
      found = -1;
    }
    if (found < 0) continue;                        // no 1 in this column at or below rpos
    // ---- swap it up, publish the pivot row's strip word ----
    if (tid == 0) {
      word *a = base + (long long)rpos * pitch, *b = base + (long long)found * pitch;
      U128 const va = ld128(a), vb = ld128(b);
      st128(a, vb);
      st128(b, va);
      s_pw[0] = vb.lo;
      s_pw[1] = vb.hi;
      out->P[rpos] = found;
      out->Q[rpos] = j;
      s_Q[rpos] = j;
    }
    __syncthreads();
    // ---- eliminate below: the pivot row is added from column j + 1 on (the 1 in column j stays: it is L) ----
    U128 const m = above(j);
    unsigned long long const plo = s_pw[0] & m.lo, phi = s_pw[1] & m.hi;
    for (int l = rpos + 1 + tid; l < nr; l += kStripThreads) {
      word *p = base + (long long)l * pitch;
      U128 v = ld128(p);
      if (bit128(v, j)) {
        v.lo ^= plo;
        v.hi ^= phi;
        st128(p, v);
      }
    }
    ++rpos;
    __syncthreads();
  }
  // ---- compress L: for j < rank, columns Q[j] and j are exchanged in rows j.. (in this order) ----
  int const rank = rpos;
  for (int l = tid; l < nr; l += kStripThreads) {
    word *p = base + (long long)l * pitch;
    U128 v = ld128(p);
    int const upto = l < rank - 1 ? l : rank - 1;
    bool changed = false;
    for (int j = 0; j <= upto; ++j) {
      int const q = s_Q[j];
      if (q > j) { v = swap_bits(v, j, q); changed = true; }
    }
    if (changed) st128(p, v);
  }
  if (tid == 0) out->rank = rank;
}

// The same factorisation with the strip held in REGISTERS by a thread-block cluster: position p of the current row
// order lives in slot p % kSlots of thread (p / kSlots) % kClThreads of CTA p / (kSlots * kClThreads), so an
// elimination step is a few predicated XORs per thread instead of a sweep over global memory (the one-CTA kernel
// above spends ~20 us per pivot on 65536 rows: every thread walks 64 strided 16-byte words through L2).  Per pivot
// column: local first candidate -> warp / CTA minimum -> the owner of every CTA's candidate stores (position, word)
// into EVERY CTA's shared memory (DSMEM), the owner of position rpos its word -> ONE cluster barrier -> every thread
// picks the global first candidate, swaps and eliminates in registers.  Up to 8 x 512 x 16 = 65536 rows.
// Measured, 65536^2: 3306 ms with the one-CTA kernel, 486 ms with two barriers per column.
constexpr int kClThreads = 512, kSlots = 16, kClMax = 8;
constexpr int kInf = 0x7fffffff;

__global__ void __launch_bounds__(kClThreads, 1) ple_strip_cluster_kernel(word *base, long long pitch, int nr, int nc, StripResult *out) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  int const crank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
  __shared__ int s_warp[kClThreads / 32];
  struct Cand {
    unsigned long long lo, hi;
    int pos;
    int pad;
  };
  __shared__ Cand s_cand[2][kClMax];                 // [parity][CTA]: each CTA's first candidate and its strip word
  __shared__ unsigned long long s_rw[2][2];          // [parity]: the word at position rpos
  __shared__ int s_Q[kStripCols];
  int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int const p0 = (crank * kClThreads + tid) * kSlots;
  U128 w[kSlots];
#pragma unroll
  for (int k = 0; k < kSlots; ++k) w[k] = p0 + k < nr ? ld128(base + (long long)(p0 + k) * pitch) : U128{0ull, 0ull};
  int rpos = 0, par = 0;
  for (int j = 0; j < nc && rpos < nr; ++j) {
    // ---- ONE cluster barrier per column: every CTA publishes (first candidate, its word) to every CTA, the owner
    //      of position rpos publishes that word as well; after the barrier each thread knows pivot and both words ----
    int cand = kInf;
#pragma unroll
    for (int k = kSlots - 1; k >= 0; --k)
      if (p0 + k >= rpos && bit128(w[k], j)) cand = p0 + k;
    int const mine = cand;
#pragma unroll
    for (int off = 16; off; off >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, off));
    if (lane == 0) s_warp[warp] = cand;
    __syncthreads();
    int bm = kInf;
#pragma unroll
    for (int i = 0; i < kClThreads / 32; ++i) bm = min(bm, s_warp[i]);
    bool const own_r = rpos >= p0 && rpos < p0 + kSlots;
    if ((bm != kInf && mine == bm) || own_r || (bm == kInf && tid == 0)) {
      U128 vg{0ull, 0ull}, vr{0ull, 0ull};
#pragma unroll
      for (int k = 0; k < kSlots; ++k) {
        if (p0 + k == bm) vg = w[k];
        if (p0 + k == rpos) vr = w[k];
      }
      for (int r = 0; r < csize; ++r) {
        if (mine == bm || (bm == kInf && tid == 0)) {
          Cand *d = cluster.map_shared_rank(&s_cand[par][crank], r);
          d->lo = vg.lo;
          d->hi = vg.hi;
          d->pos = bm;
        }
        if (own_r) {
          unsigned long long *d = cluster.map_shared_rank(s_rw[par], r);
          d[0] = vr.lo;
          d[1] = vr.hi;
        }
      }
    }
    cluster.sync();
    int g = kInf, gi = 0;
    for (int r = 0; r < csize; ++r)
      if (s_cand[par][r].pos < g) { g = s_cand[par][r].pos; gi = r; }
    U128 const pw{s_cand[par][gi].lo, s_cand[par][gi].hi}, rw{s_rw[par][0], s_rw[par][1]};
    par ^= 1;
    if (g == kInf) continue;                         // no 1 in this column at or below rpos (every CTA agrees)
    U128 const m = above(j);
    unsigned long long const plo = pw.lo & m.lo, phi = pw.hi & m.hi;
#pragma unroll
    for (int k = 0; k < kSlots; ++k) {
      int const p = p0 + k;
      if (p == rpos) w[k] = pw;                      // the pivot row moves up ...
      else if (p == g) w[k] = rw;                    // ... the row it displaces takes its place
      if (p > rpos && bit128(w[k], j)) {
        w[k].lo ^= plo;
        w[k].hi ^= phi;
      }
    }
    if (tid == 0) {
      s_Q[rpos] = j;
      if (crank == 0) {
        out->P[rpos] = g;
        out->Q[rpos] = j;
      }
    }
    ++rpos;
  }
  __syncthreads();
  int const rank = rpos;
#pragma unroll
  for (int k = 0; k < kSlots; ++k) {
    int const p = p0 + k;
    if (p >= nr) continue;
    U128 v = w[k];
    int const upto = p < rank - 1 ? p : rank - 1;
    for (int jj = 0; jj <= upto; ++jj) {
      int const q = s_Q[jj];
      if (q > jj) v = swap_bits(v, jj, q);
    }
    st128(base + (long long)p * pitch, v);
  }
  if (crank == 0 && tid == 0) out->rank = rank;
  cluster.sync();                                    // no CTA leaves while a peer could still address its shared memory
}

// rows: T[k] = M[src[k]] (GATHER) / M[dst[k]] = T[k] (scatter) over the words [w0, w0 + nw) — two launches apply a
// row permutation to a column range
template <bool GATHER>
__global__ void __launch_bounds__(256) permute_rows_kernel(word *M, long long pitch, int w0, int nw, int const *rows, int nmoves,
                                                           word *T) {
  long long const total = (long long)nmoves * (nw / 2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int const k = (int)(i / (nw / 2)), c = (int)(i - (long long)k * (nw / 2)) * 2;
    word *m = M + (long long)rows[k] * pitch + w0 + c, *t = T + (long long)k * nw + c;
    if (GATHER) *reinterpret_cast<ulonglong2 *>(t) = *reinterpret_cast<ulonglong2 const *>(m);
    else        *reinterpret_cast<ulonglong2 *>(m) = *reinterpret_cast<ulonglong2 const *>(t);
  }
}

// _mzd_compress_l (m4ri/mzp.c:294-397) for the rows r1.. of a window whose first word is `base`: the L columns of
// the second factorisation move from columns n1.. to columns r1.. .  For row i (relative to the window) with
// len = min(i - r1 + 1, r2):  new[r1 .. r1+len) = old[n1 .. n1+len),  [r1+len, n1+len) = 0  — the closed form of the
// reference's sequence of column swaps (rows < r1 + r2) and of its block move (rows >= r1 + r2); both rely on
// columns [r1, n1) of these rows being zero after the first factorisation.  One thread per row, words ascending:
// a word is read (as source) before the same thread overwrites it, because sources lie to the right.
__global__ void __launch_bounds__(256) compress_l_kernel(word *base, long long pitch, int nr, int r1, int n1, int r2) {
  int const i = r1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nr) return;
  int const len = (i - r1 + 1 < r2) ? i - r1 + 1 : r2, d = n1 - r1;
  word *row = base + (long long)i * pitch;
  int const mv_end = r1 + len, end = n1 + len;          // move target [r1, mv_end), zeroed [mv_end, end)
  int const dw = d >> 6, db = d & 63;
  for (int w = r1 >> 6; w * 64 < end; ++w) {
    int const lo = w * 64, hi = lo + 64;
    // mask of the bits of this word inside [r1, end) and inside [r1, mv_end)
    auto range = [&](int a, int b) -> word {
      int const x = a > lo ? a : lo, y = b < hi ? b : hi;
      if (y <= x) return 0;
      word mk = ~(word)0 >> (64 - (y - x));
      return mk << (x - lo);
    };
    word const in_range = range(r1, end), in_move = range(r1, mv_end);
    word src = 0;
    if (in_move) {                                      // bit p of src = old[p + d]
      src = row[w + dw] >> db;
      int const p_max = (mv_end < hi ? mv_end : hi) - 1;          // highest moved bit of this word
      if (db && ((p_max + d) >> 6) == w + dw + 1) src |= row[w + dw + 1] << (64 - db);
    }
    row[w] = (row[w] & ~in_range) | (src & in_move);
  }
}

// M4RI_B200_PLE_PROFILE=1: wall time per phase (the stream is synchronised after each, so the total grows a little)
struct PleProfile {
  bool   on = false;
  double t[6] = {};     // strip, permute, trsm, update, compress, sync
  static double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
  }
};
PleProfile g_ple_prof;
struct PhaseTimer {
  int phase;
  cudaStream_t s;
  double t0 = 0;
  PhaseTimer(int ph, cudaStream_t st) : phase(ph), s(st) { if (g_ple_prof.on) t0 = PleProfile::now(); }
  ~PhaseTimer() {
    if (!g_ple_prof.on) return;
    cudaStreamSynchronize(s);
    g_ple_prof.t[phase] += PleProfile::now() - t0;
  }
};

struct PleCtx {
  DView        M;           // the whole device matrix
  Workspace   *ws;
  cudaStream_t s;
  int          cutoff;
  StripResult *d_res;       // device
  StripResult *h_res;       // pinned host
  int         *d_rows;      // device scratch for row lists
  size_t       rows_cap;    // ints
};

// C ^= A*B on views; Strassen only as deep as the views' alignment allows (trsm.cu does the same)
void update(DView C, DView A, DView B, PleCtx &cx) {
  if (C.nrows <= 0 || C.ncols <= 0 || A.ncols <= 0) return;
  int levels = strassen_levels(A.nrows, A.ncols, B.ncols, cx.cutoff);
  while (levels > 0 && (A.nrows % (1 << levels) || A.ncols % (128 << levels) || B.ncols % (128 << levels))) --levels;
  strassen_mul(C, A, B, levels, false, *cx.ws, cx.s);
}

// mzd_apply_p_left (m4ri/mzp.c:65-73) on rows r0.. of the column range [c_lo, c_hi) (both multiples of 128, or the
// padded right edge): the transpositions P[0..np) in order, composed on the host into one gather / scatter
void apply_p_left(PleCtx &cx, int r0, int c_lo, int c_hi, int const *P, int np) {
  if (np <= 0 || c_hi <= c_lo) return;
  PhaseTimer pt(1, cx.s);
  std::unordered_map<int, int> cur;                    // position -> row whose content sits there now
  auto at = [&](int pos) { auto it = cur.find(pos); return it == cur.end() ? pos : it->second; };
  for (int i = 0; i < np; ++i) {
    if (P[i] == i) continue;
    int const a = at(i), b = at(P[i]);
    cur[i] = b;
    cur[P[i]] = a;
  }
  std::vector<int> rows;                               // [src..., dst...]
  std::vector<int> dst;
  for (auto const &kv : cur)
    if (kv.first != kv.second) { dst.push_back(r0 + kv.first); rows.push_back(r0 + kv.second); }
  int const nm = (int)dst.size();
  if (!nm) return;
  rows.insert(rows.end(), dst.begin(), dst.end());
  if ((size_t)2 * nm > cx.rows_cap) die("m4ri_b200: ple row list overflow\n");
  M4B_CUDA(cudaMemcpyAsync(cx.d_rows, rows.data(), sizeof(int) * 2 * nm, cudaMemcpyHostToDevice, cx.s));
  M4B_CUDA(cudaStreamSynchronize(cx.s));               // `rows` is a pageable temporary
  int const w_lo = c_lo / 64, w_hi = (int)(((long long)c_hi + 127) / 128 * 2);
  // the scratch for the gathered rows is bounded: process the column range in slabs
  int const slab = 1024;                               // words
  size_t const mark = cx.ws->mark();
  int const tw = w_hi - w_lo < slab ? w_hi - w_lo : slab;
  DView T = cx.ws->alloc(nm, tw * 64);
  for (int w0 = w_lo; w0 < w_hi; w0 += slab) {
    int const nw = w_hi - w0 < slab ? w_hi - w0 : slab;
    long long const total = (long long)nm * (nw / 2);
    unsigned const blocks = (unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    permute_rows_kernel<true><<<blocks, 256, 0, cx.s>>>(cx.M.data, cx.M.pitch, w0, nw, cx.d_rows, nm, T.data);
    permute_rows_kernel<false><<<blocks, 256, 0, cx.s>>>(cx.M.data, cx.M.pitch, w0, nw, cx.d_rows + nm, nm, T.data);
    g_kernel_launches += 2;
  }
  M4B_CUDA(cudaGetLastError());
  cx.ws->release(mark);
}

// window rows [r0, r0 + nr) x columns [c0, c0 + nc) of cx.M, c0 % 128 == 0; P (nr entries) and Q (nc entries) relative
int ple_rec(PleCtx &cx, int r0, int c0, int nr, int nc, int *P, int *Q) {
  for (int i = 0; i < nr; ++i) P[i] = i;
  for (int i = 0; i < nc; ++i) Q[i] = i;
  if (nr <= 0 || nc <= 0) return 0;
  if (nc <= kStripCols) {
    PhaseTimer pt(0, cx.s);
    word *base = cx.M.data + (long long)r0 * cx.M.pitch + c0 / 64;
    static int const strip_variant = [] {            // M4RI_B200_PLE_STRIP=0: always the one-CTA global-memory kernel
      char const *env = getenv("M4RI_B200_PLE_STRIP");
      return env && env[0] == '0' ? 0 : 1;
    }();
    int const per_cta = kClThreads * kSlots;
    if (strip_variant == 1 && nr <= kClMax * per_cta) {
      int cl = 1;
      while (cl * per_cta < nr) cl *= 2;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)cl);
      cfg.blockDim = dim3(kClThreads);
      cfg.stream = cx.s;
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = (unsigned)cl;
      attr.val.clusterDim.y = attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      M4B_CUDA(cudaLaunchKernelEx(&cfg, ple_strip_cluster_kernel, base, (long long)cx.M.pitch, nr, nc, cx.d_res));
    } else {
      ple_strip_kernel<<<1, kStripThreads, 0, cx.s>>>(base, cx.M.pitch, nr, nc, cx.d_res);
    }
    M4B_CUDA(cudaGetLastError());
    ++g_kernel_launches;
    M4B_CUDA(cudaMemcpyAsync(cx.h_res, cx.d_res, sizeof(StripResult), cudaMemcpyDeviceToHost, cx.s));
    M4B_CUDA(cudaStreamSynchronize(cx.s));
    int const r = cx.h_res->rank;
    for (int i = 0; i < r; ++i) { P[i] = cx.h_res->P[i]; Q[i] = cx.h_res->Q[i]; }
    return r;
  }
  int const n1 = ((nc + 127) / 128 / 2) * 128;           // multiple of 128, 0 < n1 < nc
  int const r1 = ple_rec(cx, r0, c0, nr, n1, P, Q);
  DView const &M = cx.M;
  int const c_end = c0 + nc;
  if (r1) {
    apply_p_left(cx, r0, c0 + n1, c_end, P, r1);          // mzd_apply_p_left(A1, P1)
    DView const A00 = M.sub(r0, c0, r0 + r1, c0 + r1), A01 = M.sub(r0, c0 + n1, r0 + r1, c_end);
    {
      PhaseTimer pt(2, cx.s);
      trsm_left(A00, A01, false, cx.cutoff, *cx.ws, cx.s);  // _mzd_trsm_lower_left(A00, A01)
    }
    if (nr > r1) {                                        // mzd_addmul(A11, A10, A01)
      PhaseTimer pt(3, cx.s);
      update(M.sub(r0 + r1, c0 + n1, r0 + nr, c_end), M.sub(r0 + r1, c0, r0 + nr, c0 + r1), A01, cx);
    }
  }
  int *P2 = P + r1, *Q2 = Q + n1;
  int const r2 = ple_rec(cx, r0 + r1, c0 + n1, nr - r1, nc - n1, P2, Q2);
  // mzd_apply_p_left(A10, P2): columns [r1, n1) of these rows are zero, so the whole aligned range may move
  if (r1 && r2) apply_p_left(cx, r0 + r1, c0, c0 + n1, P2, r2);
  for (int i = 0; i < nr - r1; ++i) P2[i] += r1;
  for (int i = 0; i < nc - n1; ++i) Q2[i] += n1;
  for (int i = n1, j = r1; i < n1 + r2; ++i, ++j) Q[j] = Q[i];
  if (r1 != n1 && r2 > 0 && nr > r1) {                    // _mzd_compress_l(A, r1, n1, r2)
    PhaseTimer pt(4, cx.s);
    word *base = M.data + (long long)r0 * M.pitch + c0 / 64;
    unsigned const blocks = (unsigned)((nr - r1 + 255) / 256);
    compress_l_kernel<<<blocks, 256, 0, cx.s>>>(base, M.pitch, nr, r1, n1, r2);
    M4B_CUDA(cudaGetLastError());
    ++g_kernel_launches;
  }
  return r1 + r2;
}

}  // namespace

size_t ple_workspace_bytes(int m, int n, int cutoff) {
  int const half = (n + 1) / 2 + 128;
  size_t const slab_cols = (size_t)(n < 1024 * 64 ? (n + 127) / 128 * 128 : 1024 * 64);
  return trsm_workspace_bytes(m < half ? m : half, m, n, cutoff) + strassen_workspace_bytes((m + 127) / 128 * 128, (n + 127) / 128 * 128,
                                                                                           (n + 127) / 128 * 128, strassen_levels(m, n, n, cutoff)) +
         Workspace::bytes_for(2 * m < n + 256 ? 2 * m : n + 256, (int)slab_cols) + (size_t)2 * m * sizeof(int) + sizeof(StripResult) +
         8192;
}

// A (device resident) -> its PLE decomposition in place; P[0..A.nrows), Q[0..A.ncols) on the host; returns the rank.
int ple_device(DView A, int *P, int *Q, int cutoff, Workspace &ws, cudaStream_t s) {
  int const m = A.nrows, n = A.ncols;
  if (m <= 0 || n <= 0) {
    for (int i = 0; i < m; ++i) P[i] = i;
    for (int i = 0; i < n; ++i) Q[i] = i;
    return 0;
  }
  size_t const mark = ws.mark();
  PleCtx cx;
  cx.M = A;
  cx.ws = &ws;
  cx.s = s;
  cx.cutoff = cutoff;
  cx.rows_cap = (size_t)2 * m + 64;
  cx.d_rows = reinterpret_cast<int *>(ws.alloc(1, (int)((cx.rows_cap * sizeof(int) + 255) / 256 * 256 * 8)).data);
  cx.d_res = reinterpret_cast<StripResult *>(ws.alloc(1, (int)((sizeof(StripResult) + 255) / 256 * 256 * 8)).data);
  M4B_CUDA(cudaMallocHost(reinterpret_cast<void **>(&cx.h_res), sizeof(StripResult)));
  g_ple_prof.on = getenv("M4RI_B200_PLE_PROFILE") != nullptr;
  for (double &t : g_ple_prof.t) t = 0;
  int const rank = ple_rec(cx, 0, 0, m, n, P, Q);
  M4B_CUDA(cudaStreamSynchronize(s));
  if (g_ple_prof.on)
    fprintf(stderr, "m4ri_b200 ple %d x %d: strip %.1f ms, permute %.1f ms, trsm %.1f ms, update %.1f ms, compress %.1f ms\n", m, n,
            g_ple_prof.t[0] * 1e3, g_ple_prof.t[1] * 1e3, g_ple_prof.t[2] * 1e3, g_ple_prof.t[3] * 1e3, g_ple_prof.t[4] * 1e3);
  cudaFreeHost(cx.h_res);
  ws.release(mark);
  return rank;
}

}  // namespace m4b
