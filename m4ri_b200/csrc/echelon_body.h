// echelon_body.h — device bodies of the reduced row echelon form (RREF) of a bit-packed matrix.
//
// Widening of the multiplication path into its callers: the job of the reference's
// mzd_echelonize_m4ri(A, full = 1, k) (m4ri/brilliantrussian.c:603-967, _mzd_gauss_submatrix_full :213-330,
// mzd_process_rows* :332-601).  The RREF of a matrix is unique, so any correct algorithm yields the
// reference's bits (the reference's own tests/test_elimination.c compares its M4RI, naive and PLUQ variants
// exactly that way); this one is organised for the GPU, not taken from the reference:
//
//   for every 64-column strip s (one word per row), with r = rank so far:
//     1. select_chunk   (grid = row chunks of [r, m))  every CTA finds a row basis (<= 64 rows) of ITS rows'
//                       strip words by forward elimination on the words alone: 64 column steps, each an
//                       election of any row with that bit (shared-memory atomicMin) + one XOR per row
//     2. select_final   (one CTA) the same selection over the <= 64 candidates of every chunk -> p pivot
//                       rows I and pivot columns Q of the strip; then the p x p inverse of S[I, Q] (64 threads)
//                       -> G (64 x 64): row q = combination of the pivot rows that is the reduced pivot row
//                       of strip column q (zero row if q is no pivot column); destinations, row moves
//     3. gather_pivots  PIV (64 x n') = the pivot rows;   Bm = G * PIV  (leaf kernel, l = 64)
//     4. build_x        X (m x 64) = A[:, strip] & Q for every row that is not a pivot row, 0 for pivot rows
//     5. A[:, c0:] ^= X * Bm      — the M4RM leaf kernel, the same one mzd_mul uses; this is where the time goes
//     6. move_rows / place_pivots: rows displaced from [r, r + p) go to the vacated pivot slots, the reduced
//        pivot rows (Bm) go to rows r .. r + p - 1 in pivot-column order;  advance: r += p
//
// No step needs the host: r, p, Q, the pivot rows and the moves live in a device State, every launch has a
// fixed shape.  "Any row with the bit" instead of the reference's "first row" is what makes step 1 parallel
// over chunks; the result is the same because the RREF does not depend on the pivot rows chosen.
//
// Like m4rm_leaf2_body.h this header is compiled by nvcc (echelon.cu) and by g++ (tests/c/emu_echelon.cpp:
// a CTA is a set of host threads), so the whole algorithm is checked on the CPU against a plain Gauss-Jordan elimination.
// The includer provides ECH_FN and a context type with: tid, ntid, sync(), atomic_min(int *, int).
#pragma once
#include <stdint.h>

namespace ech {

typedef unsigned long long u64;

constexpr int kSelThreads = 512;                       // threads of a selection CTA
constexpr int kRPT        = 16;                        // rows per thread
constexpr int kSelRows    = kSelThreads * kRPT;        // 8192: rows per chunk / candidates of the final CTA
constexpr int kNone       = 0x7fffffff;

struct State {                    // device resident, one per echelonize call
  int rank;                       // pivots found so far = row of the next pivot
  int p;                          // pivots of the current strip
  int nmove;
  int pad;
  u64 qmask;                      // pivot columns of the current strip
  int piv_row[64];                // t-th selected pivot row (global row index), t < p
  int piv_col[64];                // its strip column
  int col_dest[64];               // strip column q -> row that receives its reduced pivot row, -1 if none
  int move_src[64];               // displaced rows (in [rank, rank + p), not pivots) ...
  int move_dst[64];               // ... and the vacated pivot slots (>= rank + p) they move to
};

struct SelShared {                // shared memory of a selection CTA
  int win[64];                    // elected thread per strip column
  int win2[64];                   // elected row per pivot column (inversion)
  u64 word;                       // current word of the elected row
  int count;
  int sel_idx[64];                // source index of the t-th selected row
  int sel_col[64];
  u64 mrow[64];                   // inversion: current strip word / combination of the t-th pivot
  u64 comb[64];
  int asg[64];                    // 0: not assigned yet, q + 1: the row is the pivot row of strip column q
  int inpiv[64];
};

// ---- row sources of the selection -------------------------------------------------------------------------
struct MatrixRows {               // rows [row0, row0 + count) of the matrix, strip word wcol
  u64 const *A;
  long long pitch;
  int wcol, row0;
  ECH_FN int row(int idx) const { return row0 + idx; }
  ECH_FN u64 word(int idx) const { return A[(long long)(row0 + idx) * pitch + wcol]; }
};
struct CandidateRows {            // the candidates the chunks left behind (row < 0: empty slot, word 0)
  int const *rows;
  u64 const *words;
  ECH_FN int row(int idx) const { return rows[idx]; }
  ECH_FN u64 word(int idx) const { return words[idx]; }
};

// Row basis of the strip words src.word(0 .. count-1), count <= kSelRows: returns p and leaves the selected
// source indices / strip columns in sh->sel_idx / sel_col (selection order = increasing column).
template <class Ctx, class Src>
ECH_FN int select_basis(Ctx &cx, SelShared *sh, Src const &src, int count) {
  int const tid = cx.tid;
  u64 cur[kRPT];
  unsigned taken = 0;
#pragma unroll
  for (int k = 0; k < kRPT; ++k) {
    int const idx = k * kSelThreads + tid;
    cur[k] = idx < count ? src.word(idx) : 0ull;
  }
  if (tid < 64) sh->win[tid] = kNone;
  if (tid == 0) sh->count = 0;
  cx.sync();
  for (int q = 0; q < 64; ++q) {
    int kk = -1;
    u64 mine = 0;
#pragma unroll
    for (int k = 0; k < kRPT; ++k)
      if (kk < 0 && !((taken >> k) & 1u) && ((cur[k] >> q) & 1ull)) {
        kk = k;
        mine = cur[k];
      }
    if (kk >= 0) cx.atomic_min(&sh->win[q], tid);
    cx.sync();
    int const w = sh->win[q];
    if (w == tid) {                               // this thread's row kk is the pivot of column q
      int const t = sh->count;
      sh->word = mine;
      sh->sel_idx[t] = kk * kSelThreads + tid;
      sh->sel_col[t] = q;
      sh->count = t + 1;
      taken |= 1u << kk;
    }
    cx.sync();
    if (w != kNone) {
      u64 const ws = sh->word;
#pragma unroll
      for (int k = 0; k < kRPT; ++k)
        if (!((taken >> k) & 1u) && ((cur[k] >> q) & 1ull)) cur[k] ^= ws;
    }
  }
  return sh->count;                               // written before the last sync of the loop
}

// Step 1: one CTA per chunk of `chunk_rows` rows of [rank, m).
template <class Ctx>
ECH_FN void select_chunk(Ctx &cx, SelShared *sh, State const *st, u64 const *A, long long pitch, int m, int wcol,
                         int chunk_rows, int chunk, int *cand_row, u64 *cand_word) {
  long long const first = (long long)st->rank + (long long)chunk * chunk_rows;
  int count = 0;
  if (first < m) count = (m - first) < chunk_rows ? (int)(m - first) : chunk_rows;
  MatrixRows const src{A, pitch, wcol, (int)(first < m ? first : 0)};
  int const p = select_basis(cx, sh, src, count);
  if (cx.tid < 64) {
    int const t = cx.tid;
    cand_row[chunk * 64 + t]  = t < p ? src.row(sh->sel_idx[t]) : -1;
    cand_word[chunk * 64 + t] = t < p ? src.word(sh->sel_idx[t]) : 0ull;     // the ORIGINAL word
  }
}

// Step 2: the strip's pivots among the candidates, the inverse of their pivot-column block, destinations, moves.
// Gm = 64 x 64 bit matrix with a pitch of 2 words.
template <class Ctx>
ECH_FN void select_final(Ctx &cx, SelShared *sh, State *st, int const *cand_row, u64 const *cand_word, int ncand,
                         u64 *Gm) {
  int const tid = cx.tid;
  CandidateRows const src{cand_row, cand_word};
  int const p = select_basis(cx, sh, src, ncand);
  int const rank = st->rank;
  if (tid < 64) {
    sh->win2[tid] = kNone;
    sh->asg[tid] = 0;
    sh->inpiv[tid] = 0;
    sh->mrow[tid] = tid < p ? src.word(sh->sel_idx[tid]) : 0ull;
    sh->comb[tid] = tid < p ? 1ull << tid : 0ull;
    st->piv_row[tid] = tid < p ? src.row(sh->sel_idx[tid]) : -1;
    st->piv_col[tid] = tid < p ? sh->sel_col[tid] : -1;
    st->col_dest[tid] = -1;
    Gm[2 * tid] = 0;
    Gm[2 * tid + 1] = 0;
  }
  cx.sync();
  u64 qmask = 0;
  for (int t = 0; t < p; ++t) qmask |= 1ull << sh->sel_col[t];
  // Gauss-Jordan on [M | I], M = original strip words of the pivot rows: thread t owns pivot row t.  After the
  // step of pivot column q the elected row is e_q on the pivot columns and its comb says which pivot rows add
  // up to the reduced pivot row of q.
  for (int q = 0; q < 64; ++q) {
    if (!((qmask >> q) & 1ull)) continue;                      // uniform
    bool const has = tid < p && !sh->asg[tid] && ((sh->mrow[tid] >> q) & 1ull);
    if (has) cx.atomic_min(&sh->win2[q], tid);
    cx.sync();
    int const L = sh->win2[q];                                 // exists: the block is invertible
    u64 const lrow = sh->mrow[L], lcomb = sh->comb[L];
    cx.sync();
    if (tid == L) {
      sh->asg[tid] = q + 1;
      int below = 0;
      for (int b = 0; b < q; ++b) below += (int)((qmask >> b) & 1ull);
      st->col_dest[q] = rank + below;
    } else if (tid < p && ((sh->mrow[tid] >> q) & 1ull)) {
      sh->mrow[tid] ^= lrow;
      sh->comb[tid] ^= lcomb;
    }
    cx.sync();
  }
  // an assigned row keeps changing until the last pivot column has been cleared from it: its combination is final now
  if (tid < p) Gm[2 * (sh->asg[tid] - 1)] = sh->comb[tid];
  if (tid == 0) {
    // rows displaced from [rank, rank + p) -> pivot slots outside it (as many of one kind as of the other)
    for (int t = 0; t < p; ++t)
      if (st->piv_row[t] < rank + p) sh->inpiv[st->piv_row[t] - rank] = 1;
    int nm = 0, j = 0;
    for (int t = 0; t < p; ++t)
      if (st->piv_row[t] >= rank + p) {
        while (sh->inpiv[j]) ++j;
        st->move_src[nm] = rank + j;
        st->move_dst[nm] = st->piv_row[t];
        ++nm;
        ++j;
      }
    st->p = p;
    st->nmove = nm;
    st->qmask = qmask;
  }
}

// ---- element-wise passes: (gtid, gthreads) = global thread index / count, grid-stride ---------------------

// PIV (64 x nw words) = the pivot rows' words w0 .. w0 + nw - 1 in selection order, zero rows for t >= p
ECH_FN void gather_pivots(State const *st, u64 const *A, long long pitchA, int w0, int nw, u64 *PIV, long long pitchP,
                          long long gtid, long long gthreads) {
  for (long long e = gtid; e < 64ll * nw; e += gthreads) {
    int const t = (int)(e / nw), w = (int)(e % nw);
    PIV[t * pitchP + w] = t < st->p ? A[(long long)st->piv_row[t] * pitchA + w0 + w] : 0ull;
  }
}

// X (m x 64, pitch 2 words): a row's strip bits at the pivot columns; pivot rows get 0 (they are replaced, not updated)
ECH_FN void build_x(State const *st, u64 const *A, long long pitchA, int wcol, int m, u64 *X, long long gtid,
                    long long gthreads) {
  int const p = st->p;
  u64 const qmask = st->qmask;
  for (long long i = gtid; i < m; i += gthreads) {
    u64 x = A[i * pitchA + wcol] & qmask;
    for (int t = 0; t < p; ++t)
      if (st->piv_row[t] == (int)i) x = 0;
    X[2 * i] = x;
    X[2 * i + 1] = 0;
  }
}

ECH_FN void move_rows(State const *st, u64 *A, long long pitchA, int w0, int nw, long long gtid, long long gthreads) {
  for (long long e = gtid; e < (long long)st->nmove * nw; e += gthreads) {
    int const k = (int)(e / nw), w = (int)(e % nw);
    A[(long long)st->move_dst[k] * pitchA + w0 + w] = A[(long long)st->move_src[k] * pitchA + w0 + w];
  }
}

// rows rank .. rank + p - 1 = the reduced pivot rows (rows of Bm that belong to pivot columns), in column order
ECH_FN void place_pivots(State const *st, u64 *A, long long pitchA, int w0, int nw, u64 const *Bm, long long pitchB,
                         long long gtid, long long gthreads) {
  for (long long e = gtid; e < 64ll * nw; e += gthreads) {
    int const q = (int)(e / nw), w = (int)(e % nw);
    int const d = st->col_dest[q];
    if (d >= 0) A[(long long)d * pitchA + w0 + w] = Bm[q * pitchB + w];
  }
}

ECH_FN void advance(State *st) { st->rank += st->p; }

}  // namespace ech
