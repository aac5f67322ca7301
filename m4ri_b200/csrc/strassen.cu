// strassen.cu — host-side Strassen-Winograd scheduler over device views.
//
// Device counterpart of _mzd_mul_even / _mzd_addmul_even (m4ri/strassen.c:41-208, 367-526):
// the same Winograd/Bodrato operation sequences, but every "+" is one HBM-bound XOR kernel and
// every leaf product is one persistent M4RM launch, all enqueued on a single stream (the stream
// order is the dependency order, so the host never synchronises inside the recursion).
//
// Differences from the CPU scheduler, none of which can change a bit of the result:
//   * The caller pads the operands with zeros to multiples of 2^levels rows and 128*2^levels
//     columns (capi.cu), so every level halves exactly on 128-bit boundaries and there are no
//     edge strips (strassen.c:171-204) and no window copies (strassen.c:54-62).
//   * The recursion depth is fixed up front by strassen_levels(): the reference's leaf test
//     (strassen.c:39,51) restated by volume so that rectangular row-blocks still recurse.
//   * Temporaries come from a bump allocator over one cached device slab (no malloc per node).
#include "dev.h"
#include "workspace.h"

namespace m4b {

// Leaf test.  The reference stops when ANY dimension is "closer to cutoff than its half"
// (3*dim < 4*cutoff, strassen.c:39,51), i.e. its leaves keep every dimension >= 2/3*cutoff.  For
// square problems we reproduce exactly that depth.  A rectangular problem (the row-block a GPU owns in
// the multi-GPU split is m/G x l x n) would never recurse under that rule although each half-product
// is still big enough to run the persistent leaf at full efficiency, so the device rule is by VOLUME:
// split while the half-problem keeps at least (2/3*cutoff)^3 bit-triples and every dimension of it
// stays >= kMinLeafDim.  Depth never changes a result bit.
static constexpr int kMinLeafDim = 2048;
static constexpr int kTallRows   = 4096;   // rows of the tall-tile leaf: the recursion keeps m at or above it

int strassen_levels(int m, int k, int n, int cutoff) {
  double const min_volume = (2.0 / 3.0 * cutoff) * (2.0 / 3.0 * cutoff) * (2.0 / 3.0 * cutoff);
  int const min_dim = cutoff < kMinLeafDim ? (cutoff * 2 / 3 > 64 ? cutoff * 2 / 3 : 64) : kMinLeafDim;
  // with a production-size cutoff the leaves keep the 4096 rows the tall-tile kernel wants (a 2048-row leaf runs
  // on the 1024-row kernel at 0.88 of its rate and cannot be batched 49 to a launch)
  int const min_rows = cutoff >= kTallRows ? kTallRows : min_dim;
  int levels = 0;
  while (true) {
    int const m2 = (m + 1) / 2, k2 = (k + 1) / 2, n2 = (n + 1) / 2;
    if (m2 < min_rows || k2 < (min_dim < 128 ? 128 : min_dim) || n2 < (min_dim < 128 ? 128 : min_dim)) break;
    if ((double)m2 * k2 * n2 < min_volume) break;
    m = m2; k = k2; n = n2;
    ++levels;
  }
  // A single level is run as ONE batched launch of its seven half-size products, which stays efficient
  // one size class lower than a recursion whose leaves are launched one by one (8192^3: 0.338 ms vs 0.360).
  if (levels == 0 && cutoff >= 2 * kMinLeafDim) {
    int const m2 = (m + 1) / 2, k2 = (k + 1) / 2, n2 = (n + 1) / 2;
    if (m2 >= kMinLeafDim && k2 >= kMinLeafDim && n2 >= kMinLeafDim && (double)m2 * k2 * n2 >= min_volume / 8) levels = 1;
  }
  return levels;
}

// Workspace: every level is sized as a fused Winograd node (4 + 4 operand sums and 7 product
// temporaries); the in-place top level of the host path (3 quarter-size temporaries) needs less.
size_t strassen_workspace_bytes(int m, int k, int n, int levels) {
  if (levels <= 0) return 0;
  int const m2 = m / 2, k2 = k / 2, n2 = n / 2;
  size_t const sums = 4 * Workspace::bytes_for(m2, k2) + 4 * Workspace::bytes_for(k2, n2);
  // one level as its own node (winograd_node): 4 + 4 operand sums, 7 product temporaries, then one child at a time
  size_t const single = sums + Workspace::bytes_for(7 * m2, n2) + strassen_workspace_bytes(m2, k2, n2, levels - 1);
  if (levels < 2) return single;
  // two levels as one node (winograd_node2): the level-1 sums, the 28 + 28 level-2 sums and the 49 products live
  // together; below them one grandchild at a time (or, with M4RI_B200_NO_FUSED2, the seven level-1 results)
  int const m4 = m2 / 2, k4 = k2 / 2, n4 = n2 / 2;
  size_t below = strassen_workspace_bytes(m4, k4, n4, levels - 2);
  if (below < Workspace::bytes_for(7 * m2, n2)) below = Workspace::bytes_for(7 * m2, n2);
  size_t const pair = sums + 28 * Workspace::bytes_for(m4, k4) + 28 * Workspace::bytes_for(k4, n4) + Workspace::bytes_for(49 * m4, n4) + below;
  return single > pair ? single : pair;
}

static void quadrants(DView const &V, DView q[4]) {
  int const r2 = V.nrows / 2, c2 = V.ncols / 2;
  q[0] = V.sub(0, 0, r2, c2);
  q[1] = V.sub(0, c2, r2, 2 * c2);
  q[2] = V.sub(r2, 0, 2 * r2, c2);
  q[3] = V.sub(r2, c2, 2 * r2, 2 * c2);
}

// One Strassen-Winograd node below the top level, three kinds of launches only:
//   pre   S1..S4 / T1..T4 from the quadrants of A / B         (2 fused element-wise launches)
//   mul   P1..P7; when the children are leaves ALL SEVEN run in one persistent stream-K launch
//   post  C11 C12 C21 C22 from P1..P7                          (1 fused element-wise launch)
// Same products as strassen.c's sequence (Winograd's 7-product form), different bookkeeping: more
// temporaries (HBM is 180 GB), a third of the element-wise traffic and 3-4 launches instead of 29.
static void winograd_node2(DView C, DView A, DView B, int levels, bool clear, Workspace &ws, cudaStream_t s);

static void winograd_node(DView C, DView A, DView B, int levels, bool clear, Workspace &ws, cudaStream_t s) {
  if (levels == 0) {
    if (clear) launch_m4rm_batch_clear(1, &C, &A, &B, s);
    else       launch_m4rm(C, A, B, s);
    return;
  }
  // two levels at a time wherever the levels pair up: the bottom pair needs the 49-product launch of the tall-tile leaf,
  // the pairs above it only the fused two-level additions
  if (levels >= 2 && levels % 2 == 0 && !getenv("M4RI_B200_NO_NODE2") &&
      (levels > 2 ? getenv("M4RI_B200_NO_UPPER_PAIRS") == nullptr : m4rm_batch_limit(A.nrows / 4, A.ncols / 4, B.ncols / 4) >= 49)) {
    winograd_node2(C, A, B, levels, clear, ws, s);
    return;
  }
  DView a[4], b[4], c[4];
  quadrants(A, a);
  quadrants(B, b);
  quadrants(C, c);
  int const m2 = a[0].nrows, k2 = a[0].ncols, n2 = b[0].ncols;
  size_t const mark = ws.mark();
  DView S[4], T[4], P[7];
  for (int i = 0; i < 4; ++i) S[i] = ws.alloc(m2, k2);
  for (int i = 0; i < 4; ++i) T[i] = ws.alloc(k2, n2);
  DView const Pall = ws.alloc(7 * m2, n2);
  for (int i = 0; i < 7; ++i) P[i] = Pall.sub(i * m2, 0, (i + 1) * m2, n2);
  launch_winograd_pre_a(a, S, s);
  launch_winograd_pre_b(b, T, s);
  //                   P1     P2     P3     P4     P5     P6     P7
  DView const X[7] = {a[0], a[1], S[3], a[3], S[0], S[1], S[2]};
  DView const Y[7] = {b[0], b[2], b[3], T[3], T[0], T[1], T[2]};
  if (levels == 1) {
    launch_m4rm_batch_clear(7, P, X, Y, s);
  } else {
    for (int i = 0; i < 7; ++i) winograd_node(P[i], X[i], Y[i], levels - 1, true, ws, s);
  }
  launch_winograd_post(P, c, !clear, s);
  ws.release(mark);
}

// TWO levels as one node.  Additions: both levels of operand sums in one pass per side (16 sub-blocks in, the 4
// level-1 sums and the 28 level-2 sums out) and the 49 products straight into the 16 sub-blocks of C
// (elementwise.cu: winograd_pre2 / post2) — the level-1 operands are not re-read and the seven intermediate results
// are never written.  Bottom pair (levels == 2): 49 leaf products in ONE persistent launch — a leaf launch costs
// ~35 us of fill and drain whatever its size, which is 12 % of a 7 x 4096^3 launch but 2 % of a 49 x 4096^3 one; this
// is what makes one more Strassen level pay.  Pairs above it: the 49 products are Strassen products themselves.
static void winograd_node2(DView C, DView A, DView B, int levels, bool clear, Workspace &ws, cudaStream_t s) {
  static bool const fused = getenv("M4RI_B200_NO_FUSED2") == nullptr;   // the two-pass form, kept for A/B measurements
  DView a[4], b[4], c[4];
  quadrants(A, a);
  quadrants(B, b);
  quadrants(C, c);
  int const m2 = a[0].nrows, k2 = a[0].ncols, n2 = b[0].ncols, m4 = m2 / 2, k4 = k2 / 2, n4 = n2 / 2;
  size_t const mark = ws.mark();
  // ---- level 1 operands ----
  DView S[4], T[4], P1[7];
  for (int i = 0; i < 4; ++i) S[i] = ws.alloc(m2, k2);
  for (int i = 0; i < 4; ++i) T[i] = ws.alloc(k2, n2);
  DView const X1[7] = {a[0], a[1], S[3], a[3], S[0], S[1], S[2]};
  DView const Y1[7] = {b[0], b[2], b[3], T[3], T[0], T[1], T[2]};
  // ---- level 2 operands of all seven inner nodes ----
  DView xa[28], xs[28], yb[28], yt[28];
  for (int i = 0; i < 7; ++i) {
    quadrants(X1[i], xa + 4 * i);
    quadrants(Y1[i], yb + 4 * i);
    for (int q = 0; q < 4; ++q) xs[4 * i + q] = ws.alloc(m4, k4);
    for (int q = 0; q < 4; ++q) yt[4 * i + q] = ws.alloc(k4, n4);
  }
  if (fused) {
    // both levels of operand sums in ONE pass per side: 16 sub-blocks in, the 4 level-1 sums (16 sub-blocks) and the
    // 28 level-2 sums out
    DView asub[16], bsub[16], asum[44], bsum[44];
    for (int q1 = 0; q1 < 4; ++q1) {
      quadrants(a[q1], asub + 4 * q1);
      quadrants(b[q1], bsub + 4 * q1);
      quadrants(S[q1], asum + 4 * q1);
      quadrants(T[q1], bsum + 4 * q1);
    }
    for (int k = 0; k < 28; ++k) {
      asum[16 + k] = xs[k];
      bsum[16 + k] = yt[k];
    }
    launch_winograd_pre2(0, asub, asum, s);
    launch_winograd_pre2(1, bsub, bsum, s);
  } else {
    launch_winograd_pre_a(a, S, s);
    launch_winograd_pre_b(b, T, s);
    launch_winograd_pre_a_batch(7, xa, xs, s);
    launch_winograd_pre_b_batch(7, yb, yt, s);
  }
  DView const P2all = ws.alloc(49 * m4, n4);
  DView P2[49], X2[49], Y2[49];
  for (int i = 0; i < 7; ++i) {
    DView const *qa = xa + 4 * i, *qs = xs + 4 * i, *qb = yb + 4 * i, *qt = yt + 4 * i;
    DView const X[7] = {qa[0], qa[1], qs[3], qa[3], qs[0], qs[1], qs[2]};
    DView const Y[7] = {qb[0], qb[2], qb[3], qt[3], qt[0], qt[1], qt[2]};
    for (int j = 0; j < 7; ++j) {
      X2[7 * i + j] = X[j];
      Y2[7 * i + j] = Y[j];
      P2[7 * i + j] = P2all.sub((7 * i + j) * m4, 0, (7 * i + j + 1) * m4, n4);
    }
  }
  if (levels == 2) {
    launch_m4rm_batch_clear(49, P2, X2, Y2, s);      // P2 = products: no zero fill of the 49 temporaries
  } else {                                           // an upper pair: the 49 products are Strassen products themselves
    for (int i = 0; i < 49; ++i) winograd_node(P2[i], X2[i], Y2[i], levels - 2, true, ws, s);
  }
  if (fused) {
    DView csub[16];                                  // the 49 products straight into the 16 sub-blocks of C
    for (int q1 = 0; q1 < 4; ++q1) quadrants(c[q1], csub + 4 * q1);
    launch_winograd_post2(P2, csub, !clear, s);
  } else {
    DView const P1all = ws.alloc(7 * m2, n2);
    DView pq[28];
    for (int i = 0; i < 7; ++i) {
      P1[i] = P1all.sub(i * m2, 0, (i + 1) * m2, n2);
      quadrants(P1[i], pq + 4 * i);
    }
    launch_winograd_post_batch(7, P2, pq, s);          // P1[i] from its seven products, all i in one launch
    launch_winograd_post(P1, c, !clear, s);
  }
  ws.release(mark);
}

void strassen_mul(DView C, DView A, DView B, int levels, bool clear, Workspace &ws, cudaStream_t s, TopHooks *hooks) {
  if (C.nrows <= 0 || C.ncols <= 0) return;
  if (!hooks) {   // device-resident operands: no transfers to overlap, every level is a fused node
    winograd_node(C, A, B, levels, clear, ws, s);
    return;
  }
  TopHooks &hk = *hooks;
  if (levels == 0) {
    for (int q = 0; q < 4; ++q) { hk.need_a(q); hk.need_b(q); if (!clear) hk.need_c(q); }
    if (clear) launch_zero(C, s);
    launch_m4rm(C, A, B, s);
    for (int q = 0; q < 4; ++q) hk.done_c(q);
    return;
  }
  DView a[4], b[4], c[4];
  quadrants(A, a);
  quadrants(B, b);
  quadrants(C, c);
  strassen_mul_quads(c, a, b, levels, clear, ws, s, hk);
}

// The top level on explicit quadrants (order 11, 12, 21, 22; they need not be parts of one allocation: the
// multi-rank end-to-end path keeps every quadrant in its own contiguous buffer so that NCCL can all-gather it).
void strassen_mul_quads(DView const c[4], DView const a[4], DView const b[4], int levels, bool clear, Workspace &ws,
                        cudaStream_t s, TopHooks &hk) {
  int const m2 = a[0].nrows, k2 = a[0].ncols, n2 = b[0].ncols;
  DView const a11 = a[0], a12 = a[1], a21 = a[2], a22 = a[3];
  DView const b11 = b[0], b12 = b[1], b21 = b[2], b22 = b[3];
  DView const c11 = c[0], c12 = c[1], c21 = c[2], c22 = c[3];

  size_t const mark = ws.mark();
  int const lv = levels - 1;
  if (clear) {
    // C = A*B, 7 products, 15 adds (strassen.c:111-150)
    DView X = ws.alloc(m2, k2), Y = ws.alloc(k2, n2), P = ws.alloc(m2, n2);
    // A12*B21 only feeds C11 at the end and needs just two operand quadrants: doing it FIRST lets the
    // host path start computing after 2 of the 8 quadrant uploads (the reference computes it fifth).
    hk.need_a(1); hk.need_b(2);
    winograd_node(P, a12, b21, lv, true, ws, s);
    hk.need_b(3); hk.need_b(1);
    launch_xor(Y, b22, b12, s);
    hk.need_a(3);
    launch_xor(X, a22, a12, s);
    winograd_node(c21, X, Y, lv, true, ws, s);
    hk.need_a(2);
    launch_xor(X, a22, a21, s);
    launch_xor(Y, b22, b21, s);
    winograd_node(c22, X, Y, lv, true, ws, s);
    launch_xor(Y, Y, b12, s);
    launch_xor(X, X, a12, s);
    winograd_node(c11, X, Y, lv, true, ws, s);
    hk.need_a(0);
    launch_xor(X, X, a11, s);
    winograd_node(c12, X, b12, lv, true, ws, s);
    launch_xor(c12, c12, c22, s);
    launch_xor(c11, c11, P, s);
    launch_xor(c12, c11, c12, s);
    hk.done_c(1);
    launch_xor(c11, c21, c11, s);
    hk.need_b(0);
    launch_xor(Y, Y, b11, s);
    winograd_node(c21, a21, Y, lv, true, ws, s);
    launch_xor(c21, c11, c21, s);
    hk.done_c(2);
    launch_xor(c22, c22, c11, s);
    hk.done_c(3);
    winograd_node(c11, a11, b11, lv, true, ws, s);
    launch_xor(c11, c11, P, s);
    hk.done_c(0);
  } else {
    // C ^= A*B, 7 products, 14 adds (strassen.c:436-466)
    DView S = ws.alloc(m2, k2), T = ws.alloc(k2, n2), U = ws.alloc(m2, n2);
    hk.need_a(3); hk.need_a(2);
    launch_xor(S, a22, a21, s);
    hk.need_b(3); hk.need_b(2);
    launch_xor(T, b22, b21, s);
    winograd_node(U, S, T, lv, true, ws, s);
    hk.need_c(3);
    launch_xor(c22, U, c22, s);
    hk.need_c(1);
    launch_xor(c12, U, c12, s);
    hk.need_a(1);
    winograd_node(U, a12, b21, lv, true, ws, s);
    hk.need_c(0);
    launch_xor(c11, U, c11, s);
    hk.need_a(0); hk.need_b(0);
    winograd_node(c11, a11, b11, lv, false, ws, s);
    hk.done_c(0);
    launch_xor(S, S, a12, s);
    hk.need_b(1);
    launch_xor(T, T, b12, s);
    winograd_node(U, S, T, lv, false, ws, s);
    launch_xor(c12, c12, U, s);
    launch_xor(S, a11, S, s);
    winograd_node(c12, S, b12, lv, false, ws, s);
    hk.done_c(1);
    launch_xor(T, b11, T, s);
    hk.need_c(2);
    winograd_node(c21, a21, T, lv, false, ws, s);
    launch_xor(S, a22, a12, s);
    launch_xor(T, b22, b12, s);
    winograd_node(U, S, T, lv, false, ws, s);
    launch_xor(c21, c21, U, s);
    hk.done_c(2);
    launch_xor(c22, c22, U, s);
    hk.done_c(3);
  }
  ws.release(mark);
}

}  // namespace m4b
