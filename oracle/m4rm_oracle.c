/* oracle/m4rm_oracle.c — TEST INFRASTRUCTURE ONLY (see m4rm_oracle.h).
 *
 * CPU restatement, in plain scalar C, of the reference's dense GF(2) multiply:
 *   container / bit layout      m4ri/mzd.h:68-99,185-187,440-443 ; mzd.c:142-177
 *   random fill                 m4ri/misc.c:58-71 ; mzd.c:1270-1280
 *   Gray code book              m4ri/graycode.c:31-50
 *   table of 2^k combinations   m4ri/brilliantrussian.c:163-211
 *   M4RM multiply (8 tables)    m4ri/brilliantrussian.c:1032-1190
 *   naive small-shape multiply  m4ri/mzd.c:1141-1268 (semantics only)
 *   Strassen-Winograd           m4ri/strassen.c:41-208 (mul), 367-526 (addmul),
 *                               345-365 / 675-700 (entry points)
 * Written from the behavioural description, not copied: no SSE2, no OpenMP, no
 * allocator caches.  Results are pinned bit-for-bit against the compiled
 * reference by tests/test_oracle.py.
 */
#define _DEFAULT_SOURCE
#include "m4rm_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define RADIX 64
#define FLAG_EXCESS 0x2   /* m4ri/mzd.h:144 */
#define FLAG_WINDOW 0x4   /* m4ri/mzd.h:150 */
#define ORC_L2 2097152    /* cache sizes only steer k / cutoff, never the bits */
#define ORC_L3 314572800L

static orc_word left_mask(int nbits) { /* low `nbits` bits set; 0 -> all ones (misc.h:272) */
  return ~(orc_word)0 >> ((RADIX - nbits) % RADIX);
}

static inline orc_word *rowp(orc_mzd const *M, orc_rci r) { return M->data + M->rowstride * (orc_wi)r; }

static void fill_header(orc_mzd *M, orc_rci r, orc_rci c) {
  memset(M, 0, sizeof *M);
  M->nrows = r;
  M->ncols = c;
  M->width = c > 0 ? (c + RADIX - 1) / RADIX : 0;
  M->high_bitmask = left_mask(c % RADIX);
  if (c % RADIX) M->flags |= FLAG_EXCESS;
}

orc_mzd *orc_init(orc_rci r, orc_rci c) {
  orc_mzd *M = malloc(sizeof *M);
  fill_header(M, r, c);
  M->rowstride = M->width + (M->width & 1); /* even number of words per row (mzd.c:148) */
  if (r && c) {
    size_t bytes = (size_t)r * M->rowstride * sizeof(orc_word);
    if (posix_memalign((void **)&M->data, 64, bytes)) abort();
    memset(M->data, 0, bytes);
  }
  return M;
}

orc_mzd *orc_init_window(orc_mzd *P, orc_rci lowr, orc_rci lowc, orc_rci highr, orc_rci highc) {
  orc_mzd *W = malloc(sizeof *W);
  orc_rci nr = highr - lowr;
  if (P->nrows - lowr < nr) nr = P->nrows - lowr;
  fill_header(W, nr, highc - lowc);
  W->flags |= FLAG_WINDOW;
  W->rowstride = P->rowstride;
  W->data = P->data + (orc_wi)lowr * P->rowstride + lowc / RADIX; /* lowc % 64 == 0 */
  return W;
}

void orc_free(orc_mzd *M) {
  if (!M) return;
  if (!(M->flags & FLAG_WINDOW)) free(M->data);
  free(M);
}

/* stack window: same as orc_init_window without the heap header */
static orc_mzd view(orc_mzd const *P, orc_rci r0, orc_rci c0, orc_rci r1, orc_rci c1) {
  orc_mzd W;
  fill_header(&W, r1 - r0, c1 - c0);
  W.flags |= FLAG_WINDOW;
  W.rowstride = P->rowstride;
  W.data = P->data + (orc_wi)r0 * P->rowstride + c0 / RADIX;
  return W;
}

orc_word orc_random_word(void) { /* three 31-bit draws (misc.c:65-69) */
  orc_word a0 = (orc_word)random();
  orc_word a1 = (orc_word)random();
  orc_word a2 = (orc_word)random();
  return a0 ^ (a1 << 24) ^ (a2 << 48);
}

void orc_randomize(orc_mzd *M) { /* row-major, one draw per word, last word merged under mask */
  if (!M->width) return;
  for (orc_rci i = 0; i < M->nrows; ++i) {
    orc_word *row = rowp(M, i);
    for (orc_wi j = 0; j + 1 < M->width; ++j) row[j] = orc_random_word();
    orc_word last = row[M->width - 1];
    row[M->width - 1] = last ^ ((last ^ orc_random_word()) & M->high_bitmask);
  }
}

int orc_equal(orc_mzd const *A, orc_mzd const *B) { /* valid bits only (mzd.c:1314-1331) */
  if (A->nrows != B->nrows || A->ncols != B->ncols) return 0;
  if (!A->width) return 1;
  for (orc_rci i = 0; i < A->nrows; ++i) {
    orc_word const *a = rowp(A, i), *b = rowp(B, i);
    for (orc_wi j = 0; j + 1 < A->width; ++j)
      if (a[j] != b[j]) return 0;
    if ((a[A->width - 1] ^ b[A->width - 1]) & A->high_bitmask) return 0;
  }
  return 1;
}

/* dst(valid bits) = f(valid bits); bits of dst outside the valid area are kept */
static inline void put_last(orc_word *dst, orc_word v, orc_word mask) { *dst = (*dst & ~mask) | (v & mask); }

void orc_copy(orc_mzd *D, orc_mzd const *S) {
  for (orc_rci i = 0; i < S->nrows; ++i) {
    orc_word *d = rowp(D, i);
    orc_word const *s = rowp(S, i);
    for (orc_wi j = 0; j + 1 < S->width; ++j) d[j] = s[j];
    if (S->width) put_last(d + S->width - 1, s[S->width - 1], D->high_bitmask);
  }
}

static void clear_valid(orc_mzd *C) { /* mzd_set_ui(C,0): mzd.c:1294-1301 */
  for (orc_rci i = 0; i < C->nrows; ++i) {
    orc_word *c = rowp(C, i);
    for (orc_wi j = 0; j + 1 < C->width; ++j) c[j] = 0;
    if (C->width) c[C->width - 1] &= ~C->high_bitmask;
  }
}

void orc_add(orc_mzd *C, orc_mzd const *A, orc_mzd const *B) { /* _mzd_add: mzd.c:1471-1583 */
  for (orc_rci i = 0; i < C->nrows; ++i) {
    orc_word *c = rowp(C, i);
    orc_word const *a = rowp(A, i), *b = rowp(B, i);
    for (orc_wi j = 0; j + 1 < C->width; ++j) c[j] = a[j] ^ b[j];
    if (C->width) put_last(c + C->width - 1, a[C->width - 1] ^ b[C->width - 1], C->high_bitmask);
  }
}

/* ---- Gray code book (graycode.c:31-50) ------------------------------------ */

int orc_gray_code(int number, int length) {
  /* walk bits from the top; each output bit = input bit XOR the input bit above it */
  int above = 0, out = 0;
  for (int i = length - 1; i >= 0; --i) {
    int bit = number & (1 << i);
    out |= (above >> 1) ^ bit;
    above = bit;
  }
  return out;
}

void orc_build_code(int *ord, int *inc, int l) {
  int n = 1 << l;
  for (int i = 0; i < n; ++i) ord[i] = orc_gray_code(i, l);
  /* inc[i] = which of the l generator rows flips between ord[i] and ord[i+1]
   * (finest level last so it wins: every second slot flips row l-1, ...). */
  for (int lvl = l; lvl > 0; --lvl) {
    int step = 1 << (l - lvl);
    for (int j = 1; j <= (1 << lvl); ++j) inc[j * step - 1] = l - lvl;
  }
}

static int *codebook_ord[17], *codebook_inc[17];
static void need_code(int k) {
  if (codebook_ord[k]) return;
  codebook_ord[k] = calloc((size_t)1 << k, sizeof(int));
  codebook_inc[k] = calloc((size_t)1 << k, sizeof(int));
  orc_build_code(codebook_ord[k], codebook_inc[k], k);
}

/* T[i] = T[i-1] ^ M[r + inc[i-1]] in Gray order, L[ord[i]] = i; every table row
 * is masked to M's valid columns (brilliantrussian.c:163-211, column offset 0). */
void orc_make_table(orc_mzd const *M, orc_rci r, int k, orc_mzd *T, orc_rci *L) {
  need_code(k);
  int const *ord = codebook_ord[k], *inc = codebook_inc[k];
  orc_wi const wide = M->width;
  orc_word const mask_end = left_mask(M->ncols % RADIX);
  L[0] = 0;
  memset(rowp(T, 0), 0, (size_t)wide * sizeof(orc_word));
  for (orc_rci i = 1; i < (1 << k); ++i) {
    orc_rci src = r + inc[i - 1];
    L[ord[i]] = i;
    if (src >= M->nrows) continue;
    orc_word *t = rowp(T, i);
    orc_word const *p = rowp(T, i - 1), *m = rowp(M, src);
    for (orc_wi j = 0; j < wide; ++j) t[j] = p[j] ^ m[j];
    t[wide - 1] &= mask_end;
  }
}

/* n <= 64 bits of row x starting at column y, column y in bit 0 (mzd.h:892-901) */
static inline orc_word read_bits(orc_mzd const *M, orc_rci x, orc_rci y, int n) {
  orc_word const *row = rowp(M, x);
  int spot = y % RADIX;
  orc_wi blk = y / RADIX;
  orc_word v = row[blk] >> spot;
  if (spot + n > RADIX) v |= row[blk + 1] << (RADIX - spot);
  return n == RADIX ? v : v & (((orc_word)1 << n) - 1);
}

/* ---- naive multiply: C[i] (^)= XOR of B rows selected by the bits of A[i] --- */

orc_mzd *orc_mul_naive(orc_mzd *C, orc_mzd const *A, orc_mzd const *B, int clear) {
  if (!C) C = orc_init(A->nrows, B->ncols);
  if (clear) clear_valid(C);
  if (!C->width) return C;
  orc_word const mask_end = C->high_bitmask;
  for (orc_rci i = 0; i < A->nrows; ++i) {
    orc_word *c = rowp(C, i);
    for (orc_rci j = 0; j < A->ncols; ++j) {
      if (!((rowp(A, i)[j / RADIX] >> (j % RADIX)) & 1)) continue;
      orc_word const *b = rowp(B, j);
      for (orc_wi w = 0; w + 1 < C->width; ++w) c[w] ^= b[w];
      c[C->width - 1] ^= b[C->width - 1] & mask_end;
    }
  }
  return C;
}

/* ---- M4RM (brilliantrussian.c:1032-1190) ----------------------------------- */

#define NTABLES 8
#define MUL_BLOCKSIZE 2048 /* MIN(sqrt(4*L3)/2, 2048), mzd.h:59 */

static int floor_log2(int v) { int r = 0; while (v >>= 1) ++r; return r; }

static int auto_k(orc_mzd const *A, orc_mzd const *B) { /* brilliantrussian.c:1075-1089 */
  int k = (int)log2((ORC_L2 / 64) / (double)B->width);
  if ((ORC_L2 - 64.0 * (1 << k) * B->width) > (64.0 * (1 << (k + 1)) * B->width - ORC_L2)) k++;
  int mind = A->nrows < A->ncols ? A->nrows : A->ncols;
  if (B->ncols < mind) mind = B->ncols;
  int klog = (int)round(0.75 * floor_log2(mind));
  return klog < k ? klog : k;
}

/* one pass: C[rows r0..r1) ^= sum over nt tables of `bits` columns of A starting at col */
static void m4rm_pass(orc_mzd *C, orc_mzd const *A, orc_mzd const *B, orc_rci col, int bits, int nt,
                      orc_rci r0, orc_rci r1, orc_mzd **T, orc_rci **L) {
  orc_wi const wide = C->width;
  orc_word const bm = ((orc_word)1 << bits) - 1;
  for (int z = 0; z < nt; ++z) orc_make_table(B, col + bits * z, bits, T[z], L[z]);
  for (orc_rci j = r0; j < r1; ++j) {
    orc_word a = read_bits(A, j, col, bits * nt);
    orc_word *c = rowp(C, j);
    for (int z = 0; z < nt; ++z) {
      orc_word const *t = rowp(T[z], L[z][(a >> (z * bits)) & bm]);
      for (orc_wi w = 0; w < wide; ++w) c[w] ^= t[w]; /* table rows carry zero excess */
    }
  }
}

orc_mzd *orc_mul_m4rm(orc_mzd *C, orc_mzd const *A, orc_mzd const *B, int k, int clear) {
  if (!C) C = orc_init(A->nrows, B->ncols);
  orc_rci const m = A->nrows, l = A->ncols, n = B->ncols;
  if (n < RADIX - 10 || m < 16) return orc_mul_naive(C, A, B, clear); /* :1063-1068 */
  if (clear) clear_valid(C);
  if (k == 0) k = auto_k(A, B);
  if (k < 2) k = 2;
  if (k > 8) k = 8;

  orc_mzd *T[NTABLES];
  orc_rci *L[NTABLES];
  for (int z = 0; z < NTABLES; ++z) {
    T[z] = orc_init(1 << k, n);
    L[z] = malloc(sizeof(orc_rci) << k);
  }
  int const kk = NTABLES * k;
  orc_rci const full = l / kk;
  for (orc_rci g = 0; g < m; g += MUL_BLOCKSIZE) { /* giant step over row blocks */
    orc_rci gend = g + MUL_BLOCKSIZE < m ? g + MUL_BLOCKSIZE : m;
    for (orc_rci i = 0; i < full; ++i) m4rm_pass(C, A, B, kk * i, k, NTABLES, g, gend, T, L);
  }
  orc_rci col = full * kk;
  for (; col + k <= l; col += k) m4rm_pass(C, A, B, col, k, 1, 0, m, T, L); /* whole k-blocks */
  if (col < l) m4rm_pass(C, A, B, col, l - col, 1, 0, m, T, L);            /* leftover bits  */

  for (int z = 0; z < NTABLES; ++z) { orc_free(T[z]); free(L[z]); }
  return C;
}

/* ---- Strassen-Winograd (strassen.c) ---------------------------------------- */

static int closer(orc_rci a, int cutoff) { return 3 * a < 4 * cutoff; } /* strassen.c:39 */

static void split_sizes(orc_rci m, orc_rci k, orc_rci n, int cutoff, orc_rci *mm, orc_rci *kk, orc_rci *nn) {
  /* strassen.c:71-80: halves that stay word-aligned at every deeper level */
  orc_rci mult = RADIX, w = m < n ? m : n;
  if (k < w) w = k;
  w /= 2;
  while (w > cutoff) { w /= 2; mult *= 2; }
  *mm = (((m - m % mult) / RADIX) >> 1) * RADIX;
  *kk = (((k - k % mult) / RADIX) >> 1) * RADIX;
  *nn = (((n - n % mult) / RADIX) >> 1) * RADIX;
}

static void mul_even(orc_mzd *C, orc_mzd const *A, orc_mzd const *B, int cutoff);
static void addmul_even(orc_mzd *C, orc_mzd const *A, orc_mzd const *B, int cutoff);

static void leaf(orc_mzd *C, orc_mzd const *A, orc_mzd const *B, int clear) {
  /* strassen.c:54-65 / 379-392: windows are copied to dense temporaries first */
  if ((A->flags | B->flags | C->flags) & FLAG_WINDOW) {
    orc_mzd *a = orc_init(A->nrows, A->ncols), *b = orc_init(B->nrows, B->ncols), *c = orc_init(C->nrows, C->ncols);
    orc_copy(a, A);
    orc_copy(b, B);
    if (!clear) orc_copy(c, C);
    orc_mul_m4rm(c, a, b, 0, 0);
    orc_copy(C, c);
    orc_free(a); orc_free(b); orc_free(c);
  } else {
    orc_mul_m4rm(C, A, B, 0, clear);
  }
}

/* edge strips shared by mul and addmul (strassen.c:171-204 / 489-522) */
static void strips(orc_mzd *C, orc_mzd const *A, orc_mzd const *B, orc_rci mm2, orc_rci kk2, orc_rci nn2, int clear) {
  orc_rci m = A->nrows, k = A->ncols, n = B->ncols;
  if (n > nn2) { /* last columns of C: all of A times the last columns of B */
    orc_mzd Bc = view(B, 0, nn2, k, n), Cc = view(C, 0, nn2, m, n);
    orc_mul_m4rm(&Cc, A, &Bc, 0, clear);
  }
  if (m > mm2) { /* last rows of C (left part): last rows of A times the first columns of B */
    orc_mzd Ar = view(A, mm2, 0, m, k), Bc = view(B, 0, 0, k, nn2), Cr = view(C, mm2, 0, m, nn2);
    orc_mul_m4rm(&Cr, &Ar, &Bc, 0, clear);
  }
  if (k > kk2) { /* inner-dimension remainder, added onto the bulk */
    orc_mzd Ac = view(A, 0, kk2, mm2, k), Br = view(B, kk2, 0, k, nn2), Cb = view(C, 0, 0, mm2, nn2);
    orc_mul_m4rm(&Cb, &Ac, &Br, 0, 0);
  }
}

#define QUADS(M, P, r, c)                                                                   \
  orc_mzd P##11 = view(M, 0, 0, r, c), P##12 = view(M, 0, c, r, 2 * (c)),                   \
          P##21 = view(M, r, 0, 2 * (r), c), P##22 = view(M, r, c, 2 * (r), 2 * (c))

static void mul_even(orc_mzd *C, orc_mzd const *A, orc_mzd const *B, int cutoff) {
  orc_rci m = A->nrows, k = A->ncols, n = B->ncols, mm, kk, nn;
  if (C->nrows == 0 || C->ncols == 0) return;
  if (closer(m, cutoff) || closer(k, cutoff) || closer(n, cutoff)) { leaf(C, A, B, 1); return; }
  split_sizes(m, k, n, cutoff, &mm, &kk, &nn);
  {
    QUADS(A, a, mm, kk); QUADS(B, b, kk, nn); QUADS(C, c, mm, nn);
    orc_mzd *X = orc_init(mm, kk), *Y = orc_init(kk, nn), *P;
    /* Winograd form as sequenced by Bodrato (strassen.c:111-150) */
    orc_add(Y, &b22, &b12);
    orc_add(X, &a22, &a12);
    mul_even(&c21, X, Y, cutoff);
    orc_add(X, &a22, &a21);
    orc_add(Y, &b22, &b21);
    mul_even(&c22, X, Y, cutoff);
    orc_add(Y, Y, &b12);
    orc_add(X, X, &a12);
    mul_even(&c11, X, Y, cutoff);
    orc_add(X, X, &a11);
    mul_even(&c12, X, &b12, cutoff);
    orc_add(&c12, &c12, &c22);
    orc_free(X);
    P = orc_mul(NULL, &a12, &b21, cutoff);
    orc_add(&c11, &c11, P);
    orc_add(&c12, &c11, &c12);
    orc_add(&c11, &c21, &c11);
    orc_add(Y, Y, &b11);
    mul_even(&c21, &a21, Y, cutoff);
    orc_free(Y);
    orc_add(&c21, &c11, &c21);
    orc_add(&c22, &c22, &c11);
    mul_even(&c11, &a11, &b11, cutoff);
    orc_add(&c11, &c11, P);
    orc_free(P);
  }
  strips(C, A, B, 2 * mm, 2 * kk, 2 * nn, 1);
}

static void addmul_even(orc_mzd *C, orc_mzd const *A, orc_mzd const *B, int cutoff) {
  orc_rci m = A->nrows, k = A->ncols, n = B->ncols, mm, kk, nn;
  if (C->nrows == 0 || C->ncols == 0) return;
  if (closer(m, cutoff) || closer(k, cutoff) || closer(n, cutoff)) { leaf(C, A, B, 0); return; }
  split_sizes(m, k, n, cutoff, &mm, &kk, &nn);
  {
    QUADS(A, a, mm, kk); QUADS(B, b, kk, nn); QUADS(C, c, mm, nn);
    orc_mzd *S = orc_init(mm, kk), *T = orc_init(kk, nn), *U = orc_init(mm, nn);
    /* accumulate schedule (strassen.c:436-466) */
    orc_add(S, &a22, &a21);
    orc_add(T, &b22, &b21);
    mul_even(U, S, T, cutoff);
    orc_add(&c22, U, &c22);
    orc_add(&c12, U, &c12);
    mul_even(U, &a12, &b21, cutoff);
    orc_add(&c11, U, &c11);
    addmul_even(&c11, &a11, &b11, cutoff);
    orc_add(S, S, &a12);
    orc_add(T, T, &b12);
    addmul_even(U, S, T, cutoff);
    orc_add(&c12, &c12, U);
    orc_add(S, &a11, S);
    addmul_even(&c12, S, &b12, cutoff);
    orc_add(T, &b11, T);
    addmul_even(&c21, &a21, T, cutoff);
    orc_add(S, &a22, &a12);
    orc_add(T, &b22, &b12);
    addmul_even(U, S, T, cutoff);
    orc_add(&c21, &c21, U);
    orc_add(&c22, &c22, U);
    orc_free(S); orc_free(T); orc_free(U);
  }
  strips(C, A, B, 2 * mm, 2 * kk, 2 * nn, 0);
}

static int norm_cutoff(int cutoff) { /* strassen.c:349-354, strassen.h:133-135 */
  if (cutoff == 0) {
    int c = (int)sqrt((double)(4 * ORC_L3));
    cutoff = c < 4096 ? c : 4096;
  }
  cutoff = cutoff / RADIX * RADIX;
  return cutoff < RADIX ? RADIX : cutoff;
}

static void die(char const *msg) { fputs(msg, stderr); abort(); } /* m4ri_die: misc.c:36-42 */

orc_mzd *orc_mul(orc_mzd *C, orc_mzd const *A, orc_mzd const *B, int cutoff) {
  if (A->ncols != B->nrows) die("orc_mul: A ncols need to match B nrows.\n");
  if (cutoff < 0) die("orc_mul: cutoff must be >= 0.\n");
  cutoff = norm_cutoff(cutoff);
  if (!C) C = orc_init(A->nrows, B->ncols);
  else if (C->nrows != A->nrows || C->ncols != B->ncols) die("orc_mul: C has wrong dimensions.\n");
  mul_even(C, A, B, cutoff); /* A==B squaring path (strassen.c:210-343) gives the same bits */
  return C;
}

orc_mzd *orc_addmul(orc_mzd *C, orc_mzd const *A, orc_mzd const *B, int cutoff) {
  if (A->ncols != B->nrows) die("orc_addmul: A ncols need to match B nrows.\n");
  if (cutoff < 0) die("orc_addmul: cutoff must be >= 0.\n");
  cutoff = norm_cutoff(cutoff);
  if (!C) C = orc_init(A->nrows, B->ncols);
  else if (C->nrows != A->nrows || C->ncols != B->ncols) die("orc_addmul: C has wrong dimensions.\n");
  if (A->nrows == 0 || A->ncols == 0 || B->ncols == 0) return C;
  addmul_even(C, A, B, cutoff);
  return C;
}

/* ---- triangular solves with matrices, left variants (first "next" row) ----------
 * m4ri/triangular.c:406-455 (lower left) and :467-516 (upper left): L X = B resp. U X = B, X
 * overwrites B.  The triangular matrix is read with an implied unit diagonal and only its strict
 * triangle is used (triangular.c:413-425).  The reference recurses with mzd_addmul and a
 * "russian" base case; the result is the unique solution, restated here as plain substitution. */

static void xor_row_valid(orc_mzd *B, orc_rci dst, orc_rci src) {
  orc_word *d = rowp(B, dst);
  orc_word const *s = rowp(B, src);
  for (orc_wi j = 0; j + 1 < B->width; ++j) d[j] ^= s[j];
  if (B->width) d[B->width - 1] ^= s[B->width - 1] & B->high_bitmask;
}

void orc_trsm_lower_left(orc_mzd const *L, orc_mzd *B) {
  for (orc_rci i = 1; i < B->nrows; ++i)
    for (orc_rci k = 0; k < i; ++k)
      if ((rowp(L, i)[k / RADIX] >> (k % RADIX)) & 1) xor_row_valid(B, i, k);
}

void orc_trsm_upper_left(orc_mzd const *U, orc_mzd *B) {
  for (orc_rci i = B->nrows - 2; i >= 0; --i)
    for (orc_rci k = i + 1; k < B->nrows; ++k)
      if ((rowp(U, i)[k / RADIX] >> (k % RADIX)) & 1) xor_row_valid(B, i, k);
}

/* right variants, X U = B / X L = B (triangular.c:29-148, 300-392): once column k of X is final, every
 * row of B whose bit k is set receives row k of the triangular matrix restricted to its strict
 * triangle (columns > k for U, ascending k; columns < k for L, descending k). */
static void xor_tri_row(orc_mzd *B, orc_rci i, orc_mzd const *T, orc_rci k, int upper) {
  orc_word *d = rowp(B, i);
  orc_word const *t = rowp(T, k);
  orc_rci const n = B->ncols;
  for (orc_wi w = 0; w < B->width; ++w) {
    orc_word mask = ~(orc_word)0;
    orc_rci const lo = (orc_rci)(w * RADIX);              /* first column of this word */
    if (upper) {                                           /* keep columns > k */
      if (lo + RADIX - 1 <= k) mask = 0;
      else if (lo <= k) mask = ~(orc_word)0 << (k - lo) << 1;
    } else {                                               /* keep columns < k */
      if (lo >= k) mask = 0;
      else if (lo + RADIX > k) mask = ((orc_word)1 << (k - lo)) - 1;
    }
    if (lo + RADIX > n) mask &= left_mask(n % RADIX);      /* columns >= n do not exist */
    d[w] ^= t[w] & mask;
  }
}

void orc_trsm_upper_right(orc_mzd const *U, orc_mzd *B) {
  for (orc_rci k = 0; k < B->ncols; ++k)
    for (orc_rci i = 0; i < B->nrows; ++i)
      if ((rowp(B, i)[k / RADIX] >> (k % RADIX)) & 1) xor_tri_row(B, i, U, k, 1);
}

void orc_trsm_lower_right(orc_mzd const *L, orc_mzd *B) {
  for (orc_rci k = B->ncols - 1; k >= 0; --k)
    for (orc_rci i = 0; i < B->nrows; ++i)
      if ((rowp(B, i)[k / RADIX] >> (k % RADIX)) & 1) xor_tri_row(B, i, L, k, 0);
}

/* ---- transpose (m4ri/mzd.c:1118-1139): DST[j][i] = A[i][j], only DST's valid bits are written ---- */
orc_mzd *orc_transpose(orc_mzd *D, orc_mzd const *A) {
  if (!D) D = orc_init(A->ncols, A->nrows);
  for (orc_rci j = 0; j < A->ncols; ++j) {
    orc_word *d = rowp(D, j);
    for (orc_rci i = 0; i < A->nrows; ++i) {
      orc_word const bit = (rowp(A, i)[j / RADIX] >> (j % RADIX)) & 1;
      d[i / RADIX] = (d[i / RADIX] & ~((orc_word)1 << (i % RADIX))) | (bit << (i % RADIX));
    }
  }
  return D;
}

/* ---- echelon forms (m4ri/mzd.c:208-233 mzd_gauss_delayed / mzd_echelonize_naive) -------------------------
 * Column by column: the first row at or below `startrow` with a 1 in the column is swapped up and added to
 * every other row that has a 1 there (full != 0: rows above as well -> reduced row echelon form, which is
 * unique, so every correct algorithm — the reference's M4RI and PLUQ variants, tests/test_elimination.c —
 * yields the same bits; full == 0: only the rows below).  Returns the rank. */
orc_rci orc_echelonize(orc_mzd *M, int full) {
  orc_rci startrow = 0, pivots = 0;
  for (orc_rci i = 0; i < M->ncols; ++i) {
    for (orc_rci j = startrow; j < M->nrows; ++j) {
      if ((rowp(M, j)[i / RADIX] >> (i % RADIX)) & 1) {
        if (j != startrow) {                       /* mzd_row_swap: valid words only */
          orc_word *a = rowp(M, startrow), *b = rowp(M, j);
          for (orc_wi w = 0; w < M->width; ++w) {
            orc_word const mask = (w == M->width - 1) ? M->high_bitmask : ~(orc_word)0;
            orc_word const t = (a[w] ^ b[w]) & mask;
            a[w] ^= t;
            b[w] ^= t;
          }
        }
        ++pivots;
        orc_word const *src = rowp(M, startrow);
        for (orc_rci ii = full ? 0 : startrow + 1; ii < M->nrows; ++ii) {
          if (ii == startrow) continue;
          orc_word *dst = rowp(M, ii);
          if ((dst[i / RADIX] >> (i % RADIX)) & 1) {   /* mzd_row_add_offset: from the word of column i on */
            for (orc_wi w = i / RADIX; w < M->width; ++w) {
              orc_word const mask = (w == M->width - 1) ? M->high_bitmask : ~(orc_word)0;
              dst[w] ^= src[w] & mask;
            }
          }
        }
        ++startrow;
        break;
      }
    }
  }
  return pivots;
}

/* ---- PLE decomposition (m4ri/ple.c:222-272 _mzd_ple_naive) ------------------------------------------------
 * A = P L E in place: leftmost column with a 1 at or below row_pos, FIRST row with it (in the current, already
 * swapped order) becomes the pivot: P[row_pos] = that row, Q[row_pos] = the column, rows swapped, and the pivot
 * row is added to every lower row with a 1 in the column FROM THE NEXT COLUMN ON (the 1 stays as the L entry).
 * Afterwards P[i] = i and Q[i] = i for i >= rank and L is compressed: for j < rank column Q[j] is swapped into
 * column j in rows j.. (mzd_col_swap_in_rows).  Every PLE variant of the reference (naive, russian, recursive
 * with any cutoff) produces exactly these bits, P and Q[0..rank) — pinned by tests/test_ple_oracle.py. */
static int orc_bit(orc_mzd const *M, orc_rci r, orc_rci c) { return (int)((rowp(M, r)[c / RADIX] >> (c % RADIX)) & 1); }

orc_rci orc_ple(orc_mzd *A, orc_rci *P, orc_rci *Q) {
  orc_rci row_pos = 0, col_pos = 0;
  while (row_pos < A->nrows && col_pos < A->ncols) {
    int found = 0;
    orc_rci i = 0, j = 0;
    for (j = col_pos; j < A->ncols && !found; ++j)
      for (i = row_pos; i < A->nrows; ++i)
        if (orc_bit(A, i, j)) { found = 1; break; }
    if (!found) break;
    --j;                                                /* the loop header advanced it once more */
    P[row_pos] = i;
    Q[row_pos] = j;
    if (i != row_pos) {                                 /* mzd_row_swap */
      orc_word *a = rowp(A, row_pos), *b = rowp(A, i);
      for (orc_wi w = 0; w < A->width; ++w) {
        orc_word const mask = (w == A->width - 1) ? A->high_bitmask : ~(orc_word)0;
        orc_word const t = (a[w] ^ b[w]) & mask;
        a[w] ^= t;
        b[w] ^= t;
      }
    }
    if (j + 1 < A->ncols) {
      orc_word const *src = rowp(A, row_pos);
      for (orc_rci l = row_pos + 1; l < A->nrows; ++l) {
        if (!orc_bit(A, l, j)) continue;
        orc_word *dst = rowp(A, l);                     /* mzd_row_add_offset(A, l, row_pos, j + 1) */
        orc_rci const c0 = j + 1;
        for (orc_wi w = c0 / RADIX; w < A->width; ++w) {
          orc_word mask = (w == A->width - 1) ? A->high_bitmask : ~(orc_word)0;
          if (w == c0 / RADIX) mask &= ~(orc_word)0 << (c0 % RADIX);
          dst[w] ^= src[w] & mask;
        }
      }
    }
    ++row_pos;
    col_pos = j + 1;
  }
  for (orc_rci i = row_pos; i < A->nrows; ++i) P[i] = i;
  for (orc_rci i = row_pos; i < A->ncols; ++i) Q[i] = i;
  for (orc_rci j = 0; j < row_pos; ++j) {               /* compress L */
    if (Q[j] <= j) continue;
    for (orc_rci r = j; r < A->nrows; ++r) {
      int const a = orc_bit(A, r, j), b = orc_bit(A, r, Q[j]);
      if (a != b) {
        rowp(A, r)[j / RADIX] ^= (orc_word)1 << (j % RADIX);
        rowp(A, r)[Q[j] / RADIX] ^= (orc_word)1 << (Q[j] % RADIX);
      }
    }
  }
  return row_pos;
}
