/* m4ri_b200.h — C-ABI of libm4ri_b200.so: a B200 (sm_100a) drop-in for the dense
 * GF(2) multiplication path of M4RI (malb/m4ri @ 5d0d0ce).
 *
 * Part 1 re-declares the reference's own types and entry points, with the file:line
 * of the reference declaration each one replaces.  A program that includes
 * <m4ri/m4ri.h> and links libm4ri_b200.so before libm4ri.so (or LD_PRELOADs it)
 * gets these symbols from here; everything else (mzd_init, mzd_free, PLE, TRSM, ...)
 * keeps coming from libm4ri.  See INTEGRATION.md.
 *
 * Part 2 is the device-resident extension API (m4ri_b200_*): explicit device
 * matrices, uploads/downloads and the kernels without the host round trip.  Plain
 * pointers and sizes only; no C++/torch types.
 *
 * Error convention = the reference's: invalid arguments and CUDA failures print to
 * stderr and abort() (m4ri_die, m4ri/misc.c:36-42).  There is no CPU fallback: if no
 * CUDA device is usable the call dies.
 */
#ifndef M4RI_B200_H
#define M4RI_B200_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Part 1: reference ABI ------------------------------------------------------- */

#ifndef M4RI_MISC_H            /* m4ri/misc.h:64-87 */
typedef int      BIT;
typedef int      rci_t;        /* row/column index */
typedef int64_t  wi_t;         /* word index */
typedef uint64_t word;         /* 64 matrix entries, column j in bit j%64 (LSB first) */
#define m4ri_radix 64
#endif

#ifndef M4RI_MZD_H             /* m4ri/mzd.h:68-99 — 64-byte header, layout asserted in capi.cpp */
typedef struct mzd_t {
  rci_t   nrows;
  rci_t   ncols;
  wi_t    width;               /* ceil(ncols/64) */
  wi_t    rowstride;           /* words between consecutive rows */
  uint8_t flags;               /* 0x2 non-zero excess, 0x4 windowed (mzd.h:144,150) */
  uint8_t padding[63 - 2 * sizeof(rci_t) - 2 * sizeof(wi_t) - sizeof(word) - sizeof(void *)];
  word    high_bitmask;        /* valid bits of word width-1 */
  word   *data;                /* row i at data + i*rowstride (mzd.h:185-187) */
} mzd_t;
#endif

/* C = A*B, Strassen-Winograd above the leaf kernels (tensor-core / M4RM).  C may be NULL (allocated, caller
 * frees with mzd_free).  cutoff 0 = library default, <0 dies.  m4ri/strassen.h:52,
 * strassen.c:345-365. */
mzd_t *mzd_mul(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff);
/* C ^= A*B.  m4ri/strassen.h:68, strassen.c:675-700. */
mzd_t *mzd_addmul(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff);
/* unchecked variants.  m4ri/strassen.h:88 (_mzd_mul_even), :109 (_mzd_addmul_even), :126 (_mzd_addmul);
 * strassen.c:41,367,667. */
mzd_t *_mzd_mul_even(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff);
mzd_t *_mzd_addmul_even(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff);
mzd_t *_mzd_addmul(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff);
/* M4RM only (no Strassen).  k = table bits, 0 = auto; any k gives the same bits, the
 * device kernel always uses its own k.  m4ri/brilliantrussian.h:274,291,317;
 * brilliantrussian.c:999-1190. */
mzd_t *mzd_mul_m4rm(mzd_t *C, mzd_t const *A, mzd_t const *B, int k);
mzd_t *mzd_addmul_m4rm(mzd_t *C, mzd_t const *A, mzd_t const *B, int k);
mzd_t *_mzd_mul_m4rm(mzd_t *C, mzd_t const *A, mzd_t const *B, int k, int clear);
/* block-parallel multiply: the reference splits C 2x2 over four OpenMP sections
 * (m4ri/mp.h:47,62; mp.c:277-324); here C is cut into pr x pc blocks over the GPUs chosen with
 * m4ri_b200_set_num_devices (2 column blocks from four GPUs on — the reference's 2 x 2 shape, mp.c:179-228),
 * operands cross PCIe once per box and are exchanged over NVLink, transfers overlap the products. */
mzd_t *mzd_mul_mp(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff);
mzd_t *mzd_addmul_mp(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff);
/* the unchecked block forms behind them (m4ri/mp.h:74,86; mp.c:39-156, 158-275) */
mzd_t *_mzd_mul_mp4(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff);
mzd_t *_mzd_addmul_mp4(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff);

/* Widened rows (callers of the multiplication path): triangular solves with matrices, X overwrites B,
 * unit diagonal implied, only the strict triangle of the triangular operand is read.
 *   left:  L X = B, U X = B   m4ri/triangular.h:115,127,142,153; triangular.c:394-516
 *   right: X L = B, X U = B   m4ri/triangular.h:50,64,82,100;   triangular.c:29-148, 300-392 */
void mzd_trsm_lower_left(mzd_t const *L, mzd_t *B, const int cutoff);
void _mzd_trsm_lower_left(mzd_t const *L, mzd_t *B, const int cutoff);
void mzd_trsm_upper_left(mzd_t const *U, mzd_t *B, const int cutoff);
void _mzd_trsm_upper_left(mzd_t const *U, mzd_t *B, const int cutoff);
void mzd_trsm_lower_right(mzd_t const *L, mzd_t *B, const int cutoff);
void _mzd_trsm_lower_right(mzd_t const *L, mzd_t *B, const int cutoff);
void mzd_trsm_upper_right(mzd_t const *U, mzd_t *B, const int cutoff);
void _mzd_trsm_upper_right(mzd_t const *U, mzd_t *B, const int cutoff);

/* PLE decomposition A = P L E in place, with the matrix resident in HBM for the whole factorisation (the recursion,
 * its triangular solves and its Schur updates never leave the GPU): E above, L compressed into the first `rank`
 * columns below, P as row transpositions (P->values[i] = row swapped with row i), Q->values[i] = column of the i-th
 * pivot.  Same pivot rule as the reference, hence the same bits, P, rank and Q[0..rank).  m4ri/ple.h (mzd_ple,
 * _mzd_ple), ple.c:33-178.  The reference's _mzd_pluq / mzd_echelonize_pluq / solve call _mzd_ple through the PLT and
 * are captured with it when this library is preloaded. */
typedef struct mzp_t {     /* m4ri/mzp.h:37-49 */
  rci_t *values;
  rci_t  length;
} mzp_t;
rci_t mzd_ple(mzd_t *A, mzp_t *P, mzp_t *Q, const int cutoff);
rci_t _mzd_ple(mzd_t *A, mzp_t *P, mzp_t *Q, const int cutoff);

/* Elimination entry points served by the device RREF (m4ri/echelonform.h, brilliantrussian.h:56; echelonform.c:30-36,
 * brilliantrussian.c:603-967, 971-997).  full != 0: the reduced row echelon form (unique, bit-identical to the
 * reference).  full == 0: handed on to the libm4ri that follows in the link order when there is one (its result
 * depends on its k); stand-alone the reduced form is returned, which is a row echelon form as well.  Returns the rank.
 * mzd_inv_m4ri: B = A^-1 as the right block of the RREF of [A | I]; B may be NULL. */
rci_t  mzd_echelonize_m4ri(mzd_t *A, int full, int k);
rci_t  mzd_echelonize(mzd_t *A, int full);
mzd_t *mzd_inv_m4ri(mzd_t *B, mzd_t const *A, int k);

/* ---- Part 2: extension API --------------------------------------------------------- */

/* Library / device control. */
int         m4ri_b200_version(void);
int         m4ri_b200_device_count(void);
void        m4ri_b200_set_device(int device);          /* device used by this process (default: current) */
void        m4ri_b200_set_num_devices(int n);           /* GPUs used by mzd_mul_mp (default 1) */
void        m4ri_b200_set_default_cutoff(int cutoff);   /* Strassen leaf size used when cutoff==0 */
int         m4ri_b200_get_default_cutoff(void);
void        m4ri_b200_release(void);                    /* free cached device workspaces */
char const *m4ri_b200_last_path(void);                  /* "m4rm" / "strassen:<levels>" of the last product */
uint64_t    m4ri_b200_kernel_launches(void);            /* CUDA kernels launched by this library so far */
/* Leaf kernel: 0 = automatic (the tensor-core leaf for products in its tile units, else the M4RM leaves), 1 = M4RM
 * 1024 x 1024-bit tiles only, 2 = M4RM 4096 x 256-bit tiles for every shape, 3 = same as 0; anything else = back to
 * $M4RI_B200_LEAF / the built-in default.  No choice changes a result bit.  Returns the previous setting. */
int         m4ri_b200_set_leaf_variant(int variant);
int         m4ri_b200_last_leaf_variant(void);          /* kernel of the last leaf launch: 1, 2 (M4RM) or 3 (tensor-core); 0: none yet */

/* Live timing of the M4RM leaf launches: between begin and end every leaf launch is bracketed
 * by CUDA events on its stream.  end() (call after synchronising) returns the number of leaf
 * launches, their summed device time in ms and their summed 2*m*l*n. */
void     m4ri_b200_profile_begin(void);
uint64_t m4ri_b200_profile_end(double *leaf_ms, double *leaf_bitops);

/* Stand-alone host matrices (same layout/semantics as mzd_init / mzd_init_window /
 * mzd_free, m4ri/mzd.c:142-185) for programs that do not link libm4ri. */
mzd_t *m4ri_b200_mzd_init(rci_t r, rci_t c);
mzd_t *m4ri_b200_mzd_init_window(mzd_t *M, rci_t lowr, rci_t lowc, rci_t highr, rci_t highc);
void   m4ri_b200_mzd_free(mzd_t *M);
/* frees a result this library allocated for a NULL C / DST argument (it comes from the process' libm4ri
 * mzd_init when one is loaded, else from m4ri_b200_mzd_init) */
void   m4ri_b200_result_free(mzd_t *M);

/* Interchange (m4ri/io.h; io.c:49-68, 297-357): same formats and semantics as mzd_from_str, mzd_from_jcf and
 * mzd_fprint_row; to_jcf is the inverse of the reader (returns 2 for a matrix with an empty row, which JCF cannot
 * express); PBM "P4" is the libpng-free 1-bit image interchange (the reference's PNG pair needs libpng). */
mzd_t *m4ri_b200_from_str(rci_t m, rci_t n, char const *str);
mzd_t *m4ri_b200_from_jcf(char const *fn, int verbose);
int    m4ri_b200_to_jcf(mzd_t const *A, char const *fn);
void   m4ri_b200_fprint_row(FILE *stream, mzd_t const *M, rci_t i);
int    m4ri_b200_to_pbm(mzd_t const *A, char const *fn);
mzd_t *m4ri_b200_from_pbm(char const *fn);

/* Device-resident matrix: bit-packed rows exactly like mzd_t (64-bit words, LSB-first),
 * pitch a multiple of 2 words, base 16-byte aligned, and every bit between ncols and
 * the pitch zero. */
typedef struct m4ri_b200_dmat {
  word   *data;      /* device pointer */
  int64_t pitch;     /* words between rows */
  rci_t   nrows;
  rci_t   ncols;
  int     owner;     /* 1: data was allocated by m4ri_b200_dmat_alloc */
} m4ri_b200_dmat;

m4ri_b200_dmat *m4ri_b200_dmat_alloc(rci_t nrows, rci_t ncols);   /* zero-filled */
/* wrap caller-owned device memory (e.g. a torch tensor): ptr 16-byte aligned, pitch_words
 * even and >= ceil(ncols/128)*2, padding bits must be zero. */
m4ri_b200_dmat *m4ri_b200_dmat_wrap(void *device_ptr, int64_t pitch_words, rci_t nrows, rci_t ncols);
void            m4ri_b200_dmat_free(m4ri_b200_dmat *M);
m4ri_b200_dmat *m4ri_b200_dmat_from_jcf(char const *fn, int verbose);   /* JCF file -> device-resident matrix (NULL on error) */
/* host mzd_t (any window/stride/excess) -> device (excess bits cleared) and back (bits of
 * the host matrix outside nrows x ncols are preserved).  stream: cudaStream_t or NULL. */
void m4ri_b200_upload(m4ri_b200_dmat *dst, mzd_t const *src, void *stream);
void m4ri_b200_download(mzd_t *dst, m4ri_b200_dmat const *src, void *stream);
void m4ri_b200_sync(void *stream);

/* C (^)= A*B on device matrices, asynchronous on `stream`.
 *   dmul_m4rm : the M4RM leaf kernel only (config "mzd_mul_m4rm").
 *   dmul      : Strassen-Winograd down to `cutoff`, then the leaf (config "mzd_mul").
 * clear != 0 computes C = A*B, clear == 0 computes C ^= A*B. */
void m4ri_b200_dmul_m4rm(m4ri_b200_dmat *C, m4ri_b200_dmat const *A, m4ri_b200_dmat const *B, int clear, void *stream);
void m4ri_b200_dmul(m4ri_b200_dmat *C, m4ri_b200_dmat const *A, m4ri_b200_dmat const *B, int cutoff, int clear, void *stream);
/* C = A*B (clear) or C ^= A*B on explicit top-level quadrants (order 11, 12, 21, 22; each may be its own
 * allocation) with transfer hooks called while the Strassen-Winograd schedule is enqueued: need_a/b/c(q) before the
 * first kernel that reads quadrant q of A / B / C-as-addend, done_c(q) after the last kernel that writes quadrant q
 * of C.  Used by the multi-rank end-to-end path to overlap uploads, NVLink exchanges and downloads with compute. */
typedef struct m4ri_b200_hooks {
  void (*need_a)(void *user, int q);
  void (*need_b)(void *user, int q);
  void (*need_c)(void *user, int q);
  void (*done_c)(void *user, int q);
  void *user;
} m4ri_b200_hooks;
void m4ri_b200_dmul_quads(m4ri_b200_dmat *const C[4], m4ri_b200_dmat const *const A[4], m4ri_b200_dmat const *const B[4],
                          int cutoff, int clear, void *stream, m4ri_b200_hooks const *hooks);
/* as dmul with the Strassen depth given explicitly (0 = leaf only); for cutoff sweeps. */
/* Direct entry points of the tensor-core leaf (csrc/tc_leaf.cu) for tests and measurements; the reference-named entry
 * points reach the same kernel through the automatic leaf choice (M4RI_B200_LEAF, m4ri_b200_set_leaf_variant).
 * m4ri_b200_dmul_tc: the small output-stationary cross-check kernel, C = A*B (clear != 0) or C ^= A*B; Bt is B
 * transposed (n x l); m % 128 == 0, n % 256 == 0, l % 128 == 0. */
void m4ri_b200_dmul_tc(m4ri_b200_dmat *C, m4ri_b200_dmat const *A, m4ri_b200_dmat const *Bt, int clear, void *stream);
/* m4ri_b200_dmul_tc2: C = A*B with the production kernel (operand images by a pre-pass, B-stationary pipeline);
 * m % 128 == 0, l % 1024 == 0, n % 256 == 0, 16-byte aligned rows. */
void m4ri_b200_dmul_tc2(m4ri_b200_dmat *C, m4ri_b200_dmat const *A, m4ri_b200_dmat const *B, void *stream);
/* PLE of a device-resident matrix in place; P (nrows ints) and Q (ncols ints) are host arrays; returns the rank */
rci_t m4ri_b200_dple(m4ri_b200_dmat *A, rci_t *P, rci_t *Q, void *stream);
void m4ri_b200_dmul_levels(m4ri_b200_dmat *C, m4ri_b200_dmat const *A, m4ri_b200_dmat const *B, int levels, int clear, void *stream);
/* T X = B (left != 0) or X T = B (left == 0) on device matrices, T lower triangular or upper if
 * upper != 0; X overwrites B. */
void m4ri_b200_dtrsm(m4ri_b200_dmat const *T, m4ri_b200_dmat *B, int upper, int left, int cutoff, void *stream);
/* DST = A^T on device matrices (DST must not alias A), and the host form with mzd_transpose's semantics
 * (m4ri/mzd.c:1118-1139; DST may be NULL).  Not exported as mzd_transpose: libm4ri keeps its own. */
void   m4ri_b200_dtranspose(m4ri_b200_dmat *DST, m4ri_b200_dmat const *A, void *stream);
mzd_t *m4ri_b200_transpose(mzd_t *DST, mzd_t const *A);
/* C = A ^ B on device matrices (the device form of _mzd_add, m4ri/mzd.c:1471-1583). */
void m4ri_b200_dadd(m4ri_b200_dmat *C, m4ri_b200_dmat const *A, m4ri_b200_dmat const *B, void *stream);

/* Reduced row echelon form in place, returns the rank: the result of mzd_echelonize_m4ri(A, 1, k)
 * (m4ri/brilliantrussian.h:79 via echelonform.h, m4ri/brilliantrussian.c:603-967).  The reduced form is unique,
 * hence bit-identical to every reference variant; `full` is accepted for signature compatibility, the reduced form
 * is returned either way.  The device form synchronises the stream (it returns the rank). */
int   m4ri_b200_dechelonize(m4ri_b200_dmat *A, int full, void *stream);
rci_t m4ri_b200_echelonize(mzd_t *A, int full);
/* B = A^-1 as mzd_inv_m4ri computes it (m4ri/brilliantrussian.h, m4ri/brilliantrussian.c:971-997): the right block of the
 * reduced row echelon form of [A | I]; no invertibility test, B may be NULL (allocated). */
mzd_t *m4ri_b200_inv_m4ri(mzd_t *B, mzd_t const *A);

#ifdef __cplusplus
}
#endif
#endif /* M4RI_B200_H */
