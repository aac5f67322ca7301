/* oracle/m4rm_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's GF(2) multiplication path
 * (malb/m4ri @ 5d0d0ce: m4ri/mzd.{h,c}, graycode.c, brilliantrussian.c,
 * strassen.c).  It exists so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg can check the CUDA product path; nothing under
 * m4ri_b200/ may include, link or call it.
 *
 * Parity status: PINNED — tests/test_oracle.py checks every function
 * here bit-for-bit against the unmodified reference compiled into
 * oracle/_ref/libm4ri_ref.so on the reference's own test shape list
 * (tests/test_multiplication.c:251-322) and against the committed fixtures in
 * tests/golden/ (generated from the reference by tests/golden/make_golden.py).
 */
#ifndef M4RM_ORACLE_H
#define M4RM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t orc_word;
typedef int      orc_rci;   /* m4ri/misc.h:72  rci_t */
typedef int64_t  orc_wi;    /* m4ri/misc.h:81  wi_t  */

/* Same 64-byte layout as the reference's mzd_t (m4ri/mzd.h:68-99) so that one
 * ctypes structure describes reference, oracle and product matrices. */
typedef struct orc_mzd {
  orc_rci  nrows;
  orc_rci  ncols;
  orc_wi   width;
  orc_wi   rowstride;
  uint8_t  flags;
  uint8_t  padding[23];
  orc_word high_bitmask;
  orc_word *data;
} orc_mzd;

orc_mzd *orc_init(orc_rci r, orc_rci c);
orc_mzd *orc_init_window(orc_mzd *M, orc_rci lowr, orc_rci lowc, orc_rci highr, orc_rci highc);
void     orc_free(orc_mzd *M);

orc_word orc_random_word(void);
void     orc_randomize(orc_mzd *M);
int      orc_equal(orc_mzd const *A, orc_mzd const *B);
void     orc_copy(orc_mzd *dst, orc_mzd const *src);
void     orc_add(orc_mzd *C, orc_mzd const *A, orc_mzd const *B);

int      orc_gray_code(int number, int length);
void     orc_build_code(int *ord, int *inc, int l);
void     orc_make_table(orc_mzd const *M, orc_rci r, int k, orc_mzd *T, orc_rci *L);

orc_mzd *orc_mul_naive(orc_mzd *C, orc_mzd const *A, orc_mzd const *B, int clear);
orc_mzd *orc_mul_m4rm(orc_mzd *C, orc_mzd const *A, orc_mzd const *B, int k, int clear);
orc_mzd *orc_mul(orc_mzd *C, orc_mzd const *A, orc_mzd const *B, int cutoff);
orc_mzd *orc_addmul(orc_mzd *C, orc_mzd const *A, orc_mzd const *B, int cutoff);

/* L X = B / U X = B, X overwrites B; unit diagonal implied (m4ri/triangular.c:406-516) */
void     orc_trsm_lower_left(orc_mzd const *L, orc_mzd *B);
void     orc_trsm_upper_left(orc_mzd const *U, orc_mzd *B);
/* X L = B / X U = B (m4ri/triangular.c:29-148, 300-392) */
void     orc_trsm_lower_right(orc_mzd const *L, orc_mzd *B);
void     orc_trsm_upper_right(orc_mzd const *U, orc_mzd *B);

/* DST = A^T (m4ri/mzd.c:1118-1139) */
orc_mzd *orc_transpose(orc_mzd *DST, orc_mzd const *A);

/* m4ri/mzd.c:208-233: (reduced) row echelon form in place, returns the rank */
orc_rci  orc_echelonize(orc_mzd *M, int full);

/* m4ri/ple.c:222-272: A = P L E in place (L compressed into the first rank columns); P has nrows entries,
 * Q ncols; returns the rank */
orc_rci  orc_ple(orc_mzd *A, orc_rci *P, orc_rci *Q);

#ifdef __cplusplus
}
#endif
#endif
