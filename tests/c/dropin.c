/* Drop-in integration test: a plain C program in the style of the reference's
 * tests/test_multiplication.c, linked against libm4ri_b200.so FIRST and the unmodified reference
 * (oracle/_ref/libm4ri_ref.so) SECOND.  mzd_mul / mzd_addmul / mzd_mul_m4rm resolve to the GPU
 * library; mzd_init, mzd_randomize, mzd_equal, mzd_free and the naive multiply used as the checker
 * come from the reference.  C == NULL results are allocated by the reference's mzd_init (found with
 * dlsym) and released by the reference's mzd_free. */
#include <stdio.h>
#include <stdlib.h>

#include "m4ri_b200.h"

/* from the reference library (m4ri/mzd.h) */
mzd_t *mzd_init(rci_t r, rci_t c);
void   mzd_free(mzd_t *A);
void   mzd_randomize(mzd_t *A);
int    mzd_equal(mzd_t const *A, mzd_t const *B);
mzd_t *mzd_copy(mzd_t *D, mzd_t const *S);
mzd_t *mzd_mul_naive(mzd_t *C, mzd_t const *A, mzd_t const *B);
mzd_t *mzd_addmul_naive(mzd_t *C, mzd_t const *A, mzd_t const *B);

static int mul_test_equality(rci_t m, rci_t l, rci_t n, int k, int cutoff) {
  int ret = 0;
  printf("   mul: m: %4d, l: %4d, n: %4d, k: %2d, cutoff: %4d", m, l, n, k, cutoff);
  mzd_t *A = mzd_init(m, l), *B = mzd_init(l, n);
  mzd_randomize(A);
  mzd_randomize(B);
  mzd_t *C = mzd_mul(NULL, A, B, cutoff);      /* GPU, Strassen + M4RM */
  mzd_t *D = mzd_mul_m4rm(NULL, A, B, k);      /* GPU, M4RM only       */
  mzd_t *E = mzd_mul_naive(NULL, A, B);        /* reference, CPU       */
  if (!mzd_equal(C, D)) { printf(" Strassen != M4RM"); ret -= 1; }
  if (!mzd_equal(D, E)) { printf(" M4RM != Naive"); ret -= 1; }
  /* addmul: C ^= A*B must give zero again */
  mzd_addmul(C, A, B, cutoff);
  mzd_t *Z = mzd_init(m, n);
  if (!mzd_equal(C, Z)) { printf(" addmul did not cancel"); ret -= 1; }
  mzd_free(Z); mzd_free(E); mzd_free(D); mzd_free(C); mzd_free(B); mzd_free(A);
  printf(ret ? " ... FAILED\n" : " ... passed\n");
  return ret;
}

int main(void) {
  int status = 0;
  srandom(17);
  status += mul_test_equality(1, 1, 1, 0, 1024);
  status += mul_test_equality(3, 131, 257, 0, 0);
  status += mul_test_equality(193, 65, 65, 8, 64);
  status += mul_test_equality(1025, 1025, 1025, 3, 256);
  status += mul_test_equality(1710, 1290, 1000, 0, 256);
  status += mul_test_equality(2048, 2048, 4096, 0, 1024);
  printf("kernel launches: %llu, last path: %s\n", (unsigned long long)m4ri_b200_kernel_launches(), m4ri_b200_last_path());
  if (m4ri_b200_kernel_launches() == 0) status -= 1;
  if (status == 0) { printf("All tests passed.\n"); return 0; }
  return 1;
}
