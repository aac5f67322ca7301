// io.cu — interchange formats either side of the path (reference: m4ri/io.c).
//
// Host-side readers/writers with the reference's formats and semantics, plus loaders that end in a device-resident
// matrix.  No compute here.  The PNG pair of the reference (io.c:72-294) needs libpng, which this image — and the
// reference build the tests compare against — does not have; the 1-bit image interchange offered instead is PBM "P4",
// which is the same bit-per-pixel row layout (1 = black = set bit, most significant bit first within a byte).
//   m4ri_b200_from_str      mzd_from_str   io.c:350-357   '1' sets a bit, row-major
//   m4ri_b200_from_jcf      mzd_from_jcf   io.c:297-348   "m n p\nnonzero\n\n" then one entry per line, a negative
//                                                         number starts the next row; p must be 2
//   m4ri_b200_to_jcf        (no counterpart; the inverse of the reader, so that files round-trip)
//   m4ri_b200_fprint_row    mzd_fprint_row io.c:49-68     "[" 64-bit words as "1"/" " with ":" every 4 and "|" between
//   m4ri_b200_to_pbm / m4ri_b200_from_pbm
//   m4ri_b200_dmat_from_jcf reader + upload: the matrix lands in HBM
#include <inttypes.h>
#include <string.h>

#include "dev.h"

using namespace m4b;

namespace {
inline void set_bit(mzd_t *A, rci_t i, rci_t j, int v) {
  word *w = A->data + (int64_t)i * A->rowstride + j / 64;
  word const b = (word)1 << (j % 64);
  *w = v ? (*w | b) : (*w & ~b);
}
inline int get_bit(mzd_t const *A, rci_t i, rci_t j) { return (int)((A->data[(int64_t)i * A->rowstride + j / 64] >> (j % 64)) & 1); }
}  // namespace

extern "C" {

mzd_t *m4ri_b200_from_str(rci_t m, rci_t n, char const *str) {
  mzd_t *A = m4ri_b200_mzd_init(m, n);
  size_t idx = 0;
  for (rci_t i = 0; i < m; ++i)
    for (rci_t j = 0; j < n; ++j) set_bit(A, i, j, str[idx++] == '1');
  return A;
}

mzd_t *m4ri_b200_from_jcf(char const *fn, int verbose) {
  FILE *fh = fopen(fn, "r");
  if (!fh) {
    if (verbose) printf("Could not open file '%s' for reading\n", fn);
    return NULL;
  }
  rci_t m, n;
  int p = 0;
  int64_t nonzero = 0;
  mzd_t *A = NULL;
  if (fscanf(fh, "%d %d %d\n%" SCNd64 "\n\n", &m, &n, &p, &nonzero) != 4) {
    if (verbose) printf("File '%s' does not seem to be in JCF format.", fn);
    fclose(fh);
    return NULL;
  }
  if (p != 2) {
    if (verbose) printf("Expected p==2 but found p==%d\n", p);
    fclose(fh);
    return NULL;
  }
  if (verbose)
    printf("reading %d x %d matrix with at most %" PRId64 " non-zero entries (density at most: %6.5f)\n", m, n, nonzero,
           ((double)nonzero) / ((double)m * n));
  A = m4ri_b200_mzd_init(m, n);
  rci_t i = -1, j = 0;
  while (fscanf(fh, "%d\n", &j) == 1) {
    if (j < 0) {
      ++i;
      j = -j;
    }
    if (((j - 1) >= n) || (i >= m) || i < 0 || j < 1) die("trying to write to (%d,%d) in %d x %d matrix\n", i, j - 1, m, n);
    set_bit(A, i, j - 1, 1);
  }
  fclose(fh);
  return A;
}

int m4ri_b200_to_jcf(mzd_t const *A, char const *fn) {
  FILE *fh = fopen(fn, "w");
  if (!fh) return 1;
  int64_t nonzero = 0;
  for (rci_t i = 0; i < A->nrows; ++i)
    for (rci_t j = 0; j < A->ncols; ++j) nonzero += get_bit(A, i, j);
  fprintf(fh, "%d %d 2\n%" PRId64 "\n\n", A->nrows, A->ncols, nonzero);
  // every row must announce itself (a negative entry): an empty row cannot be expressed, so the writer refuses
  for (rci_t i = 0; i < A->nrows; ++i) {
    bool first = true;
    for (rci_t j = 0; j < A->ncols; ++j)
      if (get_bit(A, i, j)) {
        fprintf(fh, "%d\n", first ? -(j + 1) : (j + 1));
        first = false;
      }
    if (first) {
      fclose(fh);
      return 2;
    }
  }
  fclose(fh);
  return 0;
}

void m4ri_b200_fprint_row(FILE *stream, mzd_t const *M, rci_t i) {
  fprintf(stream, "[");
  word const *row = M->data + (int64_t)i * M->rowstride;
  for (wi_t w = 0; w + 1 < M->width; ++w) {
    for (int b = 0; b < 64; ++b) {
      if (b && b % 4 == 0) fputc(':', stream);
      fputc(((row[w] >> b) & 1) ? '1' : ' ', stream);
    }
    fputc('|', stream);
  }
  if (M->width > 0) {
    int const wide = (M->ncols % 64) ? M->ncols % 64 : 64;
    for (int b = 0; b < wide; ++b) {
      if (b && b % 4 == 0) fputc(':', stream);
      fputc(((row[M->width - 1] >> b) & 1) ? '1' : ' ', stream);
    }
  }
  fprintf(stream, "]\n");
}

int m4ri_b200_to_pbm(mzd_t const *A, char const *fn) {
  FILE *fh = fopen(fn, "wb");
  if (!fh) return 1;
  fprintf(fh, "P4\n%d %d\n", A->ncols, A->nrows);
  size_t const bytes = ((size_t)A->ncols + 7) / 8;
  std::vector<unsigned char> line(bytes);
  for (rci_t i = 0; i < A->nrows; ++i) {
    memset(line.data(), 0, bytes);
    for (rci_t j = 0; j < A->ncols; ++j)
      if (get_bit(A, i, j)) line[j / 8] |= (unsigned char)(0x80u >> (j % 8));
    if (fwrite(line.data(), 1, bytes, fh) != bytes) {
      fclose(fh);
      return 1;
    }
  }
  fclose(fh);
  return 0;
}

mzd_t *m4ri_b200_from_pbm(char const *fn) {
  FILE *fh = fopen(fn, "rb");
  if (!fh) return NULL;
  int w = 0, h = 0;
  char magic[3] = {0, 0, 0};
  if (fscanf(fh, "%2s %d %d", magic, &w, &h) != 3 || strcmp(magic, "P4") || w < 0 || h < 0) {
    fclose(fh);
    return NULL;
  }
  fgetc(fh);   // the single whitespace byte after the header
  mzd_t *A = m4ri_b200_mzd_init(h, w);
  size_t const bytes = ((size_t)w + 7) / 8;
  std::vector<unsigned char> line(bytes);
  for (rci_t i = 0; i < h; ++i) {
    if (fread(line.data(), 1, bytes, fh) != bytes) {
      fclose(fh);
      m4ri_b200_mzd_free(A);
      return NULL;
    }
    for (rci_t j = 0; j < w; ++j)
      if (line[j / 8] & (0x80u >> (j % 8))) set_bit(A, i, j, 1);
  }
  fclose(fh);
  return A;
}

m4ri_b200_dmat *m4ri_b200_dmat_from_jcf(char const *fn, int verbose) {
  mzd_t *A = m4ri_b200_from_jcf(fn, verbose);
  if (!A) return NULL;
  m4ri_b200_dmat *D = m4ri_b200_dmat_alloc(A->nrows, A->ncols);
  m4ri_b200_upload(D, A, NULL);
  m4ri_b200_mzd_free(A);
  return D;
}

}  // extern "C"
