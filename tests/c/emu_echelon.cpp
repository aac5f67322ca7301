// emu_echelon.cpp — CPU emulation of the device RREF (TEST INFRASTRUCTURE, never shipped).
//
// Compiles m4ri_b200/csrc/echelon_body.h — the source nvcc compiles into the echelon kernels — with g++ and
// runs the whole strip loop of echelon.cu on the CPU: a selection CTA is 512 host threads with a std::barrier,
// the element-wise passes run as a loop over "threads", the two products of a pass (Bm = G * PIV and
// A ^= X * Bm, the M4RM leaf on the GPU) are definition-level GF(2) products.  The result and the rank are
// compared with a plain Gauss-Jordan elimination.
//
//   g++ -O2 -std=c++20 -pthread -I m4ri_b200/csrc tests/c/emu_echelon.cpp -o emu_echelon && ./emu_echelon
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <barrier>
#include <thread>
#include <vector>

#define ECH_FN inline
#include "echelon_body.h"

namespace {

using ech::u64;

struct EmuCtx {
  int tid, ntid;
  std::barrier<> *bar;
  void sync() { bar->arrive_and_wait(); }
  int atomic_min(int *p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
  }
};

template <class F>
void run_cta(F f) {       // one selection CTA: kSelThreads host threads
  std::barrier<> bar(ech::kSelThreads);
  std::vector<std::thread> th;
  for (int t = 0; t < ech::kSelThreads; ++t)
    th.emplace_back([&, t] {
      EmuCtx cx{t, ech::kSelThreads, &bar};
      f(cx);
    });
  for (auto &x : th) x.join();
}

struct Mat {               // device-view layout: pitch = ceil(ncols/128)*2 words, zero padding
  int nrows, ncols;
  long long pitch;
  std::vector<u64> w;
  Mat(int r, int c) : nrows(r), ncols(c), pitch((long long)((c + 127) / 128) * 2), w((size_t)(r > 0 ? r : 1) * pitch, 0) {}
  u64 *row(int i) { return w.data() + (size_t)i * pitch; }
};

u64 g_rng = 12345;
u64 rnd() {
  u64 z = (g_rng += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// plain Gauss-Jordan (first row with the bit, swap up, clear the column everywhere): the unique RREF
int rref_definition(Mat &M) {
  int r = 0;
  for (int c = 0; c < M.ncols && r < M.nrows; ++c) {
    int piv = -1;
    for (int i = r; i < M.nrows; ++i)
      if ((M.row(i)[c / 64] >> (c % 64)) & 1) {
        piv = i;
        break;
      }
    if (piv < 0) continue;
    for (long long w = 0; w < M.pitch; ++w) std::swap(M.row(r)[w], M.row(piv)[w]);
    for (int i = 0; i < M.nrows; ++i)
      if (i != r && ((M.row(i)[c / 64] >> (c % 64)) & 1))
        for (long long w = 0; w < M.pitch; ++w) M.row(i)[w] ^= M.row(r)[w];
    ++r;
  }
  return r;
}

// the strip loop of echelon.cu
int device_rref_emulated(Mat &A, int chunk_rows) {
  int const m = A.nrows, n = A.ncols;
  int const nchunks = (m + chunk_rows - 1) / chunk_rows;
  if (nchunks * 64 > ech::kSelRows || chunk_rows > ech::kSelRows) {
    fprintf(stderr, "emu: chunking out of range\n");
    abort();
  }
  ech::State st;
  memset(&st, 0, sizeof st);
  std::vector<int> cand_row((size_t)nchunks * 64);
  std::vector<u64> cand_word((size_t)nchunks * 64);
  std::vector<u64> Gm(128), X((size_t)2 * m);
  long long const pitch = A.pitch;
  std::vector<u64> PIV((size_t)64 * pitch), Bm((size_t)64 * pitch);
  static ech::SelShared sh;
  long long const T = 1000;                           // "grid" of the element-wise passes
  for (int s = 0; s * 64 < n; ++s) {
    int const w0 = (s / 2) * 2, nw = (int)(pitch - w0);
    for (int c = 0; c < nchunks; ++c)
      run_cta([&](EmuCtx &cx) {
        ech::select_chunk(cx, &sh, &st, A.w.data(), pitch, m, s, chunk_rows, c, cand_row.data(), cand_word.data());
      });
    run_cta([&](EmuCtx &cx) {
      ech::select_final(cx, &sh, &st, cand_row.data(), cand_word.data(), nchunks * 64, Gm.data());
    });
    for (long long t = 0; t < T; ++t) ech::gather_pivots(&st, A.w.data(), pitch, w0, nw, PIV.data(), pitch, t, T);
    // Bm = G * PIV  (64 x 64 x n')
    for (int q = 0; q < 64; ++q)
      for (int w = 0; w < nw; ++w) {
        u64 acc = 0;
        for (int t = 0; t < 64; ++t)
          if ((Gm[2 * q] >> t) & 1) acc ^= PIV[(size_t)t * pitch + w];
        Bm[(size_t)q * pitch + w] = acc;
      }
    for (long long t = 0; t < T; ++t) ech::build_x(&st, A.w.data(), pitch, s, m, X.data(), t, T);
    // A[:, c0:] ^= X * Bm
    for (int i = 0; i < m; ++i)
      for (int q = 0; q < 64; ++q)
        if ((X[2 * (size_t)i] >> q) & 1)
          for (int w = 0; w < nw; ++w) A.row(i)[w0 + w] ^= Bm[(size_t)q * pitch + w];
    for (long long t = 0; t < T; ++t) ech::move_rows(&st, A.w.data(), pitch, w0, nw, t, T);
    for (long long t = 0; t < T; ++t) ech::place_pivots(&st, A.w.data(), pitch, w0, nw, Bm.data(), pitch, t, T);
    ech::advance(&st);
  }
  return st.rank;
}

bool run_case(int m, int n, int rank_bound, int chunk_rows) {
  Mat A(m, n);
  if (rank_bound <= 0) {
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < (n + 63) / 64; ++j) {
        u64 v = rnd();
        if (j == n / 64 && n % 64) v &= ~0ull >> (64 - n % 64);
        A.row(i)[j] = v;
      }
  } else {                                             // rows = random combinations of `rank_bound` random rows
    Mat Bs(rank_bound, n);
    for (int i = 0; i < rank_bound; ++i)
      for (int j = 0; j < (n + 63) / 64; ++j) {
        u64 v = rnd();
        if (j == n / 64 && n % 64) v &= ~0ull >> (64 - n % 64);
        Bs.row(i)[j] = v;
      }
    for (int i = 0; i < m; ++i)
      for (int k = 0; k < rank_bound; ++k)
        if (rnd() & 1)
          for (long long w = 0; w < A.pitch; ++w) A.row(i)[w] ^= Bs.row(k)[w];
  }
  Mat W = A;
  int const want = rref_definition(W);
  int const got = device_rref_emulated(A, chunk_rows);
  bool const ok = want == got && A.w == W.w;
  printf("%s %d x %d rank %d (emulated %d) chunk %d\n", ok ? "ok  " : "FAIL", m, n, want, got, chunk_rows);
  return ok;
}

}  // namespace

// matrix from a file of m * ceil(n/64) little-endian words (row major), e.g. written by a Python test
bool run_file(char const *path, int m, int n, int chunk_rows) {
  Mat A(m, n);
  FILE *f = fopen(path, "rb");
  if (!f) return false;
  for (int i = 0; i < m; ++i)
    if (fread(A.row(i), 8, (size_t)(n + 63) / 64, f) != (size_t)(n + 63) / 64) return false;
  fclose(f);
  Mat W = A;
  int const want = rref_definition(W), got = device_rref_emulated(A, chunk_rows);
  bool const ok = want == got && A.w == W.w;
  printf("%s file %d x %d rank %d (emulated %d)\n", ok ? "ok  " : "FAIL", m, n, want, got);
  return ok;
}

int main(int argc, char **argv) {
  bool ok = true;
  if (argc == 6 && !strcmp(argv[1], "file")) {
    ok = run_file(argv[2], atoi(argv[3]), atoi(argv[4]), atoi(argv[5]));
  } else if (argc == 5) {
    ok = run_case(atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]));
  } else {
    struct Case {
      int m, n, rank_bound, chunk;
    } const cases[] = {
        {4, 67, 0, 8192},      // the reference's shape list starts like this (tests/test_elimination.c)
        {65, 17, 0, 8192},   {100, 100, 0, 8192}, {193, 65, 0, 50},  // several chunks: candidates merged by the final CTA
        {300, 200, 0, 64},   {200, 300, 0, 37},   {257, 257, 129, 100},   // rank deficient
        {500, 130, 1, 128},  {64, 64, 0, 16},     {1, 1, 0, 8192},
    };
    for (auto const &c : cases) ok &= run_case(c.m, c.n, c.rank_bound, c.chunk);
  }
  return ok ? 0 : 1;
}
