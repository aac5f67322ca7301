// emu_leaf2.cpp — CPU emulation harness for the tall-tile M4RM leaf (TEST INFRASTRUCTURE, never shipped).
//
// Compiles m4ri_b200/csrc/m4rm_leaf2_body.h — the very source nvcc compiles into the kernel — with g++:
// a CTA is NT host threads, shared memory is an array, __syncthreads/__syncwarp are std::barriers, the
// mbarriers count bytes like the hardware ones, TMA is a host copy with zero fill outside the tensor
// (negative coordinates included), red.global.xor is an atomic XOR.  The result is compared with a
// definition-level GF(2) product.  What this checks: every index formula of the kernel body (table layout,
// Gray walk, lane -> bank-group mapping, A byte rotation, stream-K segment logic, ring parities, flush
// addressing, ragged edges).  What it cannot check: the PTX wrappers and the tensor-map encoding, which
// are the same as in the first leaf kernel and are covered by the GPU parity tests.
//
// check_bank_groups() additionally verifies the bank model the design rests on: the eight lanes of a
// quarter-warp always touch eight different 16-byte bank groups in a lookup, a table store and a B-row load.
//
//   g++ -O2 -std=c++20 -pthread -I m4ri_b200/csrc tests/c/emu_leaf2.cpp -o emu_leaf2 && ./emu_leaf2
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <barrier>
#include <memory>
#include <thread>
#include <vector>

#define L2_FN inline
#ifndef EMU_SPLIT
#define EMU_SPLIT 0   // 1: half of the warps build the tables of a step
#endif
#ifndef EMU_AWIDE
#define EMU_AWIDE 0   // which A-load variant of the kernel body to emulate
#endif

namespace leaf2 {

struct U4 {
  uint32_t x, y, z, w;
};
struct U2 {
  uint32_t x, y;
};
struct TMap {                 // what cuTensorMapEncodeTiled describes: a 2D / 3D u32 tensor and a box
  uint32_t const *base;
  long dim0, dim1, dim2;      // u32 words per row, rows (per group), row groups
  long stride1, stride2;      // u32 words between rows / between groups
  int box0, box1, box2;
};

constexpr int kEmuThreads = 256;
constexpr uint32_t kEmuSbase = 1024;          // dynamic shared memory starts 1024-byte aligned, not at 0
static uint8_t g_smem[kEmuSbase + 232448];
static std::barrier<> g_cta_barrier(kEmuThreads);
struct EmuMbar {
  std::atomic<long> tx{0};
  std::atomic<int> pending{1};
  std::atomic<unsigned> phase{0};
};
static EmuMbar g_bar[2];
static uint32_t g_bar_base = 0;

static std::atomic<long> g_n_lds128{0}, g_n_sts128{0};

static inline void check_range(uint32_t addr, uint32_t bytes) {
  if (addr < kEmuSbase || addr + bytes > sizeof g_smem || addr % bytes) {
    fprintf(stderr, "emu: bad shared access %u (+%u)\n", addr, bytes);
    abort();
  }
}

L2_FN U4 lds128(uint32_t addr) {
  check_range(addr, 16);
  U4 v;
  memcpy(&v, g_smem + addr, 16);
  g_n_lds128.fetch_add(1, std::memory_order_relaxed);
  return v;
}
template <int IMM>
L2_FN U4 lds128(uint32_t addr) {
  return lds128(addr + IMM);
}
L2_FN U2 lds64(uint32_t addr) {
  check_range(addr, 8);
  U2 v;
  memcpy(&v, g_smem + addr, 8);
  return v;
}
L2_FN void sts128(uint32_t addr, U4 const &v) {
  check_range(addr, 16);
  memcpy(g_smem + addr, &v, 16);
  g_n_sts128.fetch_add(1, std::memory_order_relaxed);
}
L2_FN uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {   // PRMT, default mode
  uint8_t src[8];
  memcpy(src, &a, 4);
  memcpy(src + 4, &b, 4);
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) {
    uint32_t const n = (sel >> (4 * i)) & 0xF;
    uint8_t byte = src[n & 7];
    if (n & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
    r |= (uint32_t)byte << (8 * i);
  }
  return r;
}
static inline EmuMbar &bar_at(uint32_t addr) {
  uint32_t const idx = (addr - g_bar_base) / 8;
  if (addr < g_bar_base || idx > 1 || (addr - g_bar_base) % 8) {
    fprintf(stderr, "emu: bad mbarrier address %u\n", addr);
    abort();
  }
  return g_bar[idx];
}
static inline void bar_try_complete(EmuMbar &b) {
  // phase completes when the (single) arrival has happened and all expected bytes have landed
  if (b.pending.load() == 0 && b.tx.load() == 0) {
    int expected = 0;
    if (b.pending.compare_exchange_strong(expected, 1)) b.phase.fetch_add(1);
  }
}
L2_FN void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  EmuMbar &b = bar_at(bar);
  b.tx.fetch_add((long)bytes);
  b.pending.fetch_sub(1);
  bar_try_complete(b);
}
L2_FN void mbar_wait(uint32_t bar, uint32_t parity) {
  EmuMbar &b = bar_at(bar);
  while ((b.phase.load() & 1u) == parity) std::this_thread::yield();
}
L2_FN void tma_load_3d(uint32_t dst, TMap const *map, int c0, int c1, int c2, uint32_t bar) {
  if (dst % 128) {
    fprintf(stderr, "emu: TMA destination %u not 128-byte aligned\n", dst);
    abort();
  }
  uint32_t const bytes = (uint32_t)map->box0 * map->box1 * map->box2 * 4;
  check_range(dst, 16);
  check_range(dst + bytes - 16, 16);
  size_t o = 0;
  for (int g = 0; g < map->box2; ++g)
    for (int r = 0; r < map->box1; ++r)
      for (int c = 0; c < map->box0; ++c, ++o) {
        long const x = (long)c0 + c, y = (long)c1 + r, z = (long)c2 + g;
        uint32_t v = 0;
        if (x >= 0 && x < map->dim0 && y >= 0 && y < map->dim1 && z >= 0 && z < map->dim2)
          v = map->base[z * map->stride2 + y * map->stride1 + x];
        memcpy(g_smem + dst + o * 4, &v, 4);
      }
  EmuMbar &b = bar_at(bar);
  b.tx.fetch_sub((long)bytes);
  bar_try_complete(b);
}
L2_FN void tma_load_2d(uint32_t dst, TMap const *map, int c0, int c1, uint32_t bar) {
  if (map->box2 != 1) {
    fprintf(stderr, "emu: 2D TMA with a 3D map\n");
    abort();
  }
  tma_load_3d(dst, map, c0, c1, 0, bar);
}
L2_FN void cp_async16(uint32_t dst, void const *src, uint32_t src_bytes) {   // completes at once here
  check_range(dst, 16);
  if (src_bytes != 0 && src_bytes != 16) abort();
  uint8_t tmp[16] = {0};
  if (src_bytes) memcpy(tmp, src, 16);
  memcpy(g_smem + dst, tmp, 16);
}
L2_FN void cp_async_wait_all() {}
L2_FN void red_xor64(unsigned long long *p, uint32_t lo, uint32_t hi) {
  __atomic_fetch_xor(p, ((unsigned long long)hi << 32) | lo, __ATOMIC_RELAXED);
}
L2_FN void stg128(unsigned long long *p, U4 const &v) {
  p[0] = ((unsigned long long)v.y << 32) | v.x;
  p[1] = ((unsigned long long)v.w << 32) | v.z;
}
L2_FN void cta_sync() { g_cta_barrier.arrive_and_wait(); }
L2_FN uint32_t gate(uint32_t a, uint32_t b, uint32_t c) { return a | (b & c); }

}  // namespace leaf2

#include "m4rm_leaf2_body.h"

namespace {

using leaf2::TMap;

struct Mat {                  // device-view layout: 64-bit words, pitch = ceil(ncols/128)*2, zero padding
  int nrows, ncols;
  long pitch;
  std::vector<uint64_t> w;
  Mat(int r, int c) : nrows(r), ncols(c), pitch((long)((c + 127) / 128) * 2), w((size_t)r * pitch, 0) {}
  int get(int i, int j) const { return (int)((w[(size_t)i * pitch + j / 64] >> (j % 64)) & 1); }
};

uint64_t g_rng = 0x9E3779B97F4A7C15ull;
uint64_t rnd() {              // splitmix64
  uint64_t z = (g_rng += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
void randomize(Mat &M) {
  for (int i = 0; i < M.nrows; ++i)
    for (int j = 0; j < (M.ncols + 63) / 64; ++j) {
      uint64_t v = rnd();
      if (j == M.ncols / 64 && M.ncols % 64) v &= ~0ull >> (64 - M.ncols % 64);
      M.w[(size_t)i * M.pitch + j] = v;
    }
}
TMap map_of(Mat const &M, int box0, int box1) {      // 2D: words x rows
  return TMap{reinterpret_cast<uint32_t const *>(M.w.data()), (long)((M.ncols + 127) / 128) * 4, M.nrows, 1,
              M.pitch * 2, 0, box0, box1, 1};
}
TMap map_row_groups(Mat const &M, int box0, int group_rows, int box_groups) {   // 3D: words x 256 rows x groups
  return TMap{reinterpret_cast<uint32_t const *>(M.w.data()), (long)((M.ncols + 127) / 128) * 4, group_rows,
              M.nrows / group_rows, M.pitch * 2, M.pitch * 2 * group_rows, box0, group_rows, box_groups};
}

// C ^= A*B by definition, row-wise (reference semantics: m4ri/mzd.c:1141-1268 _mzd_mul_naive)
void addmul_definition(Mat &C, Mat const &A, Mat const &B) {
  for (int i = 0; i < A.nrows; ++i)
    for (int k = 0; k < A.ncols; ++k)
      if (A.get(i, k))
        for (long j = 0; j < B.pitch; ++j) C.w[(size_t)i * C.pitch + j] ^= B.w[(size_t)k * B.pitch + j];
}

bool run_case(int count, int m, int l, int n, int nblocks) {
  using namespace leaf2;
  std::vector<Mat> A, B, C, W;
  for (int i = 0; i < count; ++i) {
    A.emplace_back(m, l);
    B.emplace_back(l, n);
    C.emplace_back(m, n);
    randomize(A.back());
    randomize(B.back());
    randomize(C.back());
    W.push_back(C.back());
    addmul_definition(W.back(), A.back(), B.back());
  }
  Args p;
  memset(&p, 0, sizeof p);
  p.m = m;
  p.l = l;
  p.a3d = m % kABoxRows == 0 ? 1 : 0;
  p.nwordsC = (n + 63) / 64;
  p.tiles_m = (m + kTM - 1) / kTM;
  p.tiles_n = (n + kTileBits - 1) / kTileBits;
  p.slabs = (l + kSlabBits - 1) / kSlabBits;
  p.nprob = count;
  p.units_per_problem = (long long)p.tiles_m * p.tiles_n * p.slabs;
  p.total_units = p.units_per_problem * count;
  for (int i = 0; i < count; ++i) {
    p.C[i] = reinterpret_cast<unsigned long long *>(C[i].w.data());
    p.pitchC[i] = C[i].pitch;
    p.mapA[i] = p.a3d ? map_row_groups(A[i], 4, kABoxRows, kAParts) : map_of(A[i], 4, kABoxRows);
    p.B[i] = reinterpret_cast<unsigned long long const *>(B[i].w.data());
    p.pitchB[i] = B[i].pitch;
  }
  if (nblocks > p.total_units) nblocks = (int)p.total_units;
  // the launcher's hybrid partition: whole tiles round-robin first, stream-K over the rest (EMU_HYBRID=0: pure stream-K)
  {
    char const *e = getenv("EMU_HYBRID");
    long long const tiles_total = (long long)p.tiles_m * p.tiles_n * count;
    p.dp_rounds = (e && e[0] == '0') ? 0 : (int)(tiles_total / nblocks);
    // EMU_STORE=1: the launcher's C = A*B mode — whole-tile rounds store, only the products holding tail tiles start
    // from zeros, every other product starts from GARBAGE that must be overwritten
    char const *st = getenv("EMU_STORE");
    if (st && st[0] == '1' && p.dp_rounds > 0 && n % 128 == 0) {
      p.store_dp = 1;
      long long const tpp = (long long)p.tiles_m * p.tiles_n;
      int const first_zeroed = (int)((long long)p.dp_rounds * nblocks / tpp);
      for (int i = 0; i < count; ++i) {
        Mat Z(m, n);
        W[i] = Z;
        addmul_definition(W[i], A[i], B[i]);                        // expected: exactly A*B
        if (i >= first_zeroed) std::fill(C[i].w.begin(), C[i].w.end(), 0);
      }
    }
  }
  g_bar_base = kEmuSbase + kOffBar;

  std::vector<std::thread> th;
  for (int tid = 0; tid < kEmuThreads; ++tid)
    th.emplace_back([&, tid] {
      for (int bid = 0; bid < nblocks; ++bid) {
        if (tid == 0) {                 // a fresh CTA: garbage shared memory, freshly initialised mbarriers
          for (size_t i = 0; i < sizeof g_smem; ++i) g_smem[i] = (uint8_t)(0xA5 ^ i);
          for (auto &b : g_bar) {
            b.tx = 0;
            b.pending = 1;
            b.phase = 0;
          }
        }
        cta_sync();
        cta_body<kEmuThreads, EMU_AWIDE, EMU_SPLIT>(p, kEmuSbase, tid, bid, nblocks);
        cta_sync();
      }
    });
  for (auto &t : th) t.join();

  bool ok = true;
  for (int i = 0; i < count && ok; ++i) ok = C[i].w == W[i].w;
  printf("%s count=%d %dx%dx%d blocks=%d units=%lld\n", ok ? "ok  " : "FAIL", count, m, l, n, nblocks, p.total_units);
  return ok;
}

// Static check of the bank-group claims: for every lane of a quarter-warp and ANY index bytes, the eight
// 16-byte pieces of one LDS.128 / STS.128 instruction fall into eight different bank groups.
bool check_bank_groups() {
  using namespace leaf2;
  bool ok = true;
  for (int jj = 0; jj < 4; ++jj)
    for (int second = 0; second < 2; ++second) {
      unsigned seen = 0;
      for (int i8 = 0; i8 < 8; ++i8) {
        int const hl = i8 >> 2;
        uint32_t const idx = (uint32_t)(rnd() & 0xFF);
        uint32_t ad = (uint32_t)hl * 64u + ((uint32_t)(i8 + jj) & 3u) * 16u + idx * kLineBytes;
        if (second) ad ^= 64u;
        seen |= 1u << ((ad / 16) & 7);
      }
      if (seen != 0xFF) ok = false;
    }
  // build: loads of B row b (same b in all eight lanes), stores of one line
  for (int b = 0; b < 8; ++b) {
    unsigned seen_ld = 0, seen_st = 0;
    for (int c = 0; c < 8; ++c) {
      uint32_t const src = (uint32_t)(c * 16 + b * 128);
      seen_ld |= 1u << ((src / 16) & 7);
      seen_st |= 1u << (((uint32_t)c * 16 / 16) & 7);
    }
    if (seen_ld != 0xFF || seen_st != 0xFF) ok = false;
  }
  printf("%s bank groups\n", ok ? "ok  " : "FAIL");
  return ok;
}

}  // namespace

int main(int argc, char **argv) {
  static_assert(leaf2::kSmemBytes <= 232448, "shared memory budget");
  bool ok = check_bank_groups();
  if (argc == 6) {
    ok &= run_case(atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]));
  } else {
    struct Case {
      int count, m, l, n, blocks;
    } const cases[] = {
        {1, 4096, 128, 256, 1},      // one tile, one slab
        {1, 4096, 256, 512, 3},      // stream-K: segments that start and end inside a tile's K range
        {1, 4096, 640, 256, 2},      // odd slab counts: ring parity across segments
        {1, 5000, 300, 700, 5},      // ragged m, l, n: zero fill on every edge, partial last words
        {1, 100, 64, 64, 4},         // far smaller than a tile
        {1, 8192, 128, 384, 148},    // more CTAs than units
        {3, 4096, 384, 320, 7},      // batch of products, n % 256 == 64
        {7, 4100, 130, 260, 11},     // full batch, one row / two columns / four bits past the tile edges
        {2, 4096, 1280, 256, 3},     // long K: many ring refills in one segment
    };
    for (auto const &c : cases) ok &= run_case(c.count, c.m, c.l, c.n, c.blocks);
  }
  printf("lds128 %ld sts128 %ld\n", leaf2::g_n_lds128.load(), leaf2::g_n_sts128.load());
  return ok ? 0 : 1;
}
