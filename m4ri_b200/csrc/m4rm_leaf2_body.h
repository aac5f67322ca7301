// m4rm_leaf2_body.h — body of the second-generation M4RM leaf ("tall tile": 4096 rows x 256 bits of C).
//
// Same job as m4rm_kernel.cu (the reference's _mzd_mul_m4rm hot loops, m4ri/brilliantrussian.c:1107-1178:
// mzd_make_table :163-211, mzd_read_bits mzd.h:892-901, _mzd_combine_8 xor_template.h), different tile:
// the leaf is bound by shared-memory bandwidth, and per A column a CTA pays
//     lookups  TM*W/8   +   table stores 256*W/8   +   table-build reads   +   A words
// bytes of shared-memory traffic for TM x 8W bits of C (W = tile row bytes, TM*W = 128 KB of registers).
// The first leaf (TM = 1024, W = 128) spends 20 % of its wavefronts on building tables; TM = 4096, W = 32
// cuts the build to 1/4 per C bit: 148 instead of 168 wavefronts per A column and 128 KB of C.
//
// What makes W = 32 work on 32 banks: FOUR tables (32 columns of A = one u32 per row) are interleaved
// in one 128-byte line per index value — piece (h, t) = 16-byte half h of the entry of table t sits at byte
// 64*h + 16*t — and lane i of every quarter-warp reads piece (h, t) = (i>>2, (i+jj)&3) in its jj-th load and
// the other half of the same entry in the next one: eight lanes, eight different 16-byte bank groups, for
// ANY eight index values.  Every LDS.128 wavefront therefore carries 128 useful bytes and serves 8 C rows.
//
// This header is compiled twice: by nvcc inside m4rm_leaf2.cu, and by g++ inside tests/c/emu_leaf2.cpp,
// where a CTA is 256 host threads, shared memory an array, TMA a host copy and the mbarriers/atomics are
// emulated — so the index arithmetic of this file is tested on the CPU (tests/test_leaf2_emu.py) without
// being restated.  The includer provides: L2_FN, U4, U2, TMap, lds128, lds64, sts128, prmt, mbar_wait,
// mbar_expect_tx, tma_load_2d, red_xor64, cta_sync, warp_sync; lds128 also as lds128<IMM>(addr) = [addr + IMM].
#pragma once
#include <stdint.h>

namespace leaf2 {

constexpr int kTM           = 4096;                       // C tile rows
constexpr int kTileBits     = 256;                        // C tile columns (one row = 32 B = two 16-byte pieces)
constexpr int kLineBytes    = 128;                        // one index value: 4 tables x 2 halves x 16 B
constexpr int kStepBufBytes = 256 * kLineBytes;           // k = 8: 32 KB of tables per step (32 A columns)
constexpr int kSlabBits     = 128;                        // K extent of one TMA slab
constexpr int kStepsPerSlab = kSlabBits / 32;             // 4
constexpr int kABoxRows     = 256;
constexpr int kAParts       = kTM / kABoxRows;            // 16 boxes of 256 rows x 16 B
constexpr int kASlabBytes   = kTM * 16;                   // 64 KB
// B arrives as one box per table: 8 rows x 128 B whose first column is 8*t words LEFT of the tile, so the
// 32 wanted bytes of table t sit 32*t bytes into each 128-byte box row.  That skew is what spreads the eight
// (t, h) pieces a quarter-warp reads over all eight 16-byte bank groups (TMA destinations must be 128-byte
// aligned, so the boxes themselves cannot be skewed); out-of-range (also negative) columns are zero-filled.
constexpr int kBBoxBytes    = 8 * 128;
constexpr int kBBoxes       = kStepsPerSlab * 4;          // 16 per slab
constexpr int kBSlabBytes   = kBBoxes * kBBoxBytes;       // 16 KB
constexpr int kOffTables    = 0;
constexpr int kOffA         = 2 * kStepBufBytes;
constexpr int kOffB         = kOffA + 2 * kASlabBytes;
constexpr int kOffBar       = kOffB + 2 * kBSlabBytes;
constexpr int kSmemBytes    = kOffBar + 64;               // 229 440 B  (limit 232 448)
constexpr uint32_t kSlabTxBytes = kASlabBytes + kBSlabBytes;
constexpr int kMaxBatch     = 7;

struct alignas(64) Args {
  TMap mapA[kMaxBatch];          // box 4 x u32 (16 B) x 256 rows
  TMap mapB[kMaxBatch];          // box 32 x u32 (128 B) x 8 rows
  unsigned long long *C[kMaxBatch];
  long long pitchC[kMaxBatch];   // words
  int m;                         // rows of A / C
  int nwordsC;                   // 64-bit words per C row that may be written
  int tiles_m;
  int tiles_n;
  int slabs;                     // ceil(l / 128)
  int nprob;
  long long units_per_problem;   // tiles_m * tiles_n * slabs
  long long total_units;
};

template <int N>
struct IntC {
  static constexpr int value = N;
};

L2_FN void xor4(U4 &d, U4 const &a) {
  d.x ^= a.x;
  d.y ^= a.y;
  d.z ^= a.z;
  d.w ^= a.w;
}
L2_FN void xor4(U4 &d, U4 const &a, U4 const &b) {   // one LOP3 per word
  d.x ^= a.x ^ b.x;
  d.y ^= a.y ^ b.y;
  d.z ^= a.z ^ b.z;
  d.w ^= a.w ^ b.w;
}

// Tables of one step (replaces mzd_make_table): thread -> piece column c = (h, t) = tid & 7 and a run of E
// consecutive index values; base = XOR of the B rows selected by the high index bits, then a reflected
// Gray walk over the low ones (one XOR + one STS.128 per entry, no table read-back).  A quarter-warp
// stores one complete 128-byte line per instruction and loads eight different bank groups.
template <int NT>
L2_FN void build_tables(uint32_t tbuf, uint32_t bstep, int tid) {
  constexpr int E  = 2048 / NT;
  constexpr int GB = E == 8 ? 3 : (E == 4 ? 2 : -1);
  static_assert(GB > 0, "unsupported thread count");
  int const c = tid & 7, t = c & 3, h = c >> 2, g = tid >> 3;
  uint32_t const src = bstep + t * (kBBoxBytes + 32) + h * 16;
  U4 low[GB];
#pragma unroll
  for (int b = 0; b < GB; ++b) low[b] = lds128(src + b * 128);
  U4 e = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int b = GB; b < 8; ++b) {
    uint32_t const mk = 0u - (((uint32_t)g >> (b - GB)) & 1u);
    U4 const v = lds128(src + b * 128);
    e.x ^= v.x & mk;
    e.y ^= v.y & mk;
    e.z ^= v.z & mk;
    e.w ^= v.w & mk;
  }
  uint32_t const dst = tbuf + (uint32_t)(g << GB) * kLineBytes + c * 16;
  sts128(dst, e);
#pragma unroll
  for (int i = 1; i < E; ++i) {
    xor4(e, low[(i & 1) ? 0 : ((i & 2) ? 1 : 2)]);
    sts128(dst + (i ^ (i >> 1)) * kLineBytes, e);
  }
}

// Lookups of one step for one C row (replaces mzd_read_bits + _mzd_combine_8): a = the row's 32 A bits with
// its bytes already rotated by the lane's table phase, so byte jj indexes table (i8 + jj) & 3 — the table
// whose piece this lane reads in its jj-th load (base[jj] = lane part of the address, IMM = table buffer).
// acc0 is the lane's "own" half (hl), acc1 the other one.
template <int IMM>
L2_FN void lookup_row(U4 &acc0, U4 &acc1, uint32_t a, uint32_t const (&base)[4]) {
  uint32_t ad[4];
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) ad[jj] = base[jj] + prmt(a, 0u, 0x4440u + jj) * kLineBytes;
  U4 const v0 = lds128<IMM>(ad[0]), v1 = lds128<IMM>(ad[1]), v2 = lds128<IMM>(ad[2]), v3 = lds128<IMM>(ad[3]);
  xor4(acc0, v0, v1);
  xor4(acc0, v2, v3);
  U4 const w0 = lds128<IMM>(ad[0] ^ 64u), w1 = lds128<IMM>(ad[1] ^ 64u), w2 = lds128<IMM>(ad[2] ^ 64u),
           w3 = lds128<IMM>(ad[3] ^ 64u);
  xor4(acc1, w0, w1);
  xor4(acc1, w2, w3);
}

// The persistent stream-K CTA.  sbase = shared-memory address of the dynamic segment (1024-byte aligned),
// mbarriers at sbase + kOffBar already initialised (count 1) and visible to all threads.
template <int NT, int AWIDE>
L2_FN void cta_body(Args const &p, uint32_t sbase, int tid, int bid, int nblocks) {
  constexpr int RT = kTM / NT;                      // rows per thread, each as two 16-byte pieces
  uint32_t const sTab = sbase + kOffTables;
  uint32_t const sA   = sbase + kOffA;
  uint32_t const sB   = sbase + kOffB;
  uint32_t const sBar = sbase + kOffBar;
  int const warp = tid >> 5, lane = tid & 31;
  int const i8 = lane & 7, hl = i8 >> 2;

  // lane constants: byte rotation of the A words (table phase) and the lane part of the lookup addresses
  uint32_t const rot = ((uint32_t)(i8 + 0) & 3u) | (((uint32_t)(i8 + 1) & 3u) << 4) | (((uint32_t)(i8 + 2) & 3u) << 8) |
                       (((uint32_t)(i8 + 3) & 3u) << 12);
  uint32_t base[4];
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) base[jj] = sTab + (uint32_t)hl * 64u + ((uint32_t)(i8 + jj) & 3u) * 16u;

  long long const u_begin = p.total_units * (long long)bid / nblocks;
  long long const u_end   = p.total_units * (long long)(bid + 1) / nblocks;
  uint32_t parity0 = 0, parity1 = 0;                // phase of each ring slot

  for (long long u = u_begin; u < u_end;) {
    int const prob      = (int)(u / p.units_per_problem);
    long long const v   = u - (long long)prob * p.units_per_problem;
    int const tile      = (int)(v / p.slabs);
    int const s0        = (int)(v % p.slabs);
    TMap const *mapA = &p.mapA[prob], *mapB = &p.mapB[prob];
    int nseg            = p.slabs - s0;
    if ((long long)nseg > u_end - u) nseg = (int)(u_end - u);
    int const tm = tile % p.tiles_m, tn = tile / p.tiles_m;
    int const row0 = tm * kTM;

    // warp 0: TMA for slab s0+i into ring slot i&1 — lane 0 arms the barrier, lanes 0..15 each fetch one
    // A box (256 rows) and one B box (table lane&3 of step lane>>2)
    auto issue = [&](int i) {
      uint32_t const slot = (uint32_t)i & 1u;
      uint32_t const bar  = sBar + 8u * slot;
      if (lane == 0) mbar_expect_tx(bar, kSlabTxBytes);
      warp_sync();
      if (lane < kAParts) {
        tma_load_2d(sA + slot * kASlabBytes + lane * (kABoxRows * 16), mapA, (s0 + i) * 4, row0 + lane * kABoxRows, bar);
        tma_load_2d(sB + slot * kBSlabBytes + lane * kBBoxBytes, mapB, tn * 8 - 8 * (lane & 3),
                    (s0 + i) * kSlabBits + lane * 8, bar);
      }
    };

    if (warp == 0) {
      issue(0);
      if (nseg > 1) issue(1);
    }

    U4 acc[RT][2];
#pragma unroll
    for (int j = 0; j < RT; ++j) acc[j][0] = acc[j][1] = U4{0u, 0u, 0u, 0u};

    mbar_wait(sBar, parity0);
    parity0 ^= 1;
    build_tables<NT>(sTab, sB, tid);
    cta_sync();

    for (int i = 0; i < nseg; ++i) {
      uint32_t const slot = (uint32_t)i & 1u;
      uint32_t const bS = sB + slot * kBSlabBytes;
      // A bits of SP steps per load: one u32 per row and step, bytes rotated by the lane's table phase.
      // AWIDE = 0: an LDS.64 per row and half slab (4 wavefronts per 32 rows for 2 steps);
      // AWIDE = 1: an LDS.128 per row and slab (4 wavefronts for 4 steps, but 32 more live registers).
      constexpr int SP = AWIDE ? 4 : 2;
      auto part = [&](auto P_) {
        constexpr int P = decltype(P_)::value;
        uint32_t aw[RT][SP];
        auto step = [&](auto S_) {
          constexpr int S = decltype(S_)::value;          // step within the slab; table buffer = S & 1
          uint32_t const tnext = sTab + ((S & 1) ^ 1) * kStepBufBytes;
          // ---- tables of the next step into the other buffer ----
          if constexpr (S < kStepsPerSlab - 1) {
            build_tables<NT>(tnext, bS + (S + 1) * 4 * kBBoxBytes, tid);
          } else if (i + 1 < nseg) {
            if (slot == 0) { mbar_wait(sBar + 8, parity1); parity1 ^= 1; }
            else           { mbar_wait(sBar, parity0);     parity0 ^= 1; }
            build_tables<NT>(tnext, sB + (slot ^ 1u) * kBSlabBytes, tid);
          }
          // ---- A bits of the next SP steps (after the build: its registers are dead by now) ----
          if constexpr (S % SP == 0) {
#pragma unroll
          for (int j = 0; j < RT; ++j) {
            uint32_t const arow = sA + slot * kASlabBytes + (uint32_t)(j * NT + tid) * 16u;
            if constexpr (AWIDE) {
              U4 const r = lds128(arow);
              aw[j][0] = prmt(r.x, 0u, rot);
              aw[j][1] = prmt(r.y, 0u, rot);
              aw[j][2] = prmt(r.z, 0u, rot);
              aw[j][3] = prmt(r.w, 0u, rot);
            } else {
              U2 const r = lds64(arow + P * 8u);
              aw[j][0] = prmt(r.x, 0u, rot);
              aw[j][1] = prmt(r.y, 0u, rot);
            }
          }
          }
          // ---- lookups (rows past m carry zero-filled A bits -> line 0 = zeros; no branch needed) ----
#pragma unroll
          for (int j = 0; j < RT; ++j)
            lookup_row<(S & 1) * kStepBufBytes>(acc[j][0], acc[j][1], aw[j][S % SP], base);
          cta_sync();
        };
        step(IntC<SP * P>{});
        step(IntC<SP * P + 1>{});
        if constexpr (SP == 4) {
          step(IntC<SP * P + 2>{});
          step(IntC<SP * P + 3>{});
        }
      };
      part(IntC<0>{});
      if constexpr (SP == 2) part(IntC<1>{});
      // ring slot `slot` is free again: refill it with slab i+2
      if (warp == 0 && i + 2 < nseg) issue(i + 2);
    }

    // ---- merge the partial tile into C (exact: XOR is associative and commutative) ----
#pragma unroll
    for (int j = 0; j < RT; ++j) {
      int const row = row0 + j * NT + tid;
      if (row < p.m) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          int const wcol = tn * (kTileBits / 64) + (hl ^ hh) * 2;
          unsigned long long *dst = p.C[prob] + (long long)row * p.pitchC[prob] + wcol;
          if (wcol < p.nwordsC) red_xor64(dst, acc[j][hh].x, acc[j][hh].y);
          if (wcol + 1 < p.nwordsC) red_xor64(dst + 1, acc[j][hh].z, acc[j][hh].w);
        }
      }
    }
    // all table/slab reads of this segment are complete before the next segment's prologue
    cta_sync();
    u += nseg;
  }
}

}  // namespace leaf2
