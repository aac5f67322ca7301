// m4rm_leaf2_body.h — body of the second-generation M4RM leaf ("tall tile": 4096 rows x 256 bits of C).
//
// Same job as m4rm_kernel.cu (the reference's _mzd_mul_m4rm hot loops, m4ri/brilliantrussian.c:1107-1178:
// mzd_make_table :163-211, mzd_read_bits mzd.h:892-901, _mzd_combine_8 xor_template.h), different tile:
// the leaf is bound by shared-memory bandwidth, and per A column a CTA pays
//     lookups  TM*W/8   +   table stores 256*W/8   +   table-build reads   +   A words
// bytes of shared-memory traffic for TM x 8W bits of C (W = tile row bytes, TM*W = 128 KB of registers).
// The first leaf (TM = 1024, W = 128) spends 20 % of its wavefronts on building tables; TM = 4096, W = 32
// cuts the build to 1/4 per C bit: 152 instead of 168 wavefronts per A column and 128 KB of C.
//
// What makes W = 32 work on 32 banks: FOUR tables (32 columns of A = one u32 per row) are interleaved
// in one 128-byte line per index value — piece (h, t) = 16-byte half h of the entry of table t sits at byte
// 64*h + 16*t — and lane i of every quarter-warp reads piece (h, t) = (i>>2, (i+jj)&3) in its jj-th load and
// the other half of the same entry in the next one: eight lanes, eight different 16-byte bank groups, for
// ANY eight index values.  Every LDS.128 wavefront therefore carries 128 useful bytes and serves 8 C rows.
//
// Operands: A by TMA (one 3D box per 128-column slab and 4096-row tile), B by cp.async (4 KB per slab: one
// 16-byte piece per thread, placed directly in the layout the table build reads), both zero-filled outside
// the matrices, which is what makes ragged m / l / n edges free.  Stream-K over (problem, tile, slab) units
// with red.global.xor merges, as in the first leaf.
//
// This header is compiled twice: by nvcc inside m4rm_leaf2.cu, and by g++ inside tests/c/emu_leaf2.cpp,
// where a CTA is 256 host threads, shared memory an array, TMA / cp.async host copies and the mbarriers and
// atomics are emulated — so the index arithmetic of this file is tested on the CPU (tests/test_leaf2_emu.py)
// without being restated.  The includer provides: L2_FN, U4, U2, TMap, lds128, lds64, sts128, prmt,
// mbar_wait, mbar_expect_tx, tma_load_2d, tma_load_3d, cp_async16, cp_async_wait_all, red_xor64, stg128, cta_sync,
// gate(a, b, c) = a | (b & c); lds128 also as lds128<IMM>(addr) = [addr + IMM].
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
// B is only 4 KB per slab and CTA (128 rows x 32 B), so it does not come by TMA: every thread fetches ONE
// 16-byte piece per slab with cp.async (LDGSTS), which can put it anywhere — straight into the layout the
// table build wants: piece (step s, row-in-table b, half h, table t) at s*1024 + b*128 + h*64 + t*16, i.e. the
// eight (h, t) pieces a quarter-warp reads for one b form one 128-byte line (all eight bank groups), with
// no over-fetch and no extra TMA instructions (a TMA issue costs its warp several hundred cycles).
constexpr int kBStepBytes   = 8 * kLineBytes;             // 1 KB
constexpr int kBSlabBytes   = kStepsPerSlab * kBStepBytes;   // 4 KB
constexpr int kOffTables    = 0;
constexpr int kOffA         = 2 * kStepBufBytes;
constexpr int kOffB         = kOffA + 2 * kASlabBytes;
constexpr int kOffBar       = kOffB + 2 * kBSlabBytes;
constexpr int kOffSeg       = kOffBar + 16;               // {prob, tn, row0, s0} of the current segment
constexpr int kSmemBytes    = kOffBar + 64;               // 204 864 B  (limit 232 448)
constexpr uint32_t kSlabTxBytes = kASlabBytes;
constexpr int kMaxBatch     = 49;                        // two Strassen levels in one launch (7.9 KB of kernel parameters)

struct alignas(64) Args {
  TMap mapA[kMaxBatch];          // a3d: 3D (words, 256 rows, row groups), box 4 x 256 x 16 = one slab of a tile;
                                 // else 2D (words, rows), box 4 x 256
  unsigned long long const *B[kMaxBatch];
  long long pitchB[kMaxBatch];   // words (even; bits past the last column up to the pitch are zero)
  unsigned long long *C[kMaxBatch];
  long long pitchC[kMaxBatch];   // words
  int m;                         // rows of A / C
  int l;                         // rows of B
  int a3d;                       // m % 256 == 0: a tile's A slab is ONE 3D TMA box
  int nwordsC;                   // 64-bit words per C row that may be written
  int tiles_m;
  int tiles_n;
  int slabs;                     // ceil(l / 128)
  int nprob;
  long long units_per_problem;   // tiles_m * tiles_n * slabs
  long long total_units;         // < 2^31
  int dp_rounds;                 // whole tiles per CTA handled round-robin before the stream-K tail (see cta_body)
  int store_dp;                  // 1: C = A*B — the whole-tile rounds STORE their tile (C need not be initialised there);
                                 //    only the tiles of the stream-K tail are merged with red.xor (into zeros)
  unsigned zero;                 // 0 — a value the compiler cannot know (see gate())
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
// NB = number of building threads: NT (every thread builds 8 entries) or NT/2 (the warps of one half of
// the CTA — `half`, alternating from step to step — build 16 entries each: every thread loads all 8 B rows
// of its piece column either way, so half the builders means half the B-row wavefronts).
template <int NT, int NB>
L2_FN void build_tables(uint32_t tbuf, uint32_t bstep, int tid, int half) {
  constexpr int E  = 2048 / NB;
  constexpr int GB = E == 16 ? 4 : (E == 8 ? 3 : -1);
  static_assert(GB > 0 && (NB == NT || 2 * NB == NT), "unsupported builder count");
  if (NB != NT && (tid / NB) != half) return;       // warp-uniform
  int const bt = tid % NB;
  int const c = bt & 7, g = bt >> 3;                // c = 4*h + t
  uint32_t const src = bstep + c * 16;
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
    xor4(e, low[(i & 1) ? 0 : ((i & 2) ? 1 : ((i & 4) ? 2 : 3))]);
    sts128(dst + (i ^ (i >> 1)) * kLineBytes, e);
  }
}

// `dep` chains the rows of a thread: the A word of a row is OR-ed with (dep & zero) — still the A word, but
// now data dependent on the last table line of the previous row — so a warp never has more than one row's
// eight LDS.128 (32 registers) in flight.  Without the chain ptxas schedules for single-warp latency, keeps
// ~12 loads in flight and spills accumulators to make room; with 224 KB of shared memory there is hardly
// any L1 left, so every spilled word costs an L2 round trip.  Eight warps x 8 loads still oversubscribe
// the data pipe (one LDS.128 per 4 clocks) several times.
template <int IMM>
L2_FN void lookup_row(U4 &acc0, U4 &acc1, uint32_t a, uint32_t const (&base)[4], uint32_t &dep, uint32_t zero) {
  a = gate(a, dep, zero);
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
  // every word of this row's sixteen table pieces feeds dep, so none of them can stay unconsumed
  dep = (acc0.x ^ acc0.y ^ acc0.z) ^ (acc0.w ^ acc1.x ^ acc1.y) ^ (acc1.z ^ acc1.w);
}

// A: TMA for K-slab `kslab` (128 columns) of the tile rows [row0, row0 + 4096) into ring slot `slot`, issued
// by thread 0: one 3D box when the row count allows it (the 3D view groups rows by 256, so its last group
// must be complete), else sixteen 2D boxes.  Rows / columns outside the matrix arrive as zeros.
L2_FN void issue_a(Args const &p, uint32_t sbase, int tid, int prob, int row0, int kslab, uint32_t slot) {
  if (tid != 0) return;
  uint32_t const bar = sbase + kOffBar + 8u * slot;
  uint32_t const dst = sbase + kOffA + slot * kASlabBytes;
  mbar_expect_tx(bar, kSlabTxBytes);
  if (p.a3d) {
    tma_load_3d(dst, &p.mapA[prob], kslab * 4, 0, row0 / kABoxRows, bar);
  } else {
    for (int part = 0; part < kAParts; ++part)
      tma_load_2d(dst + part * (kABoxRows * 16), &p.mapA[prob], kslab * 4, row0 + part * kABoxRows, bar);
  }
}

// B: this thread's 16-byte piece of K-slab `kslab`, tile column tn, into ring slot `slot` (see kBStepBytes).
// Rows >= l and pieces past the row pitch are zero-filled (cp.async with a source size of 0).
template <int NT>
L2_FN void prefetch_b(Args const &p, uint32_t sbase, int tid, int prob, int tn, int kslab, uint32_t slot) {
  static_assert(NT == 256, "one 16-byte piece of the 4 KB B slab per thread");
  int const h = tid & 1, t = (tid >> 1) & 3, b = (tid >> 3) & 7, st = tid >> 6;
  int const row = kslab * kSlabBits + st * 32 + t * 8 + b;
  long long const piece = (long long)tn * 2 + h;                 // 16-byte pieces along the row
  bool const valid = row < p.l && piece * 2 < p.pitchB[prob];
  unsigned long long const *src = p.B[prob] + (valid ? (long long)row * p.pitchB[prob] + piece * 2 : 0);
  cp_async16(sbase + kOffB + slot * kBSlabBytes + st * kBStepBytes + b * kLineBytes + (h * 4 + t) * 16, src, valid ? 16u : 0u);
}

// The persistent stream-K CTA.  sbase = shared-memory address of the dynamic segment (1024-byte aligned),
// mbarriers at sbase + kOffBar already initialised (count 1) and visible to all threads.
//
// Register budget: 128 registers of C and 32 of A bits leave little room, and a spilled value costs an L2
// round trip here (the 224 KB of shared memory leave almost no L1), so the per-segment scalars live in
// shared memory (kOffSeg) and the ring state is ONE counter: K-slab number n of this CTA uses ring slot
// n & 1 and mbarrier phase parity (n >> 1) & 1.
template <int NT, int AWIDE, int SPLIT>
L2_FN void cta_body(Args const &p, uint32_t sbase, int tid, int bid, int nblocks) {
  constexpr int RT = kTM / NT;                      // rows per thread, each as two 16-byte pieces
  constexpr int NB = SPLIT ? NT / 2 : NT;           // table-building threads per step
  uint32_t const sTab = sbase + kOffTables;
  uint32_t const sA   = sbase + kOffA;
  uint32_t const sB   = sbase + kOffB;
  uint32_t const sBar = sbase + kOffBar;
  uint32_t const sSeg = sbase + kOffSeg;
  int const i8 = tid & 7, hl = i8 >> 2;

  // lane constants: byte rotation of the A words (table phase) and the lane part of the lookup addresses
  uint32_t const rot = ((uint32_t)(i8 + 0) & 3u) | (((uint32_t)(i8 + 1) & 3u) << 4) | (((uint32_t)(i8 + 2) & 3u) << 8) |
                       (((uint32_t)(i8 + 3) & 3u) << 12);
  uint32_t base[4];
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) base[jj] = sTab + (uint32_t)hl * 64u + ((uint32_t)(i8 + jj) & 3u) * 16u;

  int const total = (int)p.total_units, upp = (int)p.units_per_problem;      // < 2^31, checked by the launcher
  uint32_t ring = 0;                                // K-slabs consumed so far by this CTA
  uint32_t dep = 0;                                 // see lookup_row()

  // Work partition.  Units are ordered (problem, tile, slab).  The first dp_rounds * nblocks tiles are dealt out whole
  // and round-robin — in round r CTA b owns tile b + r * nblocks, all CTAs start their tile at slab 0 together, and
  // neighbouring CTAs hold neighbouring column tiles of the SAME product, so an A panel is fetched from DRAM once
  // and then served from L2 to the other 15 tiles of its product (with pure stream-K every CTA sat at a different
  // slab of a different product and a 49-product launch read 2.1 GB for 0.4 GB of operands).  The remaining tiles
  // are cut into equal contiguous unit ranges (stream-K), whose partial tiles merge with red.xor as before.
  for (int region = 0; region <= p.dp_rounds; ++region) {
  int u, u_end;
  if (region < p.dp_rounds) {
    u = (bid + region * nblocks) * p.slabs;
    u_end = u + p.slabs;
  } else {
    int const tail0 = p.dp_rounds * nblocks * p.slabs, tail = total - tail0;
    u = tail0 + (int)((long long)tail * bid / nblocks);
    u_end = tail0 + (int)((long long)tail * (bid + 1) / nblocks);
  }

  while (u < u_end) {
    int nseg;
    {
      int const prob = u / upp, v = u - prob * upp;
      int const tile = v / p.slabs, s0 = v - tile * p.slabs;
      int const tm = tile % p.tiles_m, tn = tile / p.tiles_m;
      nseg = p.slabs - s0;
      if (nseg > u_end - u) nseg = u_end - u;
      if (tid == 0) sts128(sSeg, U4{(uint32_t)prob, (uint32_t)tn, (uint32_t)(tm * kTM), (uint32_t)s0});
      issue_a(p, sbase, tid, prob, tm * kTM, s0, ring & 1u);
      if (nseg > 1) issue_a(p, sbase, tid, prob, tm * kTM, s0 + 1, (ring + 1u) & 1u);
      prefetch_b<NT>(p, sbase, tid, prob, tn, s0, ring & 1u);
      cp_async_wait_all();
    }

    U4 acc[RT][2];
#pragma unroll
    for (int j = 0; j < RT; ++j) acc[j][0] = acc[j][1] = U4{0u, 0u, 0u, 0u};

    cta_sync();                                     // every thread's piece of the first B slab has landed
    build_tables<NT, NB>(sTab, sB + (ring & 1u) * kBSlabBytes, tid, 1);
    cta_sync();

    for (int i = 0; i < nseg; ++i) {
      uint32_t const n = ring + (uint32_t)i, slot = n & 1u;
      uint32_t const bS = sB + slot * kBSlabBytes;
      // B of the next slab: its slot was last read by the table builds of the previous slab's third step;
      // the copy is awaited before the barrier that ends step 2 and consumed by the build of step 3
      if (i + 1 < nseg) {
        U4 const seg = lds128(sSeg);
        prefetch_b<NT>(p, sbase, tid, (int)seg.x, (int)seg.y, (int)seg.w + i + 1, slot ^ 1u);
      }
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
            build_tables<NT, NB>(tnext, bS + (S + 1) * kBStepBytes, tid, S & 1);
          } else if (i + 1 < nseg) {
            build_tables<NT, NB>(tnext, sB + (slot ^ 1u) * kBSlabBytes, tid, S & 1);
          }
          // ---- A bits of the next SP steps (after the build: its registers are dead by now) ----
          if constexpr (S == 0) mbar_wait(sBar + 8u * slot, (n >> 1) & 1u);      // this slab's A has landed
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
            lookup_row<(S & 1) * kStepBufBytes>(acc[j][0], acc[j][1], aw[j][S % SP], base, dep, p.zero);
          if constexpr (S == kStepsPerSlab - 2) cp_async_wait_all();
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
      // ring slot `slot` is free again: refill it with slab i+2 of the segment
      if (i + 2 < nseg) {
        U4 const seg = lds128(sSeg);
        issue_a(p, sbase, tid, (int)seg.x, (int)seg.z, (int)seg.w + i + 2, slot);
      }
    }

    // ---- merge the partial tile into C (exact: XOR is associative and commutative) ----
    {
      U4 const seg = lds128(sSeg);
      int const prob = (int)seg.x, tn = (int)seg.y, row0 = (int)seg.z;
      if (p.store_dp && region < p.dp_rounds) {
        // this CTA has multiplied the tile's whole K range: plain 16-byte stores (the launcher only selects this mode
        // for C rows that end on a 128-bit boundary, so both words of a piece are inside the matrix)
#pragma unroll
        for (int j = 0; j < RT; ++j) {
          int const row = row0 + j * NT + tid;
          if (row < p.m) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              int const wcol = tn * (kTileBits / 64) + (hl ^ hh) * 2;
              if (wcol < p.nwordsC) stg128(p.C[prob] + (long long)row * p.pitchC[prob] + wcol, acc[j][hh]);
            }
          }
        }
      } else
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
    }
    // all table/slab/segment reads of this segment are complete before the next segment's prologue
    cta_sync();
    ring += (uint32_t)nseg;
    u += nseg;
  }
  }   // region
}

}  // namespace leaf2
