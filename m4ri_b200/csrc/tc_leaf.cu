// tc_leaf.cu — the tensor-core leaf: C (^)= A * B over GF(2) on the 5th-generation tensor cores (sm_100a).
//
// The bits of A (rows) and of B (columns) are expanded to e2m1 nibbles (0 -> 0x0, 1 -> 0x2 = 1.0), multiplied with
// tcgen05.mma.kind::mxf4.block_scale (unit ue8m0 scales, fp32 accumulators in TMEM; exact for 0/1 operands: every partial
// sum is an integer < 2^24 — profiles/r02_tensor_core_question.md), and the parity of every accumulator is packed back
// into the bit-packed C.  Replaces, for the shapes it suits, the M4RM leaf of brilliantrussian.c:999-1190 with the same
// result bits (tests/test_zz5_tensor_leaf_gpu.py, the reference digests of tests/golden/large_golden.json).
//
// Two forms live here.  FIRST a small output-stationary, single-buffered kernel (m4ri_b200_dmul_tc only) that expands in
// shared memory — it validated the data path (nibble layout, descriptors, mbarrier phases, parity epilogue) and is kept as
// an independent cross-check; shapes m % 128 == 0, n % 256 == 0, l % 128 == 0, Bt = B transposed (n x l).  Then the
// production kernel (second half of the file).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "dev.h"

namespace m4b {
namespace {

constexpr int kM = 128, kN = 256, kKBytes = 64;            // one K chunk: 128 elements = 64 bytes of e2m1 per row
constexpr int kSBO = kKBytes / 16 * 128;                   // bytes between 8-row groups of the canonical no-swizzle layout

__device__ __forceinline__ uint32_t smem_u32(void const *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;              // LBO: next core matrix along K
  d |= (uint64_t)(((uint32_t)kSBO >> 4) & 0x3FFF) << 32;    // SBO: next 8-row group
  d |= (uint64_t)1 << 46;                                    // descriptor version (Blackwell)
  return d;
}

// eight bits -> eight e2m1 nibbles (bit e -> nibble e: 0x2 where the bit is set)
__device__ __forceinline__ uint32_t expand8(uint32_t b) {
  uint32_t x = b & 0xFFu;
  x = (x | (x << 12)) & 0x000F000Fu;
  x = (x | (x << 6)) & 0x03030303u;
  x = (x | (x << 3)) & 0x11111111u;
  return x << 1;
}

// rows [row0, row0 + nrows) of a bit-packed matrix, K bits [k0, k0 + 128) -> the shared-memory operand image
__device__ __forceinline__ void expand_rows(uint8_t *simg, word const *M, long long pitch, int row0, int nrows, int k0, int tid,
                                            int nthreads) {
  for (int r = tid; r < nrows; r += nthreads) {
    word const *src = M + (long long)(row0 + r) * pitch + k0 / 64;
    unsigned long long const w0 = src[0], w1 = src[1];
    uint8_t *dst = simg + (r / 8) * kSBO + (r % 8) * 16;
#pragma unroll
    for (int beta = 0; beta < 16; ++beta) {                 // packed byte beta -> e2m1 bytes 4 beta .. 4 beta + 3
      uint32_t const b = (uint32_t)((beta < 8 ? w0 >> (8 * beta) : w1 >> (8 * (beta - 8))) & 0xFFull);
      int const kb = 4 * beta;
      *reinterpret_cast<uint32_t *>(dst + (kb / 16) * 128 + kb % 16) = expand8(b);
    }
  }
}

__global__ void __launch_bounds__(128, 1) tc_leaf_simple_kernel(word *C, long long pitchC, word const *A, long long pitchA,
                                                                word const *Bt, long long pitchB, int l, int accumulate) {
  __shared__ __align__(1024) uint8_t sA[kM * kKBytes];
  __shared__ __align__(1024) uint8_t sB[kN * kKBytes];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  int const tid = threadIdx.x, warp = tid >> 5;
  int const m0 = blockIdx.y * kM, n0 = blockIdx.x * kN;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t const tmem = tmem_base;
  {   // unit scales: every byte of TMEM columns 256..511 = 0x7F (ue8m0 2^0)
    uint32_t const v = 0x7F7F7F7Fu;
    for (int c = 256; c < 512; c += 8) {
      uint32_t const taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  uint32_t const idesc = (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | (1u << 23) | ((uint32_t)(kM >> 4) << 24);
  int const chunks = l / 128;
  for (int ch = 0; ch < chunks; ++ch) {
    expand_rows(sA, A, pitchA, m0, kM, ch * 128, tid, 128);
    expand_rows(sB, Bt, pitchB, n0, kN, ch * 128, tid, 128);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
      uint64_t const da = make_desc(smem_u32(sA)), db = make_desc(smem_u32(sB));
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint64_t const a = da + (uint64_t)((ks * 256) >> 4), b = db + (uint64_t)((ks * 256) >> 4);
        uint32_t const acc = (ch | ks) ? 1u : 0u;
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                     "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n}"
                     ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(tmem + 256u), "r"(tmem + 384u) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {   // the MMAs of this chunk have read the operand images: they may be overwritten (single buffer)
      uint32_t const b = smem_u32(&bar), parity = (uint32_t)(ch & 1);
      asm volatile(
          "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(b),
          "r"(parity)
          : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // ---- epilogue: parity of every accumulator -> one bit of C.  Warp w owns TMEM lanes (= tile rows) 32 w .. 32 w + 31.
  uint32_t *crow = reinterpret_cast<uint32_t *>(C + (long long)(m0 + tid) * pitchC) + n0 / 32;
  for (int c = 0; c < kN; c += 32) {
    uint32_t v[32];
    uint32_t const taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) bits |= ((uint32_t)__float2int_rn(__uint_as_float(v[j])) & 1u) << j;
    if (accumulate) crow[c / 32] ^= bits;
    else            crow[c / 32] = bits;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace

// C (m x n) (^)= A (m x l) * Bt^T, Bt = B transposed (n x l); experimental, see the file header
void launch_tc_leaf_simple(DView C, DView A, DView Bt, bool accumulate, cudaStream_t s) {
  if (A.nrows % kM || Bt.nrows % kN || A.ncols % 128 || A.ncols != Bt.ncols || C.nrows != A.nrows || C.ncols != Bt.nrows)
    die("m4ri_b200_dmul_tc: needs m %% 128 == 0, n %% 256 == 0, l %% 128 == 0 and Bt = B^T\n");
  dim3 const grid(Bt.nrows / kN, A.nrows / kM);
  tc_leaf_simple_kernel<<<grid, 128, 0, s>>>(C.data, C.pitch, A.data, A.pitch, Bt.data, Bt.pitch, A.ncols, accumulate ? 1 : 0);
  M4B_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}

}  // namespace m4b

// =====================================================================================================================
// The production form: B-stationary, pipelined, warp-specialised.  The operands are expanded to
// e2m1 ONCE by a streaming pre-pass into "images" whose pieces are exactly the shared-memory operand layout, so the main
// kernel moves them with plain bulk copies (cp.async.bulk + mbarrier complete_tx):
//   A image:  [row tile of 128][K chunk of 1024][sub 0..3] x 16 KB   (sub-image = 128 rows x 256 elements)
//   B image:  [column panel of 256][K chunk][sub 0..3]     x 32 KB   (rows = columns of B: the expansion transposes)
// A job = (product, K chunk, column panel): the 128 KB panel chunk stays in shared memory while EVERY row tile of A
// streams past it through a ring of three 32 KB stages (two sub-images each): 4 KB of L2 traffic per 128-cycle MMA pair =
// 31 B/clk/SM, well under the L2 slice limit (an output-stationary tile needs 3x that).  Per row tile: 16 K-steps into two
// of three 128-column TMEM regions (N = 128 halves for the first stage, N = 256 for the second, see the MMA warp); eight
// epilogue warps drain them (parity by the 2^23 trick) and XOR the result bits into C with red.global — partial sums
// over K chunks combine by XOR, so C = A*B starts from a cleared C and C ^= A*B from C itself.
// =====================================================================================================================
namespace m4b {
namespace {

constexpr int kSubBytes = 16384;      // A sub-image: 128 rows x 256 elements
constexpr int kStageBytes = 2 * kSubBytes, kBSubBytes = 32768, kSubs = 4, kAStages = 3;   // an A stage = two sub-images
constexpr int kPanelBytes = kSubs * kBSubBytes;
constexpr int kTc2Threads = 320;
constexpr int kTc2Smem = kPanelBytes + kAStages * kStageBytes + 256 + 1024;   // + barriers + alignment slack
constexpr int kTcMaxBatch = 49;

struct Tc2Args {
  word *C[kTcMaxBatch];
  uint8_t const *imgA[kTcMaxBatch];
  uint8_t const *imgB[kTcMaxBatch];
  int pitchC[kTcMaxBatch];
  int count, mtiles, nkc, npanels, flags;   // flags: experiment bits (1 skip the parity arithmetic, 2 skip the TMEM drain: timing only)
};

__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  long long const t0 = clock64();
  for (uint32_t spins = 1;; ++spins) {
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
    if ((spins & 0xFFFu) == 0 && clock64() - t0 > 4000000000ll) __trap();     // a protocol bug must not hang the box
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  if (!done) mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, void const *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}

__global__ void __launch_bounds__(kTc2Threads, 1) tc_leaf2_kernel(const __grid_constant__ Tc2Args args) {
  extern __shared__ uint8_t smem_raw[];
  uint32_t const base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint32_t const sB = base, sA = base + kPanelBytes, bars = sA + kAStages * kStageBytes;
  // barriers (8 bytes each): full_a[5] empty_a[5] full_b[4] empty_b[4] acc_full[3] acc_empty[3]
  auto full_a = [&](int s) { return bars + 8u * s; };
  auto empty_a = [&](int s) { return bars + 8u * (kAStages + s); };
  auto full_b = [&](int s) { return bars + 8u * (2 * kAStages + s); };
  auto empty_b = [&](int s) { return bars + 8u * (2 * kAStages + kSubs + s); };
  auto acc_full = [&](int b) { return bars + 8u * (2 * kAStages + 2 * kSubs + b); };
  auto acc_empty = [&](int b) { return bars + 8u * (2 * kAStages + 2 * kSubs + 3 + b); };
  __shared__ uint32_t tmem_slot;
  int const tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kAStages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full_a(s)));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(empty_a(s)));
    }
    for (int s = 0; s < kSubs; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full_b(s)));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(empty_b(s)));
    }
    for (int b = 0; b < 3; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(acc_full(b)));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(acc_empty(b)));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t const tmem = tmem_slot;
  if (warp >= 2 && warp < 6) {   // unit ue8m0 scales in columns 384..511 of every lane
    uint32_t const v = 0x7F7F7F7Fu;
    for (int c = 384; c < 512; c += 8) {
      uint32_t const taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  int const jobs_per_product = args.nkc * args.npanels, njobs = args.count * jobs_per_product;
  if (warp == 0) {
    if (lane == 0) {             // ---------------- producer ----------------
      uint32_t a_stage = 0, a_phase = 0, ji = 0;
      for (int job = blockIdx.x; job < njobs; job += gridDim.x, ++ji) {
        int const p = job / jobs_per_product, rem = job % jobs_per_product, kc = rem / args.npanels, np = rem % args.npanels;
        uint8_t const *srcB = args.imgB[p] + ((long long)(np * args.nkc + kc) * kSubs) * kBSubBytes;
        auto load_panel_half = [&](int s0) {     // two sub-images of the panel as soon as the previous job has let go of them
          for (int s = s0; s < s0 + 2; ++s) {
            mbar_wait(empty_b(s), (ji & 1u) ^ 1u);
            mbar_expect_tx(full_b(s), kBSubBytes);
            bulk_g2s(sB + s * kBSubBytes, srcB + (long long)s * kBSubBytes, kBSubBytes, full_b(s));
          }
        };
        load_panel_half(0);
        {                                        // ... and the next job's panel into L2 meanwhile
          int const nj = job + gridDim.x;
          if (nj < njobs) {
            int const p2 = nj / jobs_per_product, r2 = nj % jobs_per_product, kc2 = r2 / args.npanels, np2 = r2 % args.npanels;
            uint8_t const *nb = args.imgB[p2] + ((long long)(np2 * args.nkc + kc2) * kSubs) * kBSubBytes;
            for (int s = 0; s < kSubs; ++s)
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nb + (long long)s * kBSubBytes), "r"((uint32_t)kBSubBytes) : "memory");
          }
        }
        for (int mt = 0; mt < args.mtiles; ++mt) {
          uint8_t const *srcA = args.imgA[p] + ((long long)(mt * args.nkc + kc) * kSubs) * kSubBytes;
          if (mt + 2 < args.mtiles)       // the ring holds one and a half row tiles: warm L2 two tiles ahead
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(srcA + 2ll * args.nkc * kSubs * kSubBytes),
                         "r"((uint32_t)(kSubs * kSubBytes)) : "memory");
          for (int h = 0; h < 2; ++h) {
            uint32_t const st = a_stage;
            mbar_wait(empty_a(st), a_phase ^ 1u);
            if (++a_stage == kAStages) { a_stage = 0; a_phase ^= 1u; }
            mbar_expect_tx(full_a(st), kStageBytes);
            bulk_g2s(sA + st * kStageBytes, srcA + (long long)h * kStageBytes, kStageBytes, full_a(st));
            if (mt == 0 && h == 0) load_panel_half(2);      // (the first tile's first stage does not wait for the whole panel)
          }
        }
      }
    }
  } else if (warp == 1) {        // ---------------- MMA issue ----------------
    // The whole warp walks the loops (so that the address arithmetic stays in uniform registers — the instruction stream
    // of the one issuing lane must cost well under the time an MMA occupies the pipe); lane 0 issues.
    // TMEM: three 128-column regions R0 R1 R2 (+ the scale columns from 384).  An even row tile accumulates in R0|R1, an
    // odd one in R1|R2, so the K-steps of sub-images 1..3 are single N = 256 instructions; the first sub-image
    // is issued as N = 128 halves, the half in the region nobody else uses first, so that the drain of R1 (the previous
    // tile's other half) hides behind 256 cycles of work.
    uint32_t const idesc128 = (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | (1u << 23) | ((uint32_t)(128 >> 4) << 24);
    uint32_t const idesc256 = (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | (1u << 23) | ((uint32_t)(128 >> 4) << 24);
    uint32_t const sfa = tmem + 384u, sfb = tmem + 416u;
    uint32_t const desc_hi = (1024u >> 4) | (1u << 14), lbo = (128u >> 4) << 16;     // SBO, descriptor version | LBO
    uint32_t a_stage = 0, a_phase = 0, ji = 0, tile_ctr = 0;
    // eight K-steps (two sub-images = one A stage), issued by one elected lane, then up to three commits; the warp runs
    // this convergently.  Descriptor low words: +16 per K-step (256 B), +1024 / +2048 for the second sub-image of A / B.
#define TC_MMA_STEP(AOFF, BOFF, PRED)                                                                             \
  "add.u32 al, %1, " #AOFF ";\nadd.u32 bl, %2, " #BOFF ";\nmov.b64 da, {al, %6};\nmov.b64 db, {bl, %6};\n"        \
  "@pe tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], da, db, %3, [%4], [%5], " PRED ";\n"
#define TC_COMMIT(N) "setp.ne.and.b32 pc, %" #N ", 0, pe;\n@pc tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%" #N "];\n"
    auto mma8 = [&](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, bool first_overwrites, uint32_t c0, uint32_t c1, uint32_t c2) {
      asm volatile(
          "{\n.reg .pred pe, pt, pf, pc;\n.reg .b64 da, db;\n.reg .b32 al, bl;\n"
          "elect.sync _|pe, 0xffffffff;\n"
          "setp.eq.b32 pt, 0, 0;\nsetp.ne.b32 pf, %7, 0;\n"
          TC_MMA_STEP(0, 0, "pf") TC_MMA_STEP(16, 16, "pt") TC_MMA_STEP(32, 32, "pt") TC_MMA_STEP(48, 48, "pt")
          TC_MMA_STEP(1024, 2048, "pt") TC_MMA_STEP(1040, 2064, "pt") TC_MMA_STEP(1056, 2080, "pt") TC_MMA_STEP(1072, 2096, "pt")
          TC_COMMIT(8) TC_COMMIT(9) TC_COMMIT(10)
          "}"
          ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(sfa), "r"(sfb), "r"(desc_hi), "r"(first_overwrites ? 0u : 1u), "r"(c0), "r"(c1), "r"(c2)
          : "memory");
    };
#undef TC_MMA_STEP
    for (int job = blockIdx.x; job < njobs; job += gridDim.x, ++ji) {
      for (int mt = 0; mt < args.mtiles; ++mt, ++tile_ctr) {
        uint32_t const odd = tile_ctr & 1u, rfree = odd ? 2u : 0u;        // this tile: regions {rfree, 1}
        uint32_t const dbase = tmem + odd * 128u;                          // panel column c <-> TMEM column odd * 128 + c
        uint32_t const hfree = odd ? 1u : 0u, hmid = hfree ^ 1u;           // panel halves living in rfree / R1
        bool const first = mt == 0, last = mt + 1 == args.mtiles;    // of this job: the panel arrives / is let go piecewise
        uint32_t const b01 = ((sB) >> 4) | lbo, b23 = ((sB + 2 * kBSubBytes) >> 4) | lbo;
        // sub-images 0 and 1 (one stage) as N = 128 halves: first the half whose region nobody else uses (512 cycles of
        // work), then, once the previous tile's other half has been drained from R1, the R1 half
        uint32_t const st0 = a_stage;
        mbar_wait(full_a(st0), a_phase);
        if (++a_stage == kAStages) { a_stage = 0; a_phase ^= 1u; }
        if (first) { mbar_wait(full_b(0), ji & 1u); mbar_wait(full_b(1), ji & 1u); }
        mbar_wait(acc_empty(rfree), ((tile_ctr >> 1) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t const a0 = ((sA + st0 * kStageBytes) >> 4) | lbo;
        mma8(tmem + rfree * 128u, a0, b01 + hfree * 1024u, idesc128, true, 0u, 0u, 0u);
        mbar_wait(acc_empty(1), (tile_ctr & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        mma8(tmem + 128u, a0, b01 + hmid * 1024u, idesc128, true, empty_a(st0), last ? empty_b(0) : 0u, last ? empty_b(1) : 0u);
        // sub-images 2 and 3 as N = 256
        uint32_t const st1 = a_stage;
        mbar_wait(full_a(st1), a_phase);
        if (++a_stage == kAStages) { a_stage = 0; a_phase ^= 1u; }
        if (first) { mbar_wait(full_b(2), ji & 1u); mbar_wait(full_b(3), ji & 1u); }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        mma8(dbase, ((sA + st1 * kStageBytes) >> 4) | lbo, b23, idesc256, false, empty_a(st1), last ? empty_b(2) : 0u, last ? empty_b(3) : 0u);
        asm volatile("{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n"
                     "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                     "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%1];\n}"
                     ::"r"(acc_full(1)), "r"(acc_full(rfree)) : "memory");
      }
    }
#undef TC_COMMIT
  } else {                       // ---------------- epilogue: warps 2..9 ----------------
    // every warp drains 64 columns of R1 first — the region the next tile is waiting for —, then 64 columns of R0 / R2
    int const q = warp & 3, hsel = (warp - 2) >> 2;
    uint32_t tile_ctr = 0;
    auto drain = [&](uint32_t region, uint32_t parity, word *dst) {
      mbar_wait(acc_full(region), parity);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (args.flags & 2) {                              // timing experiment: hand the region back without reading it
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(region));
        return;
      }
      uint32_t v[64];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t const taddr = tmem + ((uint32_t)(q * 32) << 16) + region * 128u + (uint32_t)(hsel * 64 + c * 32);
        uint32_t *o = v + 32 * c;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]), "=r"(o[8]), "=r"(o[9]),
              "=r"(o[10]), "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15]), "=r"(o[16]), "=r"(o[17]), "=r"(o[18]),
              "=r"(o[19]), "=r"(o[20]), "=r"(o[21]), "=r"(o[22]), "=r"(o[23]), "=r"(o[24]), "=r"(o[25]), "=r"(o[26]), "=r"(o[27]),
              "=r"(o[28]), "=r"(o[29]), "=r"(o[30]), "=r"(o[31])
            : "r"(taddr));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(region));
      if (args.flags & 1) return;
      uint32_t out[2];
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        uint32_t bits = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {                // integer value + 2^23 -> parity in mantissa bit 0 -> funnel-shifted in
          float const f = __uint_as_float(v[32 * w + j]) + 8388608.0f;
          bits = __funnelshift_r(bits, __float_as_uint(f), 1);
        }
        out[w] = bits;
      }
      asm volatile("red.global.xor.b64 [%0], %1;" ::"l"(dst), "l"((unsigned long long)out[0] | ((unsigned long long)out[1] << 32)) : "memory");
    };
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
      int const p = job / jobs_per_product, np = (job % jobs_per_product) % args.npanels;
      for (int mt = 0; mt < args.mtiles; ++mt, ++tile_ctr) {
        uint32_t const odd = tile_ctr & 1u, rfree = odd ? 2u : 0u, hfree = odd, hmid = odd ^ 1u;
        word *row = args.C[p] + (long long)(mt * 128 + q * 32 + lane) * args.pitchC[p] + np * 4 + hsel;
        drain(1u, tile_ctr & 1u, row + hmid * 2);
        drain(rfree, (tile_ctr >> 1) & 1u, row + hfree * 2);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

struct TcPrepArgs {
  word const *A[kTcMaxBatch];
  word const *B[kTcMaxBatch];
  word *C[kTcMaxBatch];
  uint8_t *imgA[kTcMaxBatch];
  uint8_t *imgB[kTcMaxBatch];
  int pitchA[kTcMaxBatch], pitchB[kTcMaxBatch], pitchC[kTcMaxBatch];
  int m, l, n, nkc;
};

// A operands: bits of A (m x l) -> the tiled e2m1 image.  One thread per (row, 256-element sub-image): it reads the 32
// bytes of its row (one sector) and writes the eight 16-byte core-matrix rows they expand to, 128 bytes apart — the eight
// rows of a group (consecutive lanes) fill whole 128-byte lines.  blockIdx.y = product.
__global__ void __launch_bounds__(256) tc_expand_a_kernel(const __grid_constant__ TcPrepArgs a) {
  int const p = blockIdx.y;
  long long const t = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // (tile, kc, s, row group, row in group)
  if (t >= (long long)a.m * a.l / 256) return;
  int const r8 = (int)(t & 7), rg = (int)((t >> 3) & 15);
  long long const sub = t >> 7;                                               // 128 rows per sub-image
  int const s = (int)(sub % kSubs), kc = (int)((sub / kSubs) % a.nkc);
  long long const tile = sub / ((long long)kSubs * a.nkc);
  long long const row = tile * 128 + rg * 8 + r8;
  uint4 const *src = reinterpret_cast<uint4 const *>(a.A[p] + row * a.pitchA[p] + kc * 16 + s * 4);
  uint4 const lo = src[0], hi = src[1];
  uint32_t const x[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  uint4 *dst = reinterpret_cast<uint4 *>(a.imgA[p] + sub * kSubBytes + rg * 1024 + r8 * 16);
#pragma unroll
  for (int kg = 0; kg < 8; ++kg)
    dst[kg * 8] = make_uint4(expand8(x[kg]), expand8(x[kg] >> 8), expand8(x[kg] >> 16), expand8(x[kg] >> 24));
}

// C = 0 for every product (the main kernel XORs partial sums into it); 16 bytes per thread, blockIdx.y = product
__global__ void __launch_bounds__(256) tc_zero_c_kernel(const __grid_constant__ TcPrepArgs a) {
  int const p = blockIdx.y, w16 = a.n / 128;
  long long const t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)a.m * w16) return;
  long long const row = t / w16;
  int const c = (int)(t % w16);
  reinterpret_cast<uint4 *>(a.C[p] + row * a.pitchC[p])[c] = make_uint4(0, 0, 0, 0);
}

// lane i holds row i of a 32 x 32 bit block (column j in bit j); returns column `lane` as a row
__device__ __forceinline__ uint32_t tc_transpose32(uint32_t x, int lane) {
  uint32_t m = 0x0000FFFFu;
#pragma unroll
  for (int j = 16; j; j >>= 1) {
    uint32_t const y = __shfl_xor_sync(0xffffffffu, x, j);
    if (lane & j) x = (x & (m << j)) | ((y >> j) & m);
    else          x = (x & m) | ((y & m) << j);
    m ^= m << (j >> 1);
  }
  return x;
}

// B operands: the image rows are the COLUMNS of B.  One CTA = one 32 KB sub-image (256 columns x 256 K); warp w owns the
// K rows 32 w .. 32 w + 31 of it: a lane reads 32 bytes of its row (one sector), eight 32 x 32 bit blocks are transposed
// in registers, and lane j ends up with 32 K-bits of column j -> one 16-byte core-matrix row of the image.
__global__ void __launch_bounds__(256) tc_expand_bt_kernel(const __grid_constant__ TcPrepArgs a) {
  int const p = blockIdx.z, np = blockIdx.x, ks = blockIdx.y;        // ks = kc * 4 + s: K rows 256 ks ..
  int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int const kc = ks / kSubs, s = ks % kSubs;
  uint4 const *src = reinterpret_cast<uint4 const *>(a.B[p] + (long long)(ks * 256 + warp * 32 + lane) * a.pitchB[p] + np * 4);
  uint4 const lo = src[0], hi = src[1];
  uint32_t const x[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  uint8_t *sub = a.imgB[p] + ((long long)(np * a.nkc + kc) * kSubs + s) * kBSubBytes;
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    uint32_t const bits = tc_transpose32(x[nb], lane);      // K bits 32 warp .. of column 32 nb + lane
    int const n = nb * 32 + lane;
    *reinterpret_cast<uint4 *>(sub + (n / 8) * 1024 + warp * 128 + (n % 8) * 16) =
        make_uint4(expand8(bits), expand8(bits >> 8), expand8(bits >> 16), expand8(bits >> 24));
  }
}

struct TcScratch { uint8_t *a = nullptr, *b = nullptr; size_t bytes_a = 0, bytes_b = 0; };
TcScratch g_tc[16];
std::mutex g_tc_mu;

}  // namespace

bool tc_leaf_suits(int m, int l, int n) { return m >= 128 && m % 128 == 0 && l >= 1024 && l % 1024 == 0 && n >= 256 && n % 256 == 0; }

// C[i] = A[i] * B[i] (C overwritten) or, with `accumulate`, C[i] ^= A[i] * B[i] — the kernel XORs its partial results into C
// either way, so the accumulating form just leaves out the clearing of C — for up to 49 products of one shape
void launch_tc_batch(int count, DView const *C, DView const *A, DView const *B, cudaStream_t s, bool accumulate) {
  int const m = A[0].nrows, l = A[0].ncols, n = B[0].ncols;
  if (count < 1 || count > kTcMaxBatch || !tc_leaf_suits(m, l, n))
    die("m4ri_b200: tensor-core leaf needs <= 49 products with m %% 128 == 0, l %% 1024 == 0, n %% 256 == 0\n");
  size_t const each_a = (size_t)m * l / 2, each_b = (size_t)n * l / 2;
  int dev = 0;
  M4B_CUDA(cudaGetDevice(&dev));
  TcScratch *sc;
  {
    std::lock_guard<std::mutex> lock(g_tc_mu);
    sc = &g_tc[dev & 15];
    if (sc->bytes_a < each_a * count || sc->bytes_b < each_b * count) {
      M4B_CUDA(cudaDeviceSynchronize());
      if (sc->a) cudaFree(sc->a);
      if (sc->b) cudaFree(sc->b);
      M4B_CUDA(cudaMalloc(&sc->a, each_a * count));
      M4B_CUDA(cudaMalloc(&sc->b, each_b * count));
      sc->bytes_a = each_a * count; sc->bytes_b = each_b * count;
    }
  }
  TcPrepArgs prep{};
  Tc2Args args{};
  for (int i = 0; i < count; ++i) {
    if (A[i].nrows != m || A[i].ncols != l || B[i].nrows != l || B[i].ncols != n || C[i].nrows != m || C[i].ncols != n)
      die("m4ri_b200: tensor-core leaf batch needs identical shapes\n");
    prep.pitchA[i] = (int)A[i].pitch; prep.pitchB[i] = (int)B[i].pitch; prep.pitchC[i] = (int)C[i].pitch;
    args.pitchC[i] = (int)C[i].pitch;
    prep.A[i] = A[i].data; prep.B[i] = B[i].data; prep.C[i] = C[i].data;
    prep.imgA[i] = sc->a + each_a * i; prep.imgB[i] = sc->b + each_b * i;
    args.C[i] = C[i].data; args.imgA[i] = prep.imgA[i]; args.imgB[i] = prep.imgB[i];
  }
  prep.m = m; prep.l = l; prep.n = n; prep.nkc = l / 1024;
  static int const reuse = getenv("M4RI_B200_TC_REUSE") ? atoi(getenv("M4RI_B200_TC_REUSE")) : 0;   // timing knob: keep the images
  static word const *last_a = nullptr, *last_b = nullptr;
  long long const units = (long long)m * l / 32;
  if (reuse && count == 1 && last_a == A[0].data && last_b == B[0].data) {
    if (!accumulate) M4B_CUDA(cudaMemset2DAsync(C[0].data, C[0].pitch * sizeof(word), 0, (size_t)(n / 64) * sizeof(word), m, s));
  } else {
    tc_expand_a_kernel<<<dim3((unsigned)((units / 8 + 255) / 256), count), 256, 0, s>>>(prep);
    if (!accumulate) tc_zero_c_kernel<<<dim3((unsigned)(((long long)m * (n / 128) + 255) / 256), count), 256, 0, s>>>(prep);
    tc_expand_bt_kernel<<<dim3(n / 256, l / 256, count), 256, 0, s>>>(prep);
    last_a = A[0].data; last_b = B[0].data;
  }
  args.count = count; args.mtiles = m / 128; args.nkc = l / 1024; args.npanels = n / 256;
  static bool attr[16] = {};
  if (!attr[dev & 15]) { M4B_CUDA(cudaFuncSetAttribute(tc_leaf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTc2Smem)); attr[dev & 15] = true; }
  static int const flags = getenv("M4RI_B200_TC_FLAGS") ? atoi(getenv("M4RI_B200_TC_FLAGS")) : 0;
  args.flags = flags;
  int const njobs = count * args.nkc * args.npanels, grid = njobs < m4rm_num_sms() ? njobs : m4rm_num_sms();
  tc_leaf2_kernel<<<grid, kTc2Threads, kTc2Smem, s>>>(args);
  M4B_CUDA(cudaGetLastError());
  g_kernel_launches += 4;
}

void launch_tc_leaf2(DView C, DView A, DView B, cudaStream_t s) { launch_tc_batch(1, &C, &A, &B, s, false); }

// frees the operand-image scratch of every device (m4ri_b200_release)
void tc_scratch_release() {
  std::lock_guard<std::mutex> lock(g_tc_mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (int d = 0; d < 16; ++d) {
    if (!g_tc[d].a && !g_tc[d].b) continue;
    cudaSetDevice(d);
    cudaDeviceSynchronize();
    cudaFree(g_tc[d].a);
    cudaFree(g_tc[d].b);
    g_tc[d] = TcScratch{};
  }
  cudaSetDevice(cur);
}

}  // namespace m4b
