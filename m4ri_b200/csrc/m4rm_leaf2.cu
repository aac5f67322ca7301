// m4rm_leaf2.cu — device side and launcher of the tall-tile M4RM leaf (body: m4rm_leaf2_body.h).
//
// The body header is written against a handful of primitives so that the same source also runs inside
// the CPU emulation harness tests/c/emu_leaf2.cpp; here they are the sm_100a instructions themselves:
// LDS.128 / STS.128, PRMT, mbarrier try_wait / arrive.expect_tx, cp.async.bulk.tensor.2d/.3d (TMA),
// cp.async (LDGSTS), red.global.xor.b64.
#include <cuda.h>
#include <cuda_runtime.h>

#include "dev.h"
#include "tmap.h"

#define L2_FN __device__ __forceinline__

namespace leaf2 {

typedef uint4 U4;
typedef uint2 U2;
typedef CUtensorMap TMap;

L2_FN uint32_t smem_u32(void const *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
L2_FN U4 lds128(uint32_t addr) {
  U4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
template <int IMM>
L2_FN U4 lds128(uint32_t addr) {
  U4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4 + %5];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(addr), "n"(IMM));
  return v;
}
L2_FN U2 lds64(uint32_t addr) {
  U2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
L2_FN void sts128(uint32_t addr, U4 const &v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
L2_FN uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
L2_FN void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
L2_FN void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
L2_FN void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// L2 residency hints (round 2): a launch of 49 products touches ~300 MB — the A panels (2 MiB each, re-read by every
// column tile of their product) compete with the B and C streams for the 126 MB L2, and ncu showed 2.3 GB of DRAM
// traffic per launch against 0.4 GB of algorithmic bytes.  A is loaded evict_last, B evict_first.
// g_l2hint bits: 1 = A by TMA with an evict_last policy from createpolicy, 2 = B by cp.async with an evict_first policy,
// 4 = A by TMA with the fixed evict_last descriptor CUTLASS uses (0x14F0000000000000).  $M4RI_B200_LEAF2_L2HINT.
__constant__ int g_l2hint = 0;

L2_FN unsigned long long policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
L2_FN unsigned long long policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

L2_FN void tma_load_2d(uint32_t dst, TMap const *map, int c0, int c1, uint32_t bar) {
  if (g_l2hint & 5) {
    unsigned long long const pol = (g_l2hint & 4) ? 0x14F0000000000000ull : policy_evict_last();
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar), "l"(pol)
        : "memory");
    return;
  }
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
L2_FN void tma_load_3d(uint32_t dst, TMap const *map, int c0, int c1, int c2, uint32_t bar) {
  if (g_l2hint & 5) {
    unsigned long long const pol = (g_l2hint & 4) ? 0x14F0000000000000ull : policy_evict_last();
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(pol)
        : "memory");
    return;
  }
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
// LDGSTS: 16 bytes global -> shared without a register round trip; src_bytes = 0 writes zeros
L2_FN void cp_async16(uint32_t dst, void const *src, uint32_t src_bytes) {
  if (g_l2hint & 2) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(dst), "l"(src), "r"(src_bytes),
                 "l"(policy_evict_first())
                 : "memory");
    return;
  }
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
L2_FN void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
L2_FN void red_xor64(unsigned long long *p, uint32_t lo, uint32_t hi) {
  unsigned long long v = (static_cast<unsigned long long>(hi) << 32) | lo;
  asm volatile("red.global.xor.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
L2_FN void stg128(unsigned long long *p, U4 const &v) {
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
L2_FN void cta_sync() { __syncthreads(); }
L2_FN uint32_t gate(uint32_t a, uint32_t b, uint32_t c) {   // a | (b & c) as ONE LOP3 the compiler cannot see through
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xF8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

}  // namespace leaf2

#include "m4rm_leaf2_body.h"

namespace leaf2 {

template <int NT, int AWIDE, int SPLIT>
__global__ void __launch_bounds__(NT, 1) m4rm_leaf2_kernel(const __grid_constant__ Args p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint32_t const sbase = smem_u32(smem);
  if (threadIdx.x == 0) {
    mbar_init(sbase + kOffBar, 1);
    mbar_init(sbase + kOffBar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  cta_body<NT, AWIDE, SPLIT>(p, sbase, (int)threadIdx.x, (int)blockIdx.x, (int)gridDim.x);
}

}  // namespace leaf2

namespace m4b {

namespace {
constexpr int kThreads = 256;
constexpr int kDefaultSplit = 0;   // 1: tables of a step built by alternating halves of the CTA (+1.2 % with LDS.64 A loads, but -1.7 % together
                                   // with LDS.128 A loads, where it makes ptxas spill): 65536^3 94.09 ms (AWIDE 1, SPLIT 0) / 94.76 (0, 1) / 95.69 (1, 1)
constexpr int kDefaultAwide = 1;   // A bits of a whole slab per row with one LDS.128 (4 wavefronts per 32 rows and slab instead
                                   // of 8: the LDS.64 form has a 2-way bank conflict at its 16-byte stride); needs 32 more live
                                   // registers, which fit since the kernel compiles without spills (round 2): +1.3 % measured
}

// The tall tile only pays when its 4096 rows are (nearly) all real rows: rows past m are zero-filled by the
// TMA unit and looked up like any other.  Everything else stays on the 1024-row leaf.
bool leaf2_suits(int m, int l, int n) {
  if (m < leaf2::kTM || l <= 0 || n <= 0) return false;
  long long const padded = ((long long)m + leaf2::kTM - 1) / leaf2::kTM * leaf2::kTM;
  return padded * 16 <= (long long)m * 17;            // at most 1/16 of the lookups wasted on padding rows
}

// C ^= A*B (overwrite == false) or C = A*B (overwrite == true: C need not be initialised) for `count` (<= 49) products of
// identical shape in one persistent launch.
void launch_m4rm_leaf2(int count, DView const *Cv, DView const *A, DView const *B, cudaStream_t stream, bool overwrite) {
  using namespace leaf2;
  // experiment knobs: M4RI_B200_LEAF2_AWIDE=1 loads the A bits of a whole slab per row with one LDS.128;
  // M4RI_B200_LEAF2_SPLIT=0|1 lets all / half of the warps build the tables of a step
  static int const variant = [] {      // bit 0: AWIDE, bit 1: SPLIT
    char const *a = getenv("M4RI_B200_LEAF2_AWIDE"), *s = getenv("M4RI_B200_LEAF2_SPLIT");
    int const awide = a ? (a[0] == '1') : kDefaultAwide, split = s ? (s[0] == '1') : kDefaultSplit;
    return awide | (split << 1);
  }();
  static int const l2hint = [] {
    char const *e = getenv("M4RI_B200_LEAF2_L2HINT");
    return e ? atoi(e) & 7 : 0;
  }();
  static bool configured[64] = {};   // the opt-in shared-memory size is a per-device attribute
  auto kern = variant == 3 ? m4rm_leaf2_kernel<kThreads, 1, 1>
            : variant == 1 ? m4rm_leaf2_kernel<kThreads, 1, 0>
            : variant == 2 ? m4rm_leaf2_kernel<kThreads, 0, 1> : m4rm_leaf2_kernel<kThreads, 0, 0>;
  int dev = 0;
  M4B_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    M4B_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    M4B_CUDA(cudaMemcpyToSymbol(leaf2::g_l2hint, &l2hint, sizeof l2hint));
    configured[dev & 63] = true;
  }
  if (count > kMaxBatch) die("m4ri_b200: batch of %d leaf products exceeds %d\n", count, kMaxBatch);
  Args p;
  p.m = A[0].nrows;
  p.l = B[0].nrows;
  p.a3d = A[0].nrows % kABoxRows == 0 ? 1 : 0;
  p.nwordsC = (Cv[0].ncols + 63) / 64;
  p.tiles_m = (A[0].nrows + kTM - 1) / kTM;
  p.tiles_n = (B[0].ncols + kTileBits - 1) / kTileBits;
  p.slabs = (A[0].ncols + kSlabBits - 1) / kSlabBits;
  p.nprob = count;
  p.units_per_problem = (long long)p.tiles_m * p.tiles_n * p.slabs;
  p.total_units = p.units_per_problem * count;
  p.zero = 0;
  if (p.total_units >= (1ll << 31)) die("m4ri_b200: leaf product too large for 32-bit unit counters\n");
  for (int i = 0; i < count; ++i) {
    if (A[i].nrows != A[0].nrows || A[i].ncols != A[0].ncols || B[i].ncols != B[0].ncols)
      die("m4ri_b200: batched leaf needs identical shapes\n");
    p.C[i] = reinterpret_cast<unsigned long long *>(Cv[i].data);
    p.pitchC[i] = Cv[i].pitch;
    p.mapA[i] = p.a3d ? make_map_row_groups(A[i], 4, kABoxRows, kAParts) : make_map(A[i], 4, kABoxRows);
    p.B[i] = reinterpret_cast<unsigned long long const *>(B[i].data);
    p.pitchB[i] = B[i].pitch;
  }
  long long grid = m4rm_num_sms();
  if (grid > p.total_units) grid = p.total_units;
  static int const hybrid = [] {       // M4RI_B200_LEAF2_HYBRID=0: pure stream-K over all units (the round-1 partition)
    char const *e = getenv("M4RI_B200_LEAF2_HYBRID");
    return e && e[0] == '0' ? 0 : 1;
  }();
  long long const tiles_total = (long long)p.tiles_m * p.tiles_n * count;
  p.dp_rounds = hybrid ? (int)(tiles_total / grid) : 0;
  p.store_dp = 0;
  if (overwrite) {
    // C = A*B: the tiles of the whole-tile rounds are stored by their single owner, so only the products that contain
    // tiles of the stream-K tail (the last ones in (product, tile) order) have to start from zeros
    static int const store = [] {
      char const *e = getenv("M4RI_B200_LEAF2_STORE");
      return e && e[0] == '0' ? 0 : 1;
    }();
    bool const can_store = store && p.dp_rounds > 0 && Cv[0].ncols % 128 == 0;
    long long const tiles_per_product = (long long)p.tiles_m * p.tiles_n;
    int const first_zeroed = can_store ? (int)((long long)p.dp_rounds * grid / tiles_per_product) : 0;
    for (int i = first_zeroed; i < count; ++i) launch_zero(Cv[i], stream);
    p.store_dp = can_store ? 1 : 0;
  }
  kern<<<(unsigned)grid, kThreads, kSmemBytes, stream>>>(p);
  M4B_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}

}  // namespace m4b
