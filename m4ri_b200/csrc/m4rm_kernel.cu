// m4rm_kernel.cu — the M4RM ("Method of the Four Russians" multiplication) leaf for sm_100a.
//
// Replaces the reference's _mzd_mul_m4rm hot loops (m4ri/brilliantrussian.c:1107-1178):
//   mzd_make_table  (brilliantrussian.c:163-211)  -> build_tables(): 2^8-entry Gray-code
//                                                    tables built directly in shared memory
//   mzd_read_bits   (mzd.h:892-901)               -> byte extraction from the TMA-staged A slab
//   _mzd_combine_8  (xor_template.h, xor.h:96-122) -> LDS.128 table-row lookups XORed into a
//                                                    register-resident C tile
//
// B200 design (not a translation of the CPU code):
//   * One persistent CTA per SM.  The (C-tile x K-slab) iteration space is split evenly over
//     the CTAs ("stream-K"); because GF(2) accumulation is XOR, partial tiles are merged into
//     C with red.global.xor — exact, order independent, no second pass.
//   * C tile = TM x 1024 bits lives in registers for the whole K range of a segment
//     (C is touched once per segment instead of once per 64 columns of A as on the CPU).
//   * A table row is exactly 128 B = all 32 banks; the lookup mapping (see the kernel) makes every
//     quarter-warp pass of an LDS.128 read 128 conflict-free bytes; one warp instruction serves 8 C rows.
//   * k = 8 bits per table, 2 tables (16 columns of A) per step, tables double buffered: the
//     tables for step i+1 are built (Gray-code walk, one STS.32 wavefront per entry, no table reads)
//     while step i's lookups run; one __syncthreads per step.
//   * A (TM x 128 bit) and B (128 x 1024 bit) slabs arrive by TMA (cp.async.bulk.tensor.2d)
//     into a 2-deep ring guarded by mbarriers; out-of-range rows/columns are zero-filled by
//     the TMA unit, which is what makes ragged m / l / n edges free.
//   * Up to seven products of identical shape (the last Strassen level) share ONE launch.
//
// Binding resource: shared-memory bandwidth (128 B/clk/SM): per step and CTA 2048 lookup wavefronts
// + 512 table-store + 64 B-row + 64 A-word wavefronts for 2*16*1024*1024 bit-ops; ncu shows the
// l1tex data pipe 93 % busy.  See DESIGN.md §4 for the roofline derived from this.
#include <cuda.h>
#include <cuda_runtime.h>

#include <utility>
#include <vector>

#include "dev.h"
#include "tmap.h"

namespace m4b {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult qr;
    M4B_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qr));
    if (!sym || qr != cudaDriverEntryPointSuccess) die("m4ri_b200: cuTensorMapEncodeTiled not available from the driver\n");
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 2D map over u32 elements of a bit-packed view: dim0 = 32-bit words of the (128-bit padded)
// row, dim1 = rows.  Reads outside [dim0) x [dim1) return zeros.
CUtensorMap make_map(DView V, int box_w32, int box_rows) {
  CUtensorMap map;
  cuuint64_t dims[2]    = {(cuuint64_t)((V.ncols + 127) / 128) * 4, (cuuint64_t)V.nrows};
  cuuint64_t strides[1] = {(cuuint64_t)V.pitch * 8};
  cuuint32_t box[2]     = {(cuuint32_t)box_w32, (cuuint32_t)box_rows};
  cuuint32_t estr[2]    = {1, 1};
  CUresult r = encode_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, V.data, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    die("m4ri_b200: cuTensorMapEncodeTiled failed (%d) for view %p pitch %lld %dx%d\n", (int)r, (void *)V.data,
        (long long)V.pitch, V.nrows, V.ncols);
  return map;
}

CUtensorMap make_map_row_groups(DView V, int box_w32, int group_rows, int box_groups) {
  if (V.nrows % group_rows) die("m4ri_b200: make_map_row_groups needs nrows %% %d == 0\n", group_rows);
  CUtensorMap map;
  cuuint64_t dims[3]    = {(cuuint64_t)((V.ncols + 127) / 128) * 4, (cuuint64_t)group_rows, (cuuint64_t)(V.nrows / group_rows)};
  cuuint64_t strides[2] = {(cuuint64_t)V.pitch * 8, (cuuint64_t)V.pitch * 8 * group_rows};
  cuuint32_t box[3]     = {(cuuint32_t)box_w32, (cuuint32_t)group_rows, (cuuint32_t)box_groups};
  cuuint32_t estr[3]    = {1, 1, 1};
  CUresult r = encode_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, V.data, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    die("m4ri_b200: cuTensorMapEncodeTiled (3D) failed (%d) for view %p pitch %lld %dx%d\n", (int)r, (void *)V.data,
        (long long)V.pitch, V.nrows, V.ncols);
  return map;
}

namespace {

constexpr int kTileBits   = 1024;               // C tile width in bits (one table row = 128 B)
constexpr int kRowBytes   = kTileBits / 8;      // 128
constexpr int kTableBytes = 256 * kRowBytes;    // k = 8 -> 32 KB per table
constexpr int kStepBufBytes = 2 * kTableBytes;  // two tables per step
constexpr int kSlabBits   = 128;                // K extent of one TMA slab
constexpr int kStepsPerSlab = kSlabBits / 16;   // 8 steps of 16 A-columns
constexpr int kBSlabBytes = kSlabBits * kRowBytes;  // 16 KB

// One launch multiplies up to kMaxBatch independent products of IDENTICAL shape (the seven products of
// the last Strassen level): the stream-K unit space simply runs over (problem, tile, slab).
constexpr int kMaxBatch = 7;

// Automatic: the tall-tile leaf (m4rm_leaf2.cu) where its 4096-row tiles are filled, the 1024-row leaf elsewhere
// (measured on B200: 16384^3 2.37 ms vs 2.62 ms, n = 65536 Strassen product 107 ms vs 117 ms).
constexpr int kDefaultLeafVariant = 0;

struct alignas(64) BatchArgs {
  CUtensorMap mapA[kMaxBatch];
  CUtensorMap mapB[kMaxBatch];
  unsigned long long *C[kMaxBatch];
  long long pitchC[kMaxBatch];   // words
  int m;                // rows of A / C
  int nwordsC;          // 64-bit words per C row that may be written
  int tiles_m;
  int tiles_n;
  int slabs;            // ceil(l / 128)
  int nprob;
  int overwrite;        // 1: C = A*B with plain stores (only legal when slabs == 1: every tile has one owner);
                        //    C may then alias A or B (the owner has consumed its operands before it stores)
  long long units_per_problem;   // tiles_m * tiles_n * slabs
  long long total_units;         // nprob * units_per_problem
};

__device__ __forceinline__ uint32_t smem_u32(void const *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, CUtensorMap const *map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void red_xor64(unsigned long long *p, uint32_t lo, uint32_t hi) {
  unsigned long long v = (static_cast<unsigned long long>(hi) << 32) | lo;
  asm volatile("red.global.xor.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

template <int TM, int NT>
struct Cfg {
  static constexpr int kWarps      = NT / 32;
  static constexpr int kRowsPerWarp = TM / kWarps;
  static constexpr int R           = kRowsPerWarp / 4;   // 2 x rows held per thread (each row as two 16-byte pieces)
  static constexpr int kASlabBytes = TM * 16;
  static constexpr int kABoxRows   = TM < 256 ? TM : 256;
  static constexpr int kOffTables  = 0;
  static constexpr int kOffA       = 2 * kStepBufBytes;
  static constexpr int kOffB       = kOffA + 2 * kASlabBytes;
  static constexpr int kOffBar     = kOffB + 2 * kBSlabBytes;
  static constexpr int kSmemBytes  = kOffBar + 64;
  static constexpr uint32_t kSlabTxBytes = kASlabBytes + kBSlabBytes;
  static_assert(R >= 2 && kRowsPerWarp % 8 == 0, "bad tile shape");
};

// Build the two 256-entry tables of one step from 16 rows of the B slab (replaces mzd_make_table).
// The 32 lanes of a warp own the 32 words of one table row (LDS.32 / STS.32 = exactly one 128-byte
// wavefront each); a warp owns 512/warps consecutive entries of one table: base = XOR of the B rows
// selected by the high index bits, then the low GB index bits are walked in reflected Gray order so
// every further entry costs one XOR and one store — the trick of mzd_make_table, but without ever
// reading the table back.  (An earlier 8-lane x 16-byte mapping paid 4 wavefronts per B-row LDS.128
// because every quarter-warp re-read the same 128 bytes.)
template <int TM, int NT>
__device__ __forceinline__ void build_tables(uint32_t tbuf, uint32_t brows16, int tid) {
  constexpr int W  = NT / 32;
  constexpr int E  = 512 / W;                                   // entries per warp and step
  constexpr int GB = E == 64 ? 6 : (E == 32 ? 5 : (E == 16 ? 4 : -1));
  static_assert(GB > 0, "unsupported warp count");
  int const warp = tid >> 5, lane = tid & 31;
  int const t = warp / (W / 2), h = warp % (W / 2);             // table, index bits GB..7
  uint32_t const src = brows16 + (t * 8) * kRowBytes + lane * 4;
  uint32_t low[GB];
#pragma unroll
  for (int b = 0; b < GB; ++b) low[b] = lds32(src + b * kRowBytes);
  uint32_t e = 0;
#pragma unroll
  for (int b = GB; b < 8; ++b) e ^= lds32(src + b * kRowBytes) & (0u - ((h >> (b - GB)) & 1u));
  uint32_t const dst = tbuf + t * kTableBytes + (h << GB) * kRowBytes + lane * 4;
  sts32(dst, e);
#pragma unroll
  for (int i = 1; i < E; ++i) {
    e ^= low[(i & 1) ? 0 : (i & 2) ? 1 : (i & 4) ? 2 : (i & 8) ? 3 : (i & 16) ? 4 : 5];
    sts32(dst + (i ^ (i >> 1)) * kRowBytes, e);
  }
}

// Lane -> C-element mapping of the lookup loop: 4 lanes x 16 B own one HALF row of C (512 bits); in each
// quarter-warp lanes 0-3 take one C row and lanes 4-7 the next, and the two halves of the rows are
// fetched by two LDS.128 with the halves swapped between the lane groups, so every quarter-warp pass
// reads the low 64 B of one table row and the high 64 B of another: all 32 banks, conflict-free.
// A warp instruction serves 8 rows; a thread holds R/2 rows of 2 x 16 B.  (The obvious mapping — 8
// lanes own one full row — needs twice the A-index registers, loads and extractions per lookup.)
template <int TM, int NT>
__global__ void __launch_bounds__(NT, 1)
m4rm_streamk_kernel(const __grid_constant__ BatchArgs p) {
  using C = Cfg<TM, NT>;
  constexpr int R = C::R;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint32_t const sbase = smem_u32(smem);
  uint32_t const sTab = sbase + C::kOffTables;
  uint32_t const sA   = sbase + C::kOffA;
  uint32_t const sB   = sbase + C::kOffB;
  uint32_t const sBar = sbase + C::kOffBar;     // two 8-byte mbarriers

  int const tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int const q = lane >> 3;
  int const hi = (lane >> 2) & 1, c4 = lane & 3;          // row parity within the quarter-warp, 16 B chunk of a half row
  constexpr int RT = R / 2;                                // rows per thread

  if (tid == 0) {
    mbar_init(sBar, 1);
    mbar_init(sBar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  long long const u_begin = p.total_units * (long long)blockIdx.x / gridDim.x;
  long long const u_end   = p.total_units * (long long)(blockIdx.x + 1) / gridDim.x;
  uint32_t parity0 = 0, parity1 = 0;             // phase of each ring slot

  uint32_t const a_row_off = (warp * C::kRowsPerWarp + q * 2 + hi) * 16;   // first A row of this lane
  uint32_t const a_row_step = 128;                         // 8 rows further per j
  uint32_t const lane_off  = hi * 64 + c4 * 16;            // this lane's half of the table row ...
  uint32_t const lane_off2 = (hi ^ 1) * 64 + c4 * 16;      // ... and the other half

  for (long long u = u_begin; u < u_end;) {
    int const prob  = (int)(u / p.units_per_problem);
    long long const v = u - (long long)prob * p.units_per_problem;
    int const tile  = (int)(v / p.slabs);
    int const s0    = (int)(v % p.slabs);
    CUtensorMap const *mapA = &p.mapA[prob], *mapB = &p.mapB[prob];
    int nseg        = p.slabs - s0;
    if ((long long)nseg > u_end - u) nseg = (int)(u_end - u);
    int const tm = tile % p.tiles_m, tn = tile / p.tiles_m;
    int const row0 = tm * TM;

    auto issue = [&](int i) {                     // thread 0: TMA for slab s0+i into ring slot i&1
      uint32_t const slot = i & 1;
      uint32_t const bar  = sBar + 8 * slot;
      mbar_expect_tx(bar, C::kSlabTxBytes);
#pragma unroll
      for (int part = 0; part < TM / C::kABoxRows; ++part)
        tma_load_2d(sA + slot * C::kASlabBytes + part * C::kABoxRows * 16, mapA, (s0 + i) * 4,
                    row0 + part * C::kABoxRows, bar);
      tma_load_2d(sB + slot * kBSlabBytes, mapB, tn * 32, (s0 + i) * kSlabBits, bar);
    };

    if (tid == 0) {
      issue(0);
      if (nseg > 1) issue(1);
    }

    uint4 acc[RT][2];
#pragma unroll
    for (int j = 0; j < RT; ++j) acc[j][0] = acc[j][1] = make_uint4(0, 0, 0, 0);

    mbar_wait(sBar, parity0);
    parity0 ^= 1;
    build_tables<TM, NT>(sTab, sB, tid);
    __syncthreads();

    for (int i = 0; i < nseg; ++i) {
      uint32_t const slot = i & 1;
      uint32_t const aS = sA + slot * C::kASlabBytes + a_row_off;
      uint32_t const bS = sB + slot * kBSlabBytes;
#pragma unroll 1
      for (int pr = 0; pr < kStepsPerSlab / 2; ++pr) {
        uint32_t aw[RT];
#pragma unroll
        for (int j = 0; j < RT; ++j) aw[j] = lds32(aS + j * a_row_step + pr * 4);
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          int const step = pr * 2 + sub;
          uint32_t const tcur = sTab + sub * kStepBufBytes;          // step parity == sub
          uint32_t const tnext = sTab + (sub ^ 1) * kStepBufBytes;
          // ---- build tables for the next step into the other buffer ----
          if (step < kStepsPerSlab - 1) {
            build_tables<TM, NT>(tnext, bS + (step + 1) * 16 * kRowBytes, tid);
          } else if (i + 1 < nseg) {
            if (slot == 0) { mbar_wait(sBar + 8, parity1); parity1 ^= 1; }
            else           { mbar_wait(sBar, parity0);     parity0 ^= 1; }
            build_tables<TM, NT>(tnext, sB + (slot ^ 1) * kBSlabBytes, tid);
          }
          // ---- lookups: acc[j] ^= T0[a byte 2*sub] ^ T1[a byte 2*sub+1] ----
          // (rows past m read zero-filled A bits -> table row 0 = zeros; no branch needed)
          uint32_t const t0 = tcur + lane_off, t1 = tcur + kTableBytes + lane_off;
          uint32_t const u0 = tcur + lane_off2, u1 = tcur + kTableBytes + lane_off2;
#pragma unroll
          for (int j = 0; j < RT; ++j) {
            uint32_t const i0 = __byte_perm(aw[j], 0, 0x4440 + 2 * sub);
            uint32_t const i1 = __byte_perm(aw[j], 0, 0x4441 + 2 * sub);
            uint4 const v0 = lds128(t0 + i0 * kRowBytes);
            uint4 const v1 = lds128(t1 + i1 * kRowBytes);
            acc[j][0].x ^= v0.x ^ v1.x;
            acc[j][0].y ^= v0.y ^ v1.y;
            acc[j][0].z ^= v0.z ^ v1.z;
            acc[j][0].w ^= v0.w ^ v1.w;
            uint4 const w0 = lds128(u0 + i0 * kRowBytes);
            uint4 const w1 = lds128(u1 + i1 * kRowBytes);
            acc[j][1].x ^= w0.x ^ w1.x;
            acc[j][1].y ^= w0.y ^ w1.y;
            acc[j][1].z ^= w0.z ^ w1.z;
            acc[j][1].w ^= w0.w ^ w1.w;
          }
          __syncthreads();
        }
      }
      // ring slot `slot` is free again: refill it with slab i+2
      if (tid == 0 && i + 2 < nseg) issue(i + 2);
    }

    // ---- merge the partial tile into C (exact: XOR is associative and commutative) ----
    {
      int const rbase = row0 + warp * C::kRowsPerWarp + q * 2 + hi;
#pragma unroll
      for (int j = 0; j < RT; ++j) {
        int const row = rbase + 8 * j;
        if (row < p.m) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            int const wcol = tn * (kTileBits / 64) + (hi ^ h) * 8 + c4 * 2;
            unsigned long long *dst = p.C[prob] + (long long)row * p.pitchC[prob] + wcol;
            if (p.overwrite) {
              if (wcol < p.nwordsC) dst[0] = ((unsigned long long)acc[j][h].y << 32) | acc[j][h].x;
              if (wcol + 1 < p.nwordsC) dst[1] = ((unsigned long long)acc[j][h].w << 32) | acc[j][h].z;
            } else {
              if (wcol < p.nwordsC) red_xor64(dst, acc[j][h].x, acc[j][h].y);
              if (wcol + 1 < p.nwordsC) red_xor64(dst + 1, acc[j][h].z, acc[j][h].w);
            }
          }
        }
      }
    }
    // all table/slab reads of this segment are complete before the next segment's prologue
    __syncthreads();
    u += nseg;
  }
}

template <int TM, int NT>
void launch_variant(int count, DView const *Cv, DView const *A, DView const *B, bool overwrite, cudaStream_t stream) {
  using C = Cfg<TM, NT>;
  static bool configured[64] = {};   // the opt-in shared-memory size is a per-device attribute
  auto kern = m4rm_streamk_kernel<TM, NT>;
  int dev = 0;
  M4B_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    M4B_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    configured[dev & 63] = true;
  }
  BatchArgs p;
  p.m = A[0].nrows;
  p.nwordsC = (Cv[0].ncols + 63) / 64;
  p.tiles_m = (A[0].nrows + TM - 1) / TM;
  p.tiles_n = (B[0].ncols + kTileBits - 1) / kTileBits;
  p.slabs = (A[0].ncols + kSlabBits - 1) / kSlabBits;
  p.nprob = count;
  p.overwrite = overwrite ? 1 : 0;
  if (overwrite && p.slabs != 1) die("m4ri_b200: overwrite mode needs a single K slab\n");
  p.units_per_problem = (long long)p.tiles_m * p.tiles_n * p.slabs;
  p.total_units = p.units_per_problem * count;
  for (int i = 0; i < count; ++i) {
    if (A[i].nrows != A[0].nrows || A[i].ncols != A[0].ncols || B[i].ncols != B[0].ncols)
      die("m4ri_b200: batched leaf needs identical shapes\n");
    p.C[i] = reinterpret_cast<unsigned long long *>(Cv[i].data);
    p.pitchC[i] = Cv[i].pitch;
    p.mapA[i] = make_map(A[i], 4, C::kABoxRows);
    p.mapB[i] = make_map(B[i], 32, kSlabBits);
  }
  long long grid = m4rm_num_sms();
  if (grid > p.total_units) grid = p.total_units;
  kern<<<(unsigned)grid, NT, C::kSmemBytes, stream>>>(p);
  M4B_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}

}  // namespace

// ---- optional live timing of the leaf launches (bench.py's roofline figure) -----------------
namespace {
struct LeafProf {
  bool on = false;
  int  device = -1;      // the events belong to one device: launches on other GPUs (multi.cu's threads) are not timed
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;
  size_t used = 0;
  double bitops = 0;
} g_prof;
}  // namespace

void leaf_profile_begin() {
  cudaGetDevice(&g_prof.device);
  g_prof.on = true;
  g_prof.used = 0;
  g_prof.bitops = 0;
}

// Returns the number of leaf launches since leaf_profile_begin(); *ms = summed device time of
// those launches (CUDA events on the launching stream), *bitops = 2*m*l*n summed over them.
// The caller must have synchronised the stream.
unsigned long long leaf_profile_end(double *ms, double *bitops) {
  double total = 0;
  for (size_t i = 0; i < g_prof.used; ++i) {
    float t = 0;
    M4B_CUDA(cudaEventElapsedTime(&t, g_prof.pool[i].first, g_prof.pool[i].second));
    total += t;
  }
  if (ms) *ms = total;
  if (bitops) *bitops = g_prof.bitops;
  g_prof.on = false;
  return g_prof.used;
}

void leaf_profile_reset() {
  for (auto &pr : g_prof.pool) {
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  g_prof.pool.clear();
  g_prof.used = 0;
  g_prof.on = false;
}

int m4rm_num_sms() {   // per device: the library may be pointed at another GPU (m4ri_b200_set_device, multi.cu)
  static int sms[64] = {};
  int dev = 0;
  M4B_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  if (!sms[dev]) M4B_CUDA(cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev));
  return sms[dev];
}

// Leaf selection: M4RI_B200_LEAF=0|1|2|3 in the environment, or m4ri_b200_set_leaf_variant() at run time
// 0 (default): the tensor-core leaf of tc_leaf.cu for the C = A*B products it suits (every dimension in its tile units:
// the Strassen leaves of large products), else the tall-tile M4RM leaf where that suits, else the 1024-row M4RM leaf;
// 1 / 2: M4RM only (1024-row tiles / tall tiles for every shape); 3: same as 0.
int g_leaf_variant = -1;
int g_last_leaf = 0;        // kernel of the most recent leaf launch: 1 = 1024 x 1024-bit tiles, 2 = 4096 x 256-bit tiles
static int leaf_variant() {
  if (g_leaf_variant < 0) {
    char const *env = getenv("M4RI_B200_LEAF");
    g_leaf_variant = env && env[0] >= '0' && env[0] <= '3' && !env[1] ? env[0] - '0' : kDefaultLeafVariant;
  }
  return g_leaf_variant;
}

static bool tall_leaf(int m, int l, int n, bool overwrite) {
  int const variant = leaf_variant();
  return !overwrite && (variant == 2 || ((variant == 0 || variant == 3) && leaf2_suits(m, l, n)));
}

// the tensor-core leaf: C = A*B and C ^= A*B (not the in-place overwrite form), shapes in its tile units, 16-byte aligned rows
static bool tensor_leaf(int count, DView const *C, DView const *A, DView const *B) {
  int const variant = leaf_variant();
  if ((variant != 0 && variant != 3) || count > 49 || !tc_leaf_suits(A[0].nrows, A[0].ncols, B[0].ncols)) return false;
  for (int i = 0; i < count; ++i)
    if (((reinterpret_cast<uintptr_t>(A[i].data) | reinterpret_cast<uintptr_t>(B[i].data) | reinterpret_cast<uintptr_t>(C[i].data)) & 15) ||
        ((A[i].pitch | B[i].pitch | C[i].pitch) & 1))
      return false;
  return true;
}

static int m4rm_only_batch_limit(int m, int l, int n) { return tall_leaf(m, l, n, false) ? 49 : kMaxBatch; }

// products per launch the scheduler may ask for: 49 where the tensor-core or the tall-tile leaf takes the shape
int m4rm_batch_limit(int m, int l, int n) {
  int const variant = leaf_variant();
  if ((variant == 0 || variant == 3) && tc_leaf_suits(m, l, n)) return 49;
  return m4rm_only_batch_limit(m, l, n);
}

static void launch_leaf(int count, DView const *C, DView const *A, DView const *B, bool overwrite, cudaStream_t stream,
                        bool clear_first = false) {
  if (count <= 0 || A[0].nrows <= 0 || A[0].ncols <= 0 || B[0].ncols <= 0) return;   // empty product: C unchanged
  bool const tensor = !overwrite && tensor_leaf(count, C, A, B);
  if (!tensor) {
    // a batch sized for the tensor-core leaf that cannot take it after all (unaligned views): in pieces
    int const limit = m4rm_only_batch_limit(A[0].nrows, A[0].ncols, B[0].ncols);
    if (count > limit) {
      for (int i = 0; i < count; i += limit)
        launch_leaf(count - i < limit ? count - i : limit, C + i, A + i, B + i, overwrite, stream, clear_first);
      return;
    }
  }
  std::pair<cudaEvent_t, cudaEvent_t> *ev = nullptr;
  int cur_dev = -1;
  if (g_prof.on) cudaGetDevice(&cur_dev);
  if (g_prof.on && cur_dev == g_prof.device) {
    if (g_prof.used == g_prof.pool.size()) {
      std::pair<cudaEvent_t, cudaEvent_t> e;
      M4B_CUDA(cudaEventCreate(&e.first));
      M4B_CUDA(cudaEventCreate(&e.second));
      g_prof.pool.push_back(e);
    }
    ev = &g_prof.pool[g_prof.used++];
    g_prof.bitops += 2.0 * count * A[0].nrows * (double)A[0].ncols * B[0].ncols;
    M4B_CUDA(cudaEventRecord(ev->first, stream));
  }
  if (tensor) {
    g_last_leaf = 3;
    launch_tc_batch(count, C, A, B, stream, !clear_first);
    if (ev) M4B_CUDA(cudaEventRecord(ev->second, stream));
    return;
  }
  bool const tall = tall_leaf(A[0].nrows, A[0].ncols, B[0].ncols, overwrite);
  g_last_leaf = tall ? 2 : 1;
  if (clear_first && !tall)
    for (int i = 0; i < count; ++i) launch_zero(C[i], stream);
  if (tall)
    launch_m4rm_leaf2(count, C, A, B, stream, clear_first);          // tall tiles: 4096 rows x 256 bits
  else if (A[0].nrows <= 256)
    launch_variant<256, 256>(count, C, A, B, overwrite, stream);     // short operands: 256-row tiles
  else
    launch_variant<1024, 256>(count, C, A, B, overwrite, stream);
  if (ev) M4B_CUDA(cudaEventRecord(ev->second, stream));
}

void launch_m4rm_batch(int count, DView const *C, DView const *A, DView const *B, cudaStream_t stream) {
  launch_leaf(count, C, A, B, false, stream);
}

void launch_m4rm_batch_clear(int count, DView const *C, DView const *A, DView const *B, cudaStream_t stream) {
  if (count <= 0 || C[0].nrows <= 0 || C[0].ncols <= 0) return;
  if (A[0].ncols <= 0) {                                             // empty inner dimension: the product is zero
    for (int i = 0; i < count; ++i) launch_zero(C[i], stream);
    return;
  }
  launch_leaf(count, C, A, B, false, stream, true);
}

void launch_m4rm(DView C, DView A, DView B, cudaStream_t stream) { launch_leaf(1, &C, &A, &B, false, stream); }

// C = A*B for an inner dimension of at most 128 (one K slab): plain stores, no zero fill needed, and C
// may be the same view as A or B (used by the triangular solves for X = inv(T_block) * B_block in place).
void launch_m4rm_overwrite(DView C, DView A, DView B, cudaStream_t stream) {
  if (A.ncols > kSlabBits) die("m4ri_b200: launch_m4rm_overwrite needs l <= 128\n");
  launch_leaf(1, &C, &A, &B, true, stream);
}

}  // namespace m4b
