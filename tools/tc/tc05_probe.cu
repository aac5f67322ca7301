// tc05_probe.cu — hand-written tcgen05 probe (measurement only, not part of the product): the pipe-level ceiling of a
// GF(2) product on the 5th-generation tensor cores with operands expanded to one byte per bit.
//   kind::i8      0/1 as uint8, int32 accumulators in TMEM          (exact)
//   kind::f8f6f4  0/1 as e4m3 (0x00 / 0x38), fp32 accumulators
//   kind::mxf4    0/1 as e2m1 (two per byte, 0x0 / 0x2), block scales 2^0 (0x7F everywhere in the scale columns of TMEM),
//                 fp32 accumulators; K = 64 elements (the same 32 bytes) per instruction
// One CTA per SM, M = 128, N = 256, K = 32 bytes per tcgen05.mma (cta_group::1, both operands from shared memory, K-major
// canonical no-swizzle layout: 8 rows x 16 bytes per core matrix, LBO = 128 B along K, SBO = 512 B between 8-row groups).
// The same 128 x 64-byte A tile and 256 x 64-byte B tile are multiplied over and over (no global traffic in the timed
// loop), one elected thread issues, completion through tcgen05.commit -> mbarrier, accumulators read back with
// tcgen05.ld (32x32b) and compared on the host for a single pass.
// -DPROBE_N=128 measures the N = 128 shape (the rate per instruction halves, the A tile is re-read twice as often).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc05_probe tc05_probe.cu ; SASS: cuobjdump -sass tc05_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#ifndef PROBE_N
#define PROBE_N 256
#endif
constexpr int kM = 128, kN = PROBE_N, kKBytes = 64, kKStep = 32;   // 8 + 16 KB of static shared memory
constexpr int kSBO = kKBytes / 16 * 128;                         // bytes between 8-row groups: all K core matrices of a group

__device__ __forceinline__ uint32_t smem_u32(void const *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);            // start address
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;       // leading byte offset: next core matrix along K
  d |= (uint64_t)(((uint32_t)kSBO >> 4) & 0x3FFF) << 32;   // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  return d;                                          // layout type 0 = no swizzle
}

__host__ __device__ inline uint8_t bit_of(int row, int k, int which) {
  uint32_t h = (uint32_t)(row * 2654435761u) ^ (uint32_t)(k * 40503u) ^ (which ? 0x9E3779B9u : 0x7F4A7C15u);
  h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
  return (uint8_t)(h & 1u);
}

// element e of row `row` (operand `which`): K index in ELEMENTS (two per byte for mxf4)
template <int KIND>
__host__ __device__ inline uint8_t packed_byte(int row, int kbyte, int which) {
  if (KIND == 2) return (uint8_t)((bit_of(row, 2 * kbyte, which) ? 0x02 : 0) | (bit_of(row, 2 * kbyte + 1, which) ? 0x20 : 0));
  return bit_of(row, kbyte, which) ? (KIND == 0 ? 1 : 0x38) : 0;
}

template <int KIND>   // 0: i8, 1: f8f6f4 (e4m3), 2: mxf4 (e2m1, block-scaled)
__global__ void __launch_bounds__(128, 1) tc05_kernel(int passes, int32_t *out /* 128 x 256 of CTA 0, may be null */) {
  __shared__ __align__(1024) uint8_t sA[kM * kKBytes];
  __shared__ __align__(1024) uint8_t sB[kN * kKBytes];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  int const tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kM * kKBytes; i += 128) {
    int const r = i / kKBytes, k = i % kKBytes;
    sA[(r / 8) * kSBO + (k / 16) * 128 + (r % 8) * 16 + (k % 16)] = packed_byte<KIND>(r, k, 0);
  }
  for (int i = tid; i < kN * kKBytes; i += 128) {
    int const r = i / kKBytes, k = i % kKBytes;
    sB[(r / 8) * kSBO + (k / 16) * 128 + (r % 8) * 16 + (k % 16)] = packed_byte<KIND>(r, k, 1);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the MMA (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t const tmem = tmem_base;
  if (KIND == 2) {   // scale factors: every byte of TMEM columns 256..511 = 0x7F (ue8m0 2^0), whatever the scale layout reads
    uint32_t const v = 0x7F7F7F7Fu;
    for (int c = 256; c < 512; c += 8) {
      uint32_t const taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // instruction descriptor (cute::UMMA::InstrDescriptor): c_format [4,6), a_format [7,10), b_format [10,13), K-major both,
  // n_dim = N >> 3 at [17,23), m_dim = M >> 4 at [24,29)
  // block-scaled form (InstrDescriptorBlockScaled): a_format / b_format = 1 (MXF4 E2M1), scale_format bit 23 = 1 (UE8M0)
  uint32_t const idesc = KIND == 2 ? ((1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | (1u << 23) | ((uint32_t)(kM >> 4) << 24))
                                   : ((KIND == 0 ? (2u << 4) : (1u << 4)) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24));
  if (tid == 0) {
    uint64_t const da = make_desc(smem_u32(sA)), db = make_desc(smem_u32(sB));
    for (int p = 0; p < passes; ++p) {
#pragma unroll
      for (int ks = 0; ks < kKBytes / kKStep; ++ks) {
        uint64_t const a = da + (uint64_t)((ks * 2 * 128) >> 4), b = db + (uint64_t)((ks * 2 * 128) >> 4);
        uint32_t const acc = (p | ks) ? 1u : 0u;
        if (KIND == 0)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
                       ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
        else if (KIND == 1)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}"
                       ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
        else
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                       "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n}"
                       ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(tmem + 256u), "r"(tmem + 384u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everybody waits for the commit (phase 0)
  {
    uint32_t const b = smem_u32(&bar);
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(b)
        : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (out != nullptr && blockIdx.x == 0) {
    // warp w reads TMEM lanes 32w .. 32w+31 (= rows of D), 8 columns per instruction
    for (int c = 0; c < kN; c += 8) {
      uint32_t v[8];
      uint32_t const taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) out[tid * kN + c + j] = KIND == 0 ? (int32_t)v[j] : (int32_t)__uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int KIND>
int run(char const *name, int sms) {
  int32_t *d_out;
  cudaMalloc(&d_out, kM * kN * 4);
  cudaMemset(d_out, 0xFF, kM * kN * 4);
  tc05_kernel<KIND><<<1, 128>>>(1, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("{\"probe\": \"%s\", \"error\": \"%s\"}\n", name, cudaGetErrorString(e));
    return 1;
  }
  std::vector<int32_t> got(kM * kN);
  cudaMemcpy(got.data(), d_out, got.size() * 4, cudaMemcpyDeviceToHost);
  int const kelems = KIND == 2 ? 2 * kKBytes : kKBytes;
  std::vector<int> want(kM * kN);
  long bad = 0;
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < kN; ++n) {
      int w = 0;
      for (int k = 0; k < kelems; ++k) w += bit_of(m, k, 0) & bit_of(n, k, 1);
      want[m * kN + n] = w;
      if (got[m * kN + n] != w) ++bad;
    }
  // accumulation over many instructions: 256 passes (sums up to 256 * kelems) must stay exact integers
  tc05_kernel<KIND><<<1, 128>>>(256, d_out);
  cudaDeviceSynchronize();
  cudaMemcpy(got.data(), d_out, got.size() * 4, cudaMemcpyDeviceToHost);
  long bad256 = 0;
  int maxsum = 0;
  for (int i = 0; i < kM * kN; ++i) {
    if (got[i] != 256 * want[i]) ++bad256;
    if (256 * want[i] > maxsum) maxsum = 256 * want[i];
  }
  int const passes = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    tc05_kernel<KIND><<<sms, 128>>>(passes, nullptr);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  e = cudaGetLastError();
  double const ops = 2.0 * kM * kN * kelems * (double)passes * sms;
  printf("{\"probe\": \"%s\", \"tile\": \"M128 N256 K32B, cta_group::1, SS\", \"single_pass_mismatches\": %ld, "
         "\"mismatches_after_256_passes\": %ld, \"max_sum\": %d, \"ms\": %.3f, \"bitops_per_s\": %.4e, \"sms\": %d, \"err\": \"%s\"}\n",
         name, bad, bad256, maxsum, best, ops / (best * 1e-3), sms, cudaGetErrorString(e));
  cudaFree(d_out);
  return 0;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<0>("tcgen05.mma kind::i8 (uint8 0/1, s32 accumulate in TMEM)", sms);
  run<1>("tcgen05.mma kind::f8f6f4 (e4m3 0/1, f32 accumulate in TMEM)", sms);
  run<2>("tcgen05.mma kind::mxf4.block_scale (e2m1 0/1, ue8m0 scales 2^0, f32 accumulate in TMEM)", sms);
  return 0;
}
