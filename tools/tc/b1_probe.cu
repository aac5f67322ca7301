// b1_probe.cu — peak rate of the legacy 1-bit tensor-core path on sm_100a (BASELINE north star:
// "a b1 AND+POPC tensor-core BMMA variant taken only if sm_100a exposes it and it wins").
// mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.{xor,and}.popc on register operands, no memory traffic:
// an upper bound on what any b1 mma.sync kernel could reach.  GF(2) product needs AND+POPC (parity = LSB).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o b1_probe b1_probe.cu ; SASS: cuobjdump -sass b1_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

template <int OP>
__global__ void __launch_bounds__(256) b1_loop(int iters, uint32_t seed, int *out) {
  uint32_t a0 = seed ^ threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u, b0 = a0 * 11u, b1 = a0 * 13u;
  int c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {   // 8 independent accumulator tiles keep the pipe full
      if (OP == 0)
        asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.xor.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    a0 += 1u;
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= c[i][0] ^ c[i][1] ^ c[i][2] ^ c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int const blocks = sms * 4, threads = 256, iters = 20000;
  int *out;
  cudaMalloc(&out, (size_t)blocks * threads * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int op = 0; op < 2; ++op) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (op == 0) b1_loop<0><<<blocks, threads>>>(iters, 1u, out);
      else         b1_loop<1><<<blocks, threads>>>(iters, 1u, out);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      // per warp and mma: 16 x 8 x 256 bit-MACs = 2 * 32768 bit-ops
      double const ops = 2.0 * 16 * 8 * 256 * 8.0 * iters * (double)blocks * (threads / 32);
      if (rep == 2)
        printf("{\"probe\": \"mma.sync.m16n8k256.b1.%s.popc\", \"ms\": %.3f, \"bitops_per_s\": %.4e, \"sms\": %d, \"err\": \"%s\"}\n",
               op == 0 ? "and" : "xor", ms, ops / (ms * 1e-3), sms, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
