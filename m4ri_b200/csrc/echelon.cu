// echelon.cu — reduced row echelon form on the device (kernels: echelon_body.h; the row update of every
// 64-column strip is a product on the M4RM leaf).  First cut of SURVEY.md §8f item 1: the job of the reference's
// mzd_echelonize_m4ri(A, 1, k) (m4ri/brilliantrussian.c:603-967).  Exposed as m4ri_b200_dechelonize /
// m4ri_b200_echelonize only (capi.cu): the libm4ri symbol itself is not interposed yet.
#include <cuda_runtime.h>

#include "dev.h"
#include "workspace.h"

#define ECH_FN __device__ __forceinline__
#include "echelon_body.h"

namespace m4b {

namespace {

struct DevCtx {
  int tid, ntid;
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ int atomic_min(int *p, int v) { return atomicMin(p, v); }
};

using ech::State;
using ech::u64;

__global__ void __launch_bounds__(ech::kSelThreads) k_select_chunk(State const *st, u64 const *A, long long pitch, int m,
                                                                   int wcol, int chunk_rows, int *cand_row, u64 *cand_word) {
  __shared__ ech::SelShared sh;
  DevCtx cx{(int)threadIdx.x, (int)blockDim.x};
  ech::select_chunk(cx, &sh, st, A, pitch, m, wcol, chunk_rows, (int)blockIdx.x, cand_row, cand_word);
}

__global__ void __launch_bounds__(ech::kSelThreads) k_select_final(State *st, int const *cand_row, u64 const *cand_word,
                                                                   int ncand, u64 *Gm) {
  __shared__ ech::SelShared sh;
  DevCtx cx{(int)threadIdx.x, (int)blockDim.x};
  ech::select_final(cx, &sh, st, cand_row, cand_word, ncand, Gm);
}

#define ECH_GRID long long const gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (long long)gridDim.x * blockDim.x

__global__ void __launch_bounds__(256) k_gather(State const *st, u64 const *A, long long pitchA, int w0, int nw, u64 *PIV,
                                                long long pitchP) {
  ECH_GRID;
  ech::gather_pivots(st, A, pitchA, w0, nw, PIV, pitchP, gtid, gthreads);
}
__global__ void __launch_bounds__(256) k_build_x(State const *st, u64 const *A, long long pitchA, int wcol, int m, u64 *X) {
  ECH_GRID;
  ech::build_x(st, A, pitchA, wcol, m, X, gtid, gthreads);
}
__global__ void __launch_bounds__(256) k_move(State const *st, u64 *A, long long pitchA, int w0, int nw) {
  ECH_GRID;
  ech::move_rows(st, A, pitchA, w0, nw, gtid, gthreads);
}
__global__ void __launch_bounds__(256) k_place(State const *st, u64 *A, long long pitchA, int w0, int nw, u64 const *Bm,
                                               long long pitchB) {
  ECH_GRID;
  ech::place_pivots(st, A, pitchA, w0, nw, Bm, pitchB, gtid, gthreads);
}
__global__ void k_advance(State *st) { ech::advance(st); }

int chunk_rows_for(int m) {
  static int const forced = [] {        // test knob: small chunks exercise the candidate merge on small matrices
    char const *env = getenv("M4RI_B200_ECH_CHUNK");
    return env && atoi(env) > 0 ? atoi(env) : 0;
  }();
  int rows = forced ? forced : ech::kSelRows;
  if (rows > ech::kSelRows) rows = ech::kSelRows;
  while ((long long)((m + rows - 1) / rows) * 64 > ech::kSelRows) {   // the final CTA takes 64 candidates per chunk
    if (rows == ech::kSelRows) die("m4ri_b200: echelonize supports at most %d rows\n", ech::kSelRows / 64 * ech::kSelRows);
    rows *= 2;
    if (rows > ech::kSelRows) rows = ech::kSelRows;
  }
  return rows;
}

size_t raw_bytes(size_t b) { return (b + 255) & ~(size_t)255; }

}  // namespace

size_t echelon_workspace_bytes(int m, int n, int64_t pitch_words) {
  int const nchunks = (m + chunk_rows_for(m) - 1) / chunk_rows_for(m);
  // PIV / Bm rows are as long as A's rows: a wrapped matrix may have a larger pitch than the minimal one
  size_t const minimal = (size_t)Workspace::pitch_for(n);
  size_t const pitch = pitch_words > (int64_t)minimal ? (size_t)pitch_words : minimal;
  return raw_bytes(sizeof(State)) + raw_bytes((size_t)nchunks * 64 * 4) + raw_bytes((size_t)nchunks * 64 * 8) +
         raw_bytes(128 * 8) + raw_bytes((size_t)m * 16) + 2 * raw_bytes(64 * pitch * 8) + 4096;
}

// A -> its reduced row echelon form, in place; returns the rank (synchronises `s`).
int echelonize_device(DView A, Workspace &ws, cudaStream_t s) {
  int const m = A.nrows, n = A.ncols;
  if (m <= 0 || n <= 0) return 0;
  int const chunk_rows = chunk_rows_for(m), nchunks = (m + chunk_rows - 1) / chunk_rows;
  size_t const mark = ws.mark();
  auto raw = [&](size_t bytes) { return reinterpret_cast<char *>(ws.alloc(1, (int)(raw_bytes(bytes) * 8)).data); };
  State *st      = reinterpret_cast<State *>(raw(sizeof(State)));
  int *cand_row  = reinterpret_cast<int *>(raw((size_t)nchunks * 64 * 4));
  u64 *cand_word = reinterpret_cast<u64 *>(raw((size_t)nchunks * 64 * 8));
  u64 *Gm        = reinterpret_cast<u64 *>(raw(128 * 8));
  u64 *X         = reinterpret_cast<u64 *>(raw((size_t)m * 16));
  long long const pitch = A.pitch;
  u64 *PIV = reinterpret_cast<u64 *>(raw((size_t)64 * pitch * 8));
  u64 *Bm  = reinterpret_cast<u64 *>(raw((size_t)64 * pitch * 8));
  u64 *Ad  = reinterpret_cast<u64 *>(A.data);
  M4B_CUDA(cudaMemsetAsync(st, 0, sizeof(State), s));
  // the leaf never writes the padding word past the last column of Bm; place_pivots copies whole rows, so it must be 0
  M4B_CUDA(cudaMemsetAsync(PIV, 0, (size_t)64 * pitch * 8, s));
  M4B_CUDA(cudaMemsetAsync(Bm, 0, (size_t)64 * pitch * 8, s));

  int const sms = m4rm_num_sms();
  auto blocks = [&](long long work) {
    long long b = (work + 255) / 256;
    if (b > 4ll * sms) b = 4ll * sms;
    return (unsigned)(b < 1 ? 1 : b);
  };
  for (int strip = 0; strip * 64 < n; ++strip) {
    int const c0 = (strip * 64) & ~127, w0 = c0 / 64, nw = (int)(pitch - w0), ncols = n - c0;
    k_select_chunk<<<nchunks, ech::kSelThreads, 0, s>>>(st, Ad, pitch, m, strip, chunk_rows, cand_row, cand_word);
    k_select_final<<<1, ech::kSelThreads, 0, s>>>(st, cand_row, cand_word, nchunks * 64, Gm);
    k_gather<<<blocks(64ll * nw), 256, 0, s>>>(st, Ad, pitch, w0, nw, PIV, pitch);
    g_kernel_launches += 3;
    DView const vG{reinterpret_cast<word *>(Gm), 2, 64, 64};
    DView const vPIV{reinterpret_cast<word *>(PIV), pitch, 64, ncols}, vBm{reinterpret_cast<word *>(Bm), pitch, 64, ncols};
    launch_m4rm_overwrite(vBm, vG, vPIV, s);                               // Bm = G * PIV
    k_build_x<<<blocks(m), 256, 0, s>>>(st, Ad, pitch, strip, m, X);
    ++g_kernel_launches;
    DView const vX{reinterpret_cast<word *>(X), 2, m, 64};
    launch_m4rm(A.sub(0, c0, m, n), vX, vBm, s);                           // A[:, c0:] ^= X * Bm
    k_move<<<blocks(64ll * nw), 256, 0, s>>>(st, Ad, pitch, w0, nw);
    k_place<<<blocks(64ll * nw), 256, 0, s>>>(st, Ad, pitch, w0, nw, Bm, pitch);
    k_advance<<<1, 1, 0, s>>>(st);
    g_kernel_launches += 3;
  }
  M4B_CUDA(cudaGetLastError());
  int rank = 0;
  M4B_CUDA(cudaMemcpyAsync(&rank, &st->rank, sizeof rank, cudaMemcpyDeviceToHost, s));
  M4B_CUDA(cudaStreamSynchronize(s));
  ws.release(mark);
  return rank;
}

}  // namespace m4b
