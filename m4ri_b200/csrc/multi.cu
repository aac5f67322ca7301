// multi.cu — mzd_mul_mp / mzd_addmul_mp on several GPUs of one box, in ONE process.
//
// The reference's block-parallel multiply cuts C 2 x 2 over four OpenMP sections (m4ri/mp.c:158-275, the
// split :179-228).  Here C is cut into pr x pc blocks, one per GPU (2 x G/2 from four GPUs on — at four GPUs the
// reference's 2 x 2 shape — so that the per-GPU product keeps half of the rows and the 4096-row leaves): GPU (gr, gc) owns
//     C[rows gr, cols gc] (^)= A[rows gr, :] * B[:, cols gc].
// Every operand bit crosses PCIe ONCE per box and the rest travels over NVLink: the pr GPUs of a column group
// each upload 1/pr of B[:, cols gc] (a row-slice), the pc GPUs of a row group each upload 1/pc of the rows of
// A[rows gr, :], and the pieces are exchanged by peer copies on the copy engines (no SM involved, so they
// overlap the products; a single-GPU user never needs NCCL — the multi-PROCESS form of the same schedule in
// bench.py uses NCCL all-gathers).  There is no K-sharding of the result: XOR over K-chunks is accumulated
// locally.
//
// Pipeline (same schedule as m4ri_b200/shard.py: pipelined_product): the K range is cut into pr * sub chunks,
// chunk (g, j) = sub-chunk j of the row-slice of B that GPU (g, gc) uploads.  One host thread per GPU
//   uploads its pieces in the order it will consume them (own slice first, then g = gr + 1, ... mod pr),
//   pulls the peers' pieces as soon as their upload events exist,
//   enqueues C (^)= A_chunk * B_chunk (Strassen-Winograd over the M4RM leaf) per chunk,
//   and downloads the C block in two row parts, the first while the second is still being computed.
// XOR-accumulation over K-chunks is exact in any order.
#include <string.h>

#include <atomic>
#include <memory>
#include <thread>

#include "dev.h"
#include "staging.h"
#include "workspace.h"

namespace m4b {
namespace {

constexpr int kTail = 2;          // row parts of the last chunk's product (download overlap)

struct Dev {
  int          id = 0;
  cudaStream_t compute = nullptr, up = nullptr, xfer = nullptr, down = nullptr;
  Workspace    ws;
  Stager       stager;        // pinned ring for pageable host rows (one per GPU: the transfers run concurrently)
  std::vector<word> tmp[kTail];
};
std::vector<std::unique_ptr<Dev>> devs;   // Dev holds a Stager (threads, mutex): not movable
Dev &dev(int g) { return *devs[g]; }
int  base_device = 0;

void setup(int G, int base) {
  if ((int)devs.size() == G && base_device == base) return;
  multi_release();
  base_device = base;
  for (int g = 0; g < G; ++g) devs.emplace_back(new Dev);
  for (int g = 0; g < G; ++g) {
    Dev &d = dev(g);
    d.id = base + g;
    M4B_CUDA(cudaSetDevice(d.id));
    for (cudaStream_t *s : {&d.compute, &d.up, &d.xfer, &d.down}) M4B_CUDA(cudaStreamCreateWithFlags(s, cudaStreamNonBlocking));
    for (int p = 0; p < G; ++p) {
      if (p == g) continue;
      int can = 0;
      M4B_CUDA(cudaDeviceCanAccessPeer(&can, d.id, base + p));
      if (can) {
        cudaError_t e = cudaDeviceEnablePeerAccess(base + p, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) M4B_CUDA(e);
        cudaGetLastError();
      }   // without peer access cudaMemcpyPeerAsync still works (staged through the host)
    }
  }
}

inline int64_t round_up64(int64_t v, int64_t mult) { return (v + mult - 1) / mult * mult; }
inline int     imin(int a, int b) { return a < b ? a : b; }
inline int     imax(int a, int b) { return a > b ? a : b; }

// a clipped window of a host matrix (header only; same words); c0 % 64 == 0
mzd_t host_window(mzd_t const *M, int r0, int c0, int r1, int c1) {
  mzd_t W = *M;
  r1 = imin(r1, M->nrows);
  c1 = imin(c1, M->ncols);
  W.nrows = imax(r1 - r0, 0);
  W.ncols = imax(c1 - c0, 0);
  W.width = (W.ncols + 63) / 64;
  W.high_bitmask = ~(word)0 >> ((64 - W.ncols % 64) % 64);
  W.flags = 0x4 | (W.ncols % 64 ? 0x2 : 0);
  W.data = M->data + (int64_t)r0 * M->rowstride + c0 / 64;
  return W;
}

// One event per (producer GPU, piece); `posted` tells a consumer thread that the record call has been issued
// (cudaStreamWaitEvent on an event that was never recorded would not wait at all).
struct Signal {
  cudaEvent_t      ev = nullptr;
  std::atomic<int> posted{0};
  void post(cudaStream_t s) {
    M4B_CUDA(cudaEventRecord(ev, s));
    posted.store(1, std::memory_order_release);
  }
  void await_on(cudaStream_t s) {
    while (!posted.load(std::memory_order_acquire)) std::this_thread::yield();
    M4B_CUDA(cudaStreamWaitEvent(s, ev, 0));
  }
};

int ksub_default() {
  static int const v = [] {
    char const *env = getenv("M4RI_B200_MP_KSUB");
    int const k = env ? atoi(env) : 0;
    return k >= 1 && k <= 16 ? k : 2;     // measured on 2 GPUs, pageable host matrices: 90.0 ms (1) vs 80.7 ms (2)
  }();
  return v;
}

}  // namespace

void multi_release() {
  for (auto &d : devs) {
    cudaSetDevice(d->id);
    for (cudaStream_t s : {d->compute, d->up, d->xfer, d->down}) {
      cudaStreamSynchronize(s);
      cudaStreamDestroy(s);
    }
    d->ws.destroy();
    d->stager.release();
  }
  devs.clear();
}

void multi_grid(int G, int n, int *pr, int *pc) {
  // two row-blocks x G/2 column blocks from four GPUs on: the local product keeps half of the rows, and with them
  // the 4096-row leaves of one more Strassen level (m4ri_b200/shard.py: grid_shape)
  *pc = (G >= 4 && G % 2 == 0 && n >= 128 * (G / 2)) ? G / 2 : 1;
  *pr = G / *pc;
}

void multi_product(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff, bool clear, int G, int base, char *path_out,
                   size_t path_len) {
  int const m = A->nrows, l = A->ncols, n = B->ncols;
  if (m == 0 || n == 0) return;
  int prev = 0;
  M4B_CUDA(cudaGetDevice(&prev));
  setup(G, base);

  int pr, pc;
  multi_grid(G, n, &pr, &pc);
  int const sub = ksub_default(), S = pr * sub;
  // geometry: row blocks of 64-row granules, column blocks and K-chunks of (128 << levels)-bit granules, so that
  // every device view is 16-byte aligned and every Strassen level of a chunk product halves exactly
  int const rb0 = (int)round_up64((m + pr - 1) / pr, 64), cb0 = (n + pc - 1) / pc, kc0 = imax((l + S - 1) / S, 1);
  int const levels = l > 0 ? strassen_levels(imin(rb0, m), kc0, cb0, cutoff) : 0;
  int const rb  = (int)round_up64(rb0, (int64_t)kTail * pc * (1 << levels));   // rows per block on the device
  int const cb  = (int)round_up64(cb0, 128LL << levels);
  int const kc  = (int)round_up64(kc0, 128LL << levels);
  int const part = rb / pc;                                                    // A rows one GPU of a row group uploads
  snprintf(path_out, path_len, levels ? "mp%d:%dx%d:k%d:strassen:%d" : "mp%d:%dx%d:k%d:m4rm", G, pr, pc, S, levels);

  // signals: A parts [producer device][chunk], B chunks [producer device][j]
  std::vector<Signal> sigA((size_t)G * S), sigB((size_t)G * sub), sigZero(G);
  auto make_event = [](Signal &s, int device) {
    M4B_CUDA(cudaSetDevice(device));
    M4B_CUDA(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
  };
  for (int d = 0; d < G; ++d) {
    for (int c = 0; c < S; ++c) make_event(sigA[(size_t)d * S + c], dev(d).id);
    for (int j = 0; j < sub; ++j) make_event(sigB[(size_t)d * sub + j], dev(d).id);
    make_event(sigZero[d], dev(d).id);
  }

  struct Bufs {
    std::vector<DView> Ac, Bc;   // per chunk index c = g * sub + j
    DView Cb;
  };
  std::vector<Bufs> bufs(G);
  // buffers first (every thread needs its peers' addresses), then the pipelines
  for (int d = 0; d < G; ++d) {
    Dev &D = dev(d);
    M4B_CUDA(cudaSetDevice(D.id));
    size_t need = Workspace::bytes_for(rb, cb) + strassen_workspace_bytes(rb, kc, cb, levels);
    need += (size_t)S * (Workspace::bytes_for(rb, kc) + Workspace::bytes_for(kc, cb));
    D.ws.reserve(need);
    Bufs &b = bufs[d];
    b.Ac.resize(S);
    b.Bc.resize(S);
    for (int c = 0; c < S; ++c) b.Ac[c] = D.ws.alloc(rb, kc);
    for (int c = 0; c < S; ++c) b.Bc[c] = D.ws.alloc(kc, cb);
    b.Cb = D.ws.alloc(rb, cb);
  }

  std::vector<std::thread> threads;
  for (int d = 0; d < G; ++d) {
    threads.emplace_back([&, d] {
      Dev &D = dev(d);
      M4B_CUDA(cudaSetDevice(D.id));
      Bufs &b = bufs[d];
      int const gr = d / pc, gc = d % pc;
      int const r0 = imin(gr * rb0, m), r1 = imin((gr + 1) * rb0, m);     // host rows of this block
      int const c0 = imin(gc * cb, n), c1 = imin((gc + 1) * cb, n);       // host columns of this block
      bool const has_c = r1 > r0 && c1 > c0;

      // padding must read as zeros: clear everything once, uploads and peer copies are ordered after it
      for (int c = 0; c < S; ++c) {
        zero_async(b.Ac[c], D.compute);
        zero_async(b.Bc[c], D.compute);
      }
      if (!clear) zero_async(b.Cb, D.compute);
      sigZero[d].post(D.compute);
      M4B_CUDA(cudaStreamWaitEvent(D.up, sigZero[d].ev, 0));
      M4B_CUDA(cudaStreamWaitEvent(D.xfer, sigZero[d].ev, 0));

      cudaEvent_t evC = nullptr, evTail[kTail] = {};
      if (!clear && has_c) {
        mzd_t Cw = host_window(C, r0, c0, r1, c1);
        upload(b.Cb.sub(0, 0, Cw.nrows, cb), &Cw, D.up, &D.stager);
        M4B_CUDA(cudaEventCreateWithFlags(&evC, cudaEventDisableTiming));
        M4B_CUDA(cudaEventRecord(evC, D.up));
      }
      std::vector<cudaEvent_t> evReady(S, nullptr);

      // schedule: own slice first, then the others in rotated order
      for (int idx = 0; idx < S; ++idx) {
        int const g = (gr + idx / sub) % pr, j = idx % sub, c = g * sub + j;
        int const k0 = c * kc;                                            // host columns of A / rows of B
        bool const own = g == gr;
        // ---- uploads of this position --------------------------------------------------------------
        if (own) {      // B sub-chunk j of the slice this GPU contributes to its column group
          if (k0 < l && c1 > c0) {
            mzd_t Bw = host_window(B, k0, c0, k0 + kc, c1);
            if (Bw.nrows > 0) upload(b.Bc[c].sub(0, 0, Bw.nrows, cb), &Bw, D.up, &D.stager);
          }
          sigB[(size_t)d * sub + j].post(D.up);
        }
        {               // this GPU's row part of A chunk c (it serves the whole row group)
          int const pr0 = r0 + gc * part, pr1 = imin(pr0 + part, r1);
          if (pr1 > pr0 && k0 < l) {
            mzd_t Aw = host_window(A, pr0, k0, pr1, k0 + kc);
            if (Aw.ncols > 0) upload(b.Ac[c].sub(gc * part, 0, gc * part + Aw.nrows, kc), &Aw, D.up, &D.stager);
          }
          sigA[(size_t)d * S + c].post(D.up);
        }
        // ---- pull the peers' pieces over NVLink (copy engines) -----------------------------------------
        if (has_c) {
          for (int q = 0; q < pc; ++q) {
            int const p = gr * pc + q;
            if (q == gc) {
              M4B_CUDA(cudaStreamWaitEvent(D.xfer, sigA[(size_t)d * S + c].ev, 0));
              continue;
            }
            sigA[(size_t)p * S + c].await_on(D.xfer);
            size_t const off = (size_t)q * part * (size_t)b.Ac[c].pitch;
            M4B_CUDA(cudaMemcpyPeerAsync(b.Ac[c].data + off, D.id, bufs[p].Ac[c].data + off, dev(p).id,
                                         (size_t)part * (size_t)b.Ac[c].pitch * 8, D.xfer));
          }
          int const pb = g * pc + gc;                                     // producer of B chunk (g, j) in this column group
          if (pb == d) {
            M4B_CUDA(cudaStreamWaitEvent(D.xfer, sigB[(size_t)d * sub + j].ev, 0));
          } else {
            sigB[(size_t)pb * sub + j].await_on(D.xfer);
            M4B_CUDA(cudaMemcpyPeerAsync(b.Bc[c].data, D.id, bufs[pb].Bc[c].data, dev(pb).id,
                                         (size_t)kc * (size_t)b.Bc[c].pitch * 8, D.xfer));
          }
          M4B_CUDA(cudaEventCreateWithFlags(&evReady[c], cudaEventDisableTiming));
          M4B_CUDA(cudaEventRecord(evReady[c], D.xfer));
          // ---- the product of this chunk ----------------------------------------------------------------
          M4B_CUDA(cudaStreamWaitEvent(D.compute, evReady[c], 0));
          if (evC) M4B_CUDA(cudaStreamWaitEvent(D.compute, evC, 0));
          bool const clr = clear && idx == 0;
          if (idx + 1 < S) {
            strassen_mul(b.Cb, b.Ac[c], b.Bc[c], levels, clr, D.ws, D.compute);
          } else {
            for (int t = 0; t < kTail; ++t) {
              int const t0 = t * (rb / kTail), t1 = (t + 1) * (rb / kTail);
              strassen_mul(b.Cb.sub(t0, 0, t1, cb), b.Ac[c].sub(t0, 0, t1, kc), b.Bc[c], levels, clr, D.ws, D.compute);
              M4B_CUDA(cudaEventCreateWithFlags(&evTail[t], cudaEventDisableTiming));
              M4B_CUDA(cudaEventRecord(evTail[t], D.compute));
            }
          }
        }
      }
      // ---- downloads, issued after every product has been enqueued (a D2H copy blocks this thread) ---------
      mzd_t parts[kTail];
      if (has_c) {
        for (int t = 0; t < kTail; ++t) {
          int const t0 = t * (rb / kTail), t1 = (t + 1) * (rb / kTail);
          parts[t] = host_window(C, r0 + t0, c0, imin(r0 + t1, r1), c1);
          M4B_CUDA(cudaStreamWaitEvent(D.down, evTail[t], 0));
          if (parts[t].nrows > 0)
            download(&parts[t], b.Cb.sub(t0, 0, t0 + parts[t].nrows, cb), D.down, D.tmp[t], &D.stager);
        }
      }
      for (cudaStream_t s : {D.down, D.compute, D.xfer, D.up}) M4B_CUDA(cudaStreamSynchronize(s));
      if (evC) cudaEventDestroy(evC);
      for (cudaEvent_t e : evTail) if (e) cudaEventDestroy(e);
      for (cudaEvent_t e : evReady) if (e) cudaEventDestroy(e);
    });
  }
  for (auto &t : threads) t.join();
  // a peer may still have been reading this GPU's buffers when its own thread finished: all threads have joined
  // (every stream of every GPU is idle) before the buffers are released and the events destroyed
  for (int d = 0; d < G; ++d) dev(d).ws.release(0);
  for (auto *v : {&sigA, &sigB, &sigZero})
    for (Signal &s : *v) cudaEventDestroy(s.ev);
  M4B_CUDA(cudaSetDevice(prev));
}

}  // namespace m4b
