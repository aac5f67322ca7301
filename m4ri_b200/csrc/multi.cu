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
// Pipeline (the in-process form of bench.py's "hooks" mode): every GPU runs the Strassen-Winograd schedule of its block
// with the TOP level on four separately allocated quadrants per operand (strassen_mul_quads) and transfer hooks.  One
// host thread per GPU; when its schedule first needs quadrant q of A (B) it
//   uploads its 1/pc (1/pr) row share of that quadrant and announces it (event + flag),
//   pulls the other shares from the GPUs of its row (column) group as soon as they are announced (copy engines),
//   and makes the compute stream wait for exactly that quadrant;
// result quadrants are downloaded as soon as they are final (three of the four before the last product starts).
// A first version cut the K range into chunks instead (m4ri_b200/shard.py: pipelined_product, still an option of
// bench.py); it overlapped as well but a chunk product is one Strassen level shallower than the whole block's:
// 2 GPUs, 65536^3, pageable host matrices: 88 ms (K-chunks) against 74 ms for the quadrant hooks in bench.py.
#include <string.h>

#include <atomic>
#include <functional>
#include <memory>
#include <thread>

#include "dev.h"
#include "staging.h"
#include "workspace.h"

namespace m4b {
namespace {

constexpr int kTail = 4;          // per-quadrant scratch for partial last words of a download

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

int64_t gcd64(int64_t a, int64_t b) { return b ? gcd64(b, a % b) : a; }

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
  // geometry: every GPU multiplies (rb x lp) * (lp x cb) with at least one Strassen level, whose top level runs on four
  // separately allocated quadrants per operand (strassen_mul_quads): a quadrant is the unit of upload, exchange and
  // download.  Quadrant rows split into pc (A) resp. pr (B) equal shares, every quadrant halves `levels - 1` more times.
  int const rb0 = (m + pr - 1) / pr, cb0 = (n + pc - 1) / pc;
  int levels = l > 0 ? strassen_levels(rb0, l, cb0, cutoff) : 1;
  if (levels < 1) levels = 1;
  int64_t const sub = 1LL << (levels - 1);
  int const rb = (int)round_up64(rb0, 2 * pc * sub * 64 / gcd64(2 * pc * sub, 64));      // multiple of 64 and of 2 * pc * sub
  int const cb = (int)round_up64(cb0, 256 * sub);
  int const lp = (int)round_up64(l > 0 ? l : 1, (int64_t)2 * pr * 128 * sub / gcd64(pr, sub) );   // >= multiple of 2*pr and 256*sub
  int const m2 = rb / 2, k2 = lp / 2, n2 = cb / 2;
  int const share_a = m2 / pc, share_b = k2 / pr;
  snprintf(path_out, path_len, "mp%d:%dx%d:strassen:%d", G, pr, pc, levels);

  // signals [producer GPU][quadrant]
  std::vector<Signal> sigA((size_t)G * 4), sigB((size_t)G * 4), sigZero(G);
  auto make_event = [](Signal &s, int device) {
    M4B_CUDA(cudaSetDevice(device));
    M4B_CUDA(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
  };
  for (int d = 0; d < G; ++d) {
    for (int q = 0; q < 4; ++q) {
      make_event(sigA[(size_t)d * 4 + q], dev(d).id);
      make_event(sigB[(size_t)d * 4 + q], dev(d).id);
    }
    make_event(sigZero[d], dev(d).id);
  }

  struct Bufs {
    DView a[4], b[4], c[4];
  };
  std::vector<Bufs> bufs(G);
  // buffers first (every thread needs its peers' addresses), then the pipelines
  for (int d = 0; d < G; ++d) {
    Dev &D = dev(d);
    M4B_CUDA(cudaSetDevice(D.id));
    size_t const need = 4 * (Workspace::bytes_for(m2, k2) + Workspace::bytes_for(k2, n2) + Workspace::bytes_for(m2, n2)) +
                        strassen_workspace_bytes(rb, lp, cb, levels) + 3 * Workspace::bytes_for(m2 > k2 ? m2 : k2, k2 > n2 ? k2 : n2);
    D.ws.reserve(need);
    Bufs &b = bufs[d];
    for (int q = 0; q < 4; ++q) b.a[q] = D.ws.alloc(m2, k2);
    for (int q = 0; q < 4; ++q) b.b[q] = D.ws.alloc(k2, n2);
    for (int q = 0; q < 4; ++q) b.c[q] = D.ws.alloc(m2, n2);
    // a fixed share of the host's copy threads per GPU (all GPUs stage pageable rows at the same time)
    int const hw = (int)std::thread::hardware_concurrency();
    D.stager.set_threads(hw / G < 2 ? 2 : hw / G);
  }

  std::vector<std::thread> threads;
  for (int d = 0; d < G; ++d) {
    threads.emplace_back([&, d] {
      Dev &D = dev(d);
      M4B_CUDA(cudaSetDevice(D.id));
      Bufs &b = bufs[d];
      int const gr = d / pc, gc = d % pc;
      int const r0 = imin(gr * rb, m), r1 = imin((gr + 1) * rb, m);       // host rows of this block
      int const c0 = imin(gc * cb, n), c1 = imin((gc + 1) * cb, n);       // host columns of this block
      bool const has_c = r1 > r0 && c1 > c0;

      // padding must read as zeros: clear everything once, uploads and peer copies are ordered after it
      for (int q = 0; q < 4; ++q) {
        zero_async(b.a[q], D.compute);
        zero_async(b.b[q], D.compute);
        if (!clear) zero_async(b.c[q], D.compute);
      }
      sigZero[d].post(D.compute);
      M4B_CUDA(cudaStreamWaitEvent(D.up, sigZero[d].ev, 0));
      M4B_CUDA(cudaStreamWaitEvent(D.xfer, sigZero[d].ev, 0));

      // Transfer hooks of the top Strassen level: a quadrant is uploaded (this GPU's share), announced, completed from
      // the peers of the row / column group over NVLink, and the compute stream waits for exactly that.
      struct Hooks : TopHooks {
        std::function<void(int)> fa, fb, fc, fd;
        void need_a(int q) override { fa(q); }
        void need_b(int q) override { fb(q); }
        void need_c(int q) override { fc(q); }
        void done_c(int q) override { fd(q); }
      } hk;
      bool upA[4] = {}, upB[4] = {}, upC[4] = {};
      std::vector<cudaEvent_t> events;
      auto event = [&] {
        cudaEvent_t e;
        M4B_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        events.push_back(e);
        return e;
      };
      auto post_a = [&](int q) {          // upload this GPU's row share of A quadrant q and announce it
        if (upA[q]) return;
        upA[q] = true;
        int const hr0 = r0 + (q >> 1) * m2 + gc * share_a, hc0 = (q & 1) * k2;
        if (hr0 < r1 && hc0 < l) {
          mzd_t W = host_window(A, hr0, hc0, imin(hr0 + share_a, imin(r0 + ((q >> 1) + 1) * m2, r1)), hc0 + k2);
          if (W.nrows > 0 && W.ncols > 0) upload(b.a[q].sub(gc * share_a, 0, gc * share_a + W.nrows, k2), &W, D.up, &D.stager);
        }
        sigA[(size_t)d * 4 + q].post(D.up);
      };
      auto post_b = [&](int q) {          // this GPU's row share of quadrant q of B[:, cols gc]
        if (upB[q]) return;
        upB[q] = true;
        int const hr0 = (q >> 1) * k2 + gr * share_b, hc0 = c0 + (q & 1) * n2;
        if (hr0 < l && hc0 < c1) {
          mzd_t W = host_window(B, hr0, hc0, imin(hr0 + share_b, l), imin(hc0 + n2, c1));
          if (W.nrows > 0 && W.ncols > 0) upload(b.b[q].sub(gr * share_b, 0, gr * share_b + W.nrows, n2), &W, D.up, &D.stager);
        }
        sigB[(size_t)d * 4 + q].post(D.up);
      };
      hk.fa = [&](int q) {
        bool const first = !upA[q];
        post_a(q);
        if (!first) return;
        for (int p = 0; p < pc; ++p) {
          int const peer = gr * pc + p;
          if (p == gc) {
            M4B_CUDA(cudaStreamWaitEvent(D.xfer, sigA[(size_t)d * 4 + q].ev, 0));
            continue;
          }
          sigA[(size_t)peer * 4 + q].await_on(D.xfer);
          size_t const off = (size_t)p * share_a * (size_t)b.a[q].pitch;
          M4B_CUDA(cudaMemcpyPeerAsync(b.a[q].data + off, D.id, bufs[peer].a[q].data + off, dev(peer).id,
                                       (size_t)share_a * (size_t)b.a[q].pitch * 8, D.xfer));
        }
        cudaEvent_t e = event();
        M4B_CUDA(cudaEventRecord(e, D.xfer));
        M4B_CUDA(cudaStreamWaitEvent(D.compute, e, 0));
      };
      hk.fb = [&](int q) {
        bool const first = !upB[q];
        post_b(q);
        if (!first) return;
        for (int p = 0; p < pr; ++p) {
          int const peer = p * pc + gc;
          if (p == gr) {
            M4B_CUDA(cudaStreamWaitEvent(D.xfer, sigB[(size_t)d * 4 + q].ev, 0));
            continue;
          }
          sigB[(size_t)peer * 4 + q].await_on(D.xfer);
          size_t const off = (size_t)p * share_b * (size_t)b.b[q].pitch;
          M4B_CUDA(cudaMemcpyPeerAsync(b.b[q].data + off, D.id, bufs[peer].b[q].data + off, dev(peer).id,
                                       (size_t)share_b * (size_t)b.b[q].pitch * 8, D.xfer));
        }
        cudaEvent_t e = event();
        M4B_CUDA(cudaEventRecord(e, D.xfer));
        M4B_CUDA(cudaStreamWaitEvent(D.compute, e, 0));
      };
      auto c_window = [&](int q) {
        int const hr0 = r0 + (q >> 1) * m2, hc0 = c0 + (q & 1) * n2;
        if (hr0 >= r1 || hc0 >= c1) {
          mzd_t W = *C;
          W.nrows = W.ncols = 0;
          return W;
        }
        return host_window(C, hr0, hc0, imin(hr0 + m2, r1), imin(hc0 + n2, c1));
      };
      hk.fc = [&](int q) {
        if (upC[q]) return;
        upC[q] = true;
        mzd_t W = c_window(q);
        if (W.nrows > 0 && W.ncols > 0) upload(b.c[q].sub(0, 0, W.nrows, n2), &W, D.up, &D.stager);
        cudaEvent_t e = event();
        M4B_CUDA(cudaEventRecord(e, D.up));
        M4B_CUDA(cudaStreamWaitEvent(D.compute, e, 0));
      };
      cudaEvent_t ready[4] = {};
      int order[4], ndone = 0;
      hk.fd = [&](int q) {
        ready[q] = event();
        M4B_CUDA(cudaEventRecord(ready[q], D.compute));
        order[ndone++] = q;
      };

      if (has_c) {
        strassen_mul_quads(b.c, b.a, b.b, levels, clear, D.ws, D.compute, hk);
        // downloads, issued after the whole schedule has been enqueued (a D2H copy blocks this thread)
        mzd_t win[4];
        for (int i = 0; i < ndone; ++i) {
          int const q = order[i];
          win[q] = c_window(q);
          M4B_CUDA(cudaStreamWaitEvent(D.down, ready[q], 0));
          if (win[q].nrows > 0 && win[q].ncols > 0)
            download(&win[q], b.c[q].sub(0, 0, win[q].nrows, n2), D.down, D.tmp[q], &D.stager);
        }
      } else {
        // nothing to compute here, but the peers still need this GPU's shares (in the order their schedules ask for them)
        static int const kOrderA[4] = {1, 3, 2, 0}, kOrderB[4] = {2, 3, 1, 0};
        for (int i = 0; i < 4; ++i) {
          post_b(kOrderB[i]);
          post_a(kOrderA[i]);
        }
      }
      for (cudaStream_t s : {D.down, D.compute, D.xfer, D.up}) M4B_CUDA(cudaStreamSynchronize(s));
      for (cudaEvent_t e : events) cudaEventDestroy(e);
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
