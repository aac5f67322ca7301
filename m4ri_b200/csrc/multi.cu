// multi.cu — mzd_mul_mp / mzd_addmul_mp on several GPUs of one box, in ONE process.
//
// The reference's block-parallel multiply splits C 2x2 over four OpenMP sections
// (m4ri/mp.c:158-275).  Rows of C are independent, so here C and A are split into G contiguous
// row-blocks, one per GPU; B is needed by everybody: each GPU uploads only its 1/G row-slice of B
// from the host and the slices are exchanged with ONE ncclAllGather over NVLink (in place, uint64
// words).  There is no K-sharding: NCCL has no XOR reduction, and none is needed.
//
//   phase 1  (one host thread per GPU)  H2D of the A row-block and the B row-slice
//   phase 2  (calling thread)           grouped ncclAllGather, then the Strassen/M4RM schedule per GPU
//   phase 3  (one host thread per GPU)  D2H of the C row-block
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): a single-GPU user never needs it.  The
// multi-PROCESS form of the same split (one rank per GPU, torch.distributed) is in bench.py.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <memory>
#include <thread>

#include "dev.h"
#include "staging.h"
#include "workspace.h"

namespace m4b {
namespace {

struct Nccl {
  void *handle = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
Nccl nccl;

void load_nccl() {
  if (nccl.handle) return;
  char const *names[] = {"libnccl.so.2", "libnccl.so"};
  for (char const *nm : names)
    if ((nccl.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break;
  if (!nccl.handle) die("m4ri_b200: multi-GPU needs NCCL but libnccl.so.2 could not be loaded (%s)\n", dlerror());
#define M4B_SYM(field, name)                                                      \
  *reinterpret_cast<void **>(&nccl.field) = dlsym(nccl.handle, name);            \
  if (!nccl.field) die("m4ri_b200: symbol %s missing from libnccl\n", name)
  M4B_SYM(CommInitAll, "ncclCommInitAll");
  M4B_SYM(CommDestroy, "ncclCommDestroy");
  M4B_SYM(GroupStart, "ncclGroupStart");
  M4B_SYM(GroupEnd, "ncclGroupEnd");
  M4B_SYM(AllGather, "ncclAllGather");
  M4B_SYM(GetErrorString, "ncclGetErrorString");
#undef M4B_SYM
}

#define M4B_NCCL(expr)                                                                                     \
  do {                                                                                                     \
    ncclResult_t r_ = (expr);                                                                              \
    if (r_ != ncclSuccess) die("m4ri_b200: NCCL error at %s:%d: %s\n", __FILE__, __LINE__, nccl.GetErrorString(r_)); \
  } while (0)

struct Dev {
  int          id = 0;
  cudaStream_t stream = nullptr;
  Workspace    ws;
  Stager       stager;        // pinned ring for pageable host rows (one per GPU: the uploads run concurrently)
  std::vector<word> tmp;
};
std::vector<std::unique_ptr<Dev>> devs;   // Dev holds a Stager (threads, mutex): not movable
Dev &dev(int g) { return *devs[g]; }
std::vector<ncclComm_t> comms;

void setup(int G) {
  if ((int)devs.size() == G) return;
  multi_release();
  load_nccl();
  for (int g = 0; g < G; ++g) devs.emplace_back(new Dev);
  std::vector<int> ids(G);
  for (int g = 0; g < G; ++g) {
    dev(g).id = ids[g] = g;
    M4B_CUDA(cudaSetDevice(g));
    M4B_CUDA(cudaStreamCreateWithFlags(&dev(g).stream, cudaStreamNonBlocking));
  }
  comms.resize(G);
  M4B_NCCL(nccl.CommInitAll(comms.data(), G, ids.data()));
}

inline int64_t round_up64(int64_t v, int64_t mult) { return (v + mult - 1) / mult * mult; }
int64_t gcd64(int64_t a, int64_t b) { return b ? gcd64(b, a % b) : a; }

template <class F>
void per_device(int G, F &&f) {
  std::vector<std::thread> th;
  for (int g = 0; g < G; ++g) th.emplace_back([&, g] { M4B_CUDA(cudaSetDevice(dev(g).id)); f(g); });
  for (auto &t : th) t.join();
}

// a window of rows [r0, r1) of a host matrix (no allocation; same words)
mzd_t row_window(mzd_t const *M, int r0, int r1) {
  mzd_t W = *M;
  W.nrows = r1 - r0;
  W.flags |= 0x4;
  W.data = M->data + (int64_t)r0 * M->rowstride;
  return W;
}

}  // namespace

void multi_release() {
  for (auto &c : comms)
    if (nccl.CommDestroy) nccl.CommDestroy(c);
  comms.clear();
  for (auto &d : devs) {
    cudaSetDevice(d->id);
    cudaStreamSynchronize(d->stream);
    d->ws.destroy();
    d->stager.release();
    cudaStreamDestroy(d->stream);
  }
  devs.clear();
}

void multi_product(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff, bool clear, int num_devices, char *path_out,
                   size_t path_len) {
  int const m = A->nrows, l = A->ncols, n = B->ncols;
  if (m == 0 || n == 0) return;
  int avail = 0;
  M4B_CUDA(cudaGetDeviceCount(&avail));
  int G = num_devices < avail ? num_devices : avail;
  int prev = 0;
  M4B_CUDA(cudaGetDevice(&prev));
  setup(G);

  // row-blocks of A/C: multiples of 64 rows, the last ones may be short or empty
  int const rb = (int)round_up64((m + G - 1) / G, 64);
  int const levels = l > 0 ? strassen_levels(rb < m ? rb : m, l, n, cutoff) : 0;
  // B's rows are padded so that G equal slices exist and every Strassen level halves on 128 bits
  int64_t const a = 64 * (int64_t)G, b = 128LL << levels;
  int64_t const lp = round_up64(l > 0 ? l : 1, a / gcd64(a, b) * b);
  int const lb = (int)(lp / G);
  int const np = (int)round_up64(n, 128LL << levels);
  int const mpb = (int)round_up64(rb, 1 << levels);
  snprintf(path_out, path_len, levels ? "mp%d:strassen:%d" : "mp%d:m4rm", G, levels);

  std::vector<DView> dA(G), dB(G), dC(G);
  per_device(G, [&](int g) {
    Dev &d = dev(g);
    d.ws.reserve(Workspace::bytes_for(mpb, (int)lp) + Workspace::bytes_for((int)lp, np) + Workspace::bytes_for(mpb, np) +
                 strassen_workspace_bytes(mpb, (int)lp, np, levels));
    dA[g] = d.ws.alloc(mpb, (int)lp);
    dB[g] = d.ws.alloc((int)lp, np);
    dC[g] = d.ws.alloc(mpb, np);
    zero_async(dA[g], d.stream);
    zero_async(dB[g], d.stream);
    int const r0 = g * rb < m ? g * rb : m, r1 = (g + 1) * rb < m ? (g + 1) * rb : m;
    if (r1 > r0) {
      mzd_t Ablk = row_window(A, r0, r1);
      upload(dA[g].sub(0, 0, r1 - r0, (int)lp), &Ablk, d.stream, &d.stager);
      if (!clear) {
        zero_async(dC[g], d.stream);
        mzd_t Cblk = row_window(C, r0, r1);
        upload(dC[g].sub(0, 0, r1 - r0, np), &Cblk, d.stream, &d.stager);
      }
    }
    int const s0 = g * lb < l ? g * lb : l, s1 = (g + 1) * lb < l ? (g + 1) * lb : l;
    if (s1 > s0) {
      mzd_t Bsl = row_window(B, s0, s1);
      upload(dB[g].sub(g * lb, 0, g * lb + (s1 - s0), np), &Bsl, d.stream, &d.stager);
    }
  });

  // the one exchange step: in-place all-gather of B's row-slices over NVLink
  size_t const slice_words = (size_t)lb * (size_t)dB[0].pitch;
  M4B_NCCL(nccl.GroupStart());
  for (int g = 0; g < G; ++g)
    M4B_NCCL(nccl.AllGather(dB[g].data + (size_t)g * slice_words, dB[g].data, slice_words, ncclUint64, comms[g],
                            dev(g).stream));
  M4B_NCCL(nccl.GroupEnd());

  for (int g = 0; g < G; ++g) {
    int const r0 = g * rb < m ? g * rb : m, r1 = (g + 1) * rb < m ? (g + 1) * rb : m;
    if (r1 <= r0) continue;
    M4B_CUDA(cudaSetDevice(dev(g).id));
    strassen_mul(dC[g], dA[g], dB[g], levels, clear, dev(g).ws, dev(g).stream);
  }

  per_device(G, [&](int g) {
    Dev &d = dev(g);
    int const r0 = g * rb < m ? g * rb : m, r1 = (g + 1) * rb < m ? (g + 1) * rb : m;
    if (r1 > r0) {
      mzd_t Cblk = row_window(C, r0, r1);
      download(&Cblk, dC[g].sub(0, 0, r1 - r0, np), d.stream, d.tmp, &d.stager);
    }
    M4B_CUDA(cudaStreamSynchronize(d.stream));
    d.ws.release(0);
  });
  M4B_CUDA(cudaSetDevice(prev));
}

}  // namespace m4b
