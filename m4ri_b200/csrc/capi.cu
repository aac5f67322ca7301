// capi.cu — the C-ABI of libm4ri_b200.so (include/m4ri_b200.h).
//
// Host side of the drop-in: argument checks and error behaviour of the reference entry points
// (m4ri/strassen.c:345-365, 675-700; m4ri/brilliantrussian.c:999-1028), then
//   host mzd_t --H2D--> zero-padded device matrices --kernels--> --D2H--> host mzd_t
// honouring rowstride, windows and excess bits exactly as the reference does (m4ri/mzd.h:117-122).
// There is no CPU compute path in this file: without a CUDA device every entry point dies.
#include <dlfcn.h>
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "dev.h"
#include "staging.h"
#include "workspace.h"

namespace m4b {

void die(char const *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  abort();
}

static_assert(sizeof(mzd_t) == 64, "mzd_t must be 64 bytes (m4ri/mzd.c:143)");
static_assert(offsetof(mzd_t, nrows) == 0 && offsetof(mzd_t, ncols) == 4 && offsetof(mzd_t, width) == 8 &&
                  offsetof(mzd_t, rowstride) == 16 && offsetof(mzd_t, flags) == 24 &&
                  offsetof(mzd_t, high_bitmask) == 48 && offsetof(mzd_t, data) == 56,
              "mzd_t layout differs from m4ri/mzd.h:68-99");

namespace {

constexpr uint8_t kFlagExcess = 0x2, kFlagWindow = 0x4;
constexpr int kBuiltinCutoff = 8192;   // device Strassen leaf size with the tensor-core leaf (49 per launch); tuned on B200, see DESIGN.md
constexpr int kBuiltinCutoffM4rm = 4096;   // ... when M4RI_B200_LEAF pins the M4RM leaves (4096-row tall tiles)

struct Ctx {
  bool         ready = false;
  int          device = -1;            // -1: whatever is current at first use
  int          num_devices = 1;
  int          default_cutoff = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // PCIe transfers of the host path, overlapped with `stream`
  Workspace    ws;
  Stager       stager;                  // pinned ring for pageable host matrices
  std::vector<word> host_tmp;
  char         last_path[64] = "none";
};
Ctx g;
// The reference is "one host thread at a time" (SURVEY §8b), but an OpenMP libm4ri calls _mzd_mul_even /
// _mzd_addmul_even from four concurrent sections (m4ri/mp.c:87-108, 206-227) and those calls bind here under
// LD_PRELOAD: every entry point that touches the context (workspace stack, staging ring, streams) holds this
// lock for its whole duration.  Recursive: mzd_mul_mp -> mzd_mul, m4ri_b200_inv_m4ri -> mzd_init helpers.
std::recursive_mutex g_mu;
#define M4B_LOCKED std::lock_guard<std::recursive_mutex> lock_(g_mu)
unsigned long long g_products = 0;   // host-path products served (reported at exit with M4RI_B200_REPORT=1)

void report_at_exit() {
  fprintf(stderr, "m4ri_b200: served %llu products with %llu CUDA kernel launches (last path %s)\n", g_products,
          g_kernel_launches, g.last_path);
}

Ctx &ctx() {
  if (!g.ready) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      die("m4ri_b200: no usable CUDA device (%s); this library has no CPU fallback\n", cudaGetErrorString(e));
    if (g.device >= 0) M4B_CUDA(cudaSetDevice(g.device));
    M4B_CUDA(cudaGetDevice(&g.device));
    M4B_CUDA(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    M4B_CUDA(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
    if (getenv("M4RI_B200_REPORT")) atexit(report_at_exit);
    if (!g.default_cutoff) {
      char const *env = getenv("M4RI_B200_CUTOFF");
      char const *leaf = getenv("M4RI_B200_LEAF");
      bool const m4rm_only = leaf && (leaf[0] == '1' || leaf[0] == '2') && !leaf[1];
      g.default_cutoff = env && atoi(env) > 0 ? atoi(env) : (m4rm_only ? kBuiltinCutoffM4rm : kBuiltinCutoff);
    }
    g.ready = true;
  }
  return g;
}

inline int round_up(int v, int mult) { return (int)(((int64_t)v + mult - 1) / mult * mult); }

word left_mask(int nbits) { return ~(word)0 >> ((64 - nbits) % 64); }

// ---- stand-alone host matrices ------------------------------------------------------
void fill_header(mzd_t *M, rci_t r, rci_t c) {
  memset(M, 0, sizeof *M);
  M->nrows = r;
  M->ncols = c;
  M->width = c > 0 ? (c + 63) / 64 : 0;
  M->high_bitmask = left_mask(c % 64);
  if (c % 64) M->flags |= kFlagExcess;
}

typedef mzd_t *(*mzd_init_fn)(rci_t, rci_t);

// A NULL result matrix must be freeable by the caller's mzd_free (m4ri/strassen.c:356-357): when a
// libm4ri is loaded in the process use ITS mzd_init, otherwise our own (m4ri_b200_mzd_free).
mzd_t *alloc_result(rci_t r, rci_t c) {
  static mzd_init_fn ref_init = reinterpret_cast<mzd_init_fn>(dlsym(RTLD_DEFAULT, "mzd_init"));
  return ref_init ? ref_init(r, c) : m4ri_b200_mzd_init(r, c);
}

}  // namespace

// ---- transfers (shared with multi.cu) ------------------------------------------------------
// below this size the driver's own pageable path is as fast as the staging ring (tunable for experiments)
static size_t stage_threshold() {
  static size_t v = 0;
  if (!v) {
    char const *env = getenv("M4RI_B200_STAGE_MIN");
    v = env && atoll(env) > 0 ? (size_t)atoll(env) : (size_t)(4u << 20);
  }
  return v;
}
#define kStageThreshold stage_threshold()

void upload(DView dst, mzd_t const *src, cudaStream_t s, Stager *st) {
  if (src->nrows == 0 || src->ncols == 0) return;
  size_t const width = (size_t)src->width * 8, rows = (size_t)src->nrows;
  if (st && width * rows >= kStageThreshold && Stager::pageable(src->data))
    st->upload2d(dst.data, (size_t)dst.pitch * 8, src->data, (size_t)src->rowstride * 8, width, rows, s);
  else
    M4B_CUDA(cudaMemcpy2DAsync(dst.data, (size_t)dst.pitch * 8, src->data, (size_t)src->rowstride * 8, width, rows,
                               cudaMemcpyHostToDevice, s));
  if (src->ncols % 64) launch_mask_excess(DView{dst.data, dst.pitch, src->nrows, src->ncols}, s);
}

// Device rows -> host matrix, touching only bits (i < nrows, j < ncols) of the host matrix.
// Whole words go straight into the host rows; a partial last word is staged in `tmp` and merged under
// high_bitmask by download_finish() once the stream has been synchronised.
void download_async(mzd_t *dst, DView src, cudaStream_t s, std::vector<word> &tmp, Stager *st = nullptr) {
  tmp.clear();
  if (dst->nrows == 0 || dst->ncols == 0) return;
  int64_t const full = dst->ncols / 64;   // whole words per row
  if (full) {
    size_t const width = (size_t)full * 8, rows = (size_t)dst->nrows;
    if (st && width * rows >= kStageThreshold && Stager::pageable(dst->data))
      st->download2d(dst->data, (size_t)dst->rowstride * 8, src.data, (size_t)src.pitch * 8, width, rows, s);
    else
      M4B_CUDA(cudaMemcpy2DAsync(dst->data, (size_t)dst->rowstride * 8, src.data, (size_t)src.pitch * 8, width, rows,
                                 cudaMemcpyDeviceToHost, s));
  }
  if (dst->ncols % 64) {
    tmp.resize((size_t)dst->nrows);
    M4B_CUDA(cudaMemcpy2DAsync(tmp.data(), 8, src.data + full, (size_t)src.pitch * 8, 8, (size_t)dst->nrows,
                               cudaMemcpyDeviceToHost, s));
  }
}

void download_finish(mzd_t *dst, std::vector<word> const &tmp) {
  if (tmp.empty()) return;
  int64_t const full = dst->ncols / 64;
  word const mask = dst->high_bitmask;
  for (rci_t i = 0; i < dst->nrows; ++i) {
    word *w = dst->data + (int64_t)i * dst->rowstride + full;
    *w = (*w & ~mask) | (tmp[(size_t)i] & mask);
  }
}

void download(mzd_t *dst, DView src, cudaStream_t s, std::vector<word> &tmp, Stager *st) {
  download_async(dst, src, s, tmp, st);
  if (!tmp.empty()) {
    M4B_CUDA(cudaStreamSynchronize(s));
    download_finish(dst, tmp);
  }
}

void zero_async(DView v, cudaStream_t s) {
  M4B_CUDA(cudaMemsetAsync(v.data, 0, (size_t)v.nrows * (size_t)v.pitch * 8, s));
}

namespace {

int norm_cutoff(int cutoff, char const *who) {   // m4ri/strassen.c:349-354
  if (cutoff < 0) die("%s: cutoff must be >= 0.\n", who);
  if (cutoff == 0) cutoff = ctx().default_cutoff;
  cutoff = cutoff / 64 * 64;
  return cutoff < 64 ? 64 : cutoff;
}

// A clipped window of a host matrix (header only; shares the words).  c0 must be a multiple of 64.
mzd_t host_window(mzd_t const *M, int r0, int c0, int r1, int c1) {
  mzd_t W = *M;
  r1 = r1 < M->nrows ? r1 : M->nrows;
  c1 = c1 < M->ncols ? c1 : M->ncols;
  W.nrows = r1 > r0 ? r1 - r0 : 0;
  W.ncols = c1 > c0 ? c1 - c0 : 0;
  W.width = (W.ncols + 63) / 64;
  W.high_bitmask = left_mask(W.ncols % 64);
  W.flags = kFlagWindow | (W.ncols % 64 ? kFlagExcess : 0);
  W.data = M->data + (int64_t)r0 * M->rowstride + c0 / 64;
  return W;
}

// Overlap of PCIe transfers with the top level of the Strassen schedule (strassen.cu: TopHooks):
// operand quadrants are uploaded on the copy stream in the order the schedule first reads them and the
// compute stream waits on per-quadrant events; result quadrants are downloaded as soon as they are
// final (three of the four are final before the last of the seven products starts).
struct HostOverlap : TopHooks {
  Ctx &c;
  mzd_t *C;
  mzd_t const *A, *B;
  DView dA, dB, dC;
  bool upA[4] = {}, upB[4] = {}, upC[4] = {};
  std::vector<cudaEvent_t> events;
  std::vector<word> tails[4];
  mzd_t cwin[4];

  HostOverlap(Ctx &c_, mzd_t *C_, mzd_t const *A_, mzd_t const *B_, DView a, DView b, DView cc)
      : c(c_), C(C_), A(A_), B(B_), dA(a), dB(b), dC(cc) {}
  ~HostOverlap() override {
    for (cudaEvent_t e : events) cudaEventDestroy(e);
  }
  cudaEvent_t event() {
    cudaEvent_t e;
    M4B_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    events.push_back(e);
    return e;
  }
  static void quad(DView const &V, int q, int &r0, int &c0, int &r1, int &c1) {
    int const rh = V.nrows / 2, ch = V.ncols / 2;
    r0 = (q & 2) ? rh : 0;  r1 = r0 + rh;
    c0 = (q & 1) ? ch : 0;  c1 = c0 + ch;
  }
  void up(mzd_t const *M, DView const &V, int q, bool *done) {
    if (done[q]) return;
    done[q] = true;
    int r0, c0, r1, c1;
    quad(V, q, r0, c0, r1, c1);
    mzd_t W = host_window(M, r0, c0, r1, c1);
    if (W.nrows > 0 && W.ncols > 0) upload(V.sub(r0, c0, r0 + W.nrows, c0 + W.ncols), &W, c.copy_stream, &c.stager);
    cudaEvent_t e = event();
    M4B_CUDA(cudaEventRecord(e, c.copy_stream));
    M4B_CUDA(cudaStreamWaitEvent(c.stream, e, 0));
  }
  void need_a(int q) override { up(A, dA, q, upA); }
  void need_b(int q) override { up(B, dB, q, upB); }
  void need_c(int q) override { up(C, dC, q, upC); }
  // D2H calls are only ISSUED in finish(), after every kernel has been enqueued: a copy into pageable
  // host memory blocks the calling thread, and blocking here would stall the rest of the schedule.
  cudaEvent_t ready[4] = {};
  int order[4], ndone = 0;
  void done_c(int q) override {
    ready[q] = event();
    M4B_CUDA(cudaEventRecord(ready[q], c.stream));
    order[ndone++] = q;
  }
  void finish() {
    for (int i = 0; i < ndone; ++i) {
      int const q = order[i];
      int r0, c0, r1, c1;
      quad(dC, q, r0, c0, r1, c1);
      cwin[q] = host_window(C, r0, c0, r1, c1);
      M4B_CUDA(cudaStreamWaitEvent(c.copy_stream, ready[q], 0));
      if (cwin[q].nrows > 0 && cwin[q].ncols > 0)
        download_async(&cwin[q], dC.sub(r0, c0, r0 + cwin[q].nrows, c0 + cwin[q].ncols), c.copy_stream, tails[q], &c.stager);
    }
    M4B_CUDA(cudaStreamSynchronize(c.copy_stream));
    M4B_CUDA(cudaStreamSynchronize(c.stream));
    for (int i = 0; i < ndone; ++i) download_finish(&cwin[order[i]], tails[order[i]]);
  }
};

// Padded operand dimensions: every Strassen level must halve exactly (rows: 2^levels, columns: 128 * 2^levels).  When
// the leaves are big enough for the tensor-core leaf (tc_leaf.cu) and it may be used, the units are ITS tile units
// (128 rows, 1024 inner bits, 256 columns per leaf) — at most a quarter more work in the worst case for a leaf that is
// 1.7x faster, and nothing at all for the power-of-two sizes.  Zero padding never changes a result bit.
void padded_dims(int m, int l, int n, int levels, int *mp, int *lp, int *np) {
  int um = 1, ul = 128, un = 128;
  bool const tensor_ok = g_leaf_variant != 1 && g_leaf_variant != 2;
  if (tensor_ok && (m >> levels) >= 512 && (l >> levels) >= 4096 && (n >> levels) >= 1024) { um = 128; ul = 1024; un = 256; }
  *mp = round_up(m, um << levels);
  *lp = round_up(l > 0 ? l : 1, ul << levels);
  *np = round_up(n, un << levels);
}

// The one host->device->host product path behind every reference-named entry point.
void host_product(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff, bool clear, bool strassen) {
  M4B_LOCKED;
  Ctx &c = ctx();
  int const m = A->nrows, l = A->ncols, n = B->ncols;
  if (m == 0 || n == 0) return;
  ++g_products;
  int const levels = (strassen && l > 0) ? strassen_levels(m, l, n, cutoff) : 0;
  int mp, lp, np;
  padded_dims(m, l, n, levels, &mp, &lp, &np);
  snprintf(c.last_path, sizeof c.last_path, levels ? "strassen:%d" : "m4rm", levels);

  size_t const need = Workspace::bytes_for(mp, lp) + Workspace::bytes_for(lp, np) + Workspace::bytes_for(mp, np) +
                      strassen_workspace_bytes(mp, lp, np, levels);
  c.ws.reserve(need);
  cudaStream_t s = c.stream;
  DView dA = c.ws.alloc(mp, lp), dB = c.ws.alloc(lp, np), dC = c.ws.alloc(mp, np);
  zero_async(dA, s);
  zero_async(dB, s);
  if (!clear) zero_async(dC, s);
  if (levels > 0 && l > 0) {
    // transfers ride the copy stream, ordered after the zero fills
    cudaEvent_t zeroed;
    M4B_CUDA(cudaEventCreateWithFlags(&zeroed, cudaEventDisableTiming));
    M4B_CUDA(cudaEventRecord(zeroed, s));
    M4B_CUDA(cudaStreamWaitEvent(c.copy_stream, zeroed, 0));
    HostOverlap ov(c, C, A, B, dA, dB, dC);
    strassen_mul(dC, dA, dB, levels, clear, c.ws, s, &ov);
    ov.finish();
    M4B_CUDA(cudaEventDestroy(zeroed));
  } else {
    upload(dA, A, s, &c.stager);
    upload(dB, B, s, &c.stager);
    if (!clear) upload(dC, C, s, &c.stager);
    strassen_mul(dC, dA, dB, levels, clear, c.ws, s);
    download(C, dC, s, c.host_tmp, &c.stager);
    M4B_CUDA(cudaStreamSynchronize(s));
  }
  c.ws.release(0);
}

mzd_t *checked(char const *who, mzd_t *C, mzd_t const *A, mzd_t const *B) {
  if (A->ncols != B->nrows) die("%s: A ncols (%d) need to match B nrows (%d).\n", who, A->ncols, B->nrows);
  if (C == NULL) return alloc_result(A->nrows, B->ncols);
  if (C->nrows != A->nrows || C->ncols != B->ncols)
    die("%s: C (%d x %d) has wrong dimensions, expected (%d x %d)\n", who, C->nrows, C->ncols, A->nrows, B->ncols);
  return C;
}

// Host path of the triangular solves: T (t x t) and B (m x n) up, recursion on the device, X down into
// B (only its valid bits).  left: T X = B (t == m), right: X T = B (t == n).
void host_trsm(mzd_t const *T, mzd_t *B, int cutoff, bool upper, bool left) {
  M4B_LOCKED;
  Ctx &c = ctx();
  int const m = B->nrows, n = B->ncols, t = T->nrows;
  if (m == 0 || n == 0) return;
  ++g_products;
  snprintf(c.last_path, sizeof c.last_path, "trsm_%s_%s", upper ? "upper" : "lower", left ? "left" : "right");
  c.ws.reserve(Workspace::bytes_for(t, t) + Workspace::bytes_for(m, n) + trsm_workspace_bytes(t, m, n, cutoff));
  cudaStream_t s = c.stream;
  DView dT = c.ws.alloc(t, t), dB = c.ws.alloc(m, n);
  zero_async(dT, s);
  zero_async(dB, s);
  upload(dT, T, s, &c.stager);
  upload(dB, B, s, &c.stager);
  if (left) trsm_left(dT, dB, upper, cutoff, c.ws, s);
  else      trsm_right(dT, dB, upper, cutoff, c.ws, s);
  download(B, dB, s, c.host_tmp, &c.stager);
  M4B_CUDA(cudaStreamSynchronize(s));
  c.ws.release(0);
}

void check_trsm(char const *who, mzd_t const *T, mzd_t const *B, bool left) {   // m4ri/triangular.c:29-36, 300-307, 394-401, 457-464
  if (left && T->ncols != B->nrows) die("%s: triangular ncols (%d) need to match B nrows (%d).\n", who, T->ncols, B->nrows);
  if (!left && T->nrows != B->ncols) die("%s: triangular nrows (%d) need to match B ncols (%d).\n", who, T->nrows, B->ncols);
  if (T->nrows != T->ncols) die("%s: triangular matrix must be square and is found to be (%d) x (%d).\n", who, T->nrows, T->ncols);
}

DView as_view(m4ri_b200_dmat const *M) { return DView{M->data, M->pitch, M->nrows, M->ncols}; }

void device_product(m4ri_b200_dmat *C, m4ri_b200_dmat const *A, m4ri_b200_dmat const *B, int levels, bool clear,
                    cudaStream_t s) {
  M4B_LOCKED;
  Ctx &c = ctx();
  if (A->ncols != B->nrows || C->nrows != A->nrows || C->ncols != B->ncols)
    die("m4ri_b200_dmul: dimension mismatch (%dx%d) * (%dx%d) -> (%dx%d)\n", A->nrows, A->ncols, B->nrows, B->ncols,
        C->nrows, C->ncols);
  int const m = A->nrows, l = A->ncols, n = B->ncols;
  if (m == 0 || n == 0) return;
  if (l == 0) {
    if (clear) launch_zero(as_view(C), s);
    return;
  }
  snprintf(c.last_path, sizeof c.last_path, levels ? "strassen:%d" : "m4rm", levels);
  int mp, lp, np;
  padded_dims(m, l, n, levels, &mp, &lp, &np);
  bool const pad = levels > 0 && (mp != m || lp != l || np != n);
  if (!pad) {
    c.ws.reserve(strassen_workspace_bytes(m, l, n, levels));
    strassen_mul(as_view(C), as_view(A), as_view(B), levels, clear, c.ws, s);
    return;
  }
  c.ws.reserve(Workspace::bytes_for(mp, lp) + Workspace::bytes_for(lp, np) + Workspace::bytes_for(mp, np) +
               strassen_workspace_bytes(mp, lp, np, levels));
  DView dA = c.ws.alloc(mp, lp), dB = c.ws.alloc(lp, np), dC = c.ws.alloc(mp, np);
  zero_async(dA, s);
  zero_async(dB, s);
  launch_copy(dA.sub(0, 0, m, l), as_view(A), s);
  launch_copy(dB.sub(0, 0, l, n), as_view(B), s);
  if (!clear) {
    zero_async(dC, s);
    launch_copy(dC.sub(0, 0, m, n), as_view(C), s);
  }
  strassen_mul(dC, dA, dB, levels, clear, c.ws, s);
  launch_copy(as_view(C), dC.sub(0, 0, m, n), s);
  c.ws.release(0);   // stream-ordered reuse: later work on `s` runs after these kernels
}

}  // namespace
}  // namespace m4b

using namespace m4b;

extern "C" {

// ---- Part 1: reference-named entry points ----------------------------------------------

mzd_t *mzd_mul(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff) {
  if (A->ncols != B->nrows) die("mzd_mul: A ncols (%d) need to match B nrows (%d).\n", A->ncols, B->nrows);
  cutoff = norm_cutoff(cutoff, "mzd_mul");
  C = checked("mzd_mul", C, A, B);
  host_product(C, A, B, cutoff, true, true);
  return C;
}

mzd_t *mzd_addmul(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff) {
  if (A->ncols != B->nrows) die("mzd_addmul: A ncols (%d) need to match B nrows (%d).\n", A->ncols, B->nrows);
  cutoff = norm_cutoff(cutoff, "mzd_addmul");
  C = checked("mzd_addmul", C, A, B);
  if (A->nrows == 0 || A->ncols == 0 || B->ncols == 0) return C;   // strassen.c:692-695
  host_product(C, A, B, cutoff, false, true);
  return C;
}

mzd_t *_mzd_addmul(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff) {
  host_product(C, A, B, cutoff < 64 ? 64 : cutoff, false, true);
  return C;
}

mzd_t *_mzd_mul_even(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff) {
  host_product(C, A, B, cutoff < 64 ? 64 : cutoff, true, true);
  return C;
}

mzd_t *_mzd_addmul_even(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff) {
  host_product(C, A, B, cutoff < 64 ? 64 : cutoff, false, true);
  return C;
}

mzd_t *_mzd_mul_m4rm(mzd_t *C, mzd_t const *A, mzd_t const *B, int k, int clear) {
  (void)k;   // any table width gives the same bits; the kernel's k is fixed at 8
  host_product(C, A, B, 0, clear != 0, false);
  return C;
}

mzd_t *mzd_mul_m4rm(mzd_t *C, mzd_t const *A, mzd_t const *B, int k) {
  C = checked("mzd_mul_m4rm", C, A, B);
  return _mzd_mul_m4rm(C, A, B, k, 1);
}

mzd_t *mzd_addmul_m4rm(mzd_t *C, mzd_t const *A, mzd_t const *B, int k) {
  // the reference dereferences C before its NULL test (brilliantrussian.c:1018 vs 1022); we test first
  if (C != NULL && (C->ncols == 0 || C->nrows == 0)) return C;
  C = checked("mzd_addmul_m4rm", C, A, B);
  return _mzd_mul_m4rm(C, A, B, k, 0);
}

// ---- widened rows: triangular solves, left variants (m4ri/triangular.h:115,127,142,153) ----
#define M4B_TRSM_ENTRY(NAME, UPPER, LEFT)                                                           \
  void mzd_##NAME(mzd_t const *T, mzd_t *B, int const cutoff) {                                     \
    check_trsm("mzd_" #NAME, T, B, LEFT);                                                           \
    host_trsm(T, B, norm_cutoff(cutoff < 0 ? 0 : cutoff, "mzd_" #NAME), UPPER, LEFT);               \
  }                                                                                                 \
  void _mzd_##NAME(mzd_t const *T, mzd_t *B, int const cutoff) {                                    \
    host_trsm(T, B, norm_cutoff(cutoff < 0 ? 0 : cutoff, "_mzd_" #NAME), UPPER, LEFT);              \
  }
M4B_TRSM_ENTRY(trsm_lower_left, false, true)
M4B_TRSM_ENTRY(trsm_upper_left, true, true)
M4B_TRSM_ENTRY(trsm_lower_right, false, false)
M4B_TRSM_ENTRY(trsm_upper_right, true, false)
#undef M4B_TRSM_ENTRY

// C blocks over the GPUs chosen with m4ri_b200_set_num_devices (multi.cu), starting at the selected device; with one
// usable GPU: same as mzd_mul / mzd_addmul.
static int mp_devices() {
  Ctx &c = ctx();
  int avail = 0;
  M4B_CUDA(cudaGetDeviceCount(&avail));
  int const G = c.num_devices < avail - c.device ? c.num_devices : avail - c.device;
  return G < 1 ? 1 : G;
}

mzd_t *mzd_mul_mp(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff) {
  M4B_LOCKED;
  int const G = mp_devices();
  if (G <= 1) return mzd_mul(C, A, B, cutoff);
  if (A->ncols != B->nrows) die("mzd_mul_mp: A ncols (%d) need to match B nrows (%d).\n", A->ncols, B->nrows);
  cutoff = norm_cutoff(cutoff, "mzd_mul_mp");
  C = checked("mzd_mul_mp", C, A, B);
  ++g_products;
  multi_product(C, A, B, cutoff, true, G, ctx().device, ctx().last_path, sizeof ctx().last_path);
  return C;
}

mzd_t *mzd_addmul_mp(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff) {
  M4B_LOCKED;
  int const G = mp_devices();
  if (G <= 1) return mzd_addmul(C, A, B, cutoff);
  if (A->ncols != B->nrows) die("mzd_addmul_mp: A ncols (%d) need to match B nrows (%d).\n", A->ncols, B->nrows);
  cutoff = norm_cutoff(cutoff, "mzd_addmul_mp");
  C = checked("mzd_addmul_mp", C, A, B);
  if (A->nrows == 0 || A->ncols == 0 || B->ncols == 0) return C;
  ++g_products;
  multi_product(C, A, B, cutoff, false, G, ctx().device, ctx().last_path, sizeof ctx().last_path);
  return C;
}

// The reference's unchecked 2 x 2 block forms (m4ri/mp.h:74,86; mp.c:39-156, 158-275).  An OpenMP libm4ri would run
// four concurrent _mzd_(add)mul_even sections here; interposed, the whole product goes to the GPU(s) at once.
mzd_t *_mzd_mul_mp4(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff) {
  M4B_LOCKED;
  int const G = mp_devices();
  if (G <= 1) return _mzd_mul_even(C, A, B, cutoff);
  ++g_products;
  multi_product(C, A, B, cutoff < 64 ? 64 : cutoff, true, G, ctx().device, ctx().last_path, sizeof ctx().last_path);
  return C;
}

mzd_t *_mzd_addmul_mp4(mzd_t *C, mzd_t const *A, mzd_t const *B, int cutoff) {
  M4B_LOCKED;
  int const G = mp_devices();
  if (G <= 1) return _mzd_addmul_even(C, A, B, cutoff);
  ++g_products;
  multi_product(C, A, B, cutoff < 64 ? 64 : cutoff, false, G, ctx().device, ctx().last_path, sizeof ctx().last_path);
  return C;
}

// ---- widened rows: PLE with the matrix resident on the device (ple.cu) -----------------------------------
rci_t _mzd_ple(mzd_t *A, mzp_t *P, mzp_t *Q, int const cutoff) {
  M4B_LOCKED;
  rci_t const m = A->nrows, n = A->ncols;
  if (m == 0 || n == 0) {
    for (rci_t i = 0; i < m; ++i) P->values[i] = i;
    for (rci_t i = 0; i < n; ++i) Q->values[i] = i;
    return 0;
  }
  Ctx &c = ctx();
  ++g_products;
  snprintf(c.last_path, sizeof c.last_path, "ple");
  int const co = norm_cutoff(cutoff < 0 ? 0 : cutoff, "_mzd_ple");
  c.ws.reserve(Workspace::bytes_for(m, n) + ple_workspace_bytes(m, n, co));
  cudaStream_t s = c.stream;
  DView dA = c.ws.alloc(m, n);
  zero_async(dA, s);
  upload(dA, A, s, &c.stager);
  rci_t const rank = ple_device(dA, P->values, Q->values, co, c.ws, s);
  download(A, dA, s, c.host_tmp, &c.stager);
  M4B_CUDA(cudaStreamSynchronize(s));
  c.ws.release(0);
  return rank;
}

rci_t mzd_ple(mzd_t *A, mzp_t *P, mzp_t *Q, int const cutoff) {   // m4ri/ple.c:33-39
  if (P->length != A->nrows) die("mzd_ple: Permutation P length (%d) must match A nrows (%d)\n", P->length, A->nrows);
  if (Q->length != A->ncols) die("mzd_ple: Permutation Q length (%d) must match A ncols (%d)\n", Q->length, A->ncols);
  return _mzd_ple(A, P, Q, cutoff);
}

// ---- widened rows: elimination entry points of libm4ri served by the device RREF (echelon.cu) -------------
// m4ri/echelonform.c:30-36, m4ri/brilliantrussian.c:603-967, 971-997.  full != 0 asks for THE reduced row echelon form
// (unique: bit-identical whatever the algorithm).  full == 0 asks for AN upper-triangular echelon form whose bits
// depend on the reference's k and block sizes: when a libm4ri follows in the link order (preload / link-first
// deployment) that call is handed on to it unchanged; stand-alone, the reduced form — itself a row echelon form — is
// returned.
typedef rci_t (*echelonize_m4ri_fn)(mzd_t *, int, int);
typedef rci_t (*echelonize_fn)(mzd_t *, int);

rci_t mzd_echelonize_m4ri(mzd_t *A, int full, int k) {
  if (!full) {
    static echelonize_m4ri_fn next = reinterpret_cast<echelonize_m4ri_fn>(dlsym(RTLD_NEXT, "mzd_echelonize_m4ri"));
    if (next) return next(A, full, k);
  }
  return m4ri_b200_echelonize(A, 1);
}

rci_t mzd_echelonize(mzd_t *A, int full) {
  if (!full) {
    static echelonize_fn next = reinterpret_cast<echelonize_fn>(dlsym(RTLD_NEXT, "mzd_echelonize"));
    if (next) return next(A, full);
  }
  return m4ri_b200_echelonize(A, 1);
}

mzd_t *mzd_inv_m4ri(mzd_t *B, mzd_t const *A, int k) {
  (void)k;
  return m4ri_b200_inv_m4ri(B, A);
}

// ---- Part 2: extension API ------------------------------------------------------------------

int m4ri_b200_version(void) { return 100; }

int m4ri_b200_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
  return count;
}

void m4ri_b200_set_device(int device) {
  M4B_LOCKED;
  if (g.ready && g.device != device) {
    // everything that belongs to the old device goes, on the old device: workspace, streams, the staging ring
    // (its events were created there) and the leaf-profile event pool
    cudaSetDevice(g.device);
    cudaStreamSynchronize(g.stream);
    cudaStreamSynchronize(g.copy_stream);
    g.stager.release();
    leaf_profile_reset();
    g.ws.destroy();
    cudaStreamDestroy(g.stream);
    cudaStreamDestroy(g.copy_stream);
    g.ready = false;
  }
  g.device = device;
}

void m4ri_b200_set_num_devices(int n) { g.num_devices = n < 1 ? 1 : n; }
void m4ri_b200_set_default_cutoff(int cutoff) { g.default_cutoff = cutoff > 0 ? cutoff : kBuiltinCutoff; }
int  m4ri_b200_get_default_cutoff(void) { return g.default_cutoff ? g.default_cutoff : kBuiltinCutoff; }

void m4ri_b200_release(void) {
  M4B_LOCKED;
  multi_release();
  tc_scratch_release();
  if (!g.ready) return;
  cudaStreamSynchronize(g.stream);
  g.ws.destroy();
  g.stager.release();
}

int m4ri_b200_set_leaf_variant(int variant) {
  int const prev = g_leaf_variant;
  g_leaf_variant = variant >= 0 && variant <= 3 ? variant : -1;    // -1: back to the environment / built-in default
  return prev;
}

int m4ri_b200_last_leaf_variant(void) { return g_last_leaf; }

char const *m4ri_b200_last_path(void) { return g.last_path; }
uint64_t    m4ri_b200_kernel_launches(void) { return g_kernel_launches; }

void m4ri_b200_profile_begin(void) { leaf_profile_begin(); }
uint64_t m4ri_b200_profile_end(double *leaf_ms, double *leaf_bitops) { return leaf_profile_end(leaf_ms, leaf_bitops); }

mzd_t *m4ri_b200_mzd_init(rci_t r, rci_t c) {
  mzd_t *M = static_cast<mzd_t *>(malloc(sizeof(mzd_t)));
  fill_header(M, r, c);
  M->rowstride = M->width + (M->width & 1);
  if (r && c) {
    size_t const bytes = (size_t)r * (size_t)M->rowstride * sizeof(word);
    if (posix_memalign(reinterpret_cast<void **>(&M->data), 64, bytes)) die("m4ri_b200_mzd_init: out of memory\n");
    memset(M->data, 0, bytes);
  }
  return M;
}

mzd_t *m4ri_b200_mzd_init_window(mzd_t *P, rci_t lowr, rci_t lowc, rci_t highr, rci_t highc) {
  if (lowc % 64) die("m4ri_b200_mzd_init_window: lowc must be a multiple of 64\n");
  mzd_t *W = static_cast<mzd_t *>(malloc(sizeof(mzd_t)));
  rci_t nr = highr - lowr;
  if (P->nrows - lowr < nr) nr = P->nrows - lowr;
  fill_header(W, nr, highc - lowc);
  W->flags |= kFlagWindow;
  W->rowstride = P->rowstride;
  W->data = P->data + (int64_t)lowr * P->rowstride + lowc / 64;
  return W;
}

// Frees a matrix that an entry point of this library allocated for a NULL result argument: it came from the
// process' libm4ri mzd_init when one is loaded (alloc_result), so it goes back through that library's mzd_free.
void m4ri_b200_result_free(mzd_t *M) {
  typedef void (*mzd_free_fn)(mzd_t *);
  static mzd_free_fn ref_free = dlsym(RTLD_DEFAULT, "mzd_init") ? reinterpret_cast<mzd_free_fn>(dlsym(RTLD_DEFAULT, "mzd_free")) : nullptr;
  if (ref_free) ref_free(M);
  else m4ri_b200_mzd_free(M);
}

void m4ri_b200_mzd_free(mzd_t *M) {
  if (!M) return;
  if (!(M->flags & kFlagWindow)) free(M->data);
  free(M);
}

m4ri_b200_dmat *m4ri_b200_dmat_alloc(rci_t nrows, rci_t ncols) {
  M4B_LOCKED;
  ctx();
  m4ri_b200_dmat *M = static_cast<m4ri_b200_dmat *>(calloc(1, sizeof *M));
  M->nrows = nrows;
  M->ncols = ncols;
  M->pitch = Workspace::pitch_for(ncols);
  M->owner = 1;
  size_t const bytes = (size_t)nrows * (size_t)M->pitch * 8;
  if (bytes) {
    M4B_CUDA(cudaMalloc(&M->data, bytes));
    // on the library stream, and complete before returning: a memset on the legacy default stream is not ordered
    // against the (non-blocking) library stream, so it could still be running when the first upload lands
    M4B_CUDA(cudaMemsetAsync(M->data, 0, bytes, g.stream));
    M4B_CUDA(cudaStreamSynchronize(g.stream));
  }
  return M;
}

m4ri_b200_dmat *m4ri_b200_dmat_wrap(void *device_ptr, int64_t pitch_words, rci_t nrows, rci_t ncols) {
  if (((uintptr_t)device_ptr & 15) || (pitch_words & 1) || pitch_words < Workspace::pitch_for(ncols))
    die("m4ri_b200_dmat_wrap: pointer must be 16-byte aligned and pitch even and >= %lld words\n",
        (long long)Workspace::pitch_for(ncols));
  m4ri_b200_dmat *M = static_cast<m4ri_b200_dmat *>(calloc(1, sizeof *M));
  M->data = static_cast<word *>(device_ptr);
  M->pitch = pitch_words;
  M->nrows = nrows;
  M->ncols = ncols;
  return M;
}

void m4ri_b200_dmat_free(m4ri_b200_dmat *M) {
  if (!M) return;
  if (M->owner && M->data) cudaFree(M->data);
  free(M);
}

void m4ri_b200_upload(m4ri_b200_dmat *dst, mzd_t const *src, void *stream) {
  M4B_LOCKED;
  if (dst->nrows != src->nrows || dst->ncols != src->ncols) die("m4ri_b200_upload: dimension mismatch\n");
  upload(as_view(dst), src, stream ? static_cast<cudaStream_t>(stream) : ctx().stream, &ctx().stager);
  if (!stream) M4B_CUDA(cudaStreamSynchronize(ctx().stream));
}

void m4ri_b200_download(mzd_t *dst, m4ri_b200_dmat const *src, void *stream) {
  M4B_LOCKED;
  if (dst->nrows != src->nrows || dst->ncols != src->ncols) die("m4ri_b200_download: dimension mismatch\n");
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx().stream;
  download(dst, as_view(src), s, ctx().host_tmp, &ctx().stager);
  M4B_CUDA(cudaStreamSynchronize(s));
}

void m4ri_b200_sync(void *stream) {
  M4B_CUDA(cudaStreamSynchronize(stream ? static_cast<cudaStream_t>(stream) : ctx().stream));
}

void m4ri_b200_dmul_m4rm(m4ri_b200_dmat *C, m4ri_b200_dmat const *A, m4ri_b200_dmat const *B, int clear, void *stream) {
  device_product(C, A, B, 0, clear != 0, stream ? static_cast<cudaStream_t>(stream) : ctx().stream);
}

void m4ri_b200_dmul(m4ri_b200_dmat *C, m4ri_b200_dmat const *A, m4ri_b200_dmat const *B, int cutoff, int clear,
                    void *stream) {
  cutoff = norm_cutoff(cutoff, "m4ri_b200_dmul");
  int const levels = A->ncols > 0 ? strassen_levels(A->nrows, A->ncols, B->ncols, cutoff) : 0;
  device_product(C, A, B, levels, clear != 0, stream ? static_cast<cudaStream_t>(stream) : ctx().stream);
}

// Device product on explicit top-level quadrants with caller-supplied transfer hooks: the multi-rank end-to-end
// path (bench.py --gpus N) keeps each quadrant of its operands in its own contiguous buffer, uploads / all-gathers
// a quadrant when the schedule first needs it (need_*) and downloads a result quadrant as soon as it is final
// (done_c).  The hooks are called on the calling thread while the schedule is being enqueued on `stream`.
namespace {
struct CallbackHooks : TopHooks {
  m4ri_b200_hooks const *h;
  explicit CallbackHooks(m4ri_b200_hooks const *h_) : h(h_) {}
  void need_a(int q) override { if (h && h->need_a) h->need_a(h->user, q); }
  void need_b(int q) override { if (h && h->need_b) h->need_b(h->user, q); }
  void need_c(int q) override { if (h && h->need_c) h->need_c(h->user, q); }
  void done_c(int q) override { if (h && h->done_c) h->done_c(h->user, q); }
};
}  // namespace

void m4ri_b200_dmul_quads(m4ri_b200_dmat *const C[4], m4ri_b200_dmat const *const A[4], m4ri_b200_dmat const *const B[4],
                          int cutoff, int clear, void *stream, m4ri_b200_hooks const *hooks) {
  M4B_LOCKED;
  Ctx &c = ctx();
  int const m2 = A[0]->nrows, k2 = A[0]->ncols, n2 = B[0]->ncols;
  for (int q = 0; q < 4; ++q)
    if (A[q]->nrows != m2 || A[q]->ncols != k2 || B[q]->nrows != k2 || B[q]->ncols != n2 || C[q]->nrows != m2 || C[q]->ncols != n2)
      die("m4ri_b200_dmul_quads: the four quadrants of an operand must have identical, matching shapes\n");
  cutoff = norm_cutoff(cutoff, "m4ri_b200_dmul_quads");
  int levels = strassen_levels(2 * m2, 2 * k2, 2 * n2, cutoff);
  if (levels < 1) levels = 1;
  if (m2 % (1 << (levels - 1)) || k2 % (128 << (levels - 1)) || n2 % (128 << (levels - 1)))
    die("m4ri_b200_dmul_quads: quadrant dimensions must be multiples of 2^(levels-1) rows and 128 * 2^(levels-1) columns\n");
  snprintf(c.last_path, sizeof c.last_path, "strassen:%d", levels);
  c.ws.reserve(strassen_workspace_bytes(2 * m2, 2 * k2, 2 * n2, levels) + 3 * Workspace::bytes_for(m2, k2 > n2 ? k2 : n2));
  DView a[4], b[4], cc[4];
  for (int q = 0; q < 4; ++q) { a[q] = as_view(A[q]); b[q] = as_view(B[q]); cc[q] = as_view(C[q]); }
  CallbackHooks hk(hooks);
  strassen_mul_quads(cc, a, b, levels, clear != 0, c.ws, stream ? static_cast<cudaStream_t>(stream) : c.stream, hk);
}

void m4ri_b200_dmul_levels(m4ri_b200_dmat *C, m4ri_b200_dmat const *A, m4ri_b200_dmat const *B, int levels, int clear,
                           void *stream) {
  if (levels < 0 || levels > 12) die("m4ri_b200_dmul_levels: levels must be in 0..12\n");
  while (levels > 0 && ((A->nrows >> levels) < 1 || (A->ncols >> levels) < 128 || (B->ncols >> levels) < 128)) --levels;
  device_product(C, A, B, levels, clear != 0, stream ? static_cast<cudaStream_t>(stream) : ctx().stream);
}

void m4ri_b200_dtrsm(m4ri_b200_dmat const *T, m4ri_b200_dmat *B, int upper, int left, int cutoff, void *stream) {
  M4B_LOCKED;
  if (T->nrows != T->ncols || (left ? T->ncols != B->nrows : T->nrows != B->ncols))
    die("m4ri_b200_dtrsm: dimension mismatch\n");
  Ctx &c = ctx();
  cutoff = norm_cutoff(cutoff, "m4ri_b200_dtrsm");
  c.ws.reserve(trsm_workspace_bytes(T->nrows, B->nrows, B->ncols, cutoff));
  snprintf(c.last_path, sizeof c.last_path, "trsm_%s_%s", upper ? "upper" : "lower", left ? "left" : "right");
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c.stream;
  if (left) trsm_left(as_view(T), as_view(B), upper != 0, cutoff, c.ws, s);
  else      trsm_right(as_view(T), as_view(B), upper != 0, cutoff, c.ws, s);
}

void m4ri_b200_dtranspose(m4ri_b200_dmat *DST, m4ri_b200_dmat const *A, void *stream) {
  M4B_LOCKED;
  if (DST->nrows != A->ncols || DST->ncols != A->nrows) die("m4ri_b200_dtranspose: Wrong size for return matrix.\n");
  if (DST->data == A->data) die("m4ri_b200_dtranspose: DST must not alias A\n");
  launch_transpose(as_view(DST), as_view(A), stream ? static_cast<cudaStream_t>(stream) : ctx().stream);
}

// Host form with the reference's semantics (mzd_transpose, m4ri/mzd.c:1118-1139): DST may be NULL
// (allocated), wrong dimensions die, only DST's valid bits are written.
mzd_t *m4ri_b200_transpose(mzd_t *DST, mzd_t const *A) {
  M4B_LOCKED;
  if (DST == NULL) DST = alloc_result(A->ncols, A->nrows);
  else if (DST->nrows != A->ncols || DST->ncols != A->nrows) die("mzd_transpose: Wrong size for return matrix.\n");
  if (A->nrows == 0 || A->ncols == 0) return DST;
  Ctx &c = ctx();
  ++g_products;
  snprintf(c.last_path, sizeof c.last_path, "transpose");
  c.ws.reserve(Workspace::bytes_for(A->nrows, A->ncols) + Workspace::bytes_for(A->ncols, A->nrows));
  cudaStream_t s = c.stream;
  DView dA = c.ws.alloc(A->nrows, A->ncols), dD = c.ws.alloc(A->ncols, A->nrows);
  zero_async(dA, s);
  upload(dA, A, s, &c.stager);
  launch_transpose(dD, dA, s);
  download(DST, dD, s, c.host_tmp, &c.stager);
  M4B_CUDA(cudaStreamSynchronize(s));
  c.ws.release(0);
  return DST;
}

// Reduced row echelon form in place (the result of the reference's mzd_echelonize_m4ri(A, 1, k),
// m4ri/brilliantrussian.c:603-967 — unique, so bit-identical); returns the rank.  full == 0 asks the reference for
// *an* upper-triangular echelon form, which depends on its k; this library always returns the reduced one.
int m4ri_b200_dechelonize(m4ri_b200_dmat *A, int full, void *stream) {
  M4B_LOCKED;
  (void)full;
  Ctx &c = ctx();
  snprintf(c.last_path, sizeof c.last_path, "echelon");
  c.ws.reserve(echelon_workspace_bytes(A->nrows, A->ncols, A->pitch));
  return echelonize_device(as_view(A), c.ws, stream ? static_cast<cudaStream_t>(stream) : c.stream);
}

rci_t m4ri_b200_echelonize(mzd_t *A, int full) {
  M4B_LOCKED;
  (void)full;
  if (A->nrows == 0 || A->ncols == 0) return 0;
  Ctx &c = ctx();
  ++g_products;
  snprintf(c.last_path, sizeof c.last_path, "echelon");
  c.ws.reserve(Workspace::bytes_for(A->nrows, A->ncols) + echelon_workspace_bytes(A->nrows, A->ncols));
  cudaStream_t s = c.stream;
  DView dA = c.ws.alloc(A->nrows, A->ncols);
  zero_async(dA, s);
  upload(dA, A, s, &c.stager);
  int const rank = echelonize_device(dA, c.ws, s);
  download(A, dA, s, c.host_tmp, &c.stager);
  M4B_CUDA(cudaStreamSynchronize(s));
  c.ws.release(0);
  return rank;
}

// B = A^-1 the way the reference computes it (mzd_inv_m4ri, m4ri/brilliantrussian.c:971-997): the right block of the
// reduced row echelon form of [A | 0 | I] (the identity starts at the next 128-column boundary).  Like the reference
// it does not test invertibility: for a singular A the result is still that (unique) block.  B may be NULL.
mzd_t *m4ri_b200_inv_m4ri(mzd_t *B, mzd_t const *A) {
  M4B_LOCKED;
  if (A->nrows != A->ncols) die("mzd_inv_m4ri: the matrix must be square.\n");
  rci_t const n = A->nrows;
  if (B == NULL) B = alloc_result(n, n);
  else if (B->nrows != n || B->ncols != n) die("mzd_inv_m4ri: B has wrong dimensions.\n");
  if (n == 0) return B;
  int const np = round_up(n, 128);
  Ctx &c = ctx();
  ++g_products;
  snprintf(c.last_path, sizeof c.last_path, "inverse");
  c.ws.reserve(Workspace::bytes_for(n, np + n) + echelon_workspace_bytes(n, np + n));
  cudaStream_t s = c.stream;
  DView dC = c.ws.alloc(n, np + n);
  zero_async(dC, s);
  mzd_t *I = m4ri_b200_mzd_init(n, n);
  for (rci_t i = 0; i < n; ++i) I->data[(int64_t)i * I->rowstride + i / 64] |= (word)1 << (i % 64);
  upload(dC.sub(0, 0, n, n), A, s, &c.stager);
  upload(dC.sub(0, np, n, np + n), I, s, &c.stager);
  echelonize_device(dC, c.ws, s);
  download(B, dC.sub(0, np, n, np + n), s, c.host_tmp, &c.stager);
  M4B_CUDA(cudaStreamSynchronize(s));
  m4ri_b200_mzd_free(I);
  c.ws.release(0);
  return B;
}

rci_t m4ri_b200_dple(m4ri_b200_dmat *A, rci_t *P, rci_t *Q, void *stream) {
  M4B_LOCKED;
  Ctx &c = ctx();
  snprintf(c.last_path, sizeof c.last_path, "ple");
  int const co = norm_cutoff(0, "m4ri_b200_dple");
  c.ws.reserve(ple_workspace_bytes(A->nrows, A->ncols, co));
  return ple_device(as_view(A), P, Q, co, c.ws, stream ? static_cast<cudaStream_t>(stream) : c.stream);
}

// direct entry points of the tensor-core leaf (tests, measurements): the cross-check kernel (Bt = B transposed) ...
void m4ri_b200_dmul_tc(m4ri_b200_dmat *C, m4ri_b200_dmat const *A, m4ri_b200_dmat const *Bt, int clear, void *stream) {
  M4B_LOCKED;
  ctx();
  launch_tc_leaf_simple(as_view(C), as_view(A), as_view(Bt), clear == 0, stream ? static_cast<cudaStream_t>(stream) : ctx().stream);
}

void m4ri_b200_dmul_tc2(m4ri_b200_dmat *C, m4ri_b200_dmat const *A, m4ri_b200_dmat const *B, void *stream) {
  M4B_LOCKED;
  ctx();
  launch_tc_leaf2(as_view(C), as_view(A), as_view(B), stream ? static_cast<cudaStream_t>(stream) : ctx().stream);
}

void m4ri_b200_dadd(m4ri_b200_dmat *C, m4ri_b200_dmat const *A, m4ri_b200_dmat const *B, void *stream) {
  M4B_LOCKED;
  if (A->nrows != B->nrows || A->ncols != B->ncols || C->nrows != A->nrows || C->ncols != A->ncols)
    die("m4ri_b200_dadd: dimension mismatch\n");
  launch_xor(as_view(C), as_view(A), as_view(B), stream ? static_cast<cudaStream_t>(stream) : ctx().stream);
}

}  // extern "C"
