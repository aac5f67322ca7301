// staging.h — fast H2D/D2H for PAGEABLE host matrices (what mzd_init gives every libm4ri user).
//
// cudaMemcpy2DAsync from pageable memory is staged by the driver on one thread (measured on the B200
// box: 10-11 GB/s up, 21 GB/s down) while the PCIe5 link moves 55 GB/s from pinned memory, and
// cudaHostRegister costs as much as the slow copy (51 ms per 512 MiB).  The Stager keeps a small
// ring of pinned chunks and a few copy threads: rows are gathered into a pinned chunk in parallel
// (the 2D -> dense repack comes for free) while the previous chunk is in flight on the copy engine.
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "dev.h"

namespace m4b {

class Stager {
 public:
  Stager() = default;
  ~Stager();
  Stager(Stager const &) = delete;
  Stager &operator=(Stager const &) = delete;

  // true if `p` is ordinary (unregistered) host memory, i.e. worth staging
  static bool pageable(void const *p);

  // dst (device, pitched) <- src (host, pitched); asynchronous w.r.t. the device on `s`, returns when the
  // last chunk has been handed to the copy engine.
  void upload2d(void *dst, size_t dpitch, void const *src, size_t spitch, size_t width, size_t rows, cudaStream_t s);
  // dst (host, pitched) <- src (device, pitched); returns when all bytes are in dst.
  void download2d(void *dst, size_t dpitch, void const *src, size_t spitch, size_t width, size_t rows, cudaStream_t s);

  void release();   // free pinned memory, stop the threads
  void set_threads(int n);   // copy threads incl. the caller (takes effect at the next start of the ring)

 private:
  static constexpr size_t kChunkBytes = 8u << 20;
  static constexpr int    kSlots = 4;
  static constexpr int    kMaxThreads = 16;
  int kThreads = 4;                 // copy threads incl. the caller ($M4RI_B200_STAGE_THREADS, read once in ensure())
  int forced_threads_ = 0;

  void ensure();
  void parallel_rows(size_t rows, std::function<void(size_t, size_t)> const &fn);
  void worker(int id);

  char       *slot_[kSlots] = {};
  cudaEvent_t done_[kSlots] = {};
  bool        ready_ = false;

  std::vector<std::thread> threads_;
  std::mutex               mu_;
  std::condition_variable  cv_work_, cv_done_;
  std::function<void(size_t, size_t)> const *job_ = nullptr;
  size_t job_rows_ = 0;
  int    generation_ = 0, pending_ = 0;
  bool   stop_ = false;
};

}  // namespace m4b
