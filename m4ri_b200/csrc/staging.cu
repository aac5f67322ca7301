// staging.cu — see staging.h.
#include "staging.h"

#include <string.h>

namespace m4b {

bool Stager::pageable(void const *p) {
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) {
    cudaGetLastError();   // older drivers report unregistered memory as an error
    return true;
  }
  return attr.type == cudaMemoryTypeUnregistered;
}

void Stager::ensure() {
  if (ready_) return;
  // measured on the 16-thread B200 box, mzd_mul 65536^3 from pageable matrices: 4 threads 125.2 ms, 8: 114.4 ms,
  // 12: 111.2 ms (pinned: 107.3 ms) — default = the host's threads minus four, between 4 and 12
  int const hw = (int)std::thread::hardware_concurrency();
  kThreads = hw - 4 < 4 ? 4 : (hw - 4 > 12 ? 12 : hw - 4);
  if (forced_threads_ > 0) kThreads = forced_threads_ > kMaxThreads ? kMaxThreads : forced_threads_;
  if (char const *env = getenv("M4RI_B200_STAGE_THREADS")) {
    int const t = atoi(env);
    if (t >= 1 && t <= kMaxThreads) kThreads = t;
  }
  for (int i = 0; i < kSlots; ++i) {
    M4B_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&slot_[i]), kChunkBytes, cudaHostAllocPortable));
    M4B_CUDA(cudaEventCreateWithFlags(&done_[i], cudaEventDisableTiming));
  }
  stop_ = false;
  // no worker is alive here: restart the job counter they compare against (a worker of a re-created ring starts at
  // generation 0 and would otherwise take the stale job pointer of the ring that was released)
  generation_ = 0;
  pending_ = 0;
  job_ = nullptr;
  for (int t = 1; t < kThreads; ++t) threads_.emplace_back([this, t] { worker(t); });
  ready_ = true;
}

void Stager::set_threads(int n) {
  if (n == forced_threads_) return;
  forced_threads_ = n;
  if (ready_) {                      // restart the ring with the new thread count at its next use
    for (int i = 0; i < kSlots; ++i) cudaEventSynchronize(done_[i]);
    release();
  }
}

void Stager::release() {
  if (!ready_) return;
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  cv_work_.notify_all();
  for (auto &t : threads_) t.join();
  threads_.clear();
  for (int i = 0; i < kSlots; ++i) {
    cudaFreeHost(slot_[i]);
    cudaEventDestroy(done_[i]);
    slot_[i] = nullptr;
  }
  ready_ = false;
}

Stager::~Stager() {
  // at process exit the CUDA context may already be gone: only stop the threads
  if (!ready_) return;
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  cv_work_.notify_all();
  for (auto &t : threads_) t.join();
}

void Stager::worker(int id) {
  int seen = 0;
  for (;;) {
    std::function<void(size_t, size_t)> const *job;
    size_t rows;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_work_.wait(lk, [&] { return stop_ || generation_ != seen; });
      if (stop_) return;
      seen = generation_;
      job = job_;
      rows = job_rows_;
    }
    size_t const per = (rows + kThreads - 1) / kThreads;
    size_t const r0 = per * id < rows ? per * id : rows, r1 = per * (id + 1) < rows ? per * (id + 1) : rows;
    if (r1 > r0) (*job)(r0, r1);
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (--pending_ == 0) cv_done_.notify_one();
    }
  }
}

// run fn(r0, r1) over [0, rows) on kThreads threads (the caller is thread 0)
void Stager::parallel_rows(size_t rows, std::function<void(size_t, size_t)> const &fn) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    job_ = &fn;
    job_rows_ = rows;
    pending_ = kThreads - 1;
    ++generation_;
  }
  cv_work_.notify_all();
  size_t const per = (rows + kThreads - 1) / kThreads;
  if (per > 0) fn(0, per < rows ? per : rows);
  std::unique_lock<std::mutex> lk(mu_);
  cv_done_.wait(lk, [&] { return pending_ == 0; });
}

void Stager::upload2d(void *dst, size_t dpitch, void const *src, size_t spitch, size_t width, size_t rows, cudaStream_t s) {
  if (!rows || !width) return;
  ensure();
  size_t const rows_per_chunk = kChunkBytes / width ? kChunkBytes / width : 1;
  if (width > kChunkBytes) {   // absurdly wide rows: let the driver do it
    M4B_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyHostToDevice, s));
    return;
  }
  int slot = 0;
  for (size_t r = 0; r < rows; r += rows_per_chunk, slot = (slot + 1) % kSlots) {
    size_t const nr = rows - r < rows_per_chunk ? rows - r : rows_per_chunk;
    M4B_CUDA(cudaEventSynchronize(done_[slot]));   // the DMA that last read this slot has finished
    char *buf = slot_[slot];
    char const *from = static_cast<char const *>(src) + r * spitch;
    parallel_rows(nr, [=](size_t a, size_t b) {
      if (spitch == width) memcpy(buf + a * width, from + a * spitch, (b - a) * width);
      else for (size_t i = a; i < b; ++i) memcpy(buf + i * width, from + i * spitch, width);
    });
    M4B_CUDA(cudaMemcpy2DAsync(static_cast<char *>(dst) + r * dpitch, dpitch, buf, width, width, nr,
                               cudaMemcpyHostToDevice, s));
    M4B_CUDA(cudaEventRecord(done_[slot], s));
  }
}

void Stager::download2d(void *dst, size_t dpitch, void const *src, size_t spitch, size_t width, size_t rows, cudaStream_t s) {
  if (!rows || !width) return;
  ensure();
  if (width > kChunkBytes) {
    M4B_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToHost, s));
    M4B_CUDA(cudaStreamSynchronize(s));
    return;
  }
  size_t const rows_per_chunk = kChunkBytes / width ? kChunkBytes / width : 1;
  size_t const nchunks = (rows + rows_per_chunk - 1) / rows_per_chunk;
  auto issue = [&](size_t c) {
    int const slot = (int)(c % kSlots);
    size_t const r = c * rows_per_chunk, nr = rows - r < rows_per_chunk ? rows - r : rows_per_chunk;
    // the slot may still be the source of an H2D copy that an earlier upload2d left in flight on another stream
    M4B_CUDA(cudaEventSynchronize(done_[slot]));
    M4B_CUDA(cudaMemcpy2DAsync(slot_[slot], width, static_cast<char const *>(src) + r * spitch, spitch, width, nr,
                               cudaMemcpyDeviceToHost, s));
    M4B_CUDA(cudaEventRecord(done_[slot], s));
  };
  for (size_t c = 0; c < nchunks && c < (size_t)kSlots; ++c) issue(c);
  for (size_t c = 0; c < nchunks; ++c) {
    int const slot = (int)(c % kSlots);
    size_t const r = c * rows_per_chunk, nr = rows - r < rows_per_chunk ? rows - r : rows_per_chunk;
    M4B_CUDA(cudaEventSynchronize(done_[slot]));
    char const *buf = slot_[slot];
    char *to = static_cast<char *>(dst) + r * dpitch;
    parallel_rows(nr, [=](size_t a, size_t b) {
      if (dpitch == width) memcpy(to + a * dpitch, buf + a * width, (b - a) * width);
      else for (size_t i = a; i < b; ++i) memcpy(to + i * dpitch, buf + i * width, width);
    });
    if (c + kSlots < nchunks) issue(c + kSlots);
  }
}

}  // namespace m4b
