// workspace.h — grow-only device slab with stack (mark/release) allocation of padded views.
#pragma once
#include "dev.h"

namespace m4b {

struct Workspace {
  char  *base = nullptr;
  size_t cap  = 0;
  size_t top  = 0;

  static int64_t pitch_for(int ncols) { return (int64_t)((ncols + 127) / 128) * 2; }
  static size_t  bytes_for(int nrows, int ncols) {
    size_t b = (size_t)nrows * (size_t)pitch_for(ncols) * sizeof(word);
    return (b + 255) & ~(size_t)255;
  }
  // Make room for `bytes` in total.  Only legal while nothing is allocated (top == 0).
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    if (top != 0) die("m4ri_b200: workspace grown while in use\n");
    if (base) {
      M4B_CUDA(cudaDeviceSynchronize());
      M4B_CUDA(cudaFree(base));
      base = nullptr;
      cap  = 0;
    }
    M4B_CUDA(cudaMalloc(&base, bytes));
    cap = bytes;
  }
  DView alloc(int nrows, int ncols) {
    size_t const b = bytes_for(nrows, ncols);
    if (top + b > cap) die("m4ri_b200: workspace exhausted (%zu + %zu > %zu)\n", top, b, cap);
    DView v{reinterpret_cast<word *>(base + top), pitch_for(ncols), nrows, ncols};
    top += b;
    return v;
  }
  size_t mark() const { return top; }
  void   release(size_t m) { top = m; }
  void   destroy() {
    if (base) cudaFree(base);
    base = nullptr;
    cap = top = 0;
  }
};

size_t strassen_workspace_bytes(int m, int k, int n, int levels);

}  // namespace m4b
