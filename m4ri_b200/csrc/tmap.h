// tmap.h — TMA tensor maps over bit-packed device views (shared by the leaf kernels).
#pragma once
#include <cuda.h>

#include "dev.h"

namespace m4b {

// 2D map over the u32 elements of a view: dim0 = 32-bit words of the (128-bit padded) row, dim1 = rows;
// box = box_w32 words x box_rows rows.  Reads outside [dim0) x [dim1) (negative coordinates included)
// return zeros.
CUtensorMap make_map(DView V, int box_w32, int box_rows);

}  // namespace m4b
