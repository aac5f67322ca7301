// tmap.h — TMA tensor maps over bit-packed device views (shared by the leaf kernels).
#pragma once
#include <cuda.h>

#include "dev.h"

namespace m4b {

// 2D map over the u32 elements of a view: dim0 = 32-bit words of the (128-bit padded) row, dim1 = rows;
// box = box_w32 words x box_rows rows.  Reads outside [dim0) x [dim1) (negative coordinates included)
// return zeros.
CUtensorMap make_map(DView V, int box_w32, int box_rows);

// 3D map of the same view with its rows grouped: dim0 = words, dim1 = `group_rows` rows, dim2 = row groups;
// box = box_w32 x group_rows x box_groups, i.e. box_groups * group_rows consecutive rows in one TMA
// instruction (a 2D box is limited to 256 rows).  V.nrows must be a multiple of group_rows; whole groups
// past the last row arrive as zeros.
CUtensorMap make_map_row_groups(DView V, int box_w32, int group_rows, int box_groups);

}  // namespace m4b
