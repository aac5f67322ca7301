"""Block sharding of C = A*B over ranks (one process per GPU): host-side logic only.

Rows of C are independent, so rank r owns rows [r0, r1) of A and C; B is needed whole: every rank
starts with a 1/world row-slice of B and the slices are all-gathered once (NCCL over NVLink on the
GPU box, gloo in the CPU tests).  The same split, inside one process, is implemented in C++ in
csrc/multi.cu (mzd_mul_mp); the reference's own block-parallel scheme is m4ri/mp.c:158-275.

From 4 ranks on the row-blocks are additionally cut into two COLUMN blocks (grid pr x 2): rank
(gr, gc) owns C[rows gr, cols gc] = A[rows gr, :] * B[:, cols gc], starts with row-slice gr of
B[:, cols gc] (still 1/world of B) and all-gathers only inside its column group (the pr ranks that
share gc).  Same bits; the local product is closer to a cube (8 GPUs: 16384 x 65536 x 32768 instead of
8192 x 65536 x 65536), so the same Strassen depth ends in 4096-row leaves — the tall-tile M4RM kernel —
and each rank receives 3/16 instead of 7/8 of B.

The local multiply and the collective are passed in, so this module carries no compute and no
device dependency.
"""
from __future__ import annotations

from typing import Callable, List, Tuple


def row_blocks(nrows: int, world: int, align: int = 64) -> List[Tuple[int, int]]:
    """`world` contiguous [r0, r1) blocks covering [0, nrows): equal sizes rounded up to `align`
    rows; trailing blocks may be short or empty (nrows not divisible by align*world)."""
    per = -(-nrows // world)
    per = -(-per // align) * align
    return [(min(nrows, r * per), min(nrows, (r + 1) * per)) for r in range(world)]


def padded_slice_rows(nrows: int, world: int, align: int = 64) -> int:
    """rows of one all-gather slice: every rank contributes the same count (zero rows pad the tail)"""
    per = -(-nrows // world)
    return -(-per // align) * align


def sharded_product(rank: int, world: int, a_block, b_slice, all_gather: Callable, local_mul: Callable):
    """One step of the sharded path on this rank.

    a_block    this rank's rows of A
    b_slice    this rank's (padded) row-slice of B
    all_gather f(b_slice) -> full (padded) B, the path's single exchange step
    local_mul  f(a_block, b_full) -> this rank's rows of C
    """
    b_full = all_gather(b_slice)
    return local_mul(a_block, b_full)


# ---- 2D grid (pr row-blocks x pc column-blocks of C) -----------------------------------------

def grid_shape(world: int, mode: str = "auto") -> Tuple[int, int]:
    """(pr, pc).  'rows': pure row-blocks; 'auto': two column blocks from 4 ranks on (even world only)."""
    if mode not in ("auto", "rows"):
        raise ValueError(mode)
    pc = 2 if mode == "auto" and world >= 4 and world % 2 == 0 else 1
    return world // pc, pc


def grid_coords(rank: int, world: int, mode: str = "auto") -> Tuple[int, int]:
    """(gr, gc) of a rank: ranks are numbered row-block major, so a column group is rank % pc."""
    _, pc = grid_shape(world, mode)
    return rank // pc, rank % pc


def column_group(rank: int, world: int, mode: str = "auto") -> List[int]:
    """The ranks that share this rank's column block of B and C, in row-block order (= gather order)."""
    pr, pc = grid_shape(world, mode)
    gc = rank % pc
    return [gr * pc + gc for gr in range(pr)]


def col_blocks(ncols: int, pc: int, align: int = 128) -> List[Tuple[int, int]]:
    """`pc` contiguous [c0, c1) column blocks, widths rounded up to `align` bits (device rows and the
    TMA boxes of the leaf are 128-bit granular)."""
    return row_blocks(ncols, pc, align)


def sharded_product_2d(rank: int, world: int, a_block, b_piece, group_all_gather: Callable, local_mul: Callable,
                       mode: str = "auto"):
    """One step of the grid path on this rank.

    a_block           rows gr of A (all columns)
    b_piece           (padded) row-slice gr of B[:, cols gc]
    group_all_gather  f(b_piece, ranks) -> B[:, cols gc] (padded rows), gathered over `ranks` in order
    local_mul         f(a_block, b_colblock) -> C[rows gr, cols gc]
    """
    b_col = group_all_gather(b_piece, column_group(rank, world, mode))
    return local_mul(a_block, b_col)
