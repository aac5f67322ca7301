"""Row-block sharding of C = A*B over ranks (one process per GPU): host-side logic only.

Rows of C are independent, so rank r owns rows [r0, r1) of A and C; B is needed whole: every rank
starts with a 1/world row-slice of B and the slices are all-gathered once (NCCL over NVLink on the
GPU box, gloo in the CPU tests).  The same split, inside one process, is implemented in C++ in
csrc/multi.cu (mzd_mul_mp); the reference's own block-parallel scheme is m4ri/mp.c:158-275.

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
