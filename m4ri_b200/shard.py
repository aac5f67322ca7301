"""Block sharding of C = A*B over ranks (one process per GPU): host-side logic only.

Rows of C are independent, so rank r owns rows [r0, r1) of A and C; B is needed whole: every rank
starts with a 1/world row-slice of B and the slices are all-gathered once (NCCL over NVLink on the
GPU box, gloo in the CPU tests).  The same split, inside one process, is implemented in C++ in
csrc/multi.cu (mzd_mul_mp); the reference's own block-parallel scheme is m4ri/mp.c:158-275.

From 4 ranks on C is cut into a grid of 2 row-blocks x world/2 COLUMN blocks: rank (gr, gc) owns
C[rows gr, cols gc] = A[rows gr, :] * B[:, cols gc], starts with row-slice gr of B[:, cols gc] (still
1/world of B) and all-gathers only inside its column group (the pr ranks that share gc).  Same bits; the
local product keeps 32768 rows (8 GPUs: 32768 x 65536 x 16384 instead of 8192 x 65536 x 65536), so the
Strassen recursion still reaches 4096-row leaves — the tall-tile M4RM kernel, 49 products per launch —
after three levels.

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
    """(pr, pc).  'rows': pure row-blocks; 'auto': from 4 ranks on (even world only) TWO row-blocks x world/2 column
    blocks: the tall-tile leaf wants 4096-row operands, so the local product keeps half of the rows and with them all
    but one of the Strassen levels of the one-GPU product (8 ranks: 32768 x 65536 x 16384 per rank, three levels,
    where 4 row-blocks x 2 column blocks — 16384 x 65536 x 32768 — would stop at two)."""
    if mode not in ("auto", "rows"):
        raise ValueError(mode)
    pc = world // 2 if mode == "auto" and world >= 4 and world % 2 == 0 else 1
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


# ---- K-chunk pipeline of the end-to-end path (host <-> device transfers overlapped with compute) ------------
#
# C_blk = XOR over K-chunks of A[rows gr, K_c] * B[K_c, cols gc] is exact in any order, so a rank
# can start multiplying as soon as ONE chunk of its operands is on the device and keep the PCIe link busy with
# the next chunks meanwhile.  The K range is cut into pr * sub chunks: chunk (g, j) is sub-chunk j of the
# row-slice of B that rank (g, gc) contributes to its column group.  A rank's own slice needs no exchange, so it
# goes first; the other slices follow in rotated order.  PCIe carries every operand bit exactly once per node:
# the pc ranks of a row group each upload 1/pc of the rows of an A chunk and all-gather the parts over NVLink
# (row group), just as the pr ranks of a column group do for B.  The last chunk's product is cut into row parts
# so that the download of one part overlaps the product of the next.

def row_group(rank: int, world: int, mode: str = "auto") -> List[int]:
    """The ranks that share this rank's row-block of A and C, in column-block order (= gather order)."""
    _, pc = grid_shape(world, mode)
    gr = rank // pc
    return [gr * pc + gc for gc in range(pc)]


def chunk_schedule(gr: int, pr: int, sub: int = 1) -> List[Tuple[int, int]]:
    """Order in which rank (gr, *) consumes the K-chunks (g, j): its own slice first, then g = gr+1, ... (mod pr)."""
    return [((gr + d) % pr, j) for d in range(pr) for j in range(sub)]


def chunk_range(l: int, pr: int, sub: int, g: int, j: int, align: int = 128) -> Tuple[int, int]:
    """[k0, k1) of K-chunk (g, j); the chunk size must be a multiple of `align` bits (device views and TMA boxes)."""
    if l % (pr * sub * align):
        raise ValueError(f"l = {l} is not a multiple of {pr * sub * align}")
    kc = l // (pr * sub)
    k0 = (g * sub + j) * kc
    return k0, k0 + kc


def pipelined_product(rank: int, world: int, ops, sub: int = 1, mode: str = "auto", tail_parts: int = 2,
                      accumulate: bool = False) -> None:
    """One end-to-end step on this rank.  `ops` supplies the transfers, exchanges and products (all asynchronous
    on the caller's streams; bench.py on the GPU, numpy + gloo in the CPU tests):

      upload_c()                 C block (accumulate only)
      upload_b(j)                this rank's sub-chunk j of its B row-slice          -> device
      gather_b(j)                all-gather of sub-chunk j inside the column group   (collective, column group)
      upload_a(g, j)             this rank's 1/pc row part of A chunk (g, j)         -> device
      gather_a(g, j)             all-gather of the parts inside the row group        (collective, row group; pc > 1)
      mul(g, j, clear, part)     C[part] (^)= A[part rows, K(g,j)] * B[K(g,j), :]    part = None (all rows) or (i, nparts)
      download(part)             C[part] -> host (waits for the last product of that part only)

    Every rank issues its collectives in the same relative order (position p of every rank's sequence is the same
    collective on the same communicator), so the schedule cannot deadlock."""
    pr, pc = grid_shape(world, mode)
    gr, _ = grid_coords(rank, world, mode)
    sched = chunk_schedule(gr, pr, sub)
    if accumulate:
        ops.upload_c()
    for idx, (g, j) in enumerate(sched):
        own = g == gr
        if own:
            ops.upload_b(j)
        ops.upload_a(g, j)
        if pc > 1:
            ops.gather_a(g, j)
        clear = idx == 0 and not accumulate
        if idx + 1 < len(sched):
            ops.mul(g, j, clear, None)
        else:
            for i in range(tail_parts):
                ops.mul(g, j, clear, (i, tail_parts))
        if own and pr > 1:
            ops.gather_b(j)          # after the own product: it must not wait for the slowest peer's upload
    # downloads are issued after every product has been enqueued: a device->host copy blocks the calling thread
    # until its part is final, and the parts that follow must already be running behind it
    for i in range(tail_parts):
        ops.download((i, tail_parts))
