"""CPU test (-m "not gpu") of the N>1 path's host logic: world_size 2 and 3 over gloo.

Each rank holds a row-block of A and a padded row-slice of B, all-gathers B with
torch.distributed (gloo here, NCCL on the GPU box — same call in bench.py), multiplies locally
(the ORACLE stands in for the GPU leaf: this test is about the partition and the exchange), and
rank 0 checks the stacked row-blocks against the oracle's product of the full matrices."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from m4ri_b200 import shard
from tests import harness as H


def test_row_blocks_cover_and_align():
    for nrows, world in [(65536, 8), (1000, 3), (64, 8), (1, 2), (129, 2), (4100, 4)]:
        blocks = shard.row_blocks(nrows, world)
        assert len(blocks) == world and blocks[0][0] == 0 and blocks[-1][1] == nrows
        for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
            assert a1 == b0 and a0 <= a1
        assert all(r0 % 64 == 0 for r0, _ in blocks if r0 < nrows)
        assert shard.padded_slice_rows(nrows, world) * world >= nrows


def _matrix_from(words_2d, ncols):
    M = H.new(words_2d.shape[0], ncols)
    H.storage(M)[:, : words_2d.shape[1]] = words_2d
    return M


def _worker(rank, world, port, m, l, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(42)                       # every rank derives the same full inputs
    wl, wn = (l + 63) // 64, (n + 63) // 64
    A = rng.integers(0, 2**64, size=(m, wl), dtype=np.uint64)
    B = rng.integers(0, 2**64, size=(l, wn), dtype=np.uint64)
    if l % 64:
        A[:, -1] &= np.uint64((1 << (l % 64)) - 1)
    if n % 64:
        B[:, -1] &= np.uint64((1 << (n % 64)) - 1)
    r0, r1 = shard.row_blocks(m, world)[rank]
    per = shard.padded_slice_rows(l, world)
    s0, s1 = min(l, rank * per), min(l, (rank + 1) * per)
    b_slice = np.zeros((per, wn), dtype=np.uint64)
    b_slice[: s1 - s0] = B[s0:s1]

    def all_gather(sl):
        out = torch.empty((world * per, wn), dtype=torch.int64)
        dist.all_gather_into_tensor(out.view(-1), torch.from_numpy(sl.view(np.int64)).reshape(-1))
        return out.numpy().view(np.uint64)

    def local_mul(a_blk, b_full):
        if a_blk.shape[0] == 0:
            return np.zeros((0, wn), dtype=np.uint64)
        Am, Bm = _matrix_from(a_blk, l), _matrix_from(b_full[:l], n)
        Cm = H.oracle().orc_mul(None, Am, Bm, 0)
        out = H.storage(Cm)[:, :wn].copy()
        H.free(Am, Bm, Cm)
        return out

    c_blk = shard.sharded_product(rank, world, A[r0:r1], b_slice, all_gather, local_mul)
    gathered = [None] * world
    dist.gather_object((r0, r1, c_blk), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        C = np.zeros((m, wn), dtype=np.uint64)
        for g0, g1, blk in gathered:
            C[g0:g1] = blk
        Am, Bm = _matrix_from(A, l), _matrix_from(B, n)
        want = H.oracle().orc_mul(None, Am, Bm, 0)
        q.put(bool(np.array_equal(C, H.storage(want)[:, :wn])))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (300, 200, 260)), (2, (130, 1000, 70)), (3, (200, 333, 129))])
def test_sharded_product_over_gloo(world, shape):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + world * 7 + shape[0]) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, *shape, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


# ---- 2D grid (world >= 4): two row-blocks x world/2 column blocks, all-gather inside each column group ----

def test_grid_helpers():
    assert [shard.grid_shape(w) for w in (1, 2, 3, 4, 6, 8)] == [(1, 1), (2, 1), (3, 1), (2, 2), (2, 3), (2, 4)]
    assert shard.grid_shape(8, "rows") == (8, 1)
    for world in (4, 6, 8):
        pr, pc = shard.grid_shape(world)
        seen = set()
        for rank in range(world):
            gr, gc = shard.grid_coords(rank, world)
            assert 0 <= gr < pr and 0 <= gc < pc and (gr, gc) not in seen
            seen.add((gr, gc))
            grp = shard.column_group(rank, world)
            assert rank in grp and len(grp) == pr and grp == sorted(grp)
            assert all(shard.grid_coords(r, world)[1] == gc for r in grp)
            assert [shard.grid_coords(r, world)[0] for r in grp] == list(range(pr))   # gather order = row-slice order
    blocks = shard.col_blocks(65536, 2)
    assert blocks == [(0, 32768), (32768, 65536)]
    assert all(c0 % 128 == 0 for c0, _ in shard.col_blocks(1000, 2)) and shard.col_blocks(1000, 2)[-1][1] == 1000


def _worker2d(rank, world, port, m, l, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pr, pc = shard.grid_shape(world)
    gr, gc = shard.grid_coords(rank, world)
    groups = [dist.new_group(shard.column_group(g, world)) for g in range(pc)]
    rng = np.random.default_rng(43)                       # every rank derives the same full inputs
    assert n % 64 == 0 and l % 64 == 0                    # whole words: column blocks are cut on word boundaries here
    wl, wn = l // 64, n // 64
    A = rng.integers(0, 2**64, size=(m, wl), dtype=np.uint64)
    B = rng.integers(0, 2**64, size=(l, wn), dtype=np.uint64)
    r0, r1 = shard.row_blocks(m, pr)[gr]
    c0, c1 = shard.col_blocks(n, pc)[gc]
    w0, w1 = c0 // 64, (c1 + 63) // 64
    per = shard.padded_slice_rows(l, pr)
    s0, s1 = min(l, gr * per), min(l, (gr + 1) * per)
    b_piece = np.zeros((per, w1 - w0), dtype=np.uint64)
    b_piece[: s1 - s0] = B[s0:s1, w0:w1]

    def group_all_gather(piece, ranks):
        assert ranks == shard.column_group(rank, world)
        out = torch.empty((len(ranks) * per, piece.shape[1]), dtype=torch.int64)
        dist.all_gather_into_tensor(out.view(-1), torch.from_numpy(piece.view(np.int64)).reshape(-1), group=groups[gc])
        return out.numpy().view(np.uint64)

    def local_mul(a_blk, b_col):
        if a_blk.shape[0] == 0 or c1 == c0:
            return np.zeros((a_blk.shape[0], w1 - w0), dtype=np.uint64)
        Am, Bm = _matrix_from(a_blk, l), _matrix_from(b_col[:l], c1 - c0)
        Cm = H.oracle().orc_mul(None, Am, Bm, 0)
        out = H.storage(Cm)[:, : w1 - w0].copy()
        H.free(Am, Bm, Cm)
        return out

    c_blk = shard.sharded_product_2d(rank, world, A[r0:r1], b_piece, group_all_gather, local_mul)
    gathered = [None] * world
    dist.gather_object((r0, r1, w0, w1, c_blk), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        C = np.zeros((m, wn), dtype=np.uint64)
        for g0, g1, v0, v1, blk in gathered:
            C[g0:g1, v0:v1] = blk
        Am, Bm = _matrix_from(A, l), _matrix_from(B, n)
        want = H.oracle().orc_mul(None, Am, Bm, 0)
        q.put(bool(np.array_equal(C, H.storage(want)[:, :wn])))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(4, (300, 256, 512)), (4, (130, 1024, 256)), (6, (200, 320, 384))])
def test_grid_product_over_gloo(world, shape):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + world * 11 + shape[0]) % 2000
    procs = [ctx.Process(target=_worker2d, args=(r, world, port, *shape, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


# ---- K-chunk pipeline (end-to-end path of bench.py --gpus N): schedule + exchanges over gloo ------------------

def test_chunk_schedule_and_groups():
    assert shard.chunk_schedule(1, 4, 2) == [(1, 0), (1, 1), (2, 0), (2, 1), (3, 0), (3, 1), (0, 0), (0, 1)]
    assert shard.chunk_schedule(0, 2) == [(0, 0), (1, 0)]
    for world in (2, 4, 8):
        pr, pc = shard.grid_shape(world)
        for rank in range(world):
            grp = shard.row_group(rank, world)
            assert rank in grp and len(grp) == pc
            assert [shard.grid_coords(r, world)[1] for r in grp] == list(range(pc))
        # the chunks of all slices tile [0, l) exactly
        l, sub = 128 * pr * 3 * 2, 3
        ranges = sorted(shard.chunk_range(l, pr, sub, g, j) for g in range(pr) for j in range(sub))
        assert ranges[0][0] == 0 and ranges[-1][1] == l and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    with pytest.raises(ValueError):
        shard.chunk_range(1000, 2, 1, 0, 0)


def _worker_pipe(rank, world, port, m, l, n, sub, accumulate, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pr, pc = shard.grid_shape(world)
    gr, gc = shard.grid_coords(rank, world)
    col_groups = [dist.new_group(shard.column_group(g, world)) for g in range(pc)]
    row_groups = [dist.new_group(shard.row_group(g * pc, world)) for g in range(pr)]
    rng = np.random.default_rng(44)
    wl, wn = l // 64, n // 64
    A = rng.integers(0, 2**64, size=(m, wl), dtype=np.uint64)
    B = rng.integers(0, 2**64, size=(l, wn), dtype=np.uint64)
    Cin = rng.integers(0, 2**64, size=(m, wn), dtype=np.uint64)
    r0, r1 = shard.row_blocks(m, pr)[gr]
    c0, c1 = shard.col_blocks(n, pc)[gc]
    w0, w1 = c0 // 64, c1 // 64
    rows, kc = r1 - r0, l // (pr * sub)
    part_rows = rows // pc
    assert rows % pc == 0 and rows % 2 == 0
    devA = {}                                                           # (g, j) -> [rows, kc/64], filled part by part
    devB = [np.zeros((pr, kc, w1 - w0), dtype=np.uint64) for _ in range(sub)]
    devC = np.zeros((rows, w1 - w0), dtype=np.uint64)
    hostC = np.full((rows, w1 - w0), 0xAAAAAAAAAAAAAAAA, dtype=np.uint64)
    log = []

    class Ops:
        def upload_c(self):
            devC[:, :] = Cin[r0:r1, w0:w1]

        def upload_b(self, j):
            k0, k1 = shard.chunk_range(l, pr, sub, gr, j)
            devB[j][gr] = B[k0:k1, w0:w1]

        def gather_b(self, j):
            log.append(("b", j))
            out = torch.empty((pr, kc, w1 - w0), dtype=torch.int64)
            dist.all_gather_into_tensor(out.view(-1), torch.from_numpy(devB[j][gr].view(np.int64)).reshape(-1), group=col_groups[gc])
            devB[j][:, :, :] = out.numpy().view(np.uint64)

        def upload_a(self, g, j):
            k0, k1 = shard.chunk_range(l, pr, sub, g, j)
            buf = np.zeros((rows, kc // 64), dtype=np.uint64)
            buf[gc * part_rows:(gc + 1) * part_rows] = A[r0 + gc * part_rows:r0 + (gc + 1) * part_rows, k0 // 64:k1 // 64]
            devA[(g, j)] = buf

        def gather_a(self, g, j):
            log.append(("a", g, j))
            buf = devA[(g, j)]
            out = torch.empty(buf.shape, dtype=torch.int64)
            mine = torch.from_numpy(buf[gc * part_rows:(gc + 1) * part_rows].view(np.int64)).reshape(-1)
            dist.all_gather_into_tensor(out.view(-1), mine, group=row_groups[gr])
            buf[:, :] = out.numpy().view(np.uint64)

        def mul(self, g, j, clear, part):
            i, nparts = part if part else (0, 1)
            a0, a1 = i * rows // nparts, (i + 1) * rows // nparts
            Am, Bm = _matrix_from(devA[(g, j)][a0:a1], kc), _matrix_from(devB[j][g], c1 - c0)
            Cm = _matrix_from(np.zeros_like(devC[a0:a1]) if clear else devC[a0:a1], c1 - c0)
            H.oracle().orc_addmul(Cm, Am, Bm, 0)
            devC[a0:a1] = H.storage(Cm)[:, : w1 - w0]
            H.free(Am, Bm, Cm)

        def download(self, part):
            i, nparts = part
            a0, a1 = i * rows // nparts, (i + 1) * rows // nparts
            hostC[a0:a1] = devC[a0:a1]

    shard.pipelined_product(rank, world, Ops(), sub=sub, accumulate=accumulate)
    gathered = [None] * world
    dist.gather_object((r0, r1, w0, w1, hostC, log), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        C = np.zeros((m, wn), dtype=np.uint64)
        for g0, g1, v0, v1, blk, _ in gathered:
            C[g0:g1, v0:v1] = blk
        Am, Bm = _matrix_from(A, l), _matrix_from(B, n)
        Cm = _matrix_from(Cin if accumulate else np.zeros_like(Cin), n)
        H.oracle().orc_addmul(Cm, Am, Bm, 0)
        ok = bool(np.array_equal(C, H.storage(Cm)[:, :wn]))
        # collectives were issued in the same communicator order by every rank (the no-deadlock argument)
        kinds = [[e[0] for e in lg] for *_, lg in gathered]
        ok = ok and all(k == kinds[0] for k in kinds)
        q.put(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,sub,accumulate", [(2, (128, 512, 256), 1, False), (2, (256, 1024, 128), 2, True),
                                                        (4, (256, 512, 256), 1, False), (4, (128, 1024, 256), 2, True)])
def test_pipelined_product_over_gloo(world, shape, sub, accumulate):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() + world * 13 + shape[1] + sub) % 2000
    procs = [ctx.Process(target=_worker_pipe, args=(r, world, port, *shape, sub, accumulate, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
