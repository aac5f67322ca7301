#!/usr/bin/env python
"""bench.py — GF(2) matmul bit-ops/s (2*n^3) at n = 65536 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 65536] [--cutoff 0]

Own arm ("ours"):  one step = one mzd_mul of two random n x n GF(2) matrices (Strassen-Winograd
over the M4RM leaf kernel, all on the GPU).
  value      inputs already resident in HBM, timed with CUDA events on the launching stream,
             max over ranks.
  e2e        the same product through the reference-facing C-ABI call mzd_mul(C, A, B, cutoff)
             with HOST (pinned) mzd_t operands: H2D of A and B and D2H of C inside the timed region.
  roofline   the dominant kernel (m4rm_streamk_kernel) timed live with CUDA events around every
             leaf launch of the timed region (library hook m4ri_b200_profile_*).
  N > 1      C's row-blocks are sharded over the ranks (A row-block local, B row-slices
             all-gathered over NCCL/NVLink every step), total work fixed -> "scaling": "strong".
Reference arm ("--impl reference"): the unmodified reference (oracle/_ref/libm4ri_ref_omp.so,
mzd_mul_mp with all host threads; serial mzd_mul if the OpenMP build is absent; the oracle port
as last resort) on a bounded square sample of the same workload, rank 0 only.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gf2_matmul_bitops_per_s"
UNIT = "bit-ops/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ---------------------------------------------------------------------------------------------
# helpers shared by both arms
# ---------------------------------------------------------------------------------------------

def make_header(MzdT, ptr, nrows, ncols, rowstride):
    """A host mzd_t header over caller-owned words (layout: m4ri/mzd.h:68-99)."""
    h = MzdT()
    h.nrows, h.ncols = nrows, ncols
    h.width = (ncols + 63) // 64
    h.rowstride = rowstride
    h.flags = 0x2 if ncols % 64 else 0
    h.high_bitmask = (1 << (ncols % 64)) - 1 if ncols % 64 else 2**64 - 1
    h.data = ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint64))
    return h


def fill_random_words(arr_u64, seed):
    """uniform random bits (density 1/2), deterministic per seed; chunked to bound temporaries"""
    rng = np.random.default_rng(seed)
    flat = arr_u64.reshape(-1)
    step = 1 << 24
    for i in range(0, flat.size, step):
        j = min(flat.size, i + step)
        flat[i:j] = rng.integers(0, 2**64, size=j - i, dtype=np.uint64)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            time.sleep(0.25)
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline
# ---------------------------------------------------------------------------------------------

def time_reference(sample_n, runs, warm):
    """Times the reference's own CPU multiply on sample_n^3 random inputs.
    Returns (seconds per run, description dict)."""
    from tests import harness as H

    threads = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    lib, kind, fn_name, cores = None, "reference", "mzd_mul_mp", threads
    if os.path.exists(H.REF_OMP_SO):
        lib = H.ref_omp()
    if lib is None and H.ref() is not None:
        lib, fn_name, cores = H.ref(), "mzd_mul", 1
    if lib is not None:
        A, B = lib.mzd_init(sample_n, sample_n), lib.mzd_init(sample_n, sample_n)
        fn = getattr(lib, fn_name)
        free = lib.mzd_free
    else:  # neither reference build present: the oracle port (scalar, one thread)
        O = H.oracle()
        kind, fn_name, cores = "port", "orc_mul", 1
        A, B = O.orc_init(sample_n, sample_n), O.orc_init(sample_n, sample_n)
        fn, free = O.orc_mul, O.orc_free
    fill_random_words(H.storage(A), 101)
    fill_random_words(H.storage(B), 102)
    times = []
    for i in range(warm + runs):
        t0 = time.perf_counter()
        C = fn(None, A, B, 0)
        dt = time.perf_counter() - t0
        free(C)
        if i >= warm:
            times.append(dt)
    free(A)
    free(B)
    sec = sum(times) / len(times)
    desc = {"kind": kind, "cores": cores,
            "sample": f"{fn_name}(C, A, B, cutoff=0) on random {sample_n}^3 (1/{(65536 // sample_n) ** 3} of the "
                      f"65536^3 workload's bit-ops; reference -O2 SSE2 build, {len(times)} timed run(s), "
                      f"wall clock around the call as in bench/bench_multiplication.c:85-107)"}
    return sec, desc


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return  # other ranks exit 0 without work
    total = args.steps + args.warmup
    sample_n = 32768 if total <= 6 else 16384
    sec, desc = time_reference(sample_n, args.steps, args.warmup)
    value = 2.0 * sample_n ** 3 / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"mzd_mul {args.n}x{args.n}x{args.n} random GF(2) (BASELINE config 3); "
                               f"CPU arm timed on a bounded {sample_n}^3 sample", "sample_n": sample_n},
        "cpu_baseline": dict(desc, value=value, unit=UNIT),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------

def run_own_arm(args):
    import torch
    import torch.distributed as dist

    import m4ri_b200
    from m4ri_b200 import MzdT

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    lib = m4ri_b200.load_library()
    if lib.m4ri_b200_device_count() < 1:
        raise SystemExit("no CUDA device: m4ri_b200 has no CPU fallback")
    lib.m4ri_b200_set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n = args.n
    if n % (128 * world):
        raise SystemExit("n must be a multiple of 128 * gpus")
    cutoff = args.cutoff
    from m4ri_b200 import shard
    # C is cut into pr row-blocks x pc column-blocks (pc = 2 from 4 ranks on, see m4ri_b200/shard.py):
    # this rank owns C[rows gr, cols gc] = A[rows gr, :] * B[:, cols gc]
    pr, pc = shard.grid_shape(world, args.grid)
    gr, gc = shard.grid_coords(rank, world, args.grid)
    r0, r1 = shard.row_blocks(n, pr)[gr]
    c0, c1 = shard.col_blocks(n, pc)[gc]
    rows, ncb = r1 - r0, c1 - c0           # this rank's block of C; A block is rows x n, B block n x ncb
    brow = shard.padded_slice_rows(n, pr)  # this rank's row-slice of B[:, cols gc] (n % (64*world) == 0: no padding)
    pitch, pitchb = n // 64, ncb // 64
    group = None
    if world > 1 and pc > 1:               # every rank creates every column group, in the same order
        groups = [dist.new_group(shard.column_group(g, world, args.grid)) for g in range(pc)]
        group = groups[gc]

    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    sh = ctypes.c_void_p(tstream.cuda_stream)

    # ---- host inputs (pinned): A row-block, B piece, C block -------------------------------------
    pin = not args.pageable
    hA = torch.empty((rows, pitch), dtype=torch.int64, pin_memory=pin)
    hB = torch.empty((brow, pitchb), dtype=torch.int64, pin_memory=pin)
    hC = torch.zeros((rows, pitchb), dtype=torch.int64, pin_memory=pin)
    fill_random_words(hA.numpy().view(np.uint64), 1000 + gr)          # ranks of one row-block share A's rows
    fill_random_words(hB.numpy().view(np.uint64), 2000 + rank)
    mA = make_header(MzdT, hA.data_ptr(), rows, n, pitch)
    mB = make_header(MzdT, hB.data_ptr(), brow, ncb, pitchb)
    mC = make_header(MzdT, hC.data_ptr(), rows, ncb, pitchb)

    # ---- device matrices (torch owns the memory; the library sees plain pointers) --------------
    tA = torch.zeros((rows, pitch), dtype=torch.int64, device="cuda")
    tBs = torch.zeros((brow, pitchb), dtype=torch.int64, device="cuda")
    tB = tBs if pr == 1 else torch.zeros((n, pitchb), dtype=torch.int64, device="cuda")
    tC = torch.zeros((rows, pitchb), dtype=torch.int64, device="cuda")
    dA = lib.m4ri_b200_dmat_wrap(tA.data_ptr(), pitch, rows, n)
    dBs = lib.m4ri_b200_dmat_wrap(tBs.data_ptr(), pitchb, brow, ncb)
    dB = lib.m4ri_b200_dmat_wrap(tB.data_ptr(), pitchb, n, ncb)
    dC = lib.m4ri_b200_dmat_wrap(tC.data_ptr(), pitchb, rows, ncb)
    lib.m4ri_b200_upload(dA, ctypes.byref(mA), sh)
    lib.m4ri_b200_upload(dBs, ctypes.byref(mB), sh)
    torch.cuda.synchronize()

    def exchange():
        if pr > 1:  # the path's one exchange step: all-gather of the row-slices of B[:, cols gc] over NVLink
            dist.all_gather_into_tensor(tB.view(-1), tBs.view(-1), group=group)

    def step_resident():
        exchange()
        lib.m4ri_b200_dmul(dC, dA, dB, cutoff, 1, sh)

    def step_e2e():
        if world == 1:
            lib.mzd_mul(ctypes.byref(mC), ctypes.byref(mA), ctypes.byref(mB), cutoff)   # the drop-in call
        else:
            lib.m4ri_b200_upload(dA, ctypes.byref(mA), sh)
            lib.m4ri_b200_upload(dBs, ctypes.byref(mB), sh)
            exchange()
            lib.m4ri_b200_dmul(dC, dA, dB, cutoff, 1, sh)
            lib.m4ri_b200_download(ctypes.byref(mC), dC, sh)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident-input timing ------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.m4ri_b200_kernel_launches()
    lib.m4ri_b200_profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    wall1 = time.time()
    leaf_ms, leaf_bitops = ctypes.c_double(0), ctypes.c_double(0)
    leaf_launches = lib.m4ri_b200_profile_end(ctypes.byref(leaf_ms), ctypes.byref(leaf_bitops))
    launches = lib.m4ri_b200_kernel_launches() - launches0
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    path = lib.m4ri_b200_last_path().decode()
    total_bitops = 2.0 * n * n * n
    value = total_bitops / (ms_step * 1e-3)

    # ---- end-to-end timing (host buffers in, host buffer out) ------------------------------------
    if args.no_e2e:
        e2e_s = float("nan")
    else:
        for _ in range(min(args.warmup, 2)):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    e2e_value = total_bitops / e2e_s
    h2d = (rows * pitch + brow * pitchb) * 8 * world
    d2h = rows * pitchb * 8 * world

    if args.verify:   # small-n correctness of this rank's block against the oracle (never at full size)
        from tests import harness as H
        step_e2e()
        col_b = torch.empty((n, pitchb), dtype=torch.int64)
        col_b.copy_(tB)
        Ao, Bo = H.new(rows, n), H.new(n, ncb)
        H.storage(Ao)[:, :] = hA.numpy().view(np.uint64)
        H.storage(Bo)[:, :] = col_b.numpy().view(np.uint64)
        want = H.oracle().orc_mul(None, Ao, Bo, 0)
        ok = bool(np.array_equal(H.storage(want), hC.numpy().view(np.uint64)))
        print(f"[verify] rank {rank}: rows {r0}:{r1} cols {c0}:{c1} {'OK' if ok else 'MISMATCH'}", file=sys.stderr, flush=True)
        if not ok:
            raise SystemExit(3)

    if args.check:    # any size, on the device: the exchange delivered the pieces in order, and Freivalds on this
        #               rank's block with 128 random vectors: C_blk * X == A_blk * (B_colblk * X)   (error 2^-128)
        step_resident()
        torch.cuda.synchronize()
        ok = True
        if pr > 1:
            mine = tBs.sum().reshape(1)
            sums = torch.empty(pr, dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(sums, mine, group=group)
            ok = bool(torch.equal(tB.view(pr, brow, pitchb).sum(dim=(1, 2)), sums))
        tX = torch.randint(-2**62, 2**62, (ncb, 2), dtype=torch.int64, device="cuda")
        tY = torch.zeros((n, 2), dtype=torch.int64, device="cuda")
        tZ = torch.zeros((rows, 2), dtype=torch.int64, device="cuda")
        tW = torch.zeros((rows, 2), dtype=torch.int64, device="cuda")
        dX, dY = lib.m4ri_b200_dmat_wrap(tX.data_ptr(), 2, ncb, 128), lib.m4ri_b200_dmat_wrap(tY.data_ptr(), 2, n, 128)
        dZ, dW = lib.m4ri_b200_dmat_wrap(tZ.data_ptr(), 2, rows, 128), lib.m4ri_b200_dmat_wrap(tW.data_ptr(), 2, rows, 128)
        torch.cuda.synchronize()
        lib.m4ri_b200_dmul_m4rm(dY, dB, dX, 1, sh)
        lib.m4ri_b200_dmul_m4rm(dZ, dA, dY, 1, sh)
        lib.m4ri_b200_dmul_m4rm(dW, dC, dX, 1, sh)
        torch.cuda.synchronize()
        ok = ok and bool(torch.equal(tZ, tW)) and bool(tZ.any())
        print(f"[check] rank {rank}: rows {r0}:{r1} cols {c0}:{c1} {'OK' if ok else 'MISMATCH'}", file=sys.stderr, flush=True)
        if not ok:
            raise SystemExit(3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    leaf_avg_ms = leaf_ms.value / max(1, leaf_launches)
    leaf_rate = leaf_bitops.value / (leaf_ms.value * 1e-3) if leaf_ms.value > 0 else 0.0
    # shared-memory roofline.  ALGORITHMIC shared-memory bytes per step of
    # 16 A-columns on a 1024x1024 C tile: 1024 rows x 2 table-row lookups x 128 B + 512 table entries
    # x 128 B written + the A (1024 x 2 B) and B (16 x 128 B) slab reads = 331776 B per 2*16*1024*1024 bit-ops.
    # The tall-tile leaf (variant 2, m4rm_leaf2.cu) per step of 32 A-columns on a 4096 x 256-bit tile: 4096 rows x 8
    # 16-byte lookups + 2048 16-byte table pieces written + A (4096 x 4 B) and B (32 x 32 B) = 574464 B per
    # 2*32*4096*256 bit-ops.
    leaf_variant = lib.m4ri_b200_last_leaf_variant()
    if leaf_variant == 2:
        smem_bytes_per_bitop = (4096 * 8 * 16 + 2048 * 16 + 4096 * 4 + 32 * 32) / (2.0 * 32 * 4096 * 256)
    else:
        smem_bytes_per_bitop = (1024 * 2 * 128 + 512 * 128 + 1024 * 2 + 16 * 128) / (2.0 * 16 * 1024 * 1024)
    smem_peak_gbs = 128.0 * 148 * sm_max_mhz * 1e6 / 1e9
    smem_achieved_gbs = leaf_rate * smem_bytes_per_bitop / 1e9
    leaf_dims = None
    try:
        lv = int(path.split(":")[1]) if ":" in path else 0
        leaf_dims = (rows >> lv, n >> lv, ncb >> lv)
    except ValueError:
        lv = 0
    # compulsory HBM bytes of one leaf launch: read A and B once, RMW C once
    if leaf_dims:
        lm, ll, ln = leaf_dims
        hbm_bytes = (lm * ll + ll * ln + 2 * lm * ln) / 8.0
    else:
        hbm_bytes = 0.0
    # the last Strassen level runs its 7 products in ONE launch: scale the per-product figures
    products_per_launch = 1
    if leaf_dims and leaf_launches:
        products_per_launch = max(1, round(leaf_bitops.value / leaf_launches / (2.0 * leaf_dims[0] * leaf_dims[1] * leaf_dims[2])))
        hbm_bytes *= products_per_launch
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "leaf_traffic.json")
    if os.path.exists(tpath) and leaf_dims:
        with open(tpath) as f:
            traffic = json.load(f).get(("leaf2:" if leaf_variant == 2 else "") + "x".join(str(d) for d in leaf_dims))
        if traffic is not None:
            traffic *= products_per_launch
    roofline = {
        "kernel": "m4rm_leaf2_kernel" if leaf_variant == 2 else "m4rm_streamk_kernel", "bound": "smem", "unit": "GB/s",
        "algorithmic_smem_bytes_per_bitop": smem_bytes_per_bitop,
        "achieved": smem_achieved_gbs, "peak": smem_peak_gbs, "frac": smem_achieved_gbs / smem_peak_gbs,
        "peak_source": f"128 B/clk/SM x 148 SMs x {sm_max_mhz:.0f} MHz (clocks.max.sm, {peak_src})",
        "traffic": traffic,
        "leaf_launches": int(leaf_launches), "leaf_avg_ms": leaf_avg_ms, "leaf_dims": leaf_dims,
        "products_per_launch": products_per_launch,
        "leaf_bitops_per_s": leaf_rate, "leaf_share_of_step": leaf_ms.value / (ms_step * args.steps),
        "hbm": {"bound": "hbm", "unit": "GB/s", "achieved": hbm_bytes / (leaf_avg_ms * 1e-3) / 1e9 if leaf_avg_ms else 0.0,
                "peak": hbm_peak, "frac": (hbm_bytes / (leaf_avg_ms * 1e-3) / 1e9) / hbm_peak if leaf_avg_ms else 0.0,
                "algorithmic_bytes_per_launch": hbm_bytes, "peak_source": peak_src},
    }

    # ---- the HBM-bound kernel of the path: the device _mzd_add (C = A ^ B) on the full operands ----------
    add_roofline = None
    if world == 1:
        for _ in range(3):
            lib.m4ri_b200_dadd(dC, dA, dB, sh)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            lib.m4ri_b200_dadd(dC, dA, dB, sh)
        a1.record()
        torch.cuda.synchronize()
        add_ms = a0.elapsed_time(a1) / 10
        add_bytes = 3.0 * rows * pitch * 8           # 2 reads + 1 write, every byte once (world == 1: all n x n)
        add_gbs = add_bytes / (add_ms * 1e-3) / 1e9
        add_roofline = {"kernel": "ew_kernel<0> (_mzd_add)", "bound": "hbm", "unit": "GB/s", "achieved": add_gbs,
                        "peak": hbm_peak, "frac": add_gbs / hbm_peak, "ms": add_ms,
                        "algorithmic_bytes_per_launch": add_bytes, "peak_source": peak_src}

    # ---- CPU baseline (rank 0, N = 1 only; bounded sample) --------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sec, desc = time_reference(16384, 3, 1)
        cpu = dict(desc, value=2.0 * 16384 ** 3 / sec, unit=UNIT)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"mzd_mul {n}x{n}x{n} random GF(2), Strassen-Winograd + M4RM leaf (BASELINE config "
                               f"{'3' if world == 1 else '4'})", "n": n, "cutoff": cutoff or lib.m4ri_b200_get_default_cutoff(),
                   "path": path,
                   "parallelism": (f"row-block x{world}" if pc == 1 else f"C blocks {pr} x {pc} (row-blocks x column-blocks)") +
                                  ("" if world == 1 else ", NCCL all-gather of B per step" if pc == 1 else
                                   f", NCCL all-gather of B's column block inside each column group ({pr} ranks) per step"),
                   "local_product": [rows, n, ncb],
                   "l2": "inputs (3 x %d MiB) exceed the 126 MB L2; no flush needed" % (n * n // 8 >> 20)},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "host_memory": "pageable" if args.pageable else "pinned",
                "api": "mzd_mul(C, A, B, cutoff) on host mzd_t" if world == 1 else
                "upload + all_gather + m4ri_b200_dmul + download per rank"},
        "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks,
    }
    if add_roofline is not None:
        line["roofline_add"] = add_roofline
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", dest="n", type=int, default=65536, help="n of the n x n x n product")
    ap.add_argument("--cutoff", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--pageable", action="store_true", help="e2e with pageable (malloc) host matrices instead of pinned")
    ap.add_argument("--grid", default="auto", choices=["auto", "rows"],
                    help="C partition over ranks: 'rows' = row-blocks only; 'auto' = two column blocks from 4 ranks on")
    ap.add_argument("--check", action="store_true", help="device-side check at any size: exchange order + Freivalds on this rank's block")
    ap.add_argument("--verify", action="store_true", help="check this rank's C block against the oracle (small --n only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
