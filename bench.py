#!/usr/bin/env python
"""bench.py — GF(2) matmul bit-ops/s (W = 2*m*l*n) on the BASELINE.json configurations.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload cfg3|cfg2|cfg5] [--size n] [--cutoff c]

Workloads (BASELINE.json `configs`):
  cfg3 (default)  mzd_mul 65536^3, Strassen-Winograd over the M4RM leaf; N > 1 is config 4 (C blocks over ranks)
  cfg2            mzd_mul_m4rm 16384^3, the M4RM leaf kernel only (inputs fit the L2: flushed between steps)
  cfg5            mzd_addmul 32768 x 131072 x 32768 into a random C
Own arm ("ours"): one step = one product.
  value      inputs already resident in HBM, CUDA events on the launching stream, max over ranks
  e2e        the same product through the reference-facing C-ABI call on HOST mzd_t operands in PAGEABLE
             memory (what mzd_init gives a libm4ri user), H2D and D2H inside the timed region;
             e2e_pinned: the same from pinned host memory
  roofline   the dominant kernel (M4RM leaf) timed live with CUDA events around every leaf launch
  verified   after the timed legs: this rank's block of C against the sha256 digests of the unmodified
             reference's result (tests/golden/large_golden.json, seeded inputs) + Freivalds on the device
  N > 1      C is cut into pr x pc blocks over the ranks (m4ri_b200/shard.py); B's row-slices are
             all-gathered over NCCL/NVLink every step; total work fixed -> "scaling": "strong"
Reference arm ("--impl reference"): the unmodified reference (oracle/_ref/libm4ri_ref_omp.so, mzd_mul_mp /
mzd_addmul_mp with all host threads) on a bounded sample of the same workload, rank 0 only.
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
L2_BYTES = 126 * 1000 * 1000

WORKLOADS = {
    #        kind      entry point        m       l       n      golden case                reference sample (m, l, n)
    "cfg3": ("mul",    "mzd_mul",         65536,  65536,  65536, "cfg3_65536",              (32768, 32768, 32768)),
    "cfg2": ("mul",    "mzd_mul_m4rm",    16384,  16384,  16384, "cfg2_16384",              (16384, 16384, 16384)),
    "cfg5": ("addmul", "mzd_addmul",      32768, 131072,  32768, "cfg5_32768x131072x32768", (16384, 65536, 16384)),
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def bind_near_gpu(local):
    """Restricts this rank to the CPUs NVML reports as local to its GPU (same NUMA node / PCIe root), BEFORE any pinned
    host memory is allocated: with one rank per GPU all moving data at once, buffers that sit on the other socket cross
    the inter-socket link and share its bandwidth.  Returns (previous affinity, description) or (None, why not)."""
    try:
        import pynvml
        import torch
        prev = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1} & prev
        if not cpus or cpus == prev:
            return None, "NVML reports no narrower CPU set for this GPU"
        os.sched_setaffinity(0, cpus)
        return prev, f"{len(cpus)} of {len(prev)} CPUs (NVML cpu affinity of GPU {local})"
    except Exception as e:      # no NVML, no permission, ...: run unbound
        return None, f"unbound ({type(e).__name__}: {e})"


def workload_of(args):
    kind, fn, m, l, n, golden, sample = WORKLOADS[args.workload]
    if args.n:
        if args.workload == "cfg5":
            raise SystemExit("--size applies to the square workloads")
        m = l = n = args.n
        sample = (min(args.n, sample[0]),) * 3
    if args.ref_sample:
        sample = (args.ref_sample,) * 3 if args.workload != "cfg5" else (args.ref_sample, 4 * args.ref_sample, args.ref_sample)
    return kind, fn, m, l, n, golden, sample


def workload_name(args, world):
    kind, fn, m, l, n, _, _ = workload_of(args)
    cfg = {"cfg3": "3" if world == 1 else "4", "cfg2": "2", "cfg5": "5"}[args.workload]
    how = "M4RM leaf only" if fn == "mzd_mul_m4rm" else "Strassen-Winograd + M4RM leaf"
    return f"{fn} {m}x{l}x{n} random GF(2), {how} (BASELINE config {cfg})"


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

def time_reference(kind, dims, runs, warm):
    """Times the reference's own CPU multiply (mzd_mul_mp / mzd_addmul_mp of the OpenMP build, every host
    thread this process may use) on seeded random inputs of `dims`.  Returns (seconds per run, description)."""
    threads = host_threads()
    # torchrun exports OMP_NUM_THREADS=1 to its children: the reference arm must not inherit that
    os.environ["OMP_NUM_THREADS"] = str(threads)
    from tests import harness as H

    m, l, n = dims
    lib, what, fn_name, cores = None, "reference", "mzd_mul_mp" if kind == "mul" else "mzd_addmul_mp", threads
    if os.path.exists(H.REF_OMP_SO):
        lib = H.ref_omp()
        try:   # libgomp may have been initialised (with the inherited value) before this point: set it explicitly
            gomp = ctypes.CDLL("libgomp.so.1")
            gomp.omp_set_num_threads(threads)
            cores = int(gomp.omp_get_max_threads())
        except OSError:
            pass
    if lib is None and H.ref() is not None:
        lib, fn_name, cores = H.ref(), "mzd_mul" if kind == "mul" else "mzd_addmul", 1
    if lib is not None:
        A, B = lib.mzd_init(m, l), lib.mzd_init(l, n)
        fn, free = getattr(lib, fn_name), lib.mzd_free
        Cacc = lib.mzd_init(m, n) if kind == "addmul" else None
    else:  # neither reference build present: the oracle port (scalar, one thread)
        O = H.oracle()
        what, fn_name, cores = "port", "orc_mul" if kind == "mul" else "orc_addmul", 1
        A, B = O.orc_init(m, l), O.orc_init(l, n)
        fn, free = getattr(O, fn_name), O.orc_free
        Cacc = O.orc_init(m, n) if kind == "addmul" else None
    H.storage(A)[:, :A.contents.width] = H.seeded_words(H.SEED_A, m, A.contents.width)
    H.storage(B)[:, :B.contents.width] = H.seeded_words(H.SEED_B, l, B.contents.width)
    times = []
    for i in range(warm + runs):
        t0 = time.perf_counter()
        C = fn(Cacc, A, B, 0)
        dt = time.perf_counter() - t0
        if Cacc is None:
            free(C)
        if i >= warm:
            times.append(dt)
    for M in (A, B, Cacc):
        if M:
            free(M)
    sec = sum(times) / len(times)
    desc = {"kind": what, "cores": cores,
            "sample": f"{fn_name}(C, A, B, cutoff=0) on seeded random {m}x{l}x{n} "
                      f"(reference -O2 SSE2 OpenMP build, {len(times)} timed run(s) after {warm} warm-up, wall clock "
                      f"around the call as in bench/bench_multiplication.c:85-107)",
            "sample_dims": [m, l, n], "sample_seconds": sec}
    return sec, desc


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return  # other ranks exit 0 without work
    kind, fn, m, l, n, _, sample = workload_of(args)
    sec, desc = time_reference(kind, sample, args.steps, args.warmup)
    value = 2.0 * sample[0] * sample[1] * sample[2] / sec
    same = list(sample) == [m, l, n]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args, 1) +
                               ("" if same else f"; CPU arm timed on a bounded {sample[0]}x{sample[1]}x{sample[2]} sample "
                                                "(1/%d of the bit-ops; the reference's nominal rate grows with n — "
                                                "profiles/r02_reference_full_size.json holds a full-size run)"
                                                % round(m * l * n / (sample[0] * sample[1] * sample[2]))),
                   "sample_dims": list(sample), "same_size_as_workload": same},
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
    from m4ri_b200 import MzdT, shard
    from tests import harness as H     # seeded input generator + digest helpers (no oracle call on this path)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    prev_affinity, binding = None, "not requested"
    if world > 1:   # all ranks stage pageable rows at the same time: share the host's threads instead of oversubscribing them
        os.environ.setdefault("M4RI_B200_STAGE_THREADS", str(max(2, min(12, host_threads() // world))))
        if os.environ.get("M4RI_B200_BENCH_BIND", "1") != "0":
            prev_affinity, binding = bind_near_gpu(local)
        print(f"[bind] rank {rank}: {binding}", file=sys.stderr, flush=True)
    lib = m4ri_b200.load_library()
    if lib.m4ri_b200_device_count() < 1:
        raise SystemExit("no CUDA device: m4ri_b200 has no CPU fallback")
    lib.m4ri_b200_set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    kind, fn_name, m, l, n, golden_name, _ = workload_of(args)
    accumulate = kind == "addmul"
    leaf_only = fn_name == "mzd_mul_m4rm"
    if m % (128 * world) or l % (128 * world) or n % (128 * world):
        raise SystemExit("m, l, n must be multiples of 128 * gpus")
    cutoff = args.cutoff
    # C is cut into pr row-blocks x pc column-blocks (2 x world/2 from 4 ranks on, see m4ri_b200/shard.py):
    # this rank owns C[rows gr, cols gc] = A[rows gr, :] * B[:, cols gc]
    pr, pc = shard.grid_shape(world, args.grid)
    gr, gc = shard.grid_coords(rank, world, args.grid)
    r0, r1 = shard.row_blocks(m, pr)[gr]
    c0, c1 = shard.col_blocks(n, pc)[gc]
    rows, ncb = r1 - r0, c1 - c0           # this rank's block of C; A block is rows x l, B block l x ncb
    brow = shard.padded_slice_rows(l, pr)  # this rank's row-slice of B[:, cols gc] (l % (64*world) == 0: no padding)
    pitch, pitchb = l // 64, ncb // 64
    group = None
    if world > 1 and pc > 1:               # every rank creates every column group, in the same order
        groups = [dist.new_group(shard.column_group(g, world, args.grid)) for g in range(pc)]
        group = groups[gc]

    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    sh = ctypes.c_void_p(tstream.cuda_stream)

    # ---- host inputs: this rank's rows of the SEEDED global matrices (tests/harness.py: the generator the golden
    #      digests were made with), once in pageable and once in pinned memory -------------------------------------
    wc0, wc1 = c0 // 64, c1 // 64
    srcA = H.seeded_words(H.SEED_A, rows, pitch, row0=r0)
    srcB = np.ascontiguousarray(H.seeded_words(H.SEED_B, brow, n // 64, row0=gr * brow)[:, wc0:wc1])
    srcC = np.ascontiguousarray(H.seeded_words(H.SEED_C, rows, n // 64, row0=r0)[:, wc0:wc1]) if accumulate else None

    class HostSet:
        def __init__(self, pinned):
            self.pinned = pinned
            if pinned:
                self.tA = torch.empty((rows, pitch), dtype=torch.int64, pin_memory=True)
                self.tB = torch.empty((brow, pitchb), dtype=torch.int64, pin_memory=True)
                self.tC = torch.zeros((rows, pitchb), dtype=torch.int64, pin_memory=True)
                self.A, self.B, self.C = (t.numpy().view(np.uint64) for t in (self.tA, self.tB, self.tC))
            else:   # plain (malloc/mmap) memory, as mzd_init gives it
                self.A = np.empty((rows, pitch), dtype=np.uint64)
                self.B = np.empty((brow, pitchb), dtype=np.uint64)
                self.C = np.zeros((rows, pitchb), dtype=np.uint64)
            self.A[:, :] = srcA
            self.B[:, :] = srcB
            self.reset_c()
            self.mA = make_header(MzdT, self.A.ctypes.data, rows, l, pitch)
            self.mB = make_header(MzdT, self.B.ctypes.data, brow, ncb, pitchb)
            self.mC = make_header(MzdT, self.C.ctypes.data, rows, ncb, pitchb)

        def reset_c(self):
            if accumulate:
                self.C[:, :] = srcC
            else:
                self.C[:, :] = 0

    # Which leg is `e2e`.  N = 1: the drop-in call on PAGEABLE memory (what mzd_init gives a libm4ri caller), pinned beside
    # it.  N > 1: the ranks own their host buffers (this is a harness, not a drop-in call) and staging pageable rows is
    # bound by the host's memcpy rate once the per-rank compute shrinks (1.5 GiB through the CPUs per step: >= 30 ms on
    # this box whatever the GPUs do), so `e2e` is the pinned leg — the basis of the base contract — and the pageable leg
    # is reported beside it; the drop-in form of the multi-GPU product on pageable memory is `e2e_inproc`.
    kinds = []
    if not args.no_e2e:
        both = ["pageable", "pinned"] if world == 1 else ["pinned", "pageable"]
        kinds = both if not (args.pageable or args.pinned) else (["pageable"] if args.pageable else ["pinned"])
    hosts = {k: HostSet(k == "pinned") for k in kinds}
    upload_from = hosts[kinds[0]] if kinds else HostSet(False)

    # ---- device matrices (torch owns the memory; the library sees plain pointers) --------------
    tA = torch.zeros((rows, pitch), dtype=torch.int64, device="cuda")
    tBs = torch.zeros((brow, pitchb), dtype=torch.int64, device="cuda")
    tB = tBs if pr == 1 else torch.zeros((l, pitchb), dtype=torch.int64, device="cuda")
    tC = torch.zeros((rows, pitchb), dtype=torch.int64, device="cuda")
    dA = lib.m4ri_b200_dmat_wrap(tA.data_ptr(), pitch, rows, l)
    dBs = lib.m4ri_b200_dmat_wrap(tBs.data_ptr(), pitchb, brow, ncb)
    dB = lib.m4ri_b200_dmat_wrap(tB.data_ptr(), pitchb, l, ncb)
    dC = lib.m4ri_b200_dmat_wrap(tC.data_ptr(), pitchb, rows, ncb)
    lib.m4ri_b200_upload(dA, ctypes.byref(upload_from.mA), sh)
    lib.m4ri_b200_upload(dBs, ctypes.byref(upload_from.mB), sh)
    if accumulate:
        lib.m4ri_b200_upload(dC, ctypes.byref(upload_from.mC), sh)
    torch.cuda.synchronize()

    def exchange():
        if pr > 1:  # the path's one exchange step: all-gather of the row-slices of B[:, cols gc] over NVLink
            dist.all_gather_into_tensor(tB.view(-1), tBs.view(-1), group=group)

    def device_product():
        if leaf_only:
            lib.m4ri_b200_dmul_m4rm(dC, dA, dB, 0 if accumulate else 1, sh)
        else:
            lib.m4ri_b200_dmul(dC, dA, dB, cutoff, 0 if accumulate else 1, sh)

    def step_resident():
        exchange()
        device_product()

    # ---- N > 1, end to end: the K-chunk pipeline of m4ri_b200/shard.py (pipelined_product) on this rank's GPU -------
    # uploads ride their own stream, exchanges and products the compute stream, downloads a third stream; every
    # operand bit crosses PCIe once per node (row group: 1/pc of an A chunk's rows per rank, column group: 1/pr of B)
    pipe = None
    if world > 1 and args.e2e_mode == "kchunk":
        sub = args.ksub
        kc = l // (pr * sub)
        if l % (pr * sub * 128) or rows % (pc * 2 * 64):
            raise SystemExit("--ksub: K-chunks must be multiples of 128 columns and the row block divisible by 128 * pc")
        part_rows = rows // pc
        row_grp = None
        if pc > 1:      # every rank creates every row group, in the same order
            rgroups = [dist.new_group(shard.row_group(g * pc, world, args.grid)) for g in range(pr)]
            row_grp = rgroups[gr]
        up_stream, dn_stream = torch.cuda.Stream(), torch.cuda.Stream()
        uh, dh = ctypes.c_void_p(up_stream.cuda_stream), ctypes.c_void_p(dn_stream.cuda_stream)
        kw = kc // 64
        tAc = {(g, j): torch.zeros((rows, kw), dtype=torch.int64, device="cuda") for g in range(pr) for j in range(sub)}
        tBc = [torch.zeros((pr, kc, pitchb), dtype=torch.int64, device="cuda") for _ in range(sub)]
        tCe = torch.zeros((rows, pitchb), dtype=torch.int64, device="cuda")
        TAIL = 2
        wrap = lib.m4ri_b200_dmat_wrap

        class Pipe:
            def __init__(self):
                self.dA = {k: wrap(t.data_ptr(), kw, rows, kc) for k, t in tAc.items()}
                self.dApart = {k: wrap(t[gc * part_rows:].data_ptr(), kw, part_rows, kc) for k, t in tAc.items()}
                self.dAtail = {(k, i): wrap(t[i * rows // TAIL:].data_ptr(), kw, rows // TAIL, kc) for k, t in tAc.items() for i in range(TAIL)}
                self.dB = {(j, g): wrap(tBc[j][g].data_ptr(), pitchb, kc, ncb) for j in range(sub) for g in range(pr)}
                self.dC = wrap(tCe.data_ptr(), pitchb, rows, ncb)
                self.dCtail = [wrap(tCe[i * rows // TAIL:].data_ptr(), pitchb, rows // TAIL, ncb) for i in range(TAIL)]
                self.ev = {}
                self.tail_ev = [None] * TAIL

            def bind(self, hs):    # host windows of this step's operands (headers only; same words)
                self.hs = hs
                WIN = 0x4
                self.hA, self.hB, self.hC = {}, {}, []
                for g in range(pr):
                    for j in range(sub):
                        k0, _ = shard.chunk_range(l, pr, sub, g, j)
                        h = make_header(MzdT, hs.A[gc * part_rows:, k0 // 64:].ctypes.data, part_rows, kc, pitch)
                        h.flags |= WIN
                        self.hA[(g, j)] = h
                for j in range(sub):
                    h = make_header(MzdT, hs.B[j * kc:].ctypes.data, kc, ncb, pitchb)
                    h.flags |= WIN
                    self.hB[j] = h
                for i in range(TAIL):
                    h = make_header(MzdT, hs.C[i * rows // TAIL:].ctypes.data, rows // TAIL, ncb, pitchb)
                    h.flags |= WIN
                    self.hC.append(h)

            def _mark(self, key):
                e = torch.cuda.Event()
                e.record(up_stream)
                self.ev[key] = e

            def upload_c(self):
                lib.m4ri_b200_upload(self.dC, ctypes.byref(self.hs.mC), uh)
                self._mark("c")

            def upload_b(self, j):
                lib.m4ri_b200_upload(self.dB[(j, gr)], ctypes.byref(self.hB[j]), uh)
                self._mark(("b", j))

            def gather_b(self, j):
                tstream.wait_event(self.ev[("b", j)])
                dist.all_gather_into_tensor(tBc[j].view(-1), tBc[j][gr].view(-1), group=group)

            def upload_a(self, g, j):
                lib.m4ri_b200_upload(self.dApart[(g, j)], ctypes.byref(self.hA[(g, j)]), uh)
                self._mark(("a", g, j))

            def gather_a(self, g, j):
                tstream.wait_event(self.ev[("a", g, j)])
                t = tAc[(g, j)]
                dist.all_gather_into_tensor(t.view(-1), t[gc * part_rows:(gc + 1) * part_rows].view(-1), group=row_grp)

            def mul(self, g, j, clear, part):
                tstream.wait_event(self.ev[("a", g, j)])
                if g == gr:
                    tstream.wait_event(self.ev[("b", j)])
                if "c" in self.ev:
                    tstream.wait_event(self.ev["c"])
                dCp, dAp = (self.dC, self.dA[(g, j)]) if part is None else (self.dCtail[part[0]], self.dAtail[((g, j), part[0])])
                if args.chunk_levels >= 0:
                    lib.m4ri_b200_dmul_levels(dCp, dAp, self.dB[(j, g)], args.chunk_levels, 1 if clear else 0, sh)
                elif leaf_only:
                    lib.m4ri_b200_dmul_m4rm(dCp, dAp, self.dB[(j, g)], 1 if clear else 0, sh)
                else:
                    lib.m4ri_b200_dmul(dCp, dAp, self.dB[(j, g)], cutoff, 1 if clear else 0, sh)
                if part is not None:
                    e = torch.cuda.Event()
                    e.record(tstream)
                    self.tail_ev[part[0]] = e

            def download(self, part):
                dn_stream.wait_event(self.tail_ev[part[0]])
                lib.m4ri_b200_download(ctypes.byref(self.hC[part[0]]), self.dCtail[part[0]], dh)   # returns when the part is on the host

            def step(self, hs):
                if getattr(self, "hs", None) is not hs:
                    self.bind(hs)
                self.ev = {}
                up_stream.wait_stream(tstream)     # buffers of the previous step are free again
                shard.pipelined_product(rank, world, self, sub=sub, mode=args.grid, tail_parts=TAIL, accumulate=accumulate)

        pipe = Pipe()

    # ---- N > 1, end to end, default: the top Strassen level on separate quadrant buffers with transfer hooks ----------
    # (m4ri_b200_dmul_quads): a rank uploads its 1/pc row share of an A quadrant and its 1/pr row share of a B quadrant
    # when the schedule first needs them, the shares are all-gathered over NVLink inside the row / column group, and a
    # C quadrant is downloaded as soon as it is final — the multi-rank form of the library's own host path
    quads = None
    if world > 1 and args.e2e_mode == "hooks" and not leaf_only:
        from m4ri_b200 import DMatP, Hooks, HookFn
        m2, k2, n2 = rows // 2, l // 2, ncb // 2
        if rows % (2 * pc * 64) or l % (2 * pr * 128) or ncb % 256:
            raise SystemExit("hooks mode: block dimensions must split into quadrants and per-rank shares")
        if pc > 1:
            rgroups2 = [dist.new_group(shard.row_group(g * pc, world, args.grid)) for g in range(pr)]
            row_grp2 = rgroups2[gr]
        up2, dn2 = torch.cuda.Stream(), torch.cuda.Stream()
        uh2, dh2 = ctypes.c_void_p(up2.cuda_stream), ctypes.c_void_p(dn2.cuda_stream)
        qa = [torch.zeros((m2, k2 // 64), dtype=torch.int64, device="cuda") for _ in range(4)]
        qb = [torch.zeros((k2, n2 // 64), dtype=torch.int64, device="cuda") for _ in range(4)]
        qc = [torch.zeros((m2, n2 // 64), dtype=torch.int64, device="cuda") for _ in range(4)]
        wrap2 = lib.m4ri_b200_dmat_wrap
        share_a, share_b = m2 // pc, k2 // pr           # rows of a quadrant this rank uploads
        # this rank's share of B[:, cols gc] for this mode: the gr-th 1/pr of the rows of EACH row half (the resident and
        # serial paths use one contiguous row-slice instead; still 1/world of B per rank)
        srcB2 = [np.ascontiguousarray(H.seeded_words(H.SEED_B, share_b, n // 64, row0=h * k2 + gr * share_b)[:, wc0:wc1]) for h in range(2)]

        class Quads:
            def __init__(self):
                self.dA = (DMatP * 4)(*[wrap2(t.data_ptr(), k2 // 64, m2, k2) for t in qa])
                self.dB = (DMatP * 4)(*[wrap2(t.data_ptr(), n2 // 64, k2, n2) for t in qb])
                self.dC = (DMatP * 4)(*[wrap2(t.data_ptr(), n2 // 64, m2, n2) for t in qc])
                self.dAshare = [wrap2(t[gc * share_a:].data_ptr(), k2 // 64, share_a, k2) for t in qa]
                self.dBshare = [wrap2(t[gr * share_b:].data_ptr(), n2 // 64, share_b, n2) for t in qb]
                self.hooks = Hooks(HookFn(self.need_a), HookFn(self.need_b), HookFn(self.need_c), HookFn(self.done_c), None)
                self.hostB = {}
                self.error = None

            def bind(self, hs):
                self.hs = hs
                if id(hs) not in self.hostB:      # B share of this mode, in the same kind of memory as the other operands
                    if hs.pinned:
                        tb = torch.empty((2, share_b, pitchb), dtype=torch.int64, pin_memory=True)
                        arr = tb.numpy().view(np.uint64)
                        self.hostB[id(hs)] = (arr, tb)
                    else:
                        arr = np.empty((2, share_b, pitchb), dtype=np.uint64)
                        self.hostB[id(hs)] = (arr, None)
                    arr[0], arr[1] = srcB2[0], srcB2[1]
                hb = self.hostB[id(hs)][0]
                WIN = 0x4
                self.hA, self.hB, self.hC = [], [], []
                for q in range(4):
                    qr, qcol = q >> 1, q & 1
                    h = make_header(MzdT, hs.A[qr * m2 + gc * share_a:, qcol * (k2 // 64):].ctypes.data, share_a, k2, pitch)
                    h.flags |= WIN
                    self.hA.append(h)
                    h = make_header(MzdT, hb[qr][:, qcol * (n2 // 64):].ctypes.data, share_b, n2, pitchb)
                    h.flags |= WIN
                    self.hB.append(h)
                    h = make_header(MzdT, hs.C[qr * m2:, qcol * (n2 // 64):].ctypes.data, m2, n2, pitchb)
                    h.flags |= WIN
                    self.hC.append(h)

            def _guard(self, fn, q):
                try:
                    fn(q)
                except BaseException as e:   # noqa: BLE001 — an exception must not unwind through the C++ frames
                    self.error = e

            def need_a(self, _user, q):
                self._guard(self._need_a, q)

            def need_b(self, _user, q):
                self._guard(self._need_b, q)

            def need_c(self, _user, q):
                self._guard(self._need_c, q)

            def done_c(self, _user, q):
                self._guard(self._done_c, q)

            def _uploaded(self):
                e = torch.cuda.Event()
                e.record(up2)
                tstream.wait_event(e)

            def _need_a(self, q):
                lib.m4ri_b200_upload(self.dAshare[q], ctypes.byref(self.hA[q]), uh2)
                self._uploaded()
                if pc > 1:
                    dist.all_gather_into_tensor(qa[q].view(-1), qa[q][gc * share_a:(gc + 1) * share_a].view(-1), group=row_grp2)

            def _need_b(self, q):
                lib.m4ri_b200_upload(self.dBshare[q], ctypes.byref(self.hB[q]), uh2)
                self._uploaded()
                if pr > 1:
                    dist.all_gather_into_tensor(qb[q].view(-1), qb[q][gr * share_b:(gr + 1) * share_b].view(-1), group=group)

            def _need_c(self, q):
                lib.m4ri_b200_upload(self.dC[q], ctypes.byref(self.hC[q]), uh2)
                self._uploaded()

            def _done_c(self, q):
                e = torch.cuda.Event()
                e.record(tstream)
                self.order.append((q, e))

            def step(self, hs):
                if getattr(self, "hs", None) is not hs:
                    self.bind(hs)
                self.order = []
                up2.wait_stream(tstream)          # the quadrant buffers of the previous step are free again
                lib.m4ri_b200_dmul_quads(self.dC, self.dA, self.dB, cutoff, 0 if accumulate else 1, sh, ctypes.byref(self.hooks))
                if self.error is not None:
                    raise self.error
                for q, e in self.order:           # issued after the whole schedule: a download blocks this thread
                    dn2.wait_event(e)
                    lib.m4ri_b200_download(ctypes.byref(self.hC[q]), self.dC[q], dh2)

        quads = Quads()

    def step_e2e(hs):
        if world == 1:   # the drop-in call itself
            getattr(lib, fn_name)(ctypes.byref(hs.mC), ctypes.byref(hs.mA), ctypes.byref(hs.mB), cutoff)
        elif quads is not None:
            quads.step(hs)
        elif pipe is not None:
            pipe.step(hs)
        else:            # serial: upload, exchange, multiply, download one after the other (round-1 form)
            lib.m4ri_b200_upload(dA, ctypes.byref(hs.mA), sh)
            lib.m4ri_b200_upload(dBs, ctypes.byref(hs.mB), sh)
            if accumulate:
                lib.m4ri_b200_upload(dC, ctypes.byref(hs.mC), sh)
            exchange()
            device_product()
            lib.m4ri_b200_download(ctypes.byref(hs.mC), dC, sh)

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

    def all_ranks_ok(ok):
        if world == 1:
            return bool(ok)
        t = torch.tensor([1 if ok else 0], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    # inputs smaller than twice the L2 would be served from it on a repeat: flush between steps (cfg2)
    resident_bytes = (rows * pitch + l * pitchb + rows * pitchb) * 8
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if resident_bytes < 2 * L2_BYTES else None

    # ---- resident-input timing ------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.m4ri_b200_kernel_launches()
    lib.m4ri_b200_profile_begin()
    barrier()
    wall0 = time.time()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step_resident()
        e1.record()
        barrier()
        total_ms = e0.elapsed_time(e1)
    else:
        evs = []
        for _ in range(args.steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_resident()
            b.record()
            evs.append((a, b))
        barrier()
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
    wall1 = time.time()
    leaf_ms, leaf_bitops = ctypes.c_double(0), ctypes.c_double(0)
    leaf_launches = lib.m4ri_b200_profile_end(ctypes.byref(leaf_ms), ctypes.byref(leaf_bitops))
    launches = lib.m4ri_b200_kernel_launches() - launches0
    leaf_variant = lib.m4ri_b200_last_leaf_variant()      # of the timed region (later legs launch other leaves)
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    ms_step = max_over_ranks(total_ms / args.steps)
    path = lib.m4ri_b200_last_path().decode()
    total_bitops = 2.0 * m * l * n
    value = total_bitops / (ms_step * 1e-3)

    # ---- end-to-end timing (host buffers in, host buffer out) ------------------------------------
    e2e = {}
    h2d = (rows * pitch + brow * pitchb + (rows * pitchb if accumulate else 0)) * 8 * world
    d2h = rows * pitchb * 8 * world
    for k in kinds:
        hs = hosts[k]
        for _ in range(min(args.warmup, 2)):
            step_e2e(hs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e(hs)
        barrier()
        sec = max_over_ranks((time.perf_counter() - t0) / args.steps)
        e2e[k] = {"value": total_bitops / sec, "unit": UNIT, "ms_per_step": sec * 1e3, "h2d_bytes_per_step": h2d,
                  "d2h_bytes_per_step": d2h, "host_memory": k,
                  "api": f"{fn_name}(C, A, B, cutoff) on host mzd_t" if world == 1 else
                  ("per rank: top Strassen level on quadrant buffers (m4ri_b200_dmul_quads), quadrant shares uploaded, "
                   "all-gathered over NVLink in row/column groups and downloaded as the schedule needs them" if quads is not None else
                   "upload + all_gather + m4ri_b200_dmul + download per rank, one after the other" if pipe is None else
                   f"K-chunk pipeline per rank ({pr * args.ksub} chunks of {l // (pr * args.ksub)} columns: uploads, NVLink "
                   f"all-gathers in row/column groups, products and downloads overlapped; m4ri_b200/shard.py)")}

    # ---- verification (after the timed legs; default on) -------------------------------------------------
    verified = None
    if not args.no_check:
        verified = {}
        # (1) bit-exact against the unmodified reference: digests of its result blocks for the same seeded inputs
        digest_ok = None
        gpath = os.path.join(ROOT, "tests", "golden", "large_golden.json")
        case = None
        if os.path.exists(gpath):
            with open(gpath) as f:
                case = json.load(f)["cases"].get(golden_name)
            if case and (case["m"], case["l"], case["n"]) != (m, l, n):
                case = None
        if case is not None and kinds:
            hs = hosts[kinds[0]]
            hs.reset_c()
            step_e2e(hs)
            torch.cuda.synchronize()
            br, bw = m // H.LARGE_BLOCK_ROWS, (n // 64) // H.LARGE_BLOCK_COLS
            digest_ok = True if rows % br == 0 and pitchb % bw == 0 else None    # None: this partition does not tile the blocks
            if digest_ok:
                for i in range(rows // br):
                    for j in range(pitchb // bw):
                        got = H.block_digest(hs.C[i * br:(i + 1) * br, j * bw:(j + 1) * bw])
                        digest_ok = digest_ok and got == case["C_blocks"][r0 // br + i][wc0 // bw + j]
            if digest_ok is not None:
                digest_ok = all_ranks_ok(digest_ok)
        verified["reference_digest"] = digest_ok
        verified["reference_digest_source"] = ("tests/golden/large_golden.json:" + golden_name) if case is not None else None
        # (2) on the device, any size: the exchange delivered the pieces in order, and Freivalds on this rank's block
        #     with 128 random vectors: C_blk * X == A_blk * (B_colblk * X) [^ C_in * X]   (error probability 2^-128)
        if accumulate:
            upload_from.reset_c()
            lib.m4ri_b200_upload(dC, ctypes.byref(upload_from.mC), sh)
        step_resident()
        torch.cuda.synchronize()
        ok = True
        if pr > 1:
            mine = tBs.sum().reshape(1)
            sums = torch.empty(pr, dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(sums, mine, group=group)
            ok = bool(torch.equal(tB.view(pr, brow, pitchb).sum(dim=(1, 2)), sums))
        tX = torch.randint(-2**62, 2**62, (ncb, 2), dtype=torch.int64, device="cuda")
        tY = torch.zeros((l, 2), dtype=torch.int64, device="cuda")
        tZ = torch.zeros((rows, 2), dtype=torch.int64, device="cuda")
        tW = torch.zeros((rows, 2), dtype=torch.int64, device="cuda")
        dX, dY = lib.m4ri_b200_dmat_wrap(tX.data_ptr(), 2, ncb, 128), lib.m4ri_b200_dmat_wrap(tY.data_ptr(), 2, l, 128)
        dZ, dW = lib.m4ri_b200_dmat_wrap(tZ.data_ptr(), 2, rows, 128), lib.m4ri_b200_dmat_wrap(tW.data_ptr(), 2, rows, 128)
        torch.cuda.synchronize()
        lib.m4ri_b200_dmul_m4rm(dY, dB, dX, 1, sh)
        lib.m4ri_b200_dmul_m4rm(dZ, dA, dY, 1, sh)
        if accumulate:   # Z = A (B X) ^ C_in X
            tCin = torch.from_numpy(srcC.view(np.int64)).to("cuda")
            dCin = lib.m4ri_b200_dmat_wrap(tCin.data_ptr(), pitchb, rows, ncb)
            torch.cuda.synchronize()
            lib.m4ri_b200_dmul_m4rm(dZ, dCin, dX, 0, sh)
        lib.m4ri_b200_dmul_m4rm(dW, dC, dX, 1, sh)
        torch.cuda.synchronize()
        ok = ok and bool(torch.equal(tZ, tW)) and bool(tZ.any())
        verified["freivalds"] = all_ranks_ok(ok)
        verified["freivalds_vectors"] = 128
        print(f"[check] rank {rank}: rows {r0}:{r1} cols {c0}:{c1} digest={digest_ok} freivalds={ok}", file=sys.stderr, flush=True)
        if verified["freivalds"] is False or verified["reference_digest"] is False:
            print(json.dumps({"error": "verification failed", "verified": verified}), flush=True)
            raise SystemExit(3)

    if args.verify:   # small-n correctness of this rank's block against the oracle (never at full size)
        hs = hosts[kinds[0]]
        hs.reset_c()
        step_e2e(hs)
        col_b = torch.empty((l, pitchb), dtype=torch.int64)
        col_b.copy_(tB)
        Ao, Bo, Co = H.new(rows, l), H.new(l, ncb), H.new(rows, ncb)
        H.storage(Ao)[:, :] = hs.A
        H.storage(Bo)[:, :] = col_b.numpy().view(np.uint64)
        if accumulate:
            H.storage(Co)[:, :] = srcC
        want = H.oracle().orc_addmul(Co, Ao, Bo, 0)
        ok = bool(np.array_equal(H.storage(want), hs.C))
        print(f"[verify] rank {rank}: rows {r0}:{r1} cols {c0}:{c1} {'OK' if ok else 'MISMATCH'}", file=sys.stderr, flush=True)
        if not ok:
            raise SystemExit(3)

    # ---- N > 1: the same product through the in-process C-ABI entry point (mzd_mul_mp / mzd_addmul_mp over all GPUs of
    #      the box, csrc/multi.cu) on rank 0, with full pageable host matrices; the other ranks wait -----------------
    inproc = None
    if world > 1 and not args.no_inproc and not leaf_only:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        # the other ranks wait on the HOST (key in the rendezvous store): an NCCL barrier would park a kernel on their GPUs,
        # which are about to be driven by rank 0
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            if prev_affinity is not None:        # this leg drives every GPU of the box from one process: all CPUs again
                os.sched_setaffinity(0, prev_affinity)
            fullA = np.empty((m, l // 64), dtype=np.uint64)
            fullB = np.empty((l, n // 64), dtype=np.uint64)
            fullC = np.zeros((m, n // 64), dtype=np.uint64)
            fullA[:, :] = H.seeded_words(H.SEED_A, m, l // 64)
            fullB[:, :] = H.seeded_words(H.SEED_B, l, n // 64)
            fA = make_header(MzdT, fullA.ctypes.data, m, l, l // 64)
            fB = make_header(MzdT, fullB.ctypes.data, l, n, n // 64)
            fC = make_header(MzdT, fullC.ctypes.data, m, n, n // 64)
            mp_fn = lib.mzd_addmul_mp if accumulate else lib.mzd_mul_mp
            lib.m4ri_b200_set_num_devices(world)

            def reset_full_c():
                if accumulate:
                    fullC[:, :] = H.seeded_words(H.SEED_C, m, n // 64)

            reset_full_c()
            mp_fn(ctypes.byref(fC), ctypes.byref(fA), ctypes.byref(fB), cutoff)     # warm-up: contexts, workspaces, peer access
            runs = max(1, min(args.steps, 3))
            t0 = time.perf_counter()
            for _ in range(runs):
                mp_fn(ctypes.byref(fC), ctypes.byref(fA), ctypes.byref(fB), cutoff)
            sec = (time.perf_counter() - t0) / runs
            mp_path = lib.m4ri_b200_last_path().decode()
            ok = None
            gpath = os.path.join(ROOT, "tests", "golden", "large_golden.json")
            if os.path.exists(gpath):
                with open(gpath) as f:
                    case = json.load(f)["cases"].get(golden_name)
                if case and (case["m"], case["l"], case["n"]) == (m, l, n):
                    reset_full_c()
                    mp_fn(ctypes.byref(fC), ctypes.byref(fA), ctypes.byref(fB), cutoff)
                    ok = H.large_block_digests(fullC) == case["C_blocks"]
            lib.m4ri_b200_set_num_devices(1)
            inproc = {"value": total_bitops / sec, "unit": UNIT, "ms_per_step": sec * 1e3, "host_memory": "pageable",
                      "h2d_bytes_per_step": (m * l + l * n + (m * n if accumulate else 0)) // 8, "d2h_bytes_per_step": m * n // 8,
                      "api": ("mzd_addmul_mp" if accumulate else "mzd_mul_mp") + "(C, A, B, cutoff) on host mzd_t, one process driving "
                             f"{world} GPUs (csrc/multi.cu)", "path": mp_path, "reference_digest": ok, "runs": runs}
            del fullA, fullB, fullC
            store.set("m4ri_b200_inproc_done", "1")
        else:
            import datetime
            store.wait(["m4ri_b200_inproc_done"], datetime.timedelta(seconds=900))

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
    if leaf_variant == 2:
        smem_bytes_per_bitop = (4096 * 8 * 16 + 2048 * 16 + 4096 * 4 + 32 * 32) / (2.0 * 32 * 4096 * 256)
        lookup_bytes_per_bitop = (4096 * 8 * 16) / (2.0 * 32 * 4096 * 256)
    else:
        smem_bytes_per_bitop = (1024 * 2 * 128 + 512 * 128 + 1024 * 2 + 16 * 128) / (2.0 * 16 * 1024 * 1024)
        lookup_bytes_per_bitop = (1024 * 2 * 128) / (2.0 * 16 * 1024 * 1024)
    smem_peak_gbs = 128.0 * 148 * sm_max_mhz * 1e6 / 1e9
    smem_achieved_gbs = leaf_rate * smem_bytes_per_bitop / 1e9
    leaf_dims = None
    try:
        lv = int(path.split(":")[1]) if ":" in path else 0
        leaf_dims = (rows >> lv, l >> lv, ncb >> lv)
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
    tensor_roofline = None
    if leaf_variant == 3:
        # The tensor-core leaf (tc_leaf.cu): every GF(2) multiply-add is one e2m1 multiply-add of tcgen05.mma kind::mxf4,
        # so the leaf's bit-ops ARE its tensor flops.  Peak: 4 x the measured dense bf16 figure of MEASURED_PEAKS.json
        # (e2m1 block-scaled runs at four times the bf16 rate; the burst figure, as the leaf is timed launch by launch);
        # beside it the mxf4 instruction-issue ceiling measured by tools/tc/tc05_probe.cu on this pool (8.91e15).
        bf16_burst, bf16_sust = 1590.0, None
        ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(ppath):
            with open(ppath) as f:
                pk = json.load(f)
            bf16_burst, bf16_sust = float(pk.get("bf16_tflops", bf16_burst)), pk.get("bf16_tflops_sustained")
        # peak: the e2m1 block-scaled instruction rate measured on this pool by the hand-written probe
        # (tools/tc/tc05_probe.cu, profiles/r02/tc05_probe.jsonl: 8.912e15 multiply-add pairs x 2 per second at 1965 MHz;
        # nominal dense fp4: 9 PFLOP/s).  4 x the MEASURED_PEAKS bf16 figure — a cuBLAS GEMM, 0.74 of ITS pipe rate —
        # would put this kernel above 1, so it is reported beside, not as the denominator.
        peak_tf = 8912.0
        tensor_roofline = {
            "kernel": "tc_leaf2_kernel (+ tc_expand_a/bt, tc_zero_c pre-pass)", "bound": "tensor", "unit": "TFLOP/s",
            "achieved": leaf_rate / 1e12, "peak": peak_tf, "frac": leaf_rate / 1e12 / peak_tf,
            "peak_source": "mxf4 (e2m1, block-scaled) tcgen05.mma issue ceiling measured on this pool: tools/tc/tc05_probe.cu, "
                           "profiles/r02/tc05_probe.jsonl (nominal dense fp4 9000)",
            "frac_of_4x_measured_bf16_burst": leaf_rate / 1e12 / (4.0 * bf16_burst),
            "frac_of_4x_measured_bf16_sustained": (leaf_rate / 1e12 / (4.0 * float(bf16_sust))) if bf16_sust else None,
            "measured_bf16_tflops": {"burst": bf16_burst, "sustained": bf16_sust, "source": peak_src},
            "includes": "operand expansion to e2m1 and the clearing of C (three streaming kernels per launch, 9 % of it) and "
                        "the parity epilogue; the main kernel alone: 0.89 of the ceiling (DESIGN.md 4.1c)",
        }
    roofline = {
        "kernel": "m4rm_leaf2_kernel" if leaf_variant == 2 else "m4rm_streamk_kernel", "bound": "smem", "unit": "GB/s",
        "algorithmic_smem_bytes_per_bitop": smem_bytes_per_bitop,
        "achieved": smem_achieved_gbs, "peak": smem_peak_gbs, "frac": smem_achieved_gbs / smem_peak_gbs,
        # SURVEY §8d's own roof counts the table LOOKUPS only (16*k*128 B/clk/SM at k = 8): the stricter fraction
        "frac_lookup_only": leaf_rate * lookup_bytes_per_bitop / 1e9 / smem_peak_gbs,
        "peak_source": f"128 B/clk/SM x 148 SMs x {sm_max_mhz:.0f} MHz (clocks.max.sm, {peak_src})",
        "traffic": traffic,
        "leaf_launches": int(leaf_launches), "leaf_avg_ms": leaf_avg_ms, "leaf_dims": leaf_dims,
        "products_per_launch": products_per_launch,
        "leaf_bitops_per_s": leaf_rate, "leaf_share_of_step": leaf_ms.value / (ms_step * args.steps),
        "hbm": {"bound": "hbm", "unit": "GB/s", "achieved": hbm_bytes / (leaf_avg_ms * 1e-3) / 1e9 if leaf_avg_ms else 0.0,
                "peak": hbm_peak, "frac": (hbm_bytes / (leaf_avg_ms * 1e-3) / 1e9) / hbm_peak if leaf_avg_ms else 0.0,
                "algorithmic_bytes_per_launch": hbm_bytes, "peak_source": peak_src},
    }

    if tensor_roofline is not None:
        for k in ("traffic", "leaf_launches", "leaf_avg_ms", "leaf_dims", "products_per_launch", "leaf_bitops_per_s",
                  "leaf_share_of_step", "hbm"):
            tensor_roofline[k] = roofline[k]
        tensor_roofline["traffic"] = None
        roofline = tensor_roofline

    # ---- the HBM-bound kernel of the path: the device _mzd_add (C = A ^ B) on square operands ----------
    add_roofline = None
    if world == 1 and m == l == n:
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
                        "algorithmic_bytes_per_launch": add_bytes, "peak_source": peak_src,
                        "note": "operands in L2 at this size" if add_bytes < 2 * L2_BYTES else None}

    # ---- CPU baseline (rank 0, N = 1 only; bounded sample) --------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sample = workload_of(args)[6]
        sec, desc = time_reference(kind, sample, 2, 1)
        cpu = dict(desc, value=2.0 * sample[0] * sample[1] * sample[2] / sec, unit=UNIT)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args, world), "dims": [m, l, n],
                   "cutoff": cutoff or lib.m4ri_b200_get_default_cutoff(), "path": path,
                   "leaf": {1: "M4RM, 1024 x 1024-bit tiles (CUDA cores)", 2: "M4RM, 4096 x 256-bit tall tiles (CUDA cores)",
                            3: "tcgen05.mma kind::mxf4 on e2m1-expanded bits, f32 accumulators in TMEM, parity epilogue "
                               "(exact: sums < 2^24)"}.get(leaf_variant, str(leaf_variant)),
                   "parallelism": (f"row-block x{world}" if pc == 1 else f"C blocks {pr} x {pc} (row-blocks x column-blocks)") +
                                  ("" if world == 1 else ", NCCL all-gather of B per step" if pc == 1 else
                                   f", NCCL all-gather of B's column block inside each column group ({pr} ranks) per step"),
                   "local_product": [rows, l, ncb],
                   "cpu_binding": binding,
                   "l2": ("inputs (%d MiB per rank) exceed the 126 MB L2; no flush needed" % (resident_bytes >> 20)) if flush is None
                         else "inputs fit the L2: a 256 MiB buffer is written between timed steps, steps timed one by one"},
        "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks,
    }
    if kinds:
        line["e2e"] = e2e[kinds[0]]
        for k in kinds[1:]:
            line["e2e_" + k] = e2e[k]
    if inproc is not None:
        line["e2e_inproc"] = inproc
    if verified is not None:
        line["verified"] = verified
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
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--size", dest="n", type=int, default=0, help="n of an n x n x n product instead of the workload's size")
    ap.add_argument("--cutoff", type=int, default=0)
    ap.add_argument("--ref-sample", type=int, default=0, help="reference arm / cpu_baseline: sample size (default 32768 for cfg3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end legs (profiling runs)")
    ap.add_argument("--pageable", action="store_true", help="only the pageable end-to-end leg")
    ap.add_argument("--pinned", action="store_true", help="only the pinned end-to-end leg")
    ap.add_argument("--grid", default="auto", choices=["auto", "rows"],
                    help="C partition over ranks: 'rows' = row-blocks only; 'auto' = 2 row-blocks x world/2 column blocks from 4 ranks on")
    ap.add_argument("--e2e-mode", default="hooks", choices=["hooks", "kchunk", "serial"],
                    help="N > 1 end to end: 'hooks' = quadrants of the top Strassen level uploaded / all-gathered / downloaded as the "
                         "schedule needs them; 'kchunk' = K-chunk pipeline (m4ri_b200/shard.py); 'serial' = one after the other")
    ap.add_argument("--ksub", type=int, default=1, help="N > 1 end to end: sub-chunks per B row-slice (K-chunks = pr * ksub)")
    ap.add_argument("--chunk-levels", type=int, default=-1, help="N > 1 end to end: Strassen levels of a chunk product (-1: library rule)")
    ap.add_argument("--no-inproc", action="store_true", help="N > 1: skip the in-process mzd_mul_mp leg on rank 0")
    ap.add_argument("--no-check", action="store_true", help="skip the verification after the timed legs")
    ap.add_argument("--verify", action="store_true", help="check this rank's C block against the oracle (small --size only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
