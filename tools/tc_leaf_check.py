"""Experimental tensor-core leaf (m4ri_b200_dmul_tc) against the M4RM leaf on device-resident random matrices: bit-exact?
and how fast."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import m4ri_b200  # noqa: E402

lib = m4ri_b200.load_library()
torch.cuda.init()
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
sh = ctypes.c_void_p(ts.cuda_stream)


def rnd(rows, cols):
    t = torch.randint(-2**62, 2**62, (rows, cols // 64), dtype=torch.int64, device="cuda")
    t ^= torch.randint(-2**62, 2**62, (rows, cols // 64), dtype=torch.int64, device="cuda") << 2
    return t, lib.m4ri_b200_dmat_wrap(t.data_ptr(), cols // 64, rows, cols)


for spec in sys.argv[1:] or ["256,128,256", "1024,1024,1024", "4096,4096,4096"]:
    m, l, n = (int(x) for x in spec.split(","))
    tA, dA = rnd(m, l)
    tB, dB = rnd(l, n)
    tBt = torch.zeros((n, l // 64), dtype=torch.int64, device="cuda")
    dBt = lib.m4ri_b200_dmat_wrap(tBt.data_ptr(), l // 64, n, l)
    tC1 = torch.zeros((m, n // 64), dtype=torch.int64, device="cuda")
    tC2 = torch.full((m, n // 64), -1, dtype=torch.int64, device="cuda")
    dC1 = lib.m4ri_b200_dmat_wrap(tC1.data_ptr(), n // 64, m, n)
    dC2 = lib.m4ri_b200_dmat_wrap(tC2.data_ptr(), n // 64, m, n)
    lib.m4ri_b200_dmul_m4rm(dC1, dA, dB, 1, sh)
    lib.m4ri_b200_dtranspose(dBt, dB, sh)
    lib.m4ri_b200_dmul_tc(dC2, dA, dBt, 1, sh)
    torch.cuda.synchronize()
    same = bool(torch.equal(tC1, tC2))
    lib.m4ri_b200_dmul_tc(dC2, dA, dBt, 0, sh)          # accumulate: C ^= A*B -> zero
    torch.cuda.synchronize()
    zero = not bool(tC2.any())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lib.m4ri_b200_dmul_tc(dC2, dA, dBt, 1, sh)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    if l % 1024 == 0:
        tC3 = torch.full((m, n // 64), -1, dtype=torch.int64, device="cuda")
        dC3 = lib.m4ri_b200_dmat_wrap(tC3.data_ptr(), n // 64, m, n)
        lib.m4ri_b200_dmul_tc2(dC3, dA, dB, sh)
        torch.cuda.synchronize()
        same2 = bool(torch.equal(tC1, tC3))
        if not same2:
            diff = (tC1 ^ tC3) != 0
            rows = diff.any(dim=1).nonzero().flatten()[:8].tolist()
            cols = diff.any(dim=0).nonzero().flatten()[:8].tolist()
            print("   tc2 differs: words", int(diff.sum()), "of", diff.numel(), "first rows", rows, "first word cols", cols)
        e0.record()
        for _ in range(5):
            lib.m4ri_b200_dmul_tc2(dC3, dA, dB, sh)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 5
        e0.record()
        for _ in range(5):
            lib.m4ri_b200_dmul_m4rm(dC1, dA, dB, 1, sh)
        e1.record()
        torch.cuda.synchronize()
        ms1 = e0.elapsed_time(e1) / 5
        print(f"tc2 leaf {m}x{l}x{n}: equal to M4RM {same2}, {ms2:.3f} ms incl. expansion = {2.0*m*l*n/ms2/1e9:.1f} Tbitops/s"
              f"  (M4RM leaf {ms1:.3f} ms = {2.0*m*l*n/ms1/1e9:.1f})", flush=True)
    print(f"tc leaf {m}x{l}x{n}: equal to M4RM {same}, accumulate-to-zero {zero}, {ms:.3f} ms = {2.0*m*l*n/ms/1e9:.1f} Tbitops/s", flush=True)
