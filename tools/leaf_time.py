"""Time the leaf (or Strassen) on device-resident data with CUDA events on torch's stream."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import m4ri_b200  # noqa: E402

lib = m4ri_b200.load_library()
torch.cuda.init()
tstream = torch.cuda.Stream()          # a real (non-default) stream: handle 0 would mean "library stream"
torch.cuda.set_stream(tstream)
stream = tstream.cuda_stream


def dev_random(rows, cols):
    if cols % 128:
        raise SystemExit("use cols % 128 == 0 here")
    pitch = cols // 64
    t = torch.randint(-2**62, 2**62, (rows, pitch), dtype=torch.int64, device="cuda")
    t ^= torch.randint(-2**62, 2**62, (rows, pitch), dtype=torch.int64, device="cuda") << 2
    return t, lib.m4ri_b200_dmat_wrap(t.data_ptr(), pitch, rows, cols)


def run(m, l, n, cutoff, iters=5):
    tA, dA = dev_random(m, l)
    tB, dB = dev_random(l, n)
    tC, dC = dev_random(m, n)

    def go():
        if cutoff < 0:
            lib.m4ri_b200_dmul_m4rm(dC, dA, dB, 1, stream)
        elif cutoff < 16:          # small values mean "this many Strassen levels"
            lib.m4ri_b200_dmul_levels(dC, dA, dB, cutoff, 1, stream)
        else:
            lib.m4ri_b200_dmul(dC, dA, dB, cutoff, 1, stream)

    for _ in range(2):
        go()
    torch.cuda.synchronize()
    import ctypes
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib.m4ri_b200_profile_begin()
    e0.record()
    for _ in range(iters):
        go()
    e1.record()
    torch.cuda.synchronize()
    lms, lops = ctypes.c_double(0), ctypes.c_double(0)
    nl = lib.m4ri_b200_profile_end(ctypes.byref(lms), ctypes.byref(lops))
    ms = e0.elapsed_time(e1) / iters
    print(f"{m}x{l}x{n} cutoff={cutoff} path={lib.m4ri_b200_last_path().decode()} "
          f"{ms:.3f} ms {2.0*m*l*n/ms/1e9:.1f} Tbitops/s | leaf launches/iter {nl//iters} "
          f"leaf ms/iter {lms.value/iters:.3f} leaf rate {lops.value/lms.value/1e9:.1f} T/s", flush=True)


if __name__ == "__main__":
    for spec in sys.argv[1:]:
        m, l, n, c = (int(x) for x in spec.split(","))
        run(m, l, n, c)
