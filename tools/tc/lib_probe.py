"""Tensor-core question, measured (SURVEY §7 step 7 / VERDICT r1 #6): how fast could GF(2) products run on
the 5th-gen tensor cores if the bit-packed operands were expanded to 8-bit / 4-bit elements?

Library GEMMs only (cuBLASLt through torch) — a MEASUREMENT of the ceiling a hand-written tcgen05 kernel could
approach, never part of the product: C = A*B over the integers on 0/1 operands, GF(2) result = LSB.
  int8  : torch._int_mm           (exact int32 accumulate)
  fp8   : torch._scaled_mm e4m3   (fp32 accumulate; 0/1 exact, sums exact below 2^24 if the accumulator keeps them)
  fp4   : torch._scaled_mm e2m1x2 with e8m0 block scales = 1.0 (MXFP4; kind::mxf4 on sm_100a)
Each line: nominal bit-ops/s = 2*n^3 / t (the same W as the bench metric), exactness of the LSB against int8.
"""
import json
import sys
import time

import torch


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda")
    sizes = [int(a) for a in sys.argv[1:]] or [8192, 16384]
    for n in sizes:
        g = torch.Generator(device=dev).manual_seed(n)
        a_bits = torch.randint(0, 2, (n, n), device=dev, dtype=torch.int8, generator=g)
        b_bits = torch.randint(0, 2, (n, n), device=dev, dtype=torch.int8, generator=g)   # stored as B^T: [n_out, k]
        ops = 2.0 * n ** 3
        ref_lsb = None
        # ---- int8 ----
        try:
            bt = b_bits.t()          # column-major view, what cuBLASLt wants for the TN kernel
            ms = timeit(lambda: torch._int_mm(a_bits, bt))
            c = torch._int_mm(a_bits, bt)
            ref_lsb = (c & 1).to(torch.uint8)
            print(json.dumps({"probe": "cublasLt int8 (torch._int_mm)", "n": n, "ms": ms, "bitops_per_s": ops / (ms * 1e-3)}), flush=True)
            del c
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"probe": "int8", "n": n, "error": repr(e)[:300]}), flush=True)
        # ---- fp8 e4m3 ----
        try:
            a8 = a_bits.to(torch.float8_e4m3fn)
            b8 = b_bits.to(torch.float8_e4m3fn)
            one = torch.ones((), device=dev, dtype=torch.float32)
            for out_dtype in (torch.float32, torch.bfloat16):
                try:
                    fn = lambda: torch._scaled_mm(a8, b8.t(), scale_a=one, scale_b=one, out_dtype=out_dtype)  # noqa: E731
                    ms = timeit(fn)
                    c = fn()
                    exact = None
                    if ref_lsb is not None and out_dtype == torch.float32:
                        exact = bool(torch.equal((c.to(torch.int64) & 1).to(torch.uint8), ref_lsb))
                    print(json.dumps({"probe": "cublasLt fp8 e4m3 (torch._scaled_mm)", "out": str(out_dtype), "n": n, "ms": ms,
                                      "bitops_per_s": ops / (ms * 1e-3), "lsb_exact_vs_int8": exact}), flush=True)
                    del c
                except Exception as e:  # noqa: BLE001
                    print(json.dumps({"probe": "fp8", "out": str(out_dtype), "n": n, "error": repr(e)[:300]}), flush=True)
            del a8, b8
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"probe": "fp8", "n": n, "error": repr(e)[:300]}), flush=True)
        # ---- fp4 e2m1 (two per byte), e8m0 scales of 1.0 per 32 elements ----
        try:
            # e2m1 code of 1.0 is 0b0010, of 0.0 is 0b0000; element 2j in the low nibble
            def pack4(x):
                x = x.to(torch.uint8) * 2
                return (x[:, 0::2] | (x[:, 1::2] << 4)).contiguous().view(torch.float4_e2m1fn_x2)
            a4, b4 = pack4(a_bits), pack4(b_bits)

            def scales(rows):
                r = (rows + 127) // 128 * 128
                c = ((n // 32) + 3) // 4 * 4
                return torch.full((r * c,), 127, device=dev, dtype=torch.uint8).view(torch.float8_e8m0fnu)   # 2^0
            sa, sb = scales(n), scales(n)
            import torch.nn.functional as F
            from torch.nn.functional import ScalingType, SwizzleType
            sa2, sb2 = sa.view(-1, n // 32), sb.view(-1, n // 32)
            variants = {
                "F.scaled_mm 1x32 swizzled flat": lambda od: F.scaled_mm(a4, b4.t(), sa, ScalingType.BlockWise1x32, sb, ScalingType.BlockWise1x32,
                                                                         swizzle_a=SwizzleType.SWIZZLE_32_4_4, swizzle_b=SwizzleType.SWIZZLE_32_4_4, output_dtype=od),
                "F.scaled_mm 1x32 swizzled 2d": lambda od: F.scaled_mm(a4, b4.t(), sa2, ScalingType.BlockWise1x32, sb2, ScalingType.BlockWise1x32,
                                                                       swizzle_a=SwizzleType.SWIZZLE_32_4_4, swizzle_b=SwizzleType.SWIZZLE_32_4_4, output_dtype=od),
                "_scaled_mm 2d scales": lambda od: torch._scaled_mm(a4, b4.t(), scale_a=sa2, scale_b=sb2, out_dtype=od),
            }
            done = set()
            for vname, call in variants.items():
                for out_dtype in (torch.float32, torch.bfloat16):
                    if out_dtype in done:
                        continue
                    try:
                        fn = lambda: call(out_dtype)  # noqa: E731
                        ms = timeit(fn)
                        c = fn()
                        exact = None
                        if ref_lsb is not None and out_dtype == torch.float32:
                            exact = bool(torch.equal((c.to(torch.int64) & 1).to(torch.uint8), ref_lsb))
                        print(json.dumps({"probe": "cublasLt mxfp4 e2m1 (" + vname + ")", "out": str(out_dtype), "n": n, "ms": ms,
                                          "bitops_per_s": ops / (ms * 1e-3), "lsb_exact_vs_int8": exact,
                                          "max": float(c.max())}), flush=True)
                        done.add(out_dtype)
                        del c
                    except Exception as e:  # noqa: BLE001
                        print(json.dumps({"probe": "fp4 " + vname, "out": str(out_dtype), "n": n, "error": repr(e)[:400]}), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"probe": "fp4", "n": n, "error": repr(e)[:300]}), flush=True)
        # ---- the expansion the tensor path would need (bit-packed -> one byte per element), HBM-bound ----
        try:
            packed = torch.randint(-2**62, 2**62, (n, n // 64), device=dev, dtype=torch.int64, generator=g)
            shifts = torch.arange(64, device=dev, dtype=torch.int64)
            ms = timeit(lambda: ((packed.unsqueeze(-1) >> shifts) & 1).to(torch.int8), iters=3, warm=1)
            print(json.dumps({"probe": "torch expansion bits->int8 (unfused, upper bound on cost)", "n": n, "ms": ms}), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"probe": "expand", "n": n, "error": repr(e)[:300]}), flush=True)
        del a_bits, b_bits
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
