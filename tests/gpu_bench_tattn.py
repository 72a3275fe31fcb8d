"""Temporal attention at the level-0..2 shapes of the T = 16 and T = 64 configs: achieved HBM GB/s (algorithmic bytes =
4 x rows x C x 2 B: q, k, v read once, o written once) for the tensor-core kernels and the generic CUDA-core kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mudg_b200._lib import test_lib, check, ptr, cur_stream   # noqa: E402
import ctypes                                                   # noqa: E402

SHAPES = [("T16 L0", 2, 16, 9216, 5), ("T16 L1", 2, 16, 2304, 10), ("T32 L0", 2, 32, 9216, 5), ("T64 L0", 2, 64, 9216, 5),
          ("T64 L1", 2, 64, 2304, 10), ("T64 L2", 2, 64, 576, 20), ("T64 init", 1, 64, 9216, 8)]


def main():
    L = test_lib()
    print(f"{'shape':10s} {'mma us':>9s} {'GB/s':>7s} {'generic us':>11s} {'GB/s':>7s}")
    for name, B, T, HW, heads in SHAPES:
        C = heads * 64
        qkv = torch.randn(B, T, HW, 3 * C, device="cuda").half()
        O = torch.empty(B, T, HW, C, device="cuda").half()
        nbytes = 4.0 * B * T * HW * C * 2
        row = f"{name:10s}"
        for generic in (0, 1):
            check(L.mudg_test_set_knob(b"tattn_generic", generic))

            def run():
                check(L.mudg_test_temporal_attn(ptr(qkv), ptr(O), B, T, HW, heads, ctypes.c_float(0.125), cur_stream()))
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / reps * 1e3
            row += f" {us:9.1f} {nbytes / us / 1e3:7.0f}" + ("  " if not generic else "")
        check(L.mudg_test_set_knob(b"reset", 0))
        print(row, flush=True)


if __name__ == "__main__":
    main()
