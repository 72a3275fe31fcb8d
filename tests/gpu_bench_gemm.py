"""Micro-benchmark of the tap-GEMM backends over the layer shapes of the MDM1024 CFG forward (N=2, T=16)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mudg_b200._lib import test_lib, check, ptr, cur_stream   # noqa: E402

# (name, B, T, H, W, Cin, N, mode, res, geglu)
SHAPES = [
    ("conv3 L0 320->320", 1, 32, 72, 128, 320, 320, 1, 1, 0),
    ("conv3 L0 960->320", 1, 32, 72, 128, 960, 320, 1, 0, 0),
    ("conv3 L1 640->640", 1, 32, 36, 64, 640, 640, 1, 1, 0),
    ("conv3 L2 1280->1280", 1, 32, 18, 32, 1280, 1280, 1, 1, 0),
    ("conv3 L3 1280->1280", 1, 32, 9, 16, 1280, 1280, 1, 1, 0),
    ("conv3 L3 2560->1280", 1, 32, 9, 16, 2560, 1280, 1, 0, 0),
    ("tconv L0 320", 2, 16, 72, 128, 320, 320, 2, 0, 0),
    ("tconv L1 640", 2, 16, 36, 64, 640, 640, 2, 0, 0),
    ("tconv L2 1280", 2, 16, 18, 32, 1280, 1280, 2, 0, 0),
    ("tconv L3 1280", 2, 16, 9, 16, 1280, 1280, 2, 1, 0),
    ("lin L0 320->320 +res", 1, 1, 1, 294912, 320, 320, 0, 1, 0),
    ("lin L0 320->960 qkv", 1, 1, 1, 294912, 320, 960, 0, 0, 0),
    ("lin L0 320->2560 geglu", 1, 1, 1, 294912, 320, 2560, 0, 0, 1),
    ("lin L0 1280->320 +res", 1, 1, 1, 294912, 1280, 320, 0, 1, 0),
    ("lin L1 640->1920 qkv", 1, 1, 1, 73728, 640, 1920, 0, 0, 0),
    ("lin L1 640->5120 geglu", 1, 1, 1, 73728, 640, 5120, 0, 0, 1),
    ("lin L1 2560->640 +res", 1, 1, 1, 73728, 2560, 640, 0, 1, 0),
    ("lin L2 1280->3840 qkv", 1, 1, 1, 18432, 1280, 3840, 0, 0, 0),
    ("lin L2 1280->10240 geglu", 1, 1, 1, 18432, 1280, 10240, 0, 0, 1),
    ("lin L2 5120->1280 +res", 1, 1, 1, 18432, 5120, 1280, 0, 1, 0),
    ("lin L3 1280->1280", 1, 1, 1, 4608, 1280, 1280, 0, 1, 0),
]


# MDM512 (config 2, N = 1, T = 16, 40x64 latent): mid-size problems around the single-CTA / CTA-pair switch-over
SHAPES_512 = [
    ("512 conv3 L0 320->320", 1, 16, 40, 64, 320, 320, 1, 1, 0),
    ("512 tconv L0 320", 1, 16, 40, 64, 320, 320, 2, 0, 0),
    ("512 lin L0 320->320 +res", 1, 1, 1, 40960, 320, 320, 0, 1, 0),
    ("512 lin L0 320->960 qkv", 1, 1, 1, 40960, 320, 960, 0, 0, 0),
    ("512 lin L0 320->2560 geglu", 1, 1, 1, 40960, 320, 2560, 0, 0, 1),
    ("512 lin L0 1280->320 +res", 1, 1, 1, 40960, 1280, 320, 0, 1, 0),
    ("512 conv3 L1 640->640", 1, 16, 20, 32, 640, 640, 1, 1, 0),
    ("512 lin L1 640->1920 qkv", 1, 1, 1, 10240, 640, 1920, 0, 0, 0),
    ("512 lin L1 640->5120 geglu", 1, 1, 1, 10240, 640, 5120, 0, 0, 1),
    ("512 lin L1 640->640 +res", 1, 1, 1, 10240, 640, 640, 0, 1, 0),
    ("512 conv3 L2 1280->1280", 1, 16, 10, 16, 1280, 1280, 1, 1, 0),
    ("512 lin L2 1280->10240 geglu", 1, 1, 1, 2560, 1280, 10240, 0, 0, 1),
    ("512 lin L2 1280->1280 +res", 1, 1, 1, 2560, 1280, 1280, 0, 1, 0),
    ("512 conv3 L3 1280->1280", 1, 16, 5, 8, 1280, 1280, 1, 1, 0),
]


# shapes whose 256-wide pair tiles leave the last wave mostly empty (knob gemm_balance: narrower N tiles)
SHAPES_BALANCE = [
    ("conv3 L3 1280->1280", 1, 32, 9, 16, 1280, 1280, 1, 1, 0),
    ("conv3 L3 2560->1280", 1, 32, 9, 16, 2560, 1280, 1, 0, 0),
    ("tconv L3 1280", 2, 16, 9, 16, 1280, 1280, 2, 1, 0),
    ("lin L3 1280->1280", 1, 1, 1, 4608, 1280, 1280, 0, 1, 0),
    ("lin L3 5120->1280 +res", 1, 1, 1, 4608, 5120, 1280, 0, 1, 0),
    ("lin L3 1280->3840 qkv", 1, 1, 1, 4608, 1280, 3840, 0, 0, 0),
    ("lin L3 1280->10240 geglu", 1, 1, 1, 4608, 1280, 10240, 0, 0, 1),
    ("conv3 L2 1280->1280", 1, 32, 18, 32, 1280, 1280, 1, 1, 0),
    ("lin L2 1280->1280 +res", 1, 1, 1, 18432, 1280, 1280, 0, 1, 0),
    ("lin L1 640->640 +res", 1, 1, 1, 73728, 640, 640, 0, 1, 0),
    ("512 conv3 L2 1280->1280", 1, 16, 10, 16, 1280, 1280, 1, 1, 0),
    ("512 tconv L2 1280", 1, 16, 10, 16, 1280, 1280, 2, 0, 0),
    ("512 lin L2 1280->1280 +res", 1, 1, 1, 2560, 1280, 1280, 0, 1, 0),
    ("512 lin L2 5120->1280 +res", 1, 1, 1, 2560, 5120, 1280, 0, 1, 0),
    ("512 lin L2 1280->3840 qkv", 1, 1, 1, 2560, 1280, 3840, 0, 0, 0),
    ("512 conv3 L1 640->640", 1, 16, 20, 32, 640, 640, 1, 1, 0),
    ("512 lin L1 640->640 +res", 1, 1, 1, 10240, 640, 640, 0, 1, 0),
    ("512 lin L1 2560->640 +res", 1, 1, 1, 10240, 2560, 640, 0, 1, 0),
]


# sub-wave launches of the single-CTA kernel (MDM512 level 3: 640 rows): knob gemm_deep = 6-stage ring, one CTA per SM
SHAPES_DEEP = [
    ("512 conv3 L3 1280->1280", 1, 16, 5, 8, 1280, 1280, 1, 1, 0),
    ("512 conv3 L3 2560->1280", 1, 16, 5, 8, 2560, 1280, 1, 0, 0),
    ("512 tconv L3 1280", 1, 16, 5, 8, 1280, 1280, 2, 1, 0),
    ("512 lin L3 1280->1280 +res", 1, 1, 1, 640, 1280, 1280, 0, 1, 0),
    ("512 lin L3 5120->1280 +res", 1, 1, 1, 640, 5120, 1280, 0, 1, 0),
    ("512 lin L3 1280->3840 qkv", 1, 1, 1, 640, 1280, 3840, 0, 0, 0),
    ("512 lin L3 1280->10240 geglu", 1, 1, 1, 640, 1280, 10240, 0, 0, 1),
    ("vae-ish conv3 512->512 16x24", 1, 1, 16, 24, 512, 512, 1, 1, 0),
]


def bench_lnfuse():
    """LayerNorm statistics from the producing GEMM's epilogue (TapGemm::ln_out + ln_finalize) against the separate ln_stats
    pass over the activation: producer without / with the partial sums, ln_stats, ln_finalize."""
    dev = "cuda"
    L = test_lib()

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3

    print(f"{'rows x C':16s} {'prod':>8s} {'prod+ln':>8s} {'ln_stats':>8s} {'finalize':>8s}   (us)")
    for rows, C in ((294912, 320), (73728, 640), (18432, 1280), (4608, 1280), (40960, 320), (10240, 640)):
        A = torch.randn(rows, C, device=dev).half()
        R = torch.randn(rows, C, device=dev).half()
        D = torch.empty(rows, C, device=dev).half()
        Wp = (torch.randn(C, C, device=dev) / C ** 0.5).half()
        bias = torch.randn(C, device=dev)
        parts = torch.empty(C // 64, rows, 2, device=dev)
        stats = torch.empty(rows, 2, device=dev)

        def gemm(ln_out=None):
            if ln_out is not None:
                check(L.mudg_test_next_gemm_ln(ptr(ln_out)))
            check(L.mudg_test_tapgemm(ptr(A), 1, 1, 1, rows, C, 0, ptr(Wp), C, ptr(D), ptr(R), ptr(bias), None,
                                      ctypes.c_int(1), ctypes.c_int(0), ctypes.c_float(1.0), 0, None, None, 0, cur_stream()))
        t_p = timed(lambda: gemm())
        t_pl = timed(lambda: gemm(parts))
        t_ln = timed(lambda: check(L.mudg_test_ln_stats(ptr(A), ptr(stats), ctypes.c_int64(rows), C, cur_stream())))
        t_f = timed(lambda: check(L.mudg_test_ln_finalize(ptr(parts), C // 64, ptr(stats), ctypes.c_int64(rows), C, cur_stream())))
        print(f"{rows:7d} x {C:5d} {t_p:8.1f} {t_pl:8.1f} {t_ln:8.1f} {t_f:8.1f}", flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "lnfuse":
        return bench_lnfuse()
    dev = "cuda"
    L = test_lib()
    global SHAPES
    knob = b"gemm_epi"
    backends = [(0, "tc")]
    vals = {"tc": 1}
    if len(sys.argv) > 1 and sys.argv[1] == "mdm512":
        # columns: single-CTA kernel forced, CTA-pair kernel forced (knob gemm_pair), the product heuristic
        SHAPES, knob = SHAPES_512, b"gemm_pair"
        backends = [(0, "single"), (0, "pair"), (0, "auto")]
        vals = {"single": 0, "pair": 1, "auto": -1}
        sys.argv = sys.argv[:1]
    if len(sys.argv) > 1 and sys.argv[1] == "deep":
        SHAPES, knob = SHAPES_DEEP, b"gemm_deep"
        backends = [(0, "3stage"), (0, "deep")]
        vals = {"3stage": 0, "deep": 1}
        sys.argv = sys.argv[:1]
    if len(sys.argv) > 1 and sys.argv[1] == "resmma":
        SHAPES = [x for x in SHAPES if x[8]] + [("lin L1 640->640 +res", 1, 1, 1, 73728, 640, 640, 0, 1, 0), ("lin L2 1280->1280 +res", 1, 1, 1, 18432, 1280, 1280, 0, 1, 0),
                                                 ("512 lin L0 320->320 +res", 1, 1, 1, 40960, 320, 320, 0, 1, 0), ("512 lin L1 640->640 +res", 1, 1, 1, 10240, 640, 640, 0, 1, 0)]
        knob = b"gemm_resmma"
        backends = [(0, "epilogue"), (0, "mma"), (0, "rule")]
        vals = {"epilogue": 0, "mma": 1, "rule": -1}
        sys.argv = sys.argv[:1]
    if len(sys.argv) > 1 and sys.argv[1] == "balance":
        SHAPES, knob = SHAPES_BALANCE, b"gemm_balance"
        backends = [(0, "wide"), (0, "balanced")]
        vals = {"wide": 0, "balanced": 1}
        sys.argv = sys.argv[:1]
    print(f"{'shape':28s} " + " ".join(f"{n:>8s}us {n:>6s}TF" for _, n in backends))
    only = sys.argv[1] if len(sys.argv) > 1 else None
    for name, B, T, H, W, Cin, N, mode, res, geglu in SHAPES:
        if only and only not in name:
            continue
        ntaps = {0: 1, 1: 9, 2: 3}[mode]
        A = torch.randn(B, T, H, W, Cin, device=dev).half()
        Wt = (torch.randn(N, ntaps * Cin, device=dev) / (ntaps * Cin) ** 0.5).half()
        n_out = N // 2 if geglu else N
        D = torch.empty(B, T, H, W, n_out, device=dev).half()
        R = torch.randn(B, T, H, W, n_out, device=dev).half() if res else None
        bias = torch.randn(N, device=dev)
        flops = 2.0 * B * T * H * W * N * ntaps * Cin
        row = f"{name:28s} "
        for backend, bn in backends:
            check(L.mudg_test_set_knob(knob, vals[bn]))

            def run():
                check(L.mudg_test_tapgemm(ptr(A), B, T, H, W, Cin, mode, ptr(Wt), N, ptr(D), ptr(R), ptr(bias), None,
                                          ctypes.c_int(1), ctypes.c_int(0), ctypes.c_float(1.0), int(geglu), None, None, backend, cur_stream()))
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / reps * 1e3
            row += f"{us:10.1f} {flops / us / 1e6:8.0f} "
        print(row, flush=True)


if __name__ == "__main__":
    main()
