"""Standalone GPU probe for the tap-GEMM kernels (run under gpurun; one case per process so a device trap
cannot poison the other cases).  usage: python tests/gpu_probe_gemm.py <case>|all"""
import ctypes
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (B,T,H,W,Cin,N,mode,res,bias,bias2,geglu)
    "lin_small": (1, 1, 1, 256, 64, 128, 0, 0, 0, 0, 0),
    "lin_k320": (1, 1, 1, 1000, 320, 320, 0, 1, 1, 0, 0),
    "lin_geglu": (1, 1, 1, 512, 320, 2560, 0, 0, 1, 0, 1),
    "conv3x3": (1, 4, 16, 16, 64, 128, 1, 1, 1, 1, 0),
    "conv3x3_l0": (1, 2, 72, 128, 320, 320, 1, 1, 1, 1, 0),
    "conv3x3_odd": (1, 3, 9, 16, 128, 192, 1, 0, 1, 0, 0),
    "tconv": (2, 4, 8, 8, 128, 128, 2, 1, 1, 0, 0),
    "tconv_l0": (1, 16, 18, 32, 640, 640, 2, 1, 1, 0, 0),
    "cin16": (1, 2, 16, 16, 16, 64, 1, 0, 1, 0, 0),
    "lin_persist": (1, 1, 1, 40000, 320, 320, 0, 1, 1, 0, 0),
    "lin_n960": (1, 1, 1, 30000, 320, 960, 0, 0, 0, 0, 0),
    "lin_n64": (1, 1, 1, 5000, 128, 64, 0, 1, 1, 0, 0),
    "geglu_big": (1, 1, 1, 30000, 320, 2560, 0, 0, 1, 0, 1),
    "conv_emb": (2, 4, 18, 32, 640, 640, 1, 1, 1, 1, 0),
}


def run_case(name):
    import torch
    import torch.nn.functional as F
    from mudg_b200._lib import lib, check, ptr, cur_stream
    B, T, H, W, Cin, N, mode, res, bias, bias2, geglu = CASES[name]
    torch.manual_seed(0)
    dev = "cuda"
    ntaps = {0: 1, 1: 9, 2: 3}[mode]
    A = torch.randn(B, T, H, W, Cin, device=dev).half()
    Wt = (torch.randn(N, ntaps, Cin, device=dev) / (ntaps * Cin) ** 0.5).half()
    n_out = N // 2 if geglu else N
    R = torch.randn(B, T, H, W, n_out, device=dev).half() if res else None
    bv = torch.randn(N, device=dev) if bias else None
    b2 = torch.randn(B, N, device=dev) if bias2 else None
    alpha = 0.75
    # fp32 reference on the fp16-rounded operands
    if mode == 0:
        y = A.float().reshape(-1, Cin) @ Wt.float().reshape(N, Cin).t()
        y = y.reshape(B, T, H, W, N)
    elif mode == 1:
        x = A.float().reshape(B * T, H, W, Cin).permute(0, 3, 1, 2)
        w = Wt.float().reshape(N, 3, 3, Cin).permute(0, 3, 1, 2)
        y = F.conv2d(x, w, padding=1).permute(0, 2, 3, 1).reshape(B, T, H, W, N)
    else:
        x = A.float().permute(0, 4, 1, 2, 3)
        w = Wt.float().reshape(N, 3, Cin).permute(0, 2, 1)[..., None, None]
        y = F.conv3d(x, w, padding=(1, 0, 0)).permute(0, 2, 3, 4, 1)
    y = y * alpha
    if bv is not None:
        y = y + bv
    if b2 is not None:
        y = y + b2[:, None, None, None, :]
    if geglu:
        yy = y.reshape(B, T, H, W, N // 128, 2, 64)
        y = (yy[..., 0, :] * F.gelu(yy[..., 1, :])).reshape(B, T, H, W, n_out)
    if R is not None:
        y = y + R.float()
    out = {}
    for backend, label in ((1, "simt"), (2, "tc_v1"), (0, "tc")):
        D = torch.full((B, T, H, W, n_out), float("nan"), device=dev).half()
        torch.cuda.synchronize()
        t0 = time.time()
        rc = lib().mudg_test_tapgemm(ptr(A), B, T, H, W, Cin, mode, ptr(Wt), N, ptr(D), ptr(R),
                                     ptr(bv), ptr(b2), ctypes.c_int(T), ctypes.c_int(B if b2 is not None else 0),
                                     ctypes.c_float(alpha), int(geglu), backend, cur_stream())
        check(rc)
        torch.cuda.synchronize()
        err = (D.float() - y).abs()
        nan = int(torch.isnan(D.float()).sum())
        out[label] = (float(err[~torch.isnan(err)].max()) if nan < err.numel() else float("nan"), nan)
        print(f"{name:14s} {label:5s} max|d|={out[label][0]:.5f} nans={nan} ref_absmax={float(y.abs().max()):.3f} "
              f"({(time.time() - t0) * 1e3:.1f} ms)", flush=True)
        if label.startswith("tc") and (out[label][0] > 0.05 or nan):
            bad = (err > 0.05) | torch.isnan(D.float())
            idx = bad.nonzero()
            print("   first bad idx:", idx[:5].tolist(), " n_bad:", int(bad.sum()), "of", err.numel(), flush=True)
            # which rows/cols are bad
            rows = bad.reshape(-1, n_out).any(dim=1).nonzero().flatten()
            cols = bad.reshape(-1, n_out).any(dim=0).nonzero().flatten()
            print("   bad rows (first 16):", rows[:16].tolist(), "count", rows.numel(),
                  " bad cols (first 16):", cols[:16].tolist(), "count", cols.numel(), flush=True)
    return out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "inproc":          # one process (fast); a device trap poisons the remaining cases
        for name in CASES:
            run_case(name)
    elif which == "all":
        for name in CASES:
            r = subprocess.run([sys.executable, __file__, name], timeout=600)
            if r.returncode != 0:
                print(f"{name}: FAILED rc={r.returncode}", flush=True)
    else:
        run_case(which)
