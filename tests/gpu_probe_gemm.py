"""Standalone GPU probe for the tap-GEMM kernels (run under gpurun; `all` runs one case per process so a device trap
cannot poison the other cases).  usage: python tests/gpu_probe_gemm.py <case>|all|inproc

Every case compares the product dispatch (tcgen05: tapgemm_tc2 / the CTA-pair tapgemm_tc3) and the CUDA-core checker with
an fp32 torch reference (F.linear / F.conv2d / F.conv3d / F.layer_norm on the fp16-rounded operands), and records WHICH
kernel the dispatch took (mudg_test_last_gemm_path) so that a test can assert the pair path really ran."""
import ctypes
import subprocess
import sys
import time
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# expected kernel: 2 = tapgemm_tc2<1>, 3 = tapgemm_tc2<2>, 4 = tapgemm_tc3 (CTA pair); None = do not care
CASES = {
    # name: (B,T,H,W,Cin,N,mode,res,bias,bias2,geglu,ln,alpha,path)
    "lin_small": (1, 1, 1, 256, 64, 128, 0, 0, 0, 0, 0, 0, 0.75, 2),
    "lin_k320": (1, 1, 1, 1000, 320, 320, 0, 1, 1, 0, 0, 0, 0.75, 2),
    "lin_geglu": (1, 1, 1, 512, 320, 2560, 0, 0, 1, 0, 1, 0, 0.75, 2),
    "conv3x3": (1, 4, 16, 16, 64, 128, 1, 1, 1, 1, 0, 0, 0.75, 2),
    "conv3x3_l0": (1, 2, 72, 128, 320, 320, 1, 1, 1, 1, 0, 0, 0.75, None),
    "conv3x3_odd": (1, 3, 9, 16, 128, 192, 1, 0, 1, 0, 0, 0, 0.75, 2),
    "tconv": (2, 4, 8, 8, 128, 128, 2, 1, 1, 0, 0, 0, 0.75, 2),
    "tconv_l0": (1, 16, 18, 32, 640, 640, 2, 1, 1, 0, 0, 0, 0.75, None),
    "cin16": (1, 2, 16, 16, 16, 64, 1, 0, 1, 0, 0, 0, 0.75, 2),
    "lin_persist": (1, 1, 1, 40000, 320, 320, 0, 1, 1, 0, 0, 0, 0.75, None),
    "lin_n960": (1, 1, 1, 30000, 320, 960, 0, 0, 0, 0, 0, 0, 0.75, 4),
    "lin_n64": (1, 1, 1, 5000, 128, 64, 0, 1, 1, 0, 0, 0, 0.75, 2),
    "geglu_big": (1, 1, 1, 30000, 320, 2560, 0, 0, 1, 0, 1, 0, 0.75, 4),
    "conv_emb": (2, 4, 18, 32, 640, 640, 1, 1, 1, 1, 0, 0, 0.75, None),
    # ---- the CTA-pair kernel's hot variants at the BASELINE level-0 shapes (MDM1024, N=2 CFG batch, T=16, 72x128)
    # EPI 3: bias + per-sample bias (ResBlock in_layers conv + emb), 9 taps
    "pair_conv_l0_emb": (2, 16, 72, 128, 320, 320, 1, 0, 1, 1, 0, 0, 1.0, 4 | (4 << 8)),
    # EPI 5: bias + residual (ResBlock out_layers conv + skip), 9 taps
    "pair_conv_l0_res": (2, 16, 72, 128, 320, 320, 1, 1, 1, 0, 0, 0, 1.0, 4 | (2 << 8)),
    # concat input (output blocks): 640 -> 320, bias + per-sample bias
    "pair_conv_l0_cat": (2, 16, 72, 128, 640, 320, 1, 0, 1, 1, 0, 0, 1.0, 4 | (4 << 8)),
    # temporal conv (3,1,1) with the TemporalConvBlock identity, 3 taps shifted along T across sample borders
    "pair_tconv_l0_res": (2, 16, 72, 128, 320, 320, 2, 1, 1, 0, 0, 0, 1.0, 4 | (2 << 8)),
    "pair_tconv_l0": (2, 16, 72, 128, 320, 320, 2, 0, 1, 0, 0, 0, 1.0, 4 | (2 << 8)),
    # Linear 320 -> 320 + residual over all M = 294 912 rows (attention to_out / proj_out)
    "pair_lin_k320_res": (1, 1, 1, 294912, 320, 320, 0, 1, 1, 0, 0, 0, 1.0, 4 | (2 << 8)),     # residual through the MMA: bias-only epilogue
    # folded LayerNorm (EPI 9 takes bias = c2): QKV 320 -> 960 and to_q 320 -> 320
    "pair_lin_ln_qkv": (1, 1, 1, 294912, 320, 960, 0, 0, 1, 0, 0, 1, 1.0, 4 | (10 << 8)),
    "pair_lin_ln_q": (1, 1, 1, 147456, 320, 320, 0, 0, 1, 0, 0, 1, 1.0, 4 | (10 << 8)),
    # GEGLU with the folded LayerNorm (run-time epilogue variant), 320 -> 2560 -> 1280 columns
    "pair_geglu_ln": (1, 1, 1, 147456, 320, 2560, 0, 0, 1, 0, 1, 1, 1.0, 4),
    # N not a multiple of 256 / odd tile splits: 640 -> 192+192+256 units, 1280, 1920-wide K, N = 448
    "pair_lin_n640": (1, 1, 1, 73728, 640, 640, 0, 1, 1, 0, 0, 0, 1.0, 4 | (2 << 8)),
    "pair_lin_n448": (1, 1, 1, 73728, 320, 448, 0, 0, 1, 0, 0, 0, 1.0, 4 | (2 << 8)),
    "pair_conv_l1_cat": (2, 16, 36, 64, 1920, 640, 1, 0, 1, 1, 0, 0, 1.0, 4 | (4 << 8)),
    # rows not a multiple of the 256-row pair tile (odd M-tile count): the second CTA of the last pair runs past the end
    "pair_lin_ragged": (1, 1, 1, 200000 + 77, 320, 320, 0, 1, 1, 0, 0, 0, 1.0, 4 | (2 << 8)),
    # plain alpha != 1 goes through the run-time variant of the pair kernel
    "pair_lin_alpha": (1, 1, 1, 147456, 512, 512, 0, 0, 0, 0, 0, 0, 0.125, 4),
    # single-CTA kernel with the folded LayerNorm (small M: the small-config UNet's path)
    "lin_ln_small": (1, 1, 1, 2048, 64, 192, 0, 0, 1, 0, 0, 1, 1.0, 2),
    "geglu_ln_small": (1, 1, 1, 2048, 64, 512, 0, 0, 1, 0, 1, 1, 1.0, 2),
}
# per-sample weight matrices (GroupNorm folded into proj_in): name -> frames per sample (1 = per frame, T = per batch sample)
CASES["pair_lin_persample_frame"] = (2, 16, 72, 128, 320, 320, 0, 0, 1, 1, 0, 0, 1.0, 4 | (4 << 8))
CASES["pair_lin_persample_batch"] = (2, 16, 36, 64, 640, 640, 0, 0, 1, 1, 0, 0, 1.0, 4 | (4 << 8))
# wave-balanced N tiles (1 280 columns as 6 x 192 + 128 instead of 5 x 256): MDM1024 level 3 and MDM512 level 2
CASES["pair_conv_l3_balanced"] = (2, 16, 9, 16, 1280, 1280, 1, 1, 1, 0, 0, 0, 1.0, 4 | (2 << 8))
CASES["pair_lin_l2_balanced"] = (1, 1, 1, 2560, 1280, 1280, 0, 1, 1, 0, 0, 0, 1.0, 4 | (2 << 8))
CASES["pair_tconv_l3_balanced"] = (2, 16, 9, 16, 1280, 1280, 2, 0, 1, 0, 0, 0, 1.0, 4 | (2 << 8))
# the residual added by the tensor core (identity k-steps) with the epilogue-side residual forced off / on: same results
CASES["pair_lin_k320_res_epi"] = (1, 1, 1, 294912, 320, 320, 0, 1, 1, 0, 0, 0, 1.0, 4 | (6 << 8))
CASES["pair_conv_l0_res_epi"] = (2, 16, 72, 128, 320, 320, 1, 1, 1, 0, 0, 0, 1.0, 4 | (6 << 8))
CASES["pair_lin_k1280_res_mma"] = (1, 1, 1, 147456, 1280, 320, 0, 1, 1, 0, 0, 0, 1.0, 4 | (2 << 8))
CASES["pair_conv_l1_res_mma"] = (2, 16, 36, 64, 640, 640, 1, 1, 1, 0, 0, 0, 1.0, 4 | (2 << 8))
RESMMA_KNOB = {"pair_lin_k320_res_epi": 0, "pair_conv_l0_res_epi": 0}
RESMMA_EXPECT = {"pair_lin_k320_res": 1, "pair_lin_n640": 1, "pair_lin_ragged": 1, "pair_lin_k320_res_epi": 0, "pair_conv_l0_res_epi": 0,
                 "pair_lin_k1280_res_mma": 1, "pair_conv_l1_res_mma": 1, "pair_conv_l0_res": 1, "pair_tconv_l0_res": 1,
                 "pair_lin_l2_balanced": 1, "pair_conv_l3_balanced": 1}
PER_SAMPLE = {"pair_lin_persample_frame": 1, "pair_lin_persample_batch": 16}
# producers whose epilogue also stores the LayerNorm partial sums of their output rows (TapGemm::ln_out), checked against
# torch sums of the fp32 reference output, and reduced by ln_finalize to the (mean, rstd) rows a consumer GEMM takes
LN_OUT_CASES = {"pair_lin_k320_res", "pair_lin_k320_res_epi", "pair_lin_n640", "pair_lin_ragged", "pair_lin_persample_frame",
                "pair_lin_persample_batch", "pair_lin_l2_balanced", "pair_lin_k1280_res_mma", "pair_lin_n448"}
PAIR_CASES = [k for k in CASES if k.startswith("pair_")]
# cases that also request the fused GroupNorm statistics of their output: name -> frames per GroupNorm sample
# (1 = per-frame norm, T = TemporalConvBlock norm over (C/32, T, H, W)); the conv modes flatten (B, T) like the model does
GN_CASES = {"pair_conv_l0_emb": 1, "pair_conv_l0_res": 16, "pair_conv_l0_cat": 1, "pair_tconv_l0_res": 1, "pair_tconv_l0": 16,
            "pair_conv_l1_cat": 1, "conv3x3_l0": 1, "pair_conv_l3_balanced": 16}


def run_case(name, backends=((1, "simt"), (0, "tc"))):
    import torch
    import torch.nn.functional as F
    from mudg_b200._lib import test_lib, check, ptr, cur_stream
    L = test_lib()
    B, T, H, W, Cin, N, mode, res, bias, bias2, geglu, ln, alpha, want_path = CASES[name]
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = "cuda"
    ntaps = {0: 1, 1: 9, 2: 3}[mode]
    rows = B * T * H * W
    A = torch.randn(B, T, H, W, Cin, device=dev)
    if ln:
        A = A * (0.5 + torch.rand(B, T, H, W, 1, device=dev)) + 0.3 * torch.randn(B, T, H, W, 1, device=dev)   # per-row mean / scale
    A = A.half()
    ws_div = PER_SAMPLE.get(name)
    n_ws = (B * T) // ws_div if ws_div else 0
    Wraw = torch.randn(N, ntaps, Cin, device=dev) / (ntaps * Cin) ** 0.5
    n_out = N // 2 if geglu else N
    R = torch.randn(B, T, H, W, n_out, device=dev).half() if res else None
    bv = torch.randn(N, device=dev) if bias else None
    b2 = torch.randn(n_ws if ws_div else B, N, device=dev) if bias2 else None
    ln_stats = ln_c1 = None
    if ln:
        # the library's own fold (mudg_finalize_weights): W *= gamma (fp16), c1 = row sums, c2 = W beta + bias
        assert mode == 0
        gamma = 1.0 + 0.1 * torch.randn(Cin, device=dev)
        beta = 0.1 * torch.randn(Cin, device=dev)
        Wt = Wraw.reshape(N, Cin).half().contiguous()
        W_unfolded = Wt.float().clone()
        ln_c1 = torch.empty(N, device=dev)
        c2 = torch.empty(N, device=dev)
        check(L.mudg_test_ln_fold(ptr(Wt), ptr(gamma), ptr(beta), ptr(bv), ptr(ln_c1), ptr(c2), N, Cin, cur_stream()))
        ln_stats = torch.empty(rows, 2, device=dev)
        check(L.mudg_test_ln_stats(ptr(A), ptr(ln_stats), ctypes.c_int64(rows), Cin, cur_stream()))
        torch.cuda.synchronize()
        # ln_stats against torch
        a32 = A.float().reshape(rows, Cin)
        mean = a32.mean(1)
        rstd = (a32.var(1, unbiased=False) + 1e-5).rsqrt()
        ds = max(float((ln_stats[:, 0] - mean).abs().max()), float(((ln_stats[:, 1] - rstd) / rstd).abs().max()))
        assert ds < 1e-4, ("ln_stats", ds)
        y = F.linear(F.layer_norm(a32, (Cin,), gamma, beta, 1e-5), W_unfolded, bv).reshape(B, T, H, W, N)
        bias_arg = c2
    else:
        Wt = Wraw.half()
        bias_arg = bv
        # fp32 reference on the fp16-rounded operands, one frame / sample at a time to bound memory
        if ws_div:
            # one weight matrix (and one bias2 row) per sample of ws_div frames
            Wt = (torch.randn(n_ws, N, Cin, device=dev) / Cin ** 0.5).half()
            a = A.float().reshape(n_ws, -1, Cin)
            y = torch.stack([a[i] @ Wt[i].float().t() + b2[i] for i in range(n_ws)]).reshape(B, T, H, W, N)
        elif mode == 0:
            y = (A.float().reshape(-1, Cin) @ Wt.float().reshape(N, Cin).t()).reshape(B, T, H, W, N)
        elif mode == 1:
            w = Wt.float().reshape(N, 3, 3, Cin).permute(0, 3, 1, 2).contiguous()
            x = A.reshape(B * T, H, W, Cin)
            y = torch.cat([F.conv2d(x[i:i + 4].float().permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1)
                           for i in range(0, B * T, 4)]).reshape(B, T, H, W, N)
        else:
            w = Wt.float().reshape(N, 3, Cin).permute(0, 2, 1)[..., None, None].contiguous()
            y = torch.cat([F.conv3d(A[i:i + 1].float().permute(0, 4, 1, 2, 3), w, padding=(1, 0, 0)).permute(0, 2, 3, 4, 1)
                           for i in range(B)])
        y = y * alpha
        if bv is not None:
            y = y + bv
    if b2 is not None and not ws_div:
        y = y + b2[:, None, None, None, :]
    if geglu:
        yy = y.reshape(B, T, H, W, N // 128, 2, 64)
        y = (yy[..., 0, :] * F.gelu(yy[..., 1, :])).reshape(B, T, H, W, n_out)
    if R is not None:
        y = y + R.float()
    out = {}
    gn_div = GN_CASES.get(name)
    n_samples = (B * T) // gn_div if gn_div else 0
    if ws_div:
        backends = ((0, "tc"),)            # the CUDA-core checker has no per-sample weights
    for backend, label in backends:
        D = torch.full((B, T, H, W, n_out), float("nan"), device=dev).half()
        if ws_div:
            check(L.mudg_test_next_gemm_per_sample(n_ws, ws_div))
        gn_sums = None
        if gn_div and backend == 0:
            gn_sums = torch.zeros(n_samples, 32, 2, device=dev, dtype=torch.float64)
            check(L.mudg_test_set_knob(b"gn_fuse", 2))      # also the 3-tap convs (the product only fuses 9-tap ones: slack)
            check(L.mudg_test_next_gemm_gn(ptr(gn_sums), gn_div))
        ln_parts_out = None
        if backend == 0 and name in LN_OUT_CASES:
            ln_parts_out = torch.full((n_out // 64, rows, 2), float("nan"), device=dev)
            check(L.mudg_test_next_gemm_ln(ptr(ln_parts_out)))
        if name in RESMMA_KNOB and backend == 0:
            check(L.mudg_test_set_knob(b"gemm_resmma", RESMMA_KNOB[name]))
        if os.environ.get("PROBE_BALANCE") and backend == 0:
            check(L.mudg_test_set_knob(b"gemm_balance", int(os.environ["PROBE_BALANCE"])))
        torch.cuda.synchronize()
        t0 = time.time()
        rc = L.mudg_test_tapgemm(ptr(A), B, T, H, W, Cin, mode, ptr(Wt), N, ptr(D), ptr(R),
                                 ptr(bias_arg), ptr(b2), ctypes.c_int(ws_div if ws_div else T),
                                 ctypes.c_int((n_ws if ws_div else B) if b2 is not None else 0),
                                 ctypes.c_float(alpha), int(geglu), ptr(ln_stats), ptr(ln_c1), backend, cur_stream())
        check(rc)
        torch.cuda.synchronize()
        path = int(L.mudg_test_last_gemm_path()) if backend == 0 else 0
        check(L.mudg_test_set_knob(b"reset", 0))
        err = (D.float() - y).abs()
        nan = int(torch.isnan(D.float()).sum())
        emax = float(err[~torch.isnan(err)].max()) if nan < err.numel() else float("nan")
        out[label] = (emax, nan)
        if backend == 0:
            out["_path"] = (path & 0xffff, want_path)
            if name in RESMMA_EXPECT:
                out["_resmma"] = ((path >> 17) & 1, RESMMA_EXPECT[name])
            if ln_parts_out is not None:
                # partial LayerNorm sums of the output rows against torch sums of the fp32 reference output
                y64 = y.reshape(rows, n_out // 64, 64)
                want = torch.stack([y64.sum(-1), (y64 * y64).sum(-1)], dim=-1).permute(1, 0, 2)
                e = (ln_parts_out - want).abs() / (want.abs() + 1e-2 * want.abs().max())
                nan_p = int(torch.isnan(ln_parts_out).sum())
                rel = float(e.nan_to_num().max())
                fin = torch.full((rows, 2), float("nan"), device=dev)
                check(L.mudg_test_ln_finalize(ptr(ln_parts_out), n_out // 64, ptr(fin), ctypes.c_int64(rows), n_out, cur_stream()))
                torch.cuda.synchronize()
                y2 = y.reshape(rows, n_out)
                mean, rstd = y2.mean(1), (y2.var(1, unbiased=False) + 1e-5).rsqrt()
                dfin = max(float((fin[:, 0] - mean).abs().max()), float(((fin[:, 1] - rstd) / rstd).abs().max()))
                out["_ln_out"] = (bool((path >> 18) & 1), rel, nan_p, dfin)
                print(f"{name:18s} LayerNorm partials of the output: taken={bool((path >> 18) & 1)} max rel err {rel:.2e} nans={nan_p}; "
                      f"ln_finalize vs torch (mean abs, rstd rel) {dfin:.2e}", flush=True)
            if gn_sums is not None:
                # fused GroupNorm statistics against fp64 sums of the fp16 output the kernel stored
                d64 = D.double().reshape(n_samples, -1, 32, n_out // 32)
                want = torch.stack([d64.sum(dim=(1, 3)), (d64 * d64).sum(dim=(1, 3))], dim=-1)
                rel = float(((gn_sums - want).abs() / (want.abs() + 1e-3 * want.abs().max())).max())
                taken = bool((path >> 16) & 1)          # bit 16 only: the epilogue-group count sits at bit 20
                out["_gn"] = (taken, rel)
                if rel > 1e-3:
                    e = (gn_sums - want).abs() / (want.abs() + 1e-3 * want.abs().max())
                    bad = (e > 1e-3).nonzero()
                    print("   bad (sample, group, sum|sumsq):", bad[:12].tolist(), "count", bad.shape[0], "of", e.numel(), flush=True)
                    for i in bad[:6].tolist():
                        print("     got", float(gn_sums[tuple(i)]), "want", float(want[tuple(i)]), flush=True)
                print(f"{name:18s} fused GroupNorm statistics: taken={taken} max rel err {rel:.2e}", flush=True)
        print(f"{name:18s} {label:5s} max|d|={emax:.5f} mean|d|={float(err.nan_to_num().mean()):.6f} nans={nan} "
              f"ref_absmax={float(y.abs().max()):.3f} path={path & 255} epi={((path >> 8) & 255) - 1} groups={(path >> 20) & 15} ({(time.time() - t0) * 1e3:.1f} ms)", flush=True)
        if label.startswith("tc") and (emax > 0.05 or nan):
            bad = (err > 0.05) | torch.isnan(D.float())
            idx = bad.nonzero()
            print("   first bad idx:", idx[:5].tolist(), " n_bad:", int(bad.sum()), "of", err.numel(), flush=True)
            rows_b = bad.reshape(-1, n_out).any(dim=1).nonzero().flatten()
            cols = bad.reshape(-1, n_out).any(dim=0).nonzero().flatten()
            print("   bad rows (first 16):", rows_b[:16].tolist(), "count", rows_b.numel(),
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
