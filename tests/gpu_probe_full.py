"""Full-size timing probe: random weights generated on the GPU, CUDA-event timing of UNet forwards + VAE decode."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mudg_b200.engine import Engine, MUDG_UNET, MUDG_VAE   # noqa: E402
from mudg_b200.layout import unet_layout, vae_layout        # noqa: E402

UNET = dict(in_channels=12, out_channels=4, model_channels=320, num_res_blocks=2, channel_mult=(1, 2, 4, 4),
            attention_resolutions=(4, 2, 1), num_head_channels=64, context_dim=1024)
VAE = dict(ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4, out_ch=3, embed_dim=4)
FLOPS = {(16, 40, 64): 12.604e12, (16, 72, 128): 52.340e12, (64, 72, 128): 212.49e12}


def gpu_weights(layout, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    sd = {}
    for k, shp in layout.items():
        if len(shp) == 1:
            sd[k] = (1.0 + 0.1 * torch.randn(shp, generator=g, device="cuda")) if k.endswith(".weight") else \
                0.02 * torch.randn(shp, generator=g, device="cuda")
        else:
            fan = 1
            for s in shp[1:]:
                fan *= s
            sd[k] = torch.randn(shp, generator=g, device="cuda") / fan ** 0.5
    return sd


def main():
    shapes = sys.argv[1:] or ["1x16x40x64", "2x16x72x128"]
    eng = Engine(UNET, VAE)
    t0 = time.time()
    eng.load_state_dict(gpu_weights(unet_layout(**UNET), 0), MUDG_UNET)
    eng.load_state_dict(gpu_weights(vae_layout(**VAE), 1), MUDG_VAE)
    torch.cuda.synchronize()
    print(f"weights: {time.time() - t0:.1f}s, mem {torch.cuda.memory_allocated() / 1e9:.1f} GB torch", flush=True)
    for s in shapes:
        N, T, h, w = (int(v) for v in s.split("x"))
        x = torch.randn(N, 12, T, h, w, device="cuda")
        ctx = torch.randn(N, 77 + (16 * T if T == 16 else 256), 1024, device="cuda")
        ts = torch.full((N,), 500, device="cuda", dtype=torch.long)
        lab = torch.zeros(N, device="cuda", dtype=torch.long)
        fs = torch.full((N,), 10, device="cuda", dtype=torch.long)
        print(f"shape {s}: workspace {eng.workspace_bytes(N, T, h, w) / 1e9:.2f} GB", flush=True)
        eng.set_context(ctx, T)
        l0 = eng.launch_count()
        y = eng.unet_forward(x, ts, lab, fs)
        torch.cuda.synchronize()
        print("  launches/forward:", eng.launch_count() - l0, " out finite:", bool(torch.isfinite(y.float()).all()),
              " absmax", float(y.float().abs().max()), flush=True)
        if os.environ.get("PROBE_ONE", "0") == "1":
            continue
        for _ in range(2):
            eng.unet_forward(x, ts, lab, fs)
        torch.cuda.synchronize()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.unet_forward(x, ts, lab, fs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = FLOPS.get((T, h, w), 0) * N
        print(f"  unet forward: {ms:.2f} ms  -> {fl / ms / 1e9:.1f} TFLOP/s algorithmic", flush=True)
        if os.environ.get("MUDG_GEMM_PROFILE_DUMP"):      # per-shape table of the tcgen05 GEMM launches (one eager forward)
            import ctypes
            from mudg_b200._lib import lib
            L = lib()
            L.mudg_profile_gemm(1)
            eng.unet_forward(x, ts, lab, fs)
            gms, gfl, gn = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
            L.mudg_profile_gemm_read(ctypes.byref(gms), ctypes.byref(gfl), ctypes.byref(gn))
            L.mudg_profile_gemm(0)
            print(f"  gemm launches {gn.value}: {gms.value:.2f} ms, {gfl.value / 1e12:.2f} TFLOP -> "
                  f"{gfl.value / gms.value / 1e9:.1f} TFLOP/s", flush=True)
        if os.environ.get("PROBE_VAE", "1") == "1":
            z = torch.randn(2, 4, h, w, device="cuda")
            eng.vae_decode(z)
            torch.cuda.synchronize()
            e0.record()
            eng.vae_decode(z)
            e1.record()
            torch.cuda.synchronize()
            print(f"  vae decode: {e0.elapsed_time(e1) / 2:.2f} ms/frame", flush=True)


if __name__ == "__main__":
    main()
