"""Sustained-load probe: the N=2 MDM1024 UNet forward back to back for ~20 s, reported in blocks of 10 (power capping
only shows after several seconds; the short probes run at boost clocks)."""
import os, subprocess, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_probe_full import UNET, VAE, gpu_weights          # noqa: E402
from mudg_b200.engine import Engine, MUDG_UNET              # noqa: E402
from mudg_b200.layout import unet_layout                    # noqa: E402


def main():
    blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    dup = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # 2: the CFG form (shared prefix), as the sampler calls it
    eng = Engine(UNET, VAE)
    eng.load_state_dict(gpu_weights(unet_layout(**UNET), 0), MUDG_UNET)
    N, T, h, w = 2, 16, 72, 128
    x = torch.randn(N // dup, 12, T, h, w, device="cuda").repeat(dup, 1, 1, 1, 1).contiguous()
    ctx = torch.randn(N, 77 + 16 * T, 1024, device="cuda")
    ts = torch.full((N,), 500, device="cuda", dtype=torch.long)
    lab = torch.zeros(N, device="cuda", dtype=torch.long)
    fs = torch.full((N,), 10, device="cuda", dtype=torch.long)
    eng.set_context(ctx, T)
    for _ in range(3):
        eng.unet_forward(x, ts, lab, fs, dup=dup)
    torch.cuda.synchronize()
    q = "clocks.sm,power.draw,clocks_event_reasons.sw_power_cap"
    for b in range(blocks):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.unet_forward(x, ts, lab, fs, dup=dup)
        e1.record()
        smi = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", "0"],
                             capture_output=True, text=True).stdout.strip()
        torch.cuda.synchronize()
        print(f"block {b:2d}: {e0.elapsed_time(e1) / 10:7.2f} ms/forward   [{smi}]", flush=True)


if __name__ == "__main__":
    main()
