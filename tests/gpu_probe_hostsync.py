"""Where does the HOST block in a sampler step?  Engine-level emulation of the DDIM loop with per-call host timings."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_probe_full import UNET, VAE, gpu_weights          # noqa: E402
from mudg_b200.engine import Engine, MUDG_UNET              # noqa: E402
from mudg_b200.layout import unet_layout                    # noqa: E402


def main():
    eng = Engine(UNET, VAE)
    eng.load_state_dict(gpu_weights(unet_layout(**UNET), 0), MUDG_UNET)
    N, T, h, w = 2, 16, 72, 128
    x = torch.randn(1, 4, T, h, w, device="cuda")
    cc = torch.randn(1, 8, T, h, w, device="cuda")
    ctx = torch.randn(N, 77 + 16 * T, 1024, device="cuda")
    lab = torch.zeros(N, device="cuda", dtype=torch.long)
    fs = torch.full((N,), 10, device="cuda", dtype=torch.long)
    eng.set_context(ctx, T)
    def step(i, rec):
        t0 = time.perf_counter()
        ts = torch.full((N,), 999 - 20 * i, device="cuda", dtype=torch.long)
        xc = torch.cat([x, cc], 1)
        x2 = torch.cat([xc, xc], 0)
        t1 = time.perf_counter()
        v = eng.unet_forward(x2, ts, lab, fs)
        t2 = time.perf_counter()
        nz = torch.randn(x.shape, device="cuda")
        t3 = time.perf_counter()
        xp, _ = eng.ddim_step(x, v[:1], v[1:], nz, cfg_scale=7.5, guidance_rescale=0.7, sqrt_ac=0.5, sqrt_1mac=0.8,
                              rescale=1.0, a_prev=0.5, sigma=0.1)
        t4 = time.perf_counter()
        rec.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
    rec = []
    for i in range(3):
        step(i, rec)
    torch.cuda.synchronize()
    rec = []
    T0 = time.perf_counter()
    for i in range(50):
        step(i, rec)
    Th = time.perf_counter() - T0
    torch.cuda.synchronize()
    Tg = time.perf_counter() - T0
    import statistics as st
    for k, name in enumerate(("cat/full", "unet_forward", "randn", "ddim_step")):
        col = [r[k] * 1e3 for r in rec]
        print(f"{name:13s} host ms: median {st.median(col):8.3f}  max {max(col):8.3f}")
    print(f"50 steps: host loop {Th * 1e3:.0f} ms, GPU done after {Tg * 1e3:.0f} ms")


if __name__ == "__main__":
    main()
