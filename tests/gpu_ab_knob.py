"""A/B of one tuning knob on the captured UNet forward: python tests/gpu_ab_knob.py <knob> <a> <b> [mdm512|mdm1024] [rounds]
The product library has no run-time switches, so the TEST library (same objects + the knob hook) is loaded as the product
library (MUDG_LIB_PATH) and one Engine per knob value captures its own CUDA graph; the two are timed alternately
(a b a b ...) in blocks of 10 forwards so clock drift under the power cap hits both alike."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# (MUDG_TEST_LIB_PATH = an alternative build of the test library: A/B of two BUILDS, one process each)
os.environ["MUDG_LIB_PATH"] = os.environ.get("MUDG_TEST_LIB_PATH") or os.path.join(ROOT, "mudg_b200", "libmudg_sm100_test.so")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch                                                     # noqa: E402
from gpu_probe_full import UNET, VAE, gpu_weights                # noqa: E402
from mudg_b200._lib import check, test_lib                       # noqa: E402
from mudg_b200.engine import Engine, MUDG_UNET                   # noqa: E402
from mudg_b200.layout import unet_layout                         # noqa: E402


def main():
    knob, va, vb = sys.argv[1].encode(), int(sys.argv[2]), int(sys.argv[3])
    cfg = sys.argv[4] if len(sys.argv) > 4 else "mdm1024"
    rounds = int(sys.argv[5]) if len(sys.argv) > 5 else 6
    N, dup, T, h, w = (1, 1, 16, 40, 64) if cfg == "mdm512" else (2, 2, 16, 72, 128)
    L = test_lib()
    sd = gpu_weights(unet_layout(**UNET), 0)
    x = torch.randn(N // dup, 12, T, h, w, device="cuda").repeat(dup, 1, 1, 1, 1).contiguous()
    ctx = torch.randn(N, 77 + 16 * T, 1024, device="cuda")
    ts = torch.full((N,), 500, device="cuda", dtype=torch.long)
    lab = torch.zeros(N, device="cuda", dtype=torch.long)
    fs = torch.full((N,), 10, device="cuda", dtype=torch.long)
    engines = []
    for v in (va, vb):
        check(L.mudg_test_set_knob(knob, v))
        e = Engine(UNET, VAE)
        e.load_state_dict(sd, MUDG_UNET)
        e.set_context(ctx, T)
        for _ in range(4):                                       # eager, capture, two replays -- all under this knob value
            y = e.unet_forward(x, ts, lab, fs, dup=dup)
        torch.cuda.synchronize()
        engines.append((v, e, y.float().clone()))
    check(L.mudg_test_set_knob(b"reset", 0))
    d = float((engines[0][2] - engines[1][2]).abs().max())
    print(f"{cfg}: knob {knob.decode()} {va} vs {vb}: output max|d| = {d:.3g}")
    tot = {va: 0.0, vb: 0.0}
    for r in range(rounds):
        for v, e, _ in engines:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                e.unet_forward(x, ts, lab, fs, dup=dup)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            tot[v] += ms
            print(f"  round {r} {knob.decode()}={v}: {ms:8.3f} ms/forward", flush=True)
    print(f"{cfg}: mean {knob.decode()}={va}: {tot[va] / rounds:.3f} ms   {knob.decode()}={vb}: {tot[vb] / rounds:.3f} ms   "
          f"ratio b/a {tot[vb] / tot[va]:.4f}")


if __name__ == "__main__":
    main()
