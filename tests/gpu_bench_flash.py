"""Micro-benchmark of the flash attention kernel on the UNet's self-attention shapes (N=2 CFG batch)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mudg_b200._lib import test_lib, check, ptr, cur_stream   # noqa: E402

def main():
    L = test_lib()
    sweep = [(0, 3, 0)] if len(sys.argv) < 2 else [(p_, s_, sp) for sp in (1, 2) for p_ in (0, 1, 2) for s_ in (0, 1, 3)]
    for poly, stagger, split in sweep:
        check(L.mudg_test_set_knob(b"flash_poly", poly))
        check(L.mudg_test_set_knob(b"flash_stagger", stagger))
        check(L.mudg_test_set_knob(b"flash_split", split))
        print(f"-- flash_split={split} flash_poly={poly} flash_stagger={stagger}", flush=True)
        bench(L)
    check(L.mudg_test_set_knob(b"reset", 0))


def bench(L):
    for name, F, Nq, heads in (("L0 9216x5", 32, 9216, 5), ("L1 2304x10", 32, 2304, 10), ("L2 576x20", 32, 576, 20)):
        C = heads * 64
        qkv = torch.randn(F, Nq, 3 * C, device="cuda").half()
        O = torch.empty(F, Nq, C, device="cuda").half()
        def run():
            check(L.mudg_test_flash(ptr(qkv), 3 * C, ptr(O), C, F, Nq, heads, ctypes.c_void_p(qkv.data_ptr() + 2 * C),
                                    ctypes.c_void_p(qkv.data_ptr() + 4 * C), 3 * C, Nq, F, 1, None, None, 0, 0, 0, 1,
                                    ctypes.c_float(0.125), 0, cur_stream()))
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        fl = 4.0 * Nq * Nq * 64 * heads * F
        print(f"{name:12s} {us:9.1f} us {fl / us / 1e6:7.0f} TFLOP/s", flush=True)

if __name__ == "__main__":
    main()
