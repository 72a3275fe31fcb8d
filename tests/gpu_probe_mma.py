"""Debug probe: how fast does one thread get tcgen05.mma instructions of the attention kernel's shapes through the tensor pipe?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mudg_b200._lib import test_lib, check, ptr, cur_stream   # noqa: E402

NAMES = {0: "SS N128 one D", 1: "SS N128 two D", 2: "SS N256 one D", 3: "SS N64 one D", 4: "SS N64 four D",
         5: "TS N64 MN-B one D", 6: "TS N64 MN-B two D", 7: "TS N64 MN-B four D", 8: "SS N128 four D",
         9: "TS N64 K-B one D", 10: "TS N128 K-B one D", 11: "TS N256 K-B one D"}


def main():
    L = test_lib()
    ctas = 148
    out = torch.zeros(ctas, 2, dtype=torch.int64, device="cuda")
    for mode in (0, 1):
        for v in sorted(NAMES):
            reps = 256
            check(L.mudg_test_mma_probe(v, reps, ctas, mode, ptr(out), cur_stream()))
            torch.cuda.synchronize()
            o = out.cpu().float()
            print(f"{'one thread' if mode == 0 else 'elect_one '}  {NAMES[v]:20s} reps {reps:3d}: issue {o[:, 0].mean() / reps:7.1f} clk/MMA   "
                  f"complete {o[:, 1].mean() / reps:7.1f} clk/MMA", flush=True)


if __name__ == "__main__":
    main()
