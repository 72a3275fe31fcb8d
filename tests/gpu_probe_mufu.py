"""Special-function-unit throughput on this GPU: is the packed fp16 exponential (ex2.approx.f16x2 -> two MUFU.EX2.F16) any
cheaper per exponential than ex2.approx.ftz.f32?  Decides whether a packed softmax exponent is worth building."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mudg_b200._lib import test_lib, check, ptr, cur_stream   # noqa: E402


def main():
    L = test_lib()
    iters = 4096
    for threads in (128, 256, 512, 1024):
        ctas = 148
        out = torch.empty(ctas * threads, device="cuda")
        clk = torch.zeros(ctas, dtype=torch.int64, device="cuda")
        row = f"{threads:5d} threads/SM:"
        for mode, name, per_op in ((0, "ex2.f32", 1), (1, "ex2.f16x2", 2), (2, "rcp.f32", 1), (3, "ffma", 1)):
            check(L.mudg_test_mufu_probe(mode, iters, ctas, threads, ptr(out), ptr(clk), cur_stream()))
            torch.cuda.synchronize()
            c = float(clk.float().median())
            warps = threads // 32
            # warp-instructions per SM = warps * iters * 8; results per clock per SM
            ops_per_clk = warps * iters * 8 * 32 * per_op / c
            row += f"  {name} {ops_per_clk:6.1f}/clk/SM"
        print(row, flush=True)


if __name__ == "__main__":
    main()
