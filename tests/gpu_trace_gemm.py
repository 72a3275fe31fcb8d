"""Debug probe: clock64 time line of CTA 0 of the CTA-pair tap-GEMM (producer, MMA issuer, the two epilogue groups)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mudg_b200._lib import test_lib, check, ptr, cur_stream   # noqa: E402

SHAPES = {"res320": (294912, 320, 320, 1, 0), "qkv": (294912, 320, 960, 0, 0), "geglu": (294912, 320, 2560, 0, 1),
          "res1280": (294912, 1280, 320, 1, 0)}


def main():
    L = test_lib()
    for name in (sys.argv[1:] or list(SHAPES)):
        M, Cin, N, res, geglu = SHAPES[name]
        A = torch.randn(1, 1, 1, M, Cin, device="cuda").half()
        Wt = (torch.randn(N, Cin, device="cuda") / Cin ** 0.5).half()
        n_out = N // 2 if geglu else N
        D = torch.empty(1, 1, 1, M, n_out, device="cuda").half()
        R = torch.randn(1, 1, 1, M, n_out, device="cuda").half() if res else None
        bias = torch.randn(N, device="cuda")
        def run():
            check(L.mudg_test_tapgemm(ptr(A), 1, 1, 1, M, Cin, 0, ptr(Wt), N, ptr(D), ptr(R), ptr(bias), None,
                                      ctypes.c_int(1), ctypes.c_int(0), ctypes.c_float(1.0), int(geglu), None, None, 0, cur_stream()))
        run(); torch.cuda.synchronize()
        tr = torch.zeros(4, 64, 8, dtype=torch.int64, device="cuda")
        check(L.mudg_test_gemm_trace(ptr(tr)))
        run(); torch.cuda.synchronize()
        check(L.mudg_test_gemm_trace(None))
        t = tr.cpu()
        t0 = int(t[0, 0, 0])
        rel = lambda v: int(v) - t0 if int(v) else -1
        print(f"== {name}: M={M} K={Cin} N={N} res={res} geglu={geglu}")
        print("tile | prod: start  issued | iss: ready   free  first   last | g0: ready   full  ch0   ch1 | g1: ready   full  ch0   ch1")
        for lt in list(range(0, 4)) + list(range(10, 20)):
            pr = [rel(t[0, lt, e]) for e in range(2)]
            iss = [rel(t[1, lt, e]) for e in range(4)]
            g0 = [rel(t[2, lt, e]) for e in range(4)]
            g1 = [rel(t[3, lt, e]) for e in range(4)]
            print(f"{lt:4d} | " + " ".join(f"{v:7d}" for v in pr) + " | " + " ".join(f"{v:7d}" for v in iss) + " | " +
                  " ".join(f"{v:7d}" for v in g0) + " | " + " ".join(f"{v:7d}" for v in g1))

        per = (int(t[1, 28, 3]) - int(t[1, 8, 3])) / 20
        print("clocks per tile (issuer, tiles 8..28):", per)


if __name__ == "__main__":
    main()
