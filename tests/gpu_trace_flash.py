"""Debug probe: clock64 time line of CTA 0 of the flash kernel (softmax groups 0/1 and the MMA issuer), level-0 shape."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mudg_b200._lib import test_lib, check, ptr, cur_stream   # noqa: E402


def main():
    L = test_lib()
    F, Nq, heads = 2, 9216, 5
    C = heads * 64
    qkv = torch.randn(F, Nq, 3 * C, device="cuda").half()
    O = torch.empty(F, Nq, C, device="cuda").half()
    def run():
        check(L.mudg_test_flash(ptr(qkv), 3 * C, ptr(O), C, F, Nq, heads, ctypes.c_void_p(qkv.data_ptr() + 2 * C),
                                ctypes.c_void_p(qkv.data_ptr() + 4 * C), 3 * C, Nq, F, 1, None, None, 0, 0, 0, 1,
                                ctypes.c_float(0.125), 0, cur_stream()))
    run(); torch.cuda.synchronize()
    tr = torch.zeros(3, 96, 8, dtype=torch.int64, device="cuda")
    check(L.mudg_test_flash_trace(ptr(tr)))
    run(); torch.cuda.synchronize()
    check(L.mudg_test_flash_trace(None))
    t = tr.cpu()
    t0 = int(t[0, 0, 0])
    rel = lambda v: int(v) - t0 if int(v) else -1
    print("blk | g0: rdy  got  ld   exp  pst | g1: rdy  got  ld   exp  pst | iss: S0   S1   PV0  PV1")
    for b in list(range(0, 14)) + list(range(60, 72)):
        g0 = [rel(t[0, b, e]) for e in range(5)]
        g1 = [rel(t[1, b, e]) for e in range(5)]
        iss = [rel(t[2, b, e]) for e in range(4)]
        print(f"{b:3d} | " + " ".join(f"{v:6d}" for v in g0) + " | " + " ".join(f"{v:6d}" for v in g1) + " | " + " ".join(f"{v:6d}" for v in iss))
    per = (int(t[0, 70, 4]) - int(t[0, 10, 4])) / 60
    print("clocks per kv block (group 0, blocks 10..70):", per)
    for g in (0, 1):
        w = sum(int(t[g, b, 1]) - int(t[g, b, 0]) for b in range(10, 70)) / 60
        ld = sum(int(t[g, b, 2]) - int(t[g, b, 1]) for b in range(10, 70)) / 60
        ex = sum(int(t[g, b, 3]) - int(t[g, b, 2]) for b in range(10, 70)) / 60
        tw = sum(int(t[g, b, 5]) - int(t[g, b, 2]) for b in range(10, 70)) / 60
        print(f"group {g}: max + wait for the MUFU turn {tw:.0f}, exponentials {ex - tw:.0f}")
        ps = sum(int(t[g, b, 4]) - int(t[g, b, 3]) for b in range(10, 70)) / 60
        print(f"group {g}: wait S {w:.0f}  load {ld:.0f}  max+exp {ex:.0f}  o_done+store {ps:.0f}")


if __name__ == "__main__":
    main()
