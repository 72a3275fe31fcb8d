"""The reference's own numbers on this box (SURVEY.md section 8d last sentence): the UNCHANGED reference
`UNetModel.forward` (openaimodel3d.py:567-628) and `AutoencoderKL.decode` under eager torch.autocast(fp16) on the B200, with
  * the einsum attention fallback the reference takes when xformers is absent (attention.py:101-125), and
  * `xformers.ops.memory_efficient_attention` stubbed by torch's fused SDPA (the reference's intended path, :146-206),
CUDA-event timed, next to this repo's forward on the same weights and inputs.  Run under gpurun:

    python tests/gpu_ref_timing.py [--out gpurun_out/r2_reference_gpu_timing.json]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FLOPS = {(16, 40, 64): 12.604e12, (16, 72, 128): 52.340e12}
VAE_FLOPS = {(40, 64): 1.5635e12, (72, 128): 5.7543e12}


def timed(fn, reps, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    import gpu_probe_full as PF
    from oracle import refimpl
    from mudg_b200.engine import Engine, MUDG_UNET, MUDG_VAE
    from mudg_b200.layout import unet_layout, vae_layout
    from oracle import mudg_oracle as O
    out_path = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    assert refimpl.ref_root() is not None, "baseline/_ref is missing (run __graft_entry__.build() in the build container)"
    cfg, vcfg = O.UNetCfg(), O.VaeCfg()
    sd = PF.gpu_weights(unet_layout(**PF.UNET), 0)
    vsd = PF.gpu_weights(vae_layout(**PF.VAE), 1)
    res = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "rows": [],
           "note": "eager PyTorch, torch.autocast(float16), no torch.compile; weights seeded random (timing does not depend on values)"}
    eng = Engine(PF.UNET, PF.VAE)
    eng.load_state_dict(sd, MUDG_UNET)
    eng.load_state_dict(vsd, MUDG_VAE)
    models = {}
    for name, sdpa in (("einsum fallback (xformers absent)", False), ("xformers stubbed by torch SDPA", True)):
        m = refimpl.reference_unet(cfg, None, device="meta", xformers_sdpa=sdpa)
        m.load_state_dict(sd, strict=True, assign=True)
        models[name] = m
    for (N, T, h, w) in ((1, 16, 40, 64), (2, 16, 40, 64), (1, 16, 72, 128), (2, 16, 72, 128)):
        g = torch.Generator(device="cuda").manual_seed(3)
        x = torch.randn(N, 12, T, h, w, device="cuda", generator=g)
        ctx = torch.randn(N, 77 + 16 * T, 1024, device="cuda", generator=g)
        ts = torch.full((N,), 499, device="cuda", dtype=torch.long)
        lab = torch.zeros(N, device="cuda", dtype=torch.long)
        fs = torch.full((N,), 10, device="cuda", dtype=torch.long)
        fl = FLOPS[(T, h, w)] * N
        eng.set_context(ctx, T)
        ms = timed(lambda: eng.unet_forward(x, ts, lab, fs), 5, warm=3)
        row = {"op": "UNetModel.forward", "shape": [N, 12, T, h, w], "algorithmic_tflop": fl / 1e12,
               "mudg_b200": {"ms": ms, "tflops": fl / ms / 1e9}}
        for name, m in models.items():
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()

            def fwd():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                    return m(x, ts, c_label=lab, context=ctx, fs=fs)
            try:
                ms_r = timed(fwd, 3, warm=1)
                row[name] = {"ms": ms_r, "tflops": fl / ms_r / 1e9, "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
                             "speedup_of_mudg_b200": ms_r / ms}
            except torch.cuda.OutOfMemoryError as e:
                row[name] = {"error": "CUDA OOM: " + str(e)[:100]}
                torch.cuda.empty_cache()
        print(json.dumps(row), flush=True)
        res["rows"].append(row)
    # VAE decode, per frame (decode_core with perframe_ae loops over frames, ddpm3d.py:646-667)
    vae = refimpl.reference_vae(vcfg, None, device="meta")
    vae.load_state_dict(vsd, strict=True, assign=True)
    for (h, w) in ((40, 64), (72, 128)):
        z = torch.randn(1, 4, h, w, device="cuda")
        fl = VAE_FLOPS[(h, w)]
        ms = timed(lambda: eng.vae_decode(z), 4, warm=2)

        def dec():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                return vae.decode(z)
        ms_r = timed(dec, 3, warm=1)
        row = {"op": "AutoencoderKL.decode (1 frame)", "latent": [1, 4, h, w], "algorithmic_tflop": fl / 1e12,
               "mudg_b200": {"ms": ms, "tflops": fl / ms / 1e9},
               "reference eager autocast-fp16": {"ms": ms_r, "tflops": fl / ms_r / 1e9, "speedup_of_mudg_b200": ms_r / ms}}
        print(json.dumps(row), flush=True)
        res["rows"].append(row)
    # derived: the reference's frames/s on the BASELINE configs if its sampler loop cost nothing else
    by = {(tuple(r["shape"]) if "shape" in r else tuple(r["latent"])): r for r in res["rows"]}
    try:
        u = by[(2, 12, 16, 72, 128)]; d = by[(1, 4, 72, 128)]
        for name in models:
            if "ms" in u.get(name, {}):
                clip_s = (50 * u[name]["ms"] + 16 * d["reference eager autocast-fp16"]["ms"]) / 1e3
                res.setdefault("derived", {})[f"mdm1024 cfg clip, reference {name}"] = {"s_per_clip": clip_s, "frames_per_s": 16 / clip_s}
        clip_s = (50 * u["mudg_b200"]["ms"] + 16 * d["mudg_b200"]["ms"]) / 1e3
        res.setdefault("derived", {})["mdm1024 cfg clip, mudg_b200 (forward + decode only)"] = {"s_per_clip": clip_s, "frames_per_s": 16 / clip_s}
    except KeyError:
        pass
    print(json.dumps(res.get("derived", {}), indent=1), flush=True)
    if out_path:
        os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
        with open(out_path, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
