"""GPU probe: small-config UNet forward + VAE decode through the C-ABI vs the oracle and the golden vectors."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mudg_oracle as O       # noqa: E402
from mudg_b200.engine import Engine, MUDG_UNET, MUDG_VAE   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def stats(name, out, ref):
    err = (out.float().cpu() - ref.float().cpu()).abs()
    nan = int(torch.isnan(out.float()).sum())
    print(f"{name}: max|d|={float(err[~torch.isnan(err)].max()):.5f} mean|d|={float(err[~torch.isnan(err)].mean()):.6f} "
          f"nans={nan} ref_absmax={float(ref.abs().max()):.3f}", flush=True)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    dev = "cuda"
    g = np.load(os.path.join(GOLD, "unet_small.npz"))
    cfg = O.UNetCfg(model_channels=64, temporal_length=4)
    vcfg = O.VaeCfg(ch=64)
    eng = Engine(dict(in_channels=12, out_channels=4, model_channels=64, num_res_blocks=2, channel_mult=(1, 2, 4, 4),
                      attention_resolutions=(4, 2, 1), num_head_channels=64, context_dim=1024),
                 dict(ch=64, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4, out_ch=3, embed_dim=4))
    if which in ("all", "unet"):
        sd = O.seeded_state_dict(O.unet_param_shapes(cfg), seed=1)
        t0 = time.time()
        eng.load_state_dict(sd, MUDG_UNET)
        print(f"unet weights loaded in {time.time() - t0:.2f}s", flush=True)
        t = lambda k: torch.from_numpy(g[k]).to(dev)
        for ck, yk in (("ctx", "y"), ("ctx2", "y2")):
            eng.set_context(t(ck), T=4)
            y = eng.unet_forward(t("x"), t("ts"), t("lab"), t("fs"))
            torch.cuda.synchronize()
            stats(f"unet_small[{ck}] vs reference golden", y, torch.from_numpy(g[yk]))
        print("workspace bytes:", eng.workspace_bytes(2, 4, 16, 16), "launches:", eng.launch_count(), flush=True)
    if which in ("all", "vae"):
        gv = np.load(os.path.join(GOLD, "vae_small.npz"))
        vsd = O.seeded_state_dict(O.vae_param_shapes(vcfg), seed=2)
        eng.load_state_dict(vsd, MUDG_VAE)
        dec = eng.vae_decode(torch.from_numpy(gv["z"]).to(dev))
        torch.cuda.synchronize()
        stats("vae_small vs reference golden", dec, torch.from_numpy(gv["dec"]))
    if which in ("all", "ddim"):
        tab = O.make_tables(base_scale=0.3)
        sch = O.make_ddim_schedule(tab, 50, "uniform_trailing", 1.0)
        gen = torch.Generator().manual_seed(3)
        x, vc, vu, nz = (torch.randn(2, 4, 4, 16, 16, generator=gen) for _ in range(4))
        vc, vu = vc.half().float(), vu.half().float()
        for index in (49, 20, 0):
            ref_prev, ref_x0 = O.ddim_step(tab, sch, index, x, vc, vu, nz, 7.5, 0.7)
            tt = int(sch.timesteps[index])
            xp, x0 = eng.ddim_step(x.to(dev), vc.to(dev), vu.to(dev), nz.to(dev), cfg_scale=7.5, guidance_rescale=0.7,
                                   sqrt_ac=float(tab.sqrt_alphas_cumprod[tt]), sqrt_1mac=float(tab.sqrt_one_minus_alphas_cumprod[tt]),
                                   rescale=float(sch.scale_arr_prev[index] / sch.scale_arr[index]),
                                   a_prev=float(sch.alphas_prev[index]), sigma=float(sch.sigmas[index]))
            torch.cuda.synchronize()
            stats(f"ddim_step[index={index}] x_prev", xp, ref_prev)
            stats(f"ddim_step[index={index}] pred_x0", x0, ref_x0)


if __name__ == "__main__":
    main()
