"""Generate tests/golden/schedules.npz: every host-side schedule helper of the path, run from the UNCHANGED reference
(lvdm/models/utils_diffusion.py) over all variants the samplers can be called with.  Build container only:
    python oracle/make_golden_schedules.py
"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "schedules.npz")


def main():
    spec = importlib.util.spec_from_file_location("ref_utils_diffusion", "/root/reference/lvdm/models/utils_diffusion.py")
    R = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(R)
    out = {}
    for sched in ("linear", "cosine", "sqrt_linear", "sqrt"):
        out[f"betas_{sched}"] = np.asarray(R.make_beta_schedule(sched, 1000, linear_start=0.00085, linear_end=0.012), dtype=np.float64)
    betas = R.make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012)
    out["betas_zero_snr"] = np.asarray(R.rescale_zero_terminal_snr(betas), dtype=np.float64)
    ac = torch.tensor(np.cumprod(1.0 - out["betas_zero_snr"], axis=0), dtype=torch.float32)
    for method in ("uniform", "quad", "uniform_trailing"):
        for S in (50, 25, 7):
            ts = R.make_ddim_timesteps(method, S, 1000, verbose=False)
            out[f"ts_{method}_{S}"] = np.asarray(ts)
            for eta in (0.0, 1.0):
                sig, a, ap = R.make_ddim_sampling_parameters(ac.cpu(), ts, eta, verbose=False)
                out[f"sig_{method}_{S}_{eta}"] = np.asarray(sig, dtype=np.float64)
                out[f"a_{method}_{S}_{eta}"] = np.asarray(a, dtype=np.float64)
                out[f"ap_{method}_{S}_{eta}"] = np.asarray(ap, dtype=np.float64)
    t = torch.tensor([0, 1, 19, 500, 999], dtype=torch.long)
    for dim in (320, 64, 7):
        out[f"temb_{dim}"] = R.timestep_embedding(t, dim).numpy()
    out["temb_repeat"] = R.timestep_embedding(t, 8, repeat_only=True).numpy()
    g = torch.Generator().manual_seed(2)
    cfg, txt = torch.randn(2, 4, 3, 5, 6, generator=g), torch.randn(2, 4, 3, 5, 6, generator=g)
    out["rescale_in_cfg"], out["rescale_in_txt"] = cfg.numpy(), txt.numpy()
    out["rescale_out"] = R.rescale_noise_cfg(cfg, txt, guidance_rescale=0.7).numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, len(out), "arrays")


if __name__ == "__main__":
    main()
