"""Generate tests/golden/host_small.npz: the pure-tensor host helpers of the sampler path, run from the UNCHANGED
reference classes -- DDPM.q_sample / predict_start_from_z_and_v / predict_eps_from_z_and_v (ddpm3d.py:239-262) on the
shipped schedule (zero-terminal-SNR linear, v-parameterisation) and DiagonalGaussianDistribution (distributions.py:24-61)
with the CPU generator.  Build container only:  python oracle/make_golden_host.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "host_small.npz")


def main():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
    MG = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(MG)
    MG.install_shims()
    sys.path[:] = [REF] + [p for p in sys.path if os.path.abspath(p or os.getcwd()) not in (ROOT, HERE)]
    for m in [m for m in sys.modules if m.split(".")[0] in ("lvdm", "utils")]:
        del sys.modules[m]
    from lvdm.models.ddpm3d import DDPM
    from lvdm.distributions import DiagonalGaussianDistribution
    import lvdm.models.ddpm3d as _m
    assert _m.__file__.startswith(REF)
    torch.set_grad_enabled(False)
    ident = MG.to_attr(dict(target="torch.nn.Identity", params=dict(temporal_length=4)))
    import torch.nn as _nn
    _nn.Identity.__init__ = lambda self, *a, **k: _nn.Module.__init__(self)      # accept (and ignore) the params
    ddpm = DDPM(unet_config=ident, timesteps=1000, linear_start=0.00085, linear_end=0.012, parameterization="v",
                rescale_betas_zero_snr=True, use_ema=False, conditioning_key=None, image_size=[8, 8], channels=4)
    g = torch.Generator().manual_seed(17)
    x0, nz, v = (torch.randn(3, 4, 2, 5, 6, generator=g) for _ in range(3))
    t = torch.tensor([0, 499, 999], dtype=torch.long)
    out = dict(x0=x0.numpy(), noise=nz.numpy(), v=v.numpy(), t=t.numpy(),
               q_sample=ddpm.q_sample(x0, t, nz).numpy(),
               pred_start=ddpm.predict_start_from_z_and_v(x0, t, v).numpy(),
               pred_eps=ddpm.predict_eps_from_z_and_v(x0, t, v).numpy())
    mom = torch.randn(2, 8, 4, 6, generator=g) * 3
    mom[0, 4:, 0, 0] = torch.tensor([-40.0, 25.0, 0.0, 1.0])            # logvar clamp limits
    d = DiagonalGaussianDistribution(mom)
    torch.manual_seed(5)
    out.update(moments=mom.numpy(), sample=d.sample().numpy(), mode=d.mode().numpy())
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
