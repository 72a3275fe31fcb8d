"""`AutoencoderKL` with the reference's parameter layout (lvdm/models/autoencoder.py:13-107).

decode() -- on the sampler hot path -- is one call into libmudg_sm100.so (Decoder.forward, ae_modules.py:539-578).
encode() is the step *before* the path (get_latent_z) and runs on the same kernels
(Encoder.forward, ae_modules.py:432-463); only the posterior sampling stays in Python (CPU-generator RNG parity).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from lvdm.distributions import DiagonalGaussianDistribution
from mudg_b200._lib import MudgError
from mudg_b200.layout import vae_layout
from mudg_b200.paramtree import build_param_tree, mark_dirty_on_load


class AutoencoderKL(nn.Module):
    def __init__(self, ddconfig, lossconfig=None, embed_dim=4, ckpt_path=None, ignore_keys=(), image_key="image",
                 colorize_nlabels=None, monitor=None, test=False, logdir=None, input_dim=4, test_args=None):
        super().__init__()
        dd = dict(ddconfig)
        assert dd["double_z"]
        if dd.get("attn_resolutions"):
            raise NotImplementedError("attn_resolutions must be empty (MuDG first-stage config)")
        self.image_key, self.embed_dim, self.input_dim = image_key, embed_dim, input_dim
        self._cfg = dict(ch=dd["ch"], ch_mult=list(dd["ch_mult"]), num_res_blocks=dd["num_res_blocks"],
                         z_channels=dd["z_channels"], out_ch=dd["out_ch"], in_channels=dd["in_channels"],
                         embed_dim=embed_dim)
        build_param_tree(self, vae_layout(**self._cfg))
        if monitor is not None:
            self.monitor = monitor
        self._engine = None
        mark_dirty_on_load(self)
        if ckpt_path is not None:
            sd = torch.load(ckpt_path, map_location="cpu")
            sd = sd.get("state_dict", sd)
            self.load_state_dict({k: v for k, v in sd.items() if not any(k.startswith(i) for i in ignore_keys)}, strict=False)

    @property
    def device(self):
        return next(self.parameters()).device

    def _apply(self, fn, *a, **k):
        self._engine_dirty = True
        return super()._apply(fn, *a, **k)

    def engine(self):
        p = next(self.parameters())
        if not p.is_cuda:
            raise MudgError("AutoencoderKL.decode runs only on a CUDA (B200) device; no CPU fallback exists")
        if self._engine is None:
            from mudg_b200.engine import Engine
            dummy_unet = dict(in_channels=8, out_channels=4, model_channels=64, num_res_blocks=1, channel_mult=(1,),
                              attention_resolutions=(1,), num_head_channels=64, context_dim=64)
            self._engine = Engine(dummy_unet, self._cfg, device=p.device.index)
        if self._engine_dirty:
            from mudg_b200.engine import MUDG_VAE
            self._engine.load_state_dict(self.state_dict(), MUDG_VAE)
            self._engine_dirty = False
        return self._engine

    # ------------------------------------------------------------------ decode (hot path)
    @torch.no_grad()
    def decode(self, z, **kwargs):
        return self.engine().vae_decode(z)

    # ------------------------------------------------------------------ encode (the step before the path; also native)
    @torch.no_grad()
    def encode(self, x, **kwargs):
        """Encoder.forward (ae_modules.py:432-463) + quant_conv (autoencoder.py:97-102) -> DiagonalGaussianDistribution."""
        return DiagonalGaussianDistribution(self.engine().vae_encode_moments(x))

    def forward(self, input, sample_posterior=True):
        posterior = self.encode(input)
        z = posterior.sample() if sample_posterior else posterior.mode()
        return self.decode(z), posterior
