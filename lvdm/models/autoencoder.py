"""`AutoencoderKL` with the reference's parameter layout (lvdm/models/autoencoder.py:13-107).

decode() -- on the sampler hot path -- is one call into libmudg_sm100.so (Decoder.forward, ae_modules.py:539-578).
encode() is the step *before* the path; it is kept as plain PyTorch ops here until its own
kernels land, and is never used by the denoising loop or the benchmark's timed region.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

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
            sd = {k: v for k, v in self.state_dict().items() if k.startswith(("decoder.", "post_quant_conv."))}
            self._engine.load_state_dict(sd, MUDG_VAE)
            self._engine_dirty = False
        return self._engine

    # ------------------------------------------------------------------ decode (hot path)
    @torch.no_grad()
    def decode(self, z, **kwargs):
        return self.engine().vae_decode(z)

    # ------------------------------------------------------------------ encode (not on the path; PyTorch ops)
    def _p(self, key):
        mod = self
        for seg in key.split("."):
            mod = getattr(mod, seg) if not seg.isdigit() else mod._modules[seg]
        return mod

    def _conv(self, x, p, **kw):
        m = self._p(p)
        return F.conv2d(x, m.weight.to(x.dtype), m.bias.to(x.dtype), **kw)

    def _gn(self, x, p):
        m = self._p(p)
        return F.group_norm(x.float(), 32, m.weight.float(), m.bias.float(), 1e-6).to(x.dtype)

    def _res(self, x, p):
        h = self._conv(F.silu(self._gn(x, p + ".norm1")), p + ".conv1", padding=1)
        h = self._conv(F.silu(self._gn(h, p + ".norm2")), p + ".conv2", padding=1)
        if "nin_shortcut" in self._p(p)._modules:
            x = self._conv(x, p + ".nin_shortcut")
        return x + h

    def _attn(self, x, p):
        b, c, hh, ww = x.shape
        y = self._gn(x, p + ".norm")
        q = self._conv(y, p + ".q").reshape(b, c, -1).transpose(1, 2)
        k = self._conv(y, p + ".k").reshape(b, c, -1).transpose(1, 2)
        v = self._conv(y, p + ".v").reshape(b, c, -1).transpose(1, 2)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        return x + self._conv(o.transpose(1, 2).reshape(b, c, hh, ww), p + ".proj_out")

    @torch.no_grad()
    def encode(self, x, **kwargs):
        """Encoder.forward (ae_modules.py:432-463) + quant_conv (autoencoder.py:97-102)."""
        cfg = self._cfg
        h = self._conv(x, "encoder.conv_in", padding=1)
        n = len(cfg["ch_mult"])
        for lvl in range(n):
            for ib in range(cfg["num_res_blocks"]):
                h = self._res(h, f"encoder.down.{lvl}.block.{ib}")
            if lvl != n - 1:
                h = self._conv(F.pad(h, (0, 1, 0, 1)), f"encoder.down.{lvl}.downsample.conv", stride=2)
        h = self._res(h, "encoder.mid.block_1")
        h = self._attn(h, "encoder.mid.attn_1")
        h = self._res(h, "encoder.mid.block_2")
        h = self._conv(F.silu(self._gn(h, "encoder.norm_out")), "encoder.conv_out", padding=1)
        return DiagonalGaussianDistribution(self._conv(h, "quant_conv"))

    def forward(self, input, sample_posterior=True):
        posterior = self.encode(input)
        z = posterior.sample() if sample_posterior else posterior.mode()
        return self.decode(z), posterior
