"""TEST INFRASTRUCTURE (like everything under oracle/): access to the UNCHANGED reference modules (heiheishuang/MuDG `lvdm/`, `utils/`,
`virtual_render/`) next to this repo's drop-in packages of the same names.

The reference tree is copied verbatim by `__graft_entry__.build()` into the git-ignored `baseline/_ref/` (it travels to the
GPU box with gpurun; /root/reference does not exist there).  `reference_modules()` swaps the `lvdm` / `utils` /
`virtual_render` entries of sys.modules so that inside the `with` block imports resolve to the reference; classes
obtained there keep working afterwards (their functions hold their own module globals).  Shims, none of which changes
arithmetic on the path: a `pytorch_lightning` stub (absent in the image), `omegaconf` / `megfile`
stand-ins from mudg_b200.compat, and -- only when asked -- an `xformers.ops.memory_efficient_attention` stub that calls
torch's fused SDPA, to time the reference's *intended* attention path (attention.py:146-206) as well as its einsum
fallback (:101-125).
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_TOPS = ("lvdm", "utils", "virtual_render")
_REF_LOADED: dict = {}


def ref_root():
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(cand, "lvdm", "modules", "networks", "openaimodel3d.py")):
            return cand
    return None


def install_shims(xformers_sdpa: bool = False):
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(torch.nn.Module):
            @property
            def device(self):
                return next(self.parameters()).device
        pl.LightningModule = LightningModule
        pl.seed_everything = lambda s: torch.manual_seed(s)
        util = types.ModuleType("pytorch_lightning.utilities")
        util.rank_zero_only = lambda f: f
        pl.utilities = util
        sys.modules["pytorch_lightning"] = pl
        sys.modules["pytorch_lightning.utilities"] = util
    if xformers_sdpa and "xformers" not in sys.modules:
        xf = types.ModuleType("xformers")
        ops = types.ModuleType("xformers.ops")

        def memory_efficient_attention(q, k, v, attn_bias=None, op=None):
            # the reference passes [(b heads), tokens, d] (attention.py:168-176); torch's fused (flash) kernels want 4-D
            assert attn_bias is None and q.dim() == 3
            return torch.nn.functional.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        ops.memory_efficient_attention = memory_efficient_attention
        xf.ops = ops
        sys.modules["xformers"] = xf
        sys.modules["xformers.ops"] = ops


@contextlib.contextmanager
def reference_modules(xformers_sdpa: bool = False):
    root = ref_root()
    if root is None:
        raise FileNotFoundError("reference tree not found: run __graft_entry__.build() in the build container "
                                "(copies /root/reference/{lvdm,utils,virtual_render,configs} to baseline/_ref/)")
    mine = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in _TOPS}
    saved_path = list(sys.path)
    sys.path[:] = [root] + [p for p in sys.path if os.path.abspath(p or os.getcwd()) != ROOT]
    key = "sdpa" if xformers_sdpa else "einsum"
    had_xf = {k: sys.modules.get(k) for k in ("xformers", "xformers.ops")}
    if not xformers_sdpa:                                  # the einsum fallback needs `import xformers` to FAIL
        for k in ("xformers", "xformers.ops"):
            sys.modules[k] = None
    install_shims(xformers_sdpa)
    sys.modules.update(_REF_LOADED.get(key, {}))
    try:
        yield root
    finally:
        _REF_LOADED[key] = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in _TOPS}
        for k, v in had_xf.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        sys.modules.update(mine)
        sys.path[:] = saved_path


UNET_KW = dict(dropout=0.1, transformer_depth=1, use_linear=True, use_checkpoint=False, temporal_conv=True,
               temporal_attention=True, temporal_selfatt_only=True, use_relative_position=False, use_causal_attention=False,
               addition_attention=True, image_cross_attention=True, default_fs=24, fs_condition=True, class_label_condition=True)


def reference_unet(cfg, sd=None, device="cpu", xformers_sdpa: bool = False):
    """The reference `UNetModel` (openaimodel3d.py:281-628) built like configs/*_infer.yaml:26-56 for the oracle config
    `cfg` (oracle.mudg_oracle.UNetCfg), optionally loaded (strict) with a state dict of the oracle's key layout."""
    with reference_modules(xformers_sdpa):
        from lvdm.modules.networks.openaimodel3d import UNetModel
        import lvdm.modules.attention as A
        assert A.XFORMERS_IS_AVAILBLE == bool(xformers_sdpa)
        kw = dict(in_channels=cfg.in_channels, out_channels=cfg.out_channels, model_channels=cfg.model_channels,
                  attention_resolutions=list(cfg.attention_resolutions), num_res_blocks=cfg.num_res_blocks,
                  channel_mult=list(cfg.channel_mult), num_head_channels=cfg.num_head_channels, context_dim=cfg.context_dim,
                  temporal_length=cfg.temporal_length, **UNET_KW)
        with torch.device(device):
            m = UNetModel(**kw)
    if sd is not None:
        m.load_state_dict(sd, strict=True)
    return m.eval()


def reference_vae(cfg, sd=None, device="cpu"):
    with reference_modules():
        from lvdm.models.autoencoder import AutoencoderKL
        dd = dict(double_z=True, z_channels=cfg.z_channels, resolution=256, in_channels=cfg.in_channels, out_ch=cfg.out_ch,
                  ch=cfg.ch, ch_mult=list(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks, attn_resolutions=[], dropout=0.0)
        with torch.device(device):
            m = AutoencoderKL(ddconfig=dd, lossconfig=dict(target="torch.nn.Identity"), embed_dim=cfg.embed_dim)
    if sd is not None:
        m.load_state_dict(sd, strict=True)
    return m.eval()
