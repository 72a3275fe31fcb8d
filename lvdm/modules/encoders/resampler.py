"""Drop-in for lvdm/modules/encoders/resampler.py: `Resampler` with the reference's constructor and state-dict layout
(resampler.py:104-129); `forward` is ONE call into libmudg_sm100.so (mudg_resampler_forward: tcgen05 GEMMs, flash
attention over the [x ; latents] keys, LayerNorm, erf GELU) -- SURVEY.md section 8f row 3.  Selected by the YAML
`image_proj_stage_config.target: lvdm.modules.encoders.resampler.Resampler` exactly as in the reference configs."""
from __future__ import annotations

import torch
import torch.nn as nn

from mudg_b200._lib import MudgError


class _PerceiverAttention(nn.Module):          # parameter holder (resampler.py:49-60)
    def __init__(self, dim, dim_head, heads):
        super().__init__()
        inner = dim_head * heads
        self.norm1, self.norm2 = nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)


def _feed_forward(dim, mult):                  # parameter holder (resampler.py:31-37): keys 0 (LN), 1, 3 (Linear)
    inner = int(dim * mult)
    return nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, inner, bias=False), nn.GELU(), nn.Linear(inner, dim, bias=False))


class Resampler(nn.Module):
    def __init__(self, dim=1024, depth=8, dim_head=64, heads=16, num_queries=8, embedding_dim=768, output_dim=1024,
                 ff_mult=4, video_length=None):
        super().__init__()
        if dim_head != 64:
            raise NotImplementedError("mudg_b200 Resampler: the attention kernel is specialised for dim_head == 64")
        self.num_queries, self.video_length = num_queries, video_length
        if video_length is not None:
            num_queries = num_queries * video_length
        self.latents = nn.Parameter(torch.randn(1, num_queries, dim) / dim ** 0.5)
        self.proj_in = nn.Linear(embedding_dim, dim)
        self.proj_out = nn.Linear(dim, output_dim)
        self.norm_out = nn.LayerNorm(output_dim)
        self.layers = nn.ModuleList(
            nn.ModuleList([_PerceiverAttention(dim, dim_head, heads), _feed_forward(dim, ff_mult)]) for _ in range(depth))
        self._n_out, self._out_dim = num_queries, output_dim
        self._engine, self._engine_dirty = None, True
        self._register_load_state_dict_pre_hook(lambda *a, **k: setattr(self, "_engine_dirty", True))

    def _apply(self, fn, *a, **k):             # .cuda() / .to() move the parameters: re-pack lazily
        self._engine_dirty = True
        return super()._apply(fn, *a, **k)

    def engine(self):
        p = self.latents
        if not p.is_cuda:
            raise MudgError("Resampler runs only on a CUDA (B200) device: call .cuda() first; no CPU fallback exists")
        if self._engine is None:
            from mudg_b200.engine import Engine
            self._engine = Engine(None, None, device=p.device.index)
        if self._engine_dirty:
            from mudg_b200.engine import MUDG_RESAMPLER
            self._engine.load_state_dict(self.state_dict(), MUDG_RESAMPLER)
            self._engine_dirty = False
        return self._engine

    @torch.no_grad()
    def forward(self, x):
        y = self.engine().resampler_forward(x, self._n_out, self._out_dim)
        return y if x.dtype == torch.float32 else y.to(x.dtype)
