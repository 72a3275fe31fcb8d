"""Seeded stand-ins for the once-per-clip conditioning encoders, for runs WITHOUT weights (bench.py's synthetic conditioning,
tests of the driver flow).  The real towers are lvdm/modules/encoders/condition.py (csrc/clip.cu) and resampler.py; they need the
OpenCLIP ViT-H/14 weights a MuDG checkpoint carries.  Point `cond_stage_config.target`,
`img_cond_stage_config.target` and `image_proj_stage_config.target` at these in a YAML to run the sampler with
synthetic conditioning of the right shapes: text [B,77,1024], image tokens [B,257,1280] -> [B,16*T,1024]."""
from __future__ import annotations

import hashlib

import torch
import torch.nn as nn


class SeededTextEmbedder(nn.Module):
    def __init__(self, dim=1024, max_length=77, **_):
        super().__init__()
        self.dim, self.max_length = dim, max_length
        self.register_buffer("anchor", torch.zeros(1), persistent=False)

    def encode(self, text):
        return self(text)

    def forward(self, text):
        outs = []
        for s in text:
            seed = int.from_bytes(hashlib.sha256(s.encode()).digest()[:4], "little")
            g = torch.Generator().manual_seed(seed)
            outs.append(torch.randn(self.max_length, self.dim, generator=g))
        return torch.stack(outs).to(self.anchor.device)


class SeededImageEmbedder(nn.Module):
    def __init__(self, tokens=257, dim=1280, **_):
        super().__init__()
        g = torch.Generator().manual_seed(11)
        self.register_buffer("proj", torch.randn(3, tokens * 4, generator=g) * 0.1, persistent=False)
        self.tokens, self.dim = tokens, dim

    def forward(self, image):
        feat = image.float().mean(dim=(2, 3)) @ self.proj.to(image.device)          # [B, tokens*4]
        base = feat.reshape(image.shape[0], self.tokens, 4)
        return base.repeat(1, 1, self.dim // 4)


class SeededResampler(nn.Module):
    """[B,257,1280] -> [B, num_queries*video_length, output_dim]"""

    def __init__(self, dim=1024, num_queries=16, embedding_dim=1280, output_dim=1024, video_length=16, **_):
        super().__init__()
        g = torch.Generator().manual_seed(12)
        self.n = num_queries * video_length
        self.w = nn.Parameter(torch.randn(embedding_dim, output_dim, generator=g) / embedding_dim ** 0.5, requires_grad=False)
        self.q = nn.Parameter(torch.randn(self.n, output_dim, generator=g), requires_grad=False)

    def forward(self, x):
        pooled = x.float().mean(dim=1) @ self.w                                     # [B, out]
        return torch.nn.functional.layer_norm(pooled[:, None, :] + self.q[None], (self.q.shape[1],))
