"""Drop-in for the two conditioning encoders the shipped configs select from lvdm/modules/encoders/condition.py:
`FrozenOpenCLIPEmbedder` (text, layer "penultimate"; condition.py:174-234) and `FrozenOpenCLIPImageEmbedderV2`
(token-level image features; condition.py:295-372) -- SURVEY.md section 8f row 3.

Same constructor arguments and the same state-dict layout as the reference modules (which hold an `open_clip` CLIP under
`self.model` with `visual` or `transformer` deleted), so a MuDG checkpoint's `cond_stage_model.*` / `embedder.*` entries
load with strict=True.  The modules here only HOLD the parameters; `forward` is one call into libmudg_sm100.so
(mudg_clip_text_forward / mudg_clip_image_forward: tcgen05 GEMMs, fused bias / residual, fp32-softmax attention; the
image path includes the reference's kornia resize + CLIP normalisation as a CUDA kernel).  No CPU fallback.

What stays outside: the pretrained weights (the reference downloads `laion2b_s32b_b79k` through open_clip at construction;
here they arrive with the checkpoint's state dict) and the BPE tokenizer (open_clip.tokenize needs open_clip's vocabulary
file): `forward(text)` uses `open_clip.tokenize` when that package is importable or a `tokenizer` callable given to the
constructor; `encode_with_transformer(tokens)` takes token ids directly."""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn

from mudg_b200._lib import MudgError
from mudg_b200.engine import MUDG_CLIP_IMAGE, MUDG_CLIP_TEXT

# open_clip model_configs/ViT-H-14.json
_ARCH = {
    "ViT-H-14": dict(embed_dim=1024,
                     vision=dict(width=1280, layers=32, heads=16, mlp=5120, image_size=224, patch=14),
                     text=dict(width=1024, layers=24, heads=16, mlp=4096, vocab=49408, ctx=77)),
}


def _arch(arch):
    if hasattr(arch, "items"):                 # a dict / OmegaConf node spelling the towers out (tests use small ones)
        return {k: (_arch(v) if hasattr(v, "items") else v) for k, v in arch.items()}
    if arch not in _ARCH:
        raise NotImplementedError(f"mudg_b200 OpenCLIP towers: unknown arch {arch!r} (known: {sorted(_ARCH)})")
    return _ARCH[arch]


def _zeros(*shape):
    return nn.Parameter(torch.zeros(shape), requires_grad=False)


class _Linear(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.weight, self.bias = _zeros(o, i), _zeros(o)


class _Norm(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.weight, self.bias = nn.Parameter(torch.ones(n), requires_grad=False), _zeros(n)


class _Attention(nn.Module):                   # nn.MultiheadAttention's parameter names
    def __init__(self, width):
        super().__init__()
        self.in_proj_weight, self.in_proj_bias = _zeros(3 * width, width), _zeros(3 * width)
        self.out_proj = _Linear(width, width)


class _Block(nn.Module):                       # open_clip ResidualAttentionBlock
    def __init__(self, width, mlp):
        super().__init__()
        self.ln_1, self.attn, self.ln_2 = _Norm(width), _Attention(width), _Norm(width)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", _Linear(width, mlp)), ("gelu", nn.Identity()), ("c_proj", _Linear(mlp, width))]))


class _Transformer(nn.Module):
    def __init__(self, width, layers, mlp):
        super().__init__()
        self.resblocks = nn.ModuleList(_Block(width, mlp) for _ in range(layers))


class _Visual(nn.Module):                      # open_clip VisionTransformer
    def __init__(self, width, layers, mlp, image_size, patch, embed_dim, **_):
        super().__init__()
        grid = image_size // patch
        self.conv1 = nn.Module()
        self.conv1.weight = _zeros(width, 3, patch, patch)
        self.class_embedding = _zeros(width)
        self.positional_embedding = _zeros(grid * grid + 1, width)
        self.ln_pre = _Norm(width)
        self.transformer = _Transformer(width, layers, mlp)
        self.ln_post = _Norm(width)
        self.proj = _zeros(width, embed_dim)


class _Clip(nn.Module):                        # open_clip CLIP with one of `visual` / `transformer` deleted
    def __init__(self, arch, keep):
        super().__init__()
        t = arch["text"]
        if keep == "visual":
            self.visual = _Visual(embed_dim=arch["embed_dim"], **arch["vision"])
        else:
            self.transformer = _Transformer(t["width"], t["layers"], t["mlp"])
        self.token_embedding = nn.Module()
        self.token_embedding.weight = _zeros(t["vocab"], t["width"])
        self.positional_embedding = _zeros(t["ctx"], t["width"])
        self.ln_final = _Norm(t["width"])
        self.text_projection = _zeros(t["width"], arch["embed_dim"])
        self.logit_scale = _zeros()


class _EngineHolder(nn.Module):
    """Lazily packs the held parameters into a mudg_b200 engine on the parameters' CUDA device; re-packs after
    load_state_dict / .to()."""
    _which = None
    _prefix = ""

    def __init__(self):
        super().__init__()
        self._engine, self._engine_dirty = None, True
        self._register_load_state_dict_pre_hook(lambda *a, **k: setattr(self, "_engine_dirty", True))

    def _apply(self, fn, *a, **k):
        self._engine_dirty = True
        return super()._apply(fn, *a, **k)

    def engine(self):
        p = self.model.positional_embedding
        if not p.is_cuda:
            raise MudgError(f"{type(self).__name__} runs only on a CUDA (B200) device: call .cuda() first; no CPU fallback exists")
        if self._engine is None:
            from mudg_b200.engine import Engine
            self._engine = Engine(None, None, device=p.device.index)
        if self._engine_dirty:
            self._engine.load_state_dict(self.state_dict(), self._which, prefix=self._prefix)
            self._engine_dirty = False
        return self._engine

    def freeze(self):
        self.model = self.model.eval()
        for param in self.parameters():
            param.requires_grad = False


class AbstractEncoder(nn.Module):
    def encode(self, *args, **kwargs):
        raise NotImplementedError


class FrozenOpenCLIPEmbedder(_EngineHolder, AbstractEncoder):
    """OpenCLIP text tower (condition.py:174-234)."""
    LAYERS = ["last", "penultimate"]
    _which = MUDG_CLIP_TEXT
    _prefix = "model."

    def __init__(self, arch="ViT-H-14", version="laion2b_s32b_b79k", device="cuda", max_length=77, freeze=True, layer="last",
                 tokenizer=None):
        super().__init__()
        assert layer in self.LAYERS
        self._cfg = _arch(arch)
        self.model = _Clip(self._cfg, keep="transformer")
        self.device, self.max_length, self.layer = device, max_length, layer
        self.layer_idx = {"last": 0, "penultimate": 1}[layer]
        self.tokenizer = tokenizer
        if freeze:
            self.freeze()

    def _tokenize(self, text):
        if self.tokenizer is not None:
            return self.tokenizer(text)
        try:
            import open_clip
        except ImportError as e:
            raise MudgError("FrozenOpenCLIPEmbedder.forward(text) needs the CLIP BPE tokenizer: install open_clip or pass "
                            "tokenizer=callable(list[str]) -> LongTensor [B, 77]; encode_with_transformer(tokens) takes ids") from e
        return open_clip.tokenize(text)

    def forward(self, text):
        return self.encode_with_transformer(self._tokenize(text).to(self.model.positional_embedding.device))

    @torch.no_grad()
    def encode_with_transformer(self, text):
        t = self._cfg["text"]
        if text.dim() != 2 or text.shape[1] > t["ctx"]:
            raise MudgError(f"tokens {tuple(text.shape)}: expected [B, L <= {t['ctx']}]")
        if text.numel() and (int(text.min()) < 0 or int(text.max()) >= t["vocab"]):
            raise MudgError(f"token id out of range [0, {t['vocab']})")
        return self.engine().clip_text_forward(text.to(self.model.positional_embedding.device), t["width"], t["heads"], self.layer_idx)

    def encode(self, text):
        return self(text)


class FrozenOpenCLIPImageEmbedderV2(_EngineHolder, AbstractEncoder):
    """OpenCLIP vision tower, token-level output (condition.py:295-372)."""
    _which = MUDG_CLIP_IMAGE
    _prefix = "model.visual."

    def __init__(self, arch="ViT-H-14", version="laion2b_s32b_b79k", device="cuda", freeze=True, layer="pooled", antialias=True):
        super().__init__()
        if layer == "penultimate":
            raise NotImplementedError()
        if not antialias:
            raise NotImplementedError("mudg_b200 FrozenOpenCLIPImageEmbedderV2: antialias=False is not built (no shipped config uses it)")
        self._cfg = _arch(arch)
        self.model = _Clip(self._cfg, keep="visual")
        self.device, self.layer, self.antialias = device, layer, antialias
        self.register_buffer("mean", torch.Tensor([0.48145466, 0.4578275, 0.40821073]), persistent=False)
        self.register_buffer("std", torch.Tensor([0.26862954, 0.26130258, 0.27577711]), persistent=False)
        if freeze:
            self.freeze()

    @torch.no_grad()
    def forward(self, image, no_dropout=False):
        return self.encode_with_vision_transformer(image)

    @torch.no_grad()
    def encode_with_vision_transformer(self, x):
        v = self._cfg["vision"]
        tokens = (v["image_size"] // v["patch"]) ** 2 + 1
        y = self.engine().clip_image_forward(x.to(self.model.positional_embedding.device), tokens, v["width"], v["heads"], resize=True)
        return y if x.dtype == torch.float32 else y.to(x.dtype)

    def encode(self, text):
        return self(text)
