"""CPU/GPU fp32 restatement of the OpenCLIP ViT-H/14 towers as the reference drives them.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and oracle/make_golden_clip.py; the product
(mudg_b200/, lvdm/) never imports it.

The reference does not contain this arithmetic: `lvdm/modules/encoders/condition.py` calls the third-party package
`open_clip` (open_clip_torch, requirements.txt pins 2.22.0) for the model and `kornia` (unpinned in requirements.txt) for the resize; neither is
vendored under /root/reference nor installed in this image.  What is restated here is therefore
  * the reference's OWN call sites, line by line:
      FrozenOpenCLIPImageEmbedderV2.encode_with_vision_transformer   condition.py:339-372   -> clip_image_tokens
      FrozenOpenCLIPImageEmbedderV2.preprocess                       condition.py:318-326   -> clip_preprocess
      FrozenOpenCLIPEmbedder.encode_with_transformer /
        text_transformer_forward (layer = "penultimate")             condition.py:214-232   -> clip_text_encode
  * open_clip's published graph for what those call: `VisionTransformer` / `Transformer` / `ResidualAttentionBlock`
    (pre-LN block: x += out_proj(MHA(ln_1 x)); x += c_proj(gelu(c_fc(ln_2 x))), torch.nn.MultiheadAttention with the
    packed in_proj, erf GELU for ViT-H-14, additive causal mask in the text tower), with open_clip's state-dict names.
PINNING: the restatement is checked against an independent implementation of the same published model that IS in this
image -- HuggingFace transformers' CLIPVisionModel / CLIPTextModel (the classes the HF port of the laion2b ViT-H/14
checkpoint loads into) -- on seeded weights, small and full size, by oracle/make_golden_clip.py, which also writes
tests/golden/clip_small.npz from the transformers outputs.  The kornia resize (gaussian anti-alias blur + bicubic,
align_corners=True) has no second implementation here: clip_preprocess is "parity unpinned" (restated from kornia's
published `geometry.transform.resize` / `filters.gaussian_blur2d`) and the tests say so.
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Tuple

import torch
import torch.nn.functional as F

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)      # condition.py:311-312
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


# ------------------------------------------------------------------ parameter layouts (open_clip state-dict names)
def _block_shapes(p: str, width: int, mlp: int) -> Dict[str, Tuple[int, ...]]:
    return {
        p + ".ln_1.weight": (width,), p + ".ln_1.bias": (width,),
        p + ".attn.in_proj_weight": (3 * width, width), p + ".attn.in_proj_bias": (3 * width,),
        p + ".attn.out_proj.weight": (width, width), p + ".attn.out_proj.bias": (width,),
        p + ".ln_2.weight": (width,), p + ".ln_2.bias": (width,),
        p + ".mlp.c_fc.weight": (mlp, width), p + ".mlp.c_fc.bias": (mlp,),
        p + ".mlp.c_proj.weight": (width, mlp), p + ".mlp.c_proj.bias": (width,),
    }


def clip_vision_param_shapes(width=1280, layers=32, mlp=5120, image_size=224, patch=14, embed_dim=1024):
    """`model.visual.*` of open_clip's CLIP (ViT-H-14: width 1280, 32 layers, 16 heads, mlp 5120, 224 / 14)."""
    g = image_size // patch
    s = {"conv1.weight": (width, 3, patch, patch), "class_embedding": (width,), "positional_embedding": (g * g + 1, width),
         "ln_pre.weight": (width,), "ln_pre.bias": (width,), "ln_post.weight": (width,), "ln_post.bias": (width,),
         "proj": (width, embed_dim)}
    for i in range(layers):
        s.update(_block_shapes(f"transformer.resblocks.{i}", width, mlp))
    return s


def clip_text_param_shapes(width=1024, layers=24, mlp=4096, vocab=49408, ctx=77, embed_dim=1024):
    """`model.*` of open_clip's CLIP text tower (ViT-H-14: width 1024, 24 layers, 16 heads)."""
    s = {"token_embedding.weight": (vocab, width), "positional_embedding": (ctx, width), "ln_final.weight": (width,),
         "ln_final.bias": (width,), "text_projection": (width, embed_dim)}
    for i in range(layers):
        s.update(_block_shapes(f"transformer.resblocks.{i}", width, mlp))
    return s


def seeded_clip_state_dict(shapes: Mapping[str, Tuple[int, ...]], seed: int) -> Dict[str, torch.Tensor]:
    """Seeded stand-in weights with the statistics of a trained tower: LN gains near 1, matrices ~ 1/sqrt(fan_in)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in shapes.items():
        if len(shp) == 1 and k.endswith(".weight"):                    # every 1-D "weight" is a LayerNorm gain
            sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif len(shp) == 1:
            sd[k] = 0.05 * torch.randn(shp, generator=g)
        elif k in ("positional_embedding", "token_embedding.weight"):
            sd[k] = 0.3 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            sd[k] = torch.randn(shp, generator=g) / math.sqrt(fan_in)
    return sd


# ------------------------------------------------------------------ open_clip graph
def _residual_attention_block(sd, p, x, heads, attn_mask):
    """open_clip ResidualAttentionBlock.forward on x [B, L, W] (batch-first; the reference permutes to LND and back)."""
    B, L, W = x.shape
    d = W // heads
    h = F.layer_norm(x, (W,), sd[p + ".ln_1.weight"], sd[p + ".ln_1.bias"], 1e-5)
    qkv = F.linear(h, sd[p + ".attn.in_proj_weight"], sd[p + ".attn.in_proj_bias"])
    q, k, v = (t.reshape(B, L, heads, d).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
    s = (q * d ** -0.5) @ k.transpose(-1, -2)                        # nn.MultiheadAttention scales q
    if attn_mask is not None:
        s = s + attn_mask
    o = (s.softmax(dim=-1) @ v).transpose(1, 2).reshape(B, L, W)
    x = x + F.linear(o, sd[p + ".attn.out_proj.weight"], sd[p + ".attn.out_proj.bias"])
    h = F.layer_norm(x, (W,), sd[p + ".ln_2.weight"], sd[p + ".ln_2.bias"], 1e-5)
    h = F.gelu(F.linear(h, sd[p + ".mlp.c_fc.weight"], sd[p + ".mlp.c_fc.bias"]))
    return x + F.linear(h, sd[p + ".mlp.c_proj.weight"], sd[p + ".mlp.c_proj.bias"])


def _n_blocks(sd) -> int:
    n = 0
    while f"transformer.resblocks.{n}.ln_1.weight" in sd:
        n += 1
    return n


def clip_image_tokens(sd: Mapping[str, torch.Tensor], img: torch.Tensor, heads: int) -> torch.Tensor:
    """condition.py:339-372 after `preprocess`: img [B, 3, S, S] (already CLIP-normalised) -> tokens [B, 1 + (S/p)^2, W].
    No ln_post / proj: the V2 embedder returns the transformer output."""
    w = sd["conv1.weight"]
    x = F.conv2d(img, w, None, stride=w.shape[-1])                   # :350
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)       # :351-352
    cls = sd["class_embedding"].to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype, device=x.device)
    x = torch.cat([cls, x], dim=1) + sd["positional_embedding"]      # :355-359
    x = F.layer_norm(x, (x.shape[-1],), sd["ln_pre.weight"], sd["ln_pre.bias"], 1e-5)   # :363
    for i in range(_n_blocks(sd)):                                   # :365-367
        x = _residual_attention_block(sd, f"transformer.resblocks.{i}", x, heads, None)
    return x


def clip_text_encode(sd: Mapping[str, torch.Tensor], tokens: torch.Tensor, heads: int, layer_idx: int = 1) -> torch.Tensor:
    """condition.py:214-232: tokens [B, 77] int64 -> [B, 77, W]; layer_idx 1 = "penultimate" (the last block is skipped)."""
    x = sd["token_embedding.weight"][tokens] + sd["positional_embedding"]          # :215-216
    L = x.shape[1]
    mask = torch.full((L, L), float("-inf"), device=x.device, dtype=x.dtype).triu_(1)   # open_clip build_attention_mask
    for i in range(_n_blocks(sd) - layer_idx):                                     # :223-226
        x = _residual_attention_block(sd, f"transformer.resblocks.{i}", x, heads, mask)
    return F.layer_norm(x, (x.shape[-1],), sd["ln_final.weight"], sd["ln_final.bias"], 1e-5)   # :220


# ------------------------------------------------------------------ kornia resize + normalise (parity unpinned)
def _gaussian_kernel1d(ks: int, sigma: float, device, dtype) -> torch.Tensor:
    x = torch.arange(ks, device=device, dtype=dtype) - ks // 2
    if ks % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2) / (2.0 * sigma * sigma))
    return g / g.sum()


def clip_preprocess(x: torch.Tensor, size: int = 224, antialias: bool = True) -> torch.Tensor:
    """condition.py:318-326: kornia.geometry.resize(x, (224, 224), 'bicubic', align_corners=True, antialias) -> (x + 1) / 2 ->
    kornia.enhance.normalize(mean, std).  kornia's resize blurs before down-scaling only: per axis
    sigma = max((factor - 1) / 2, 0.001), kernel = int(max(4 sigma, 3)) made odd, separable gaussian with reflect padding,
    then F.interpolate(mode='bicubic', align_corners=True)."""
    B, C, H, W = x.shape
    fy, fx = H / size, W / size
    if antialias and max(fy, fx) > 1:
        sy, sx = max((fy - 1.0) / 2.0, 0.001), max((fx - 1.0) / 2.0, 0.001)
        ky, kx = int(max(2.0 * 2 * sy, 3)), int(max(2.0 * 2 * sx, 3))
        ky += 1 - ky % 2
        kx += 1 - kx % 2
        gy = _gaussian_kernel1d(ky, sy, x.device, x.dtype)
        gx = _gaussian_kernel1d(kx, sx, x.device, x.dtype)
        k2 = (gy[:, None] * gx[None, :])[None, None].expand(C, 1, ky, kx)
        x = F.conv2d(F.pad(x, (kx // 2, kx // 2, ky // 2, ky // 2), mode="reflect"), k2, groups=C)
    x = F.interpolate(x, size=(size, size), mode="bicubic", align_corners=True)
    x = (x + 1.0) / 2.0
    mean = torch.tensor(CLIP_MEAN, device=x.device, dtype=x.dtype)[None, :, None, None]
    std = torch.tensor(CLIP_STD, device=x.device, dtype=x.dtype)[None, :, None, None]
    return (x - mean) / std


# ------------------------------------------------------------------ transformers (HF) key mapping, used by the pin script
def open_clip_to_hf(sd: Mapping[str, torch.Tensor], tower: str) -> Dict[str, torch.Tensor]:
    """Rename an open_clip-named tower state dict to transformers' CLIPVisionModel / CLIPTextModel names (the packed
    in_proj split into q / k / v) -- the inverse of the conversion the HF port of the laion2b checkpoints applied."""
    root = "vision_model." if tower == "vision" else "text_model."
    out: Dict[str, torch.Tensor] = {}
    if tower == "vision":
        out[root + "embeddings.patch_embedding.weight"] = sd["conv1.weight"]
        out[root + "embeddings.class_embedding"] = sd["class_embedding"]
        out[root + "embeddings.position_embedding.weight"] = sd["positional_embedding"]
        for wb in ("weight", "bias"):
            out[root + "pre_layrnorm." + wb] = sd["ln_pre." + wb]
            out[root + "post_layernorm." + wb] = sd["ln_post." + wb]
    else:
        out[root + "embeddings.token_embedding.weight"] = sd["token_embedding.weight"]
        out[root + "embeddings.position_embedding.weight"] = sd["positional_embedding"]
        for wb in ("weight", "bias"):
            out[root + "final_layer_norm." + wb] = sd["ln_final." + wb]
    for i in range(_n_blocks(sd)):
        d, s = f"{root}encoder.layers.{i}.", f"transformer.resblocks.{i}."
        for wb in ("weight", "bias"):
            out[d + "layer_norm1." + wb] = sd[s + "ln_1." + wb]
            out[d + "layer_norm2." + wb] = sd[s + "ln_2." + wb]
            for n, part in zip("qkv", sd[s + "attn.in_proj_" + wb].chunk(3, dim=0)):
                out[d + f"self_attn.{n}_proj." + wb] = part
            out[d + "self_attn.out_proj." + wb] = sd[s + "attn.out_proj." + wb]
            out[d + "mlp.fc1." + wb] = sd[s + "mlp.c_fc." + wb]
            out[d + "mlp.fc2." + wb] = sd[s + "mlp.c_proj." + wb]
    return out
