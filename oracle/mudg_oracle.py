"""CPU oracle for the MuDG sampler hot path -- TEST INFRASTRUCTURE ONLY.

This file is a functional (state-dict in, tensor out) fp32 restatement of the
reference's algorithm for the path BASELINE.json names:

    DDIMSampler.sample -> LatentVisualDiffusion.apply_model
        -> openaimodel3d.UNetModel.forward -> AutoencoderKL.decode

It is the checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it; the product
(`mudg_b200`, `lvdm`) never does.

Pinning: the reference ships no tests / golden vectors, so the
oracle is pinned against the *unchanged reference modules imported from
/root/reference* by oracle/make_golden.py, which loads the oracle's seeded
state dict into the reference classes with strict=True (pins key names and
shapes) and stores the reference's outputs under tests/golden/.  tests/
test_oracle_golden.py replays them.  The arithmetic lives in torch==2.0.0 in
the reference's requirements.txt (un-vendored third party); here it runs on the
torch in this image.

Every function cites the reference file:line it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# ----------------------------------------------------------------------------
# configs
# ----------------------------------------------------------------------------
@dataclass
class UNetCfg:
    """Mirror of unet_config.params in configs/stage2-1024_mdm_waymo_infer.yaml:26-56."""
    in_channels: int = 12
    out_channels: int = 4
    model_channels: int = 320
    attention_resolutions: Sequence[int] = (4, 2, 1)
    num_res_blocks: int = 2
    channel_mult: Sequence[int] = (1, 2, 4, 4)
    num_head_channels: int = 64
    context_dim: int = 1024
    temporal_length: int = 16
    init_attn_heads: int = 8            # openaimodel3d.py:408 (hard-coded n_heads=8)
    text_context_len: int = 77          # attention.py:45

    @property
    def time_embed_dim(self) -> int:
        return 4 * self.model_channels


@dataclass
class VaeCfg:
    """Mirror of first_stage_config.params.ddconfig (infer yaml :63-77)."""
    ch: int = 128
    ch_mult: Sequence[int] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    z_channels: int = 4
    out_ch: int = 3
    in_channels: int = 3
    embed_dim: int = 4


@dataclass
class BlockPlan:
    """One TimestepEmbedSequential, flattened: list of (kind, prefix, meta)."""
    layers: List[Tuple[str, str, dict]] = field(default_factory=list)


def unet_plan(cfg: UNetCfg):
    """Enumerate the UNet's blocks exactly as the constructor does
    (openaimodel3d.py:398-565).  Returns (input_blocks, middle, output_blocks)."""
    mc = cfg.model_channels
    inputs: List[BlockPlan] = []
    inputs.append(BlockPlan([("conv", "input_blocks.0.0", dict(cin=cfg.in_channels, cout=mc))]))
    chans = [mc]
    ch, ds = mc, 1
    idx = 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            bp = BlockPlan()
            bp.layers.append(("res", f"input_blocks.{idx}.0", dict(cin=ch, cout=mult * mc, tconv=True)))
            ch = mult * mc
            if ds in cfg.attention_resolutions:
                heads = ch // cfg.num_head_channels
                bp.layers.append(("spatial", f"input_blocks.{idx}.1", dict(ch=ch, heads=heads)))
                bp.layers.append(("temporal", f"input_blocks.{idx}.2", dict(ch=ch, heads=heads, linear=True)))
            inputs.append(bp)
            chans.append(ch)
            idx += 1
        if level != len(cfg.channel_mult) - 1:
            inputs.append(BlockPlan([("down", f"input_blocks.{idx}.0", dict(ch=ch))]))
            chans.append(ch)
            idx += 1
            ds *= 2
    heads = ch // cfg.num_head_channels
    middle = BlockPlan([
        ("res", "middle_block.0", dict(cin=ch, cout=ch, tconv=True)),
        ("spatial", "middle_block.1", dict(ch=ch, heads=heads)),
        ("temporal", "middle_block.2", dict(ch=ch, heads=heads, linear=True)),
        ("res", "middle_block.3", dict(cin=ch, cout=ch, tconv=True)),
    ])
    outputs: List[BlockPlan] = []
    oidx = 0
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            bp = BlockPlan()
            li = 0
            bp.layers.append(("res", f"output_blocks.{oidx}.{li}", dict(cin=ch + ich, cout=mc * mult, tconv=True)))
            li += 1
            ch = mc * mult
            if ds in cfg.attention_resolutions:
                heads = ch // cfg.num_head_channels
                bp.layers.append(("spatial", f"output_blocks.{oidx}.{li}", dict(ch=ch, heads=heads)))
                li += 1
                bp.layers.append(("temporal", f"output_blocks.{oidx}.{li}", dict(ch=ch, heads=heads, linear=True)))
                li += 1
            if level and i == cfg.num_res_blocks:
                bp.layers.append(("up", f"output_blocks.{oidx}.{li}", dict(ch=ch)))
                ds //= 2
            outputs.append(bp)
            oidx += 1
    return inputs, middle, outputs


# ----------------------------------------------------------------------------
# parameter shapes (state-dict key layout == the reference's; pinned by
# make_golden.py through load_state_dict(strict=True))
# ----------------------------------------------------------------------------
def _attn_shapes(p: str, dim: int, ctx_dim: Optional[int], image_ca: bool) -> Dict[str, Tuple[int, ...]]:
    kv = dim if ctx_dim is None else ctx_dim
    s = {f"{p}.to_q.weight": (dim, dim), f"{p}.to_k.weight": (dim, kv), f"{p}.to_v.weight": (dim, kv),
         f"{p}.to_out.0.weight": (dim, dim), f"{p}.to_out.0.bias": (dim,)}
    if image_ca:
        s[f"{p}.to_k_ip.weight"] = (dim, kv)
        s[f"{p}.to_v_ip.weight"] = (dim, kv)
    return s


def _tblock_shapes(p: str, dim: int, ctx_dim: Optional[int], image_ca: bool):
    s = {}
    s.update(_attn_shapes(f"{p}.attn1", dim, None, False))
    s.update(_attn_shapes(f"{p}.attn2", dim, ctx_dim, image_ca))
    s[f"{p}.ff.net.0.proj.weight"] = (8 * dim, dim)
    s[f"{p}.ff.net.0.proj.bias"] = (8 * dim,)
    s[f"{p}.ff.net.2.weight"] = (dim, 4 * dim)
    s[f"{p}.ff.net.2.bias"] = (dim,)
    for n in ("norm1", "norm2", "norm3"):
        s[f"{p}.{n}.weight"] = (dim,)
        s[f"{p}.{n}.bias"] = (dim,)
    return s


def _mlp_shapes(p: str, mc: int, ted: int):
    return {f"{p}.0.weight": (ted, mc), f"{p}.0.bias": (ted,), f"{p}.2.weight": (ted, ted), f"{p}.2.bias": (ted,)}


def unet_param_shapes(cfg: UNetCfg) -> Dict[str, Tuple[int, ...]]:
    mc, ted = cfg.model_channels, cfg.time_embed_dim
    s: Dict[str, Tuple[int, ...]] = {}
    s.update(_mlp_shapes("time_embed", mc, ted))
    s.update(_mlp_shapes("class_embed", mc, ted))
    s.update(_mlp_shapes("fps_embedding", mc, ted))
    inputs, middle, outputs = unet_plan(cfg)

    def add(kind, p, m):
        if kind == "conv":
            s[f"{p}.weight"] = (m["cout"], m["cin"], 3, 3)
            s[f"{p}.bias"] = (m["cout"],)
        elif kind == "res":
            ci, co = m["cin"], m["cout"]
            s[f"{p}.in_layers.0.weight"] = (ci,); s[f"{p}.in_layers.0.bias"] = (ci,)
            s[f"{p}.in_layers.2.weight"] = (co, ci, 3, 3); s[f"{p}.in_layers.2.bias"] = (co,)
            s[f"{p}.emb_layers.1.weight"] = (co, ted); s[f"{p}.emb_layers.1.bias"] = (co,)
            s[f"{p}.out_layers.0.weight"] = (co,); s[f"{p}.out_layers.0.bias"] = (co,)
            s[f"{p}.out_layers.3.weight"] = (co, co, 3, 3); s[f"{p}.out_layers.3.bias"] = (co,)
            if ci != co:
                s[f"{p}.skip_connection.weight"] = (co, ci, 1, 1); s[f"{p}.skip_connection.bias"] = (co,)
            if m["tconv"]:
                for j, ci_ in ((1, 2), (2, 3), (3, 3), (4, 3)):     # conv1 has no Dropout -> index 2
                    q = f"{p}.temopral_conv.conv{j}"             # sic: reference typo openaimodel3d.py:190
                    s[f"{q}.0.weight"] = (co,); s[f"{q}.0.bias"] = (co,)
                    s[f"{q}.{ci_}.weight"] = (co, co, 3, 1, 1); s[f"{q}.{ci_}.bias"] = (co,)
        elif kind == "spatial":
            ch = m["ch"]
            s[f"{p}.norm.weight"] = (ch,); s[f"{p}.norm.bias"] = (ch,)
            s[f"{p}.proj_in.weight"] = (ch, ch); s[f"{p}.proj_in.bias"] = (ch,)
            s.update(_tblock_shapes(f"{p}.transformer_blocks.0", ch, cfg.context_dim, True))
            s[f"{p}.proj_out.weight"] = (ch, ch); s[f"{p}.proj_out.bias"] = (ch,)
        elif kind == "temporal":
            ch = m["ch"]
            inner = m.get("inner", ch)
            s[f"{p}.norm.weight"] = (ch,); s[f"{p}.norm.bias"] = (ch,)
            if m["linear"]:
                s[f"{p}.proj_in.weight"] = (inner, ch); s[f"{p}.proj_out.weight"] = (ch, inner)
            else:                                                    # Conv1d k=1 (attention.py:491,517)
                s[f"{p}.proj_in.weight"] = (inner, ch, 1); s[f"{p}.proj_out.weight"] = (ch, inner, 1)
            s[f"{p}.proj_in.bias"] = (inner,); s[f"{p}.proj_out.bias"] = (ch,)
            s.update(_tblock_shapes(f"{p}.transformer_blocks.0", inner, None, False))
        elif kind == "down":
            s[f"{p}.op.weight"] = (m["ch"], m["ch"], 3, 3); s[f"{p}.op.bias"] = (m["ch"],)
        elif kind == "up":
            s[f"{p}.conv.weight"] = (m["ch"], m["ch"], 3, 3); s[f"{p}.conv.bias"] = (m["ch"],)

    for bp in inputs:
        for l in bp.layers:
            add(*l)
    # init_attn: TemporalTransformer(mc, n_heads=8, d_head=num_head_channels), Conv1d proj (openaimodel3d.py:404-414)
    add("temporal", "init_attn.0", dict(ch=mc, heads=cfg.init_attn_heads,
                                        inner=cfg.init_attn_heads * cfg.num_head_channels, linear=False))
    for l in middle.layers:
        add(*l)
    for bp in outputs:
        for l in bp.layers:
            add(*l)
    s["out.0.weight"] = (mc,); s["out.0.bias"] = (mc,)
    s["out.2.weight"] = (cfg.out_channels, mc, 3, 3); s["out.2.bias"] = (cfg.out_channels,)
    return s


def vae_param_shapes(cfg: VaeCfg, decoder_only: bool = False) -> Dict[str, Tuple[int, ...]]:
    """Keys of AutoencoderKL (autoencoder.py:27-32) + Decoder/Encoder (ae_modules.py:364-537)."""
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(p, co, ci, k):
        s[f"{p}.weight"] = (co, ci, k, k); s[f"{p}.bias"] = (co,)

    def norm(p, c):
        s[f"{p}.weight"] = (c,); s[f"{p}.bias"] = (c,)

    def res(p, ci, co):
        norm(f"{p}.norm1", ci); conv(f"{p}.conv1", co, ci, 3)
        norm(f"{p}.norm2", co); conv(f"{p}.conv2", co, co, 3)
        if ci != co:
            conv(f"{p}.nin_shortcut", co, ci, 1)

    def attn(p, c):
        norm(f"{p}.norm", c)
        for n in ("q", "k", "v", "proj_out"):
            conv(f"{p}.{n}", c, c, 1)

    nres = len(cfg.ch_mult)
    # decoder
    block_in = cfg.ch * cfg.ch_mult[-1]
    conv("decoder.conv_in", block_in, cfg.z_channels, 3)
    res("decoder.mid.block_1", block_in, block_in)
    attn("decoder.mid.attn_1", block_in)
    res("decoder.mid.block_2", block_in, block_in)
    for lvl in reversed(range(nres)):
        block_out = cfg.ch * cfg.ch_mult[lvl]
        for ib in range(cfg.num_res_blocks + 1):
            res(f"decoder.up.{lvl}.block.{ib}", block_in, block_out)
            block_in = block_out
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", block_in, block_in, 3)
    norm("decoder.norm_out", block_in)
    conv("decoder.conv_out", cfg.out_ch, block_in, 3)
    conv("post_quant_conv", cfg.z_channels, cfg.embed_dim, 1)
    if decoder_only:
        return s
    # encoder
    conv("encoder.conv_in", cfg.ch, cfg.in_channels, 3)
    in_mult = (1,) + tuple(cfg.ch_mult)
    block_in = cfg.ch
    for lvl in range(nres):
        block_in = cfg.ch * in_mult[lvl]
        block_out = cfg.ch * cfg.ch_mult[lvl]
        for ib in range(cfg.num_res_blocks):
            res(f"encoder.down.{lvl}.block.{ib}", block_in, block_out)
            block_in = block_out
        if lvl != nres - 1:
            conv(f"encoder.down.{lvl}.downsample.conv", block_in, block_in, 3)
    res("encoder.mid.block_1", block_in, block_in)
    attn("encoder.mid.attn_1", block_in)
    res("encoder.mid.block_2", block_in, block_in)
    norm("encoder.norm_out", block_in)
    conv("encoder.conv_out", 2 * cfg.z_channels, block_in, 3)
    conv("quant_conv", 2 * cfg.embed_dim, 2 * cfg.z_channels, 1)
    return s


def seeded_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int, dtype=torch.float32) -> SD:
    """Deterministic synthetic weights (no checkpoint is reachable).

    Matrices/convs ~ N(0, 1/fan_in) scaled so activations stay O(1); norm weights
    ~ 1 + 0.1 N(0,1); biases ~ 0.02 N(0,1).  Nothing is left exactly zero, so the
    reference's zero-initialised layers (App. D #1) cannot make parity vacuous.
    Values depend only on (key order, shape, seed)."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for k, shp in shapes.items():
        if len(shp) == 1:
            if k.endswith(".weight"):
                v = 1.0 + 0.1 * torch.randn(shp, generator=g)
            else:
                v = 0.02 * torch.randn(shp, generator=g)
        else:
            fan_in = int(np.prod(shp[1:]))
            v = torch.randn(shp, generator=g) / math.sqrt(fan_in)
        sd[k] = v.to(dtype)
    return sd


# ----------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------
def timestep_embedding(t: Tensor, dim: int, max_period: float = 10000.0) -> Tensor:
    """[cos | sin] sinusoid, freqs = exp(-ln(max_period) * i / half)  (utils_diffusion.py:8-28)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _lin(sd: SD, p: str, x: Tensor, bias: bool = True) -> Tensor:
    return F.linear(x, sd[f"{p}.weight"], sd[f"{p}.bias"] if bias else None)


def _gn(sd: SD, p: str, x: Tensor, eps: float) -> Tensor:
    return F.group_norm(x, 32, sd[f"{p}.weight"], sd[f"{p}.bias"], eps)


def _ln(sd: SD, p: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[f"{p}.weight"], sd[f"{p}.bias"], 1e-5)


def _mlp(sd: SD, p: str, x: Tensor) -> Tensor:
    return _lin(sd, f"{p}.2", F.silu(_lin(sd, f"{p}.0", x)))


# Scores of one attention call are materialised like the reference's einsum path does; above this many bytes the batch
# (frames) is processed in slices -- same arithmetic per (frame, head), only the peak memory changes (a full-size level-0
# call is 32 frames x 5 heads x 9216^2 fp32 = 54 GB at once).  tests/test_oracle_golden.py pins sliced == unsliced.
ATTN_SCORE_BYTES_MAX = 2 << 30


def _softmax_attend(q: Tensor, k: Tensor, v: Tensor, heads: int) -> Tensor:
    """einsum path of CrossAttention.forward (attention.py:101-125): (b h) n d split, scale d^-0.5."""
    b, n, c = q.shape
    d = c // heads

    def split(t):
        return t.reshape(t.shape[0], t.shape[1], heads, d).permute(0, 2, 1, 3)

    def attend(qq, kk, vv):
        qh, kh, vh = split(qq), split(kk), split(vv)
        sim = torch.matmul(qh, kh.transpose(-1, -2)) * (d ** -0.5)
        o = torch.matmul(sim.softmax(dim=-1), vh)
        return o.permute(0, 2, 1, 3).reshape(qq.shape[0], n, c)

    per_frame = heads * n * k.shape[1] * q.element_size()
    step = max(1, int(ATTN_SCORE_BYTES_MAX // max(1, per_frame)))
    if step >= b:
        return attend(q, k, v)
    return torch.cat([attend(q[i:i + step], k[i:i + step], v[i:i + step]) for i in range(0, b, step)], dim=0)


def cross_attention(sd: SD, p: str, x: Tensor, heads: int, context: Optional[Tensor],
                    image_ca: bool, text_len: int = 77) -> Tensor:
    """CrossAttention.forward (attention.py:81-144).  context None -> self-attention;
    image_ca -> text (first 77 tokens) and image (rest) branches with separate
    softmaxes, summed with scale 1.0 (:89-94,129-142)."""
    q = _lin(sd, f"{p}.to_q", x, bias=False)
    if context is None:
        k = _lin(sd, f"{p}.to_k", x, bias=False)
        v = _lin(sd, f"{p}.to_v", x, bias=False)
        out = _softmax_attend(q, k, v, heads)
    else:
        ctx_t, ctx_i = context[:, :text_len], context[:, text_len:]
        out = _softmax_attend(q, _lin(sd, f"{p}.to_k", ctx_t, False), _lin(sd, f"{p}.to_v", ctx_t, False), heads)
        if image_ca:
            out = out + _softmax_attend(q, _lin(sd, f"{p}.to_k_ip", ctx_i, False),
                                        _lin(sd, f"{p}.to_v_ip", ctx_i, False), heads)
    return _lin(sd, f"{p}.to_out.0", out)


def feed_forward(sd: SD, p: str, x: Tensor) -> Tensor:
    """GEGLU FF (attention.py:579-606): proj -> chunk(value, gate) -> value * gelu_erf(gate) -> Linear."""
    val, gate = _lin(sd, f"{p}.net.0.proj", x).chunk(2, dim=-1)
    return _lin(sd, f"{p}.net.2", val * F.gelu(gate))


def transformer_block(sd: SD, p: str, x: Tensor, heads: int, context: Optional[Tensor], image_ca: bool) -> Tensor:
    """BasicTransformerBlock._forward (attention.py:392-400), disable_self_attn=False."""
    x = cross_attention(sd, f"{p}.attn1", _ln(sd, f"{p}.norm1", x), heads, None, False) + x
    x = cross_attention(sd, f"{p}.attn2", _ln(sd, f"{p}.norm2", x), heads, context, image_ca) + x
    x = feed_forward(sd, f"{p}.ff", _ln(sd, f"{p}.norm3", x)) + x
    return x


def spatial_transformer(sd: SD, p: str, x: Tensor, heads: int, context: Tensor) -> Tensor:
    """SpatialTransformer.forward, use_linear=True (attention.py:451-467). x: [(b t), c, h, w]."""
    bt, c, h, w = x.shape
    y = _gn(sd, f"{p}.norm", x, 1e-6)
    y = y.permute(0, 2, 3, 1).reshape(bt, h * w, c)
    y = _lin(sd, f"{p}.proj_in", y)
    y = transformer_block(sd, f"{p}.transformer_blocks.0", y, heads, context, True)
    y = _lin(sd, f"{p}.proj_out", y)
    return y.reshape(bt, h, w, c).permute(0, 3, 1, 2) + x


def temporal_transformer(sd: SD, p: str, x: Tensor, heads: int, b: int) -> Tensor:
    """TemporalTransformer.forward (attention.py:529-576) on x: [(b t), c, h, w];
    GroupNorm statistics span (C/32, T, H, W) (5-D input, :532); both attentions are
    self-attention over T (only_self_att, :504-505,551-554)."""
    bt, c, h, w = x.shape
    t = bt // b
    x5 = x.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4)            # b c t h w
    y = _gn(sd, f"{p}.norm", x5, 1e-6)
    y = y.permute(0, 3, 4, 2, 1).reshape(b * h * w, t, c)            # (b h w) t c
    wi = sd[f"{p}.proj_in.weight"]
    y = F.linear(y, wi.reshape(wi.shape[0], wi.shape[1]), sd[f"{p}.proj_in.bias"])
    y = transformer_block(sd, f"{p}.transformer_blocks.0", y, heads, None, False)
    wo = sd[f"{p}.proj_out.weight"]
    y = F.linear(y, wo.reshape(wo.shape[0], wo.shape[1]), sd[f"{p}.proj_out.bias"])
    y = y.reshape(b, h, w, t, c).permute(0, 4, 3, 1, 2)              # b c t h w
    y = y + x5
    return y.permute(0, 2, 1, 3, 4).reshape(bt, c, h, w)


def temporal_conv_block(sd: SD, p: str, x: Tensor, b: int) -> Tensor:
    """TemporalConvBlock.forward (openaimodel3d.py:272-279): 4 x [GN32 over (C/32,T,H,W),
    SiLU, Conv3d(3,1,1) pad (1,0,0)] + identity."""
    bt, c, h, w = x.shape
    t = bt // b
    x5 = x.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4)
    y = x5
    for j, ci in ((1, 2), (2, 3), (3, 3), (4, 3)):
        q = f"{p}.conv{j}"
        y = F.silu(_gn(sd, f"{q}.0", y, 1e-5))
        y = F.conv3d(y, sd[f"{q}.{ci}.weight"], sd[f"{q}.{ci}.bias"], padding=(1, 0, 0))
    y = x5 + y
    return y.permute(0, 2, 1, 3, 4).reshape(bt, c, h, w)


def res_block(sd: SD, p: str, x: Tensor, emb: Tensor, b: int, tconv: bool) -> Tensor:
    """ResBlock._forward (openaimodel3d.py:210-236), use_scale_shift_norm=False, no up/down."""
    h = F.conv2d(F.silu(_gn(sd, f"{p}.in_layers.0", x, 1e-5)), sd[f"{p}.in_layers.2.weight"],
                 sd[f"{p}.in_layers.2.bias"], padding=1)
    h = h + _lin(sd, f"{p}.emb_layers.1", F.silu(emb))[:, :, None, None]
    h = F.conv2d(F.silu(_gn(sd, f"{p}.out_layers.0", h, 1e-5)), sd[f"{p}.out_layers.3.weight"],
                 sd[f"{p}.out_layers.3.bias"], padding=1)
    if f"{p}.skip_connection.weight" in sd:
        x = F.conv2d(x, sd[f"{p}.skip_connection.weight"], sd[f"{p}.skip_connection.bias"])
    h = x + h
    if tconv:
        h = temporal_conv_block(sd, f"{p}.temopral_conv", h, b)
    return h


def split_context(context: Tensor, t: int, text_len: int = 77) -> Tensor:
    """Per-frame context (openaimodel3d.py:580-587): 77 + 16*t tokens -> text repeated per
    frame + that frame's 16 image tokens; any other length -> whole context per frame."""
    b, l, _ = context.shape
    if l == text_len + t * 16:
        text = context[:, :text_len].repeat_interleave(t, dim=0)
        img = context[:, text_len:].reshape(b * t, 16, -1)
        return torch.cat([text, img], dim=1)
    return context.repeat_interleave(t, dim=0)


def unet_embedding(sd: SD, cfg: UNetCfg, timesteps: Tensor, c_label: Tensor, fs: Tensor) -> Tensor:
    """time_embed(t) + class_embed(label) + fps_embedding(fs)   (openaimodel3d.py:569-576,594-602). [B, 4mc]"""
    mc = cfg.model_channels
    emb = _mlp(sd, "time_embed", timestep_embedding(timesteps, mc))
    emb = emb + _mlp(sd, "class_embed", timestep_embedding(c_label, mc))
    emb = emb + _mlp(sd, "fps_embedding", timestep_embedding(fs, mc))
    return emb


def _run_layers(sd: SD, cfg: UNetCfg, bp: BlockPlan, h: Tensor, emb: Tensor, ctx: Tensor, b: int) -> Tensor:
    """TimestepEmbedSequential dispatch (openaimodel3d.py:36-48)."""
    for kind, p, m in bp.layers:
        if kind == "conv":
            h = F.conv2d(h, sd[f"{p}.weight"], sd[f"{p}.bias"], padding=1)
        elif kind == "res":
            h = res_block(sd, p, h, emb, b, m["tconv"])
        elif kind == "spatial":
            h = spatial_transformer(sd, p, h, m["heads"], ctx)
        elif kind == "temporal":
            h = temporal_transformer(sd, p, h, m["heads"], b)
        elif kind == "down":
            h = F.conv2d(h, sd[f"{p}.op.weight"], sd[f"{p}.op.bias"], stride=2, padding=1)
        elif kind == "up":
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[f"{p}.conv.weight"], sd[f"{p}.conv.bias"], padding=1)
    return h


@torch.no_grad()
def unet_forward(sd: SD, cfg: UNetCfg, x: Tensor, timesteps: Tensor, c_label: Tensor,
                 context: Tensor, fs: Tensor) -> Tensor:
    """UNetModel.forward (openaimodel3d.py:567-628). x: [B, Cin, T, H, W] -> [B, Cout, T, H, W]."""
    b, _, t, hh, ww = x.shape
    emb = unet_embedding(sd, cfg, timesteps, c_label, fs).repeat_interleave(t, dim=0)
    ctx = split_context(context, t, cfg.text_context_len)
    h = x.permute(0, 2, 1, 3, 4).reshape(b * t, -1, hh, ww)
    inputs, middle, outputs = unet_plan(cfg)
    hs = []
    for i, bp in enumerate(inputs):
        h = _run_layers(sd, cfg, bp, h, emb, ctx, b)
        if i == 0:
            h = temporal_transformer(sd, "init_attn.0", h, cfg.init_attn_heads, b)
        hs.append(h)
    h = _run_layers(sd, cfg, middle, h, emb, ctx, b)
    for bp in outputs:
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_layers(sd, cfg, bp, h, emb, ctx, b)
    y = F.conv2d(F.silu(_gn(sd, "out.0", h, 1e-5)), sd["out.2.weight"], sd["out.2.bias"], padding=1)
    return y.reshape(b, t, -1, hh, ww).permute(0, 2, 1, 3, 4)


# ----------------------------------------------------------------------------
# VAE decoder
# ----------------------------------------------------------------------------
def _swish(x):
    return x * torch.sigmoid(x)


def _vae_res(sd: SD, p: str, x: Tensor) -> Tensor:
    """ResnetBlock.forward with temb=None (ae_modules.py:190-210); GN eps 1e-6 (:15-16)."""
    h = F.conv2d(_swish(_gn(sd, f"{p}.norm1", x, 1e-6)), sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], padding=1)
    h = F.conv2d(_swish(_gn(sd, f"{p}.norm2", h, 1e-6)), sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"], padding=1)
    if f"{p}.nin_shortcut.weight" in sd:
        x = F.conv2d(x, sd[f"{p}.nin_shortcut.weight"], sd[f"{p}.nin_shortcut.bias"])
    return x + h


def _vae_attn(sd: SD, p: str, x: Tensor) -> Tensor:
    """AttnBlock.forward (ae_modules.py:53-78): single head, d = C, scale C^-0.5."""
    b, c, h, w = x.shape
    y = _gn(sd, f"{p}.norm", x, 1e-6)
    q = F.conv2d(y, sd[f"{p}.q.weight"], sd[f"{p}.q.bias"]).reshape(b, c, h * w).permute(0, 2, 1)
    k = F.conv2d(y, sd[f"{p}.k.weight"], sd[f"{p}.k.bias"]).reshape(b, c, h * w)
    v = F.conv2d(y, sd[f"{p}.v.weight"], sd[f"{p}.v.bias"]).reshape(b, c, h * w)
    w_ = torch.softmax(torch.bmm(q, k) * (c ** -0.5), dim=2)
    o = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + F.conv2d(o, sd[f"{p}.proj_out.weight"], sd[f"{p}.proj_out.bias"])


@torch.no_grad()
def vae_decode(sd: SD, cfg: VaeCfg, z: Tensor) -> Tensor:
    """AutoencoderKL.decode (autoencoder.py:104-107) + Decoder.forward (ae_modules.py:539-578).
    z: [F, 4, h, w] (already divided by scale_factor) -> [F, 3, 8h, 8w]; no clamp/tanh."""
    h = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    h = F.conv2d(h, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    h = _vae_res(sd, "decoder.mid.block_1", h)
    h = _vae_attn(sd, "decoder.mid.attn_1", h)
    h = _vae_res(sd, "decoder.mid.block_2", h)
    for lvl in reversed(range(len(cfg.ch_mult))):
        for ib in range(cfg.num_res_blocks + 1):
            h = _vae_res(sd, f"decoder.up.{lvl}.block.{ib}", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"decoder.up.{lvl}.upsample.conv.weight"], sd[f"decoder.up.{lvl}.upsample.conv.bias"], padding=1)
    h = _swish(_gn(sd, "decoder.norm_out", h, 1e-6))
    return F.conv2d(h, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)


@torch.no_grad()
def vae_encode_moments(sd: SD, cfg: VaeCfg, x: Tensor) -> Tensor:
    """Encoder.forward (ae_modules.py:432-463) + quant_conv (autoencoder.py:97-102): x [F,3,H,W] -> moments
    [F, 2*z, H/8, W/8] (mean | logvar).  Downsample = pad (0,1,0,1) then 3x3 stride-2 conv (ae_modules.py:103-106)."""
    h = F.conv2d(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    n = len(cfg.ch_mult)
    for lvl in range(n):
        for ib in range(cfg.num_res_blocks):
            h = _vae_res(sd, f"encoder.down.{lvl}.block.{ib}", h)
        if lvl != n - 1:
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[f"encoder.down.{lvl}.downsample.conv.weight"],
                         sd[f"encoder.down.{lvl}.downsample.conv.bias"], stride=2)
    h = _vae_res(sd, "encoder.mid.block_1", h)
    h = _vae_attn(sd, "encoder.mid.attn_1", h)
    h = _vae_res(sd, "encoder.mid.block_2", h)
    h = F.conv2d(_swish(_gn(sd, "encoder.norm_out", h, 1e-6)), sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


@torch.no_grad()
def decode_first_stage(sd: SD, cfg: VaeCfg, z: Tensor, scale_factor: float = 0.18215) -> Tensor:
    """LatentDiffusion.decode_core, perframe_ae=True (ddpm3d.py:646-667): z [B,4,T,h,w] -> [B,3,T,8h,8w]."""
    b, c, t, h, w = z.shape
    frames = z.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    outs = [vae_decode(sd, cfg, (1.0 / scale_factor) * frames[i:i + 1]) for i in range(b * t)]
    out = torch.cat(outs, dim=0)
    return out.reshape(b, t, *out.shape[1:]).permute(0, 2, 1, 3, 4)


# ----------------------------------------------------------------------------
# schedules + DDIM
# ----------------------------------------------------------------------------
def linear_betas(n: int = 1000, linear_start: float = 0.00085, linear_end: float = 0.012) -> np.ndarray:
    """'linear' schedule = linspace(sqrt(s), sqrt(e), n)^2 in f64 (utils_diffusion.py:32-35)."""
    return np.linspace(linear_start ** 0.5, linear_end ** 0.5, n, dtype=np.float64) ** 2


def zero_terminal_snr(betas: np.ndarray) -> np.ndarray:
    """rescale_zero_terminal_snr (utils_diffusion.py:112-144)."""
    abar_sqrt = np.sqrt(np.cumprod(1.0 - betas))
    a0, aT = abar_sqrt[0], abar_sqrt[-1]
    abar_sqrt = (abar_sqrt - aT) * (a0 / (a0 - aT))
    abar = abar_sqrt ** 2
    alphas = np.concatenate([abar[:1], abar[1:] / abar[:-1]])
    return 1.0 - alphas


@dataclass
class DiffusionTables:
    """fp32 buffers the sampler reads from the model (ddpm3d.py:123-186,522-527)."""
    alphas_cumprod: Tensor
    sqrt_alphas_cumprod: Tensor
    sqrt_one_minus_alphas_cumprod: Tensor
    scale_arr: Tensor


def make_tables(timesteps=1000, linear_start=0.00085, linear_end=0.012, zero_snr=True,
                base_scale=0.3, turning_step=400) -> DiffusionTables:
    betas = linear_betas(timesteps, linear_start, linear_end)
    if zero_snr:
        betas = zero_terminal_snr(betas)
    ac = np.cumprod(1.0 - betas)
    scale = np.concatenate([np.linspace(1.0, base_scale, turning_step), np.full(timesteps, base_scale)])
    f = lambda a: torch.tensor(a, dtype=torch.float32)
    return DiffusionTables(f(ac), f(np.sqrt(ac)), f(np.sqrt(1.0 - ac)), f(scale))


def ddim_timesteps(method: str, n_ddim: int, n_ddpm: int) -> np.ndarray:
    """make_ddim_timesteps (utils_diffusion.py:56-76)."""
    if method == "uniform":
        return np.arange(0, n_ddpm, n_ddpm // n_ddim) + 1
    if method == "uniform_trailing":
        c = n_ddpm / n_ddim
        return np.flip(np.round(np.arange(n_ddpm, 0, -c))).astype(np.int64) - 1
    if method == "quad":
        return (np.linspace(0, np.sqrt(n_ddpm * 0.8), n_ddim) ** 2).astype(int) + 1
    raise NotImplementedError(method)


@dataclass
class DDIMSchedule:
    timesteps: np.ndarray
    alphas: np.ndarray
    alphas_prev: np.ndarray
    sigmas: np.ndarray
    sqrt_one_minus_alphas: np.ndarray
    scale_arr: Tensor
    scale_arr_prev: Tensor


def make_ddim_schedule(tab: DiffusionTables, S: int, spacing: str, eta: float) -> DDIMSchedule:
    """DDIMSampler.make_schedule (ddim.py:24-57) + make_ddim_sampling_parameters (utils_diffusion.py:79-91)."""
    ts = ddim_timesteps(spacing, S, tab.alphas_cumprod.shape[0])
    ac = tab.alphas_cumprod.cpu()
    alphas = ac[ts].numpy()                                # fp32 values, as the reference indexes a fp32 tensor
    alphas_prev = np.asarray([float(ac[0])] + ac[ts[:-1]].tolist())        # f64 array of fp32 values
    # The reference evaluates `ndarray_f64 / tensor_f32` (utils_diffusion.py:86), which numpy defers to
    # Tensor.__rtruediv__ = reciprocal(tensor) * other: an fp32 reciprocal of the fp32 (1 - alphas),
    # everything else in f64.  Reproduced bit-for-bit (pinned by make_golden.py).
    recip = (np.float32(1.0) / (np.float32(1.0) - alphas)).astype(np.float64)
    sigmas = eta * np.sqrt(recip * (1 - alphas_prev) * (1 - alphas.astype(np.float64) / alphas_prev))
    sc = tab.scale_arr[ts]
    return DDIMSchedule(ts, alphas, alphas_prev, sigmas, np.sqrt(1.0 - alphas), sc, torch.cat([sc[0:1], sc[:-1]]))


def rescale_noise_cfg(cfg_out: Tensor, cond_out: Tensor, phi: float) -> Tensor:
    """utils_diffusion.py:147-158."""
    dims = list(range(1, cond_out.ndim))
    s_text = cond_out.std(dim=dims, keepdim=True)
    s_cfg = cfg_out.std(dim=dims, keepdim=True)
    return phi * (cfg_out * (s_text / s_cfg)) + (1 - phi) * cfg_out


def ddim_step(tab: DiffusionTables, sch: DDIMSchedule, index: int, x: Tensor, v_cond: Tensor,
              v_uncond: Optional[Tensor], noise: Tensor, cfg_scale: float, guidance_rescale: float,
              dynamic_rescale: bool = True) -> Tuple[Tensor, Tensor]:
    """DDIMSampler.p_sample_ddim after the UNet calls (ddim.py:226-277), v-parameterisation."""
    t = int(sch.timesteps[index])
    if v_uncond is None or cfg_scale == 1.0:
        v = v_cond
    else:
        v = v_uncond + cfg_scale * (v_cond - v_uncond)
        if guidance_rescale > 0.0:
            v = rescale_noise_cfg(v, v_cond, guidance_rescale)
    sa, s1 = tab.sqrt_alphas_cumprod[t], tab.sqrt_one_minus_alphas_cumprod[t]
    e_t = sa * v + s1 * x                                  # ddpm3d.py:247-251
    pred_x0 = sa * x - s1 * v                              # ddpm3d.py:239-245
    if dynamic_rescale:
        pred_x0 = pred_x0 * (sch.scale_arr_prev[index] / sch.scale_arr[index])
    a_prev = torch.tensor(sch.alphas_prev[index], dtype=torch.float32, device=x.device)
    sigma = torch.tensor(sch.sigmas[index], dtype=torch.float32, device=x.device)
    dir_xt = (1.0 - a_prev - sigma ** 2).sqrt() * e_t
    x_prev = a_prev.sqrt() * pred_x0 + dir_xt + sigma * noise
    return x_prev, pred_x0


@torch.no_grad()
def ddim_sample(unet_sd: SD, ucfg: UNetCfg, tab: DiffusionTables, *, S: int, shape, c_concat: Tensor,
                context: Tensor, uc_context: Optional[Tensor], class_label: Tensor, fs: Tensor,
                cfg_scale: float = 1.0, guidance_rescale: float = 0.0, eta: float = 1.0,
                spacing: str = "uniform_trailing", generator: Optional[torch.Generator] = None,
                noises: Optional[List[Tensor]] = None, device="cpu", unet_fn=None, mask: Optional[Tensor] = None,
                x0: Optional[Tensor] = None, clean_cond: bool = False) -> Tensor:
    """DDIMSampler.sample/ddim_sampling (ddim.py:60-203) with the hybrid DiffusionWrapper
    (ddpm3d.py:1320-1324).  RNG order: x_T first, then one draw per step (App. D #10);
    `noises` (len S+1) overrides the generator so CUDA/CPU runs can share draws.
    mask / x0 (ddim.py:173-180): before every step the known latent x0 -- noised to the step's level by q_sample
    (ddpm3d.py:305-308, ONE EXTRA draw per step, taken before the step's own) or clean -- replaces x where mask == 1.
    `unet_fn(xc, ts, class_label, context, fs)` replaces the oracle's own UNet forward (used to put the reference's
    autocast-fp16 UNetModel into the same loop when calibrating tolerances)."""
    if unet_fn is not None:
        unet_forward = lambda _sd, _cfg, xc, ts, lab, ctx, fs_: unet_fn(xc, ts, lab, ctx, fs_).float()   # noqa: E731
    else:
        unet_forward = globals()["unet_forward"]
    sch = make_ddim_schedule(tab, S, spacing, eta)
    draw = (lambda i: noises[i]) if noises is not None else (lambda i: torch.randn(shape, generator=generator, device=device))
    x = draw(0)
    B = shape[0]
    for i, step in enumerate(np.flip(sch.timesteps)):
        index = S - i - 1
        ts = torch.full((B,), int(step), dtype=torch.long, device=x.device)
        if mask is not None:
            assert x0 is not None and noises is None, "mask branch: draws come from the generator in the reference's order"
            orig = x0
            if not clean_cond:
                a = tab.sqrt_alphas_cumprod.to(x.device)[int(step)]
                b1 = tab.sqrt_one_minus_alphas_cumprod.to(x.device)[int(step)]
                orig = a * x0 + b1 * torch.randn(shape, generator=generator, device=device)
            x = orig * mask + (1.0 - mask) * x
        xc = torch.cat([x, c_concat], dim=1)
        v_c = unet_forward(unet_sd, ucfg, xc, ts, class_label, context, fs)
        v_u = None
        if uc_context is not None and cfg_scale != 1.0:
            v_u = unet_forward(unet_sd, ucfg, xc, ts, class_label, uc_context, fs)
        x, _ = ddim_step(tab, sch, index, x, v_c, v_u, draw(i + 1), cfg_scale, guidance_rescale)
    return x


@torch.no_grad()
def ddim_sample_multicond(unet_sd: SD, ucfg: UNetCfg, tab: DiffusionTables, *, S: int, shape, c_concat: Tensor,
                          context: Tensor, uc_context: Tensor, uc_img_context: Tensor, class_label: Tensor, fs: Tensor,
                          cfg_scale: float, cfg_img: Optional[float], guidance_rescale: float = 0.0, eta: float = 1.0,
                          spacing: str = "uniform_trailing", noises: Optional[List[Tensor]] = None) -> Tensor:
    """DDIMSampler_multicond (ddim_multiplecond.py:210-237): three UNet evaluations per step,
    v = v_u + cfg_img (v_ui - v_u) + s (v_c - v_ui), then the single-guidance update (ddim_multiplecond.py:238-284)."""
    sch = make_ddim_schedule(tab, S, spacing, eta)
    draw = (lambda i: noises[i]) if noises is not None else (lambda i: torch.randn(shape))
    if cfg_img is None:
        cfg_img = cfg_scale
    x = draw(0)
    B = shape[0]
    for i, step in enumerate(np.flip(sch.timesteps)):
        index = S - i - 1
        ts = torch.full((B,), int(step), dtype=torch.long)
        xc = torch.cat([x, c_concat], dim=1)
        v_c = unet_forward(unet_sd, ucfg, xc, ts, class_label, context, fs)
        v_u = unet_forward(unet_sd, ucfg, xc, ts, class_label, uc_context, fs)
        v_ui = unet_forward(unet_sd, ucfg, xc, ts, class_label, uc_img_context, fs)
        v = v_u + cfg_img * (v_ui - v_u) + cfg_scale * (v_c - v_ui)
        if guidance_rescale > 0.0:
            v = rescale_noise_cfg(v, v_c, guidance_rescale)
        x, _ = ddim_step(tab, sch, index, x, v, None, draw(i + 1), 1.0, 0.0)
    return x


# ================================================================ Resampler
# Perceiver resampler that turns the image-encoder tokens into the UNet's image context
# (lvdm/modules/encoders/resampler.py:48-144): [B, n1, embedding_dim] -> [B, num_queries * video_length, output_dim].
def resampler_param_shapes(dim=1024, depth=4, dim_head=64, heads=12, num_queries=16, embedding_dim=1280, output_dim=1024,
                           ff_mult=4, video_length=16) -> Dict[str, Tuple[int, ...]]:
    nq = num_queries * (video_length if video_length is not None else 1)
    inner = dim_head * heads
    s: Dict[str, Tuple[int, ...]] = {"latents": (1, nq, dim), "proj_in.weight": (dim, embedding_dim), "proj_in.bias": (dim,),
                                     "proj_out.weight": (output_dim, dim), "proj_out.bias": (output_dim,),
                                     "norm_out.weight": (output_dim,), "norm_out.bias": (output_dim,)}
    for i in range(depth):
        a, f = f"layers.{i}.0", f"layers.{i}.1"
        for n in ("norm1", "norm2"):
            s[f"{a}.{n}.weight"] = (dim,)
            s[f"{a}.{n}.bias"] = (dim,)
        s[f"{a}.to_q.weight"] = (inner, dim)
        s[f"{a}.to_kv.weight"] = (2 * inner, dim)
        s[f"{a}.to_out.weight"] = (dim, inner)
        s[f"{f}.0.weight"] = (dim,)
        s[f"{f}.0.bias"] = (dim,)
        s[f"{f}.1.weight"] = (int(dim * ff_mult), dim)
        s[f"{f}.3.weight"] = (dim, int(dim * ff_mult))
    return s


@torch.no_grad()
def resampler_forward(sd: SD, x: Tensor, heads: int, dim_head: int = 64) -> Tensor:
    """Resampler.forward (resampler.py:131-144) with PerceiverAttention.forward (:66-101) and FeedForward (:31-37)."""
    F = torch.nn.functional
    depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("layers."))
    b = x.shape[0]
    lat = sd["latents"].repeat(b, 1, 1)
    x = F.linear(x, sd["proj_in.weight"], sd["proj_in.bias"])
    scale = 1.0 / math.sqrt(math.sqrt(dim_head))
    for i in range(depth):
        a, f = f"layers.{i}.0", f"layers.{i}.1"
        xn = F.layer_norm(x, x.shape[-1:], sd[f"{a}.norm1.weight"], sd[f"{a}.norm1.bias"])
        ln = F.layer_norm(lat, lat.shape[-1:], sd[f"{a}.norm2.weight"], sd[f"{a}.norm2.bias"])
        q = F.linear(ln, sd[f"{a}.to_q.weight"])
        k, v = F.linear(torch.cat((xn, ln), dim=-2), sd[f"{a}.to_kv.weight"]).chunk(2, dim=-1)
        split = lambda t: t.view(b, t.shape[1], heads, -1).transpose(1, 2)
        q, k, v = split(q), split(k), split(v)
        w = torch.softmax(((q * scale) @ (k * scale).transpose(-2, -1)).float(), dim=-1).type(q.dtype)
        o = (w @ v).permute(0, 2, 1, 3).reshape(b, lat.shape[1], -1)
        lat = F.linear(o, sd[f"{a}.to_out.weight"]) + lat
        h = F.layer_norm(lat, lat.shape[-1:], sd[f"{f}.0.weight"], sd[f"{f}.0.bias"])
        h = F.linear(F.gelu(F.linear(h, sd[f"{f}.1.weight"])), sd[f"{f}.3.weight"])
        lat = h + lat
    lat = F.linear(lat, sd["proj_out.weight"], sd["proj_out.bias"])
    return F.layer_norm(lat, lat.shape[-1:], sd["norm_out.weight"], sd["norm_out.bias"])
