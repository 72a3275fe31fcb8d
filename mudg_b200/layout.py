"""State-dict layouts of the modules on the hot path (key -> shape), so the drop-in nn.Modules expose exactly the
reference's parameter names (strict load_state_dict of real checkpoints) without re-stating its layer classes.

UNet: lvdm/modules/networks/openaimodel3d.py:376-565 (+ attention.py:42-79,348-365,413-449,477-520);
VAE : lvdm/models/autoencoder.py:27-32, lvdm/modules/networks/ae_modules.py:151-188,364-537.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Iterable, Optional, Sequence, Tuple

Shape = Tuple[int, ...]


class _Spec(OrderedDict):
    def conv(self, p, cout, cin, *k):
        self[p + ".weight"] = (cout, cin) + tuple(k)
        self[p + ".bias"] = (cout,)

    def lin(self, p, cout, cin, bias=True):
        self[p + ".weight"] = (cout, cin)
        if bias:
            self[p + ".bias"] = (cout,)

    def norm(self, p, c):
        self[p + ".weight"] = (c,)
        self[p + ".bias"] = (c,)


def _transformer_block(s: _Spec, p: str, dim: int, ctx: Optional[int], image_ca: bool):
    for name, kv, ip in (("attn1", dim, False), ("attn2", ctx or dim, image_ca)):
        q = f"{p}.{name}"
        s.lin(q + ".to_q", dim, dim, bias=False)
        s.lin(q + ".to_k", dim, kv, bias=False)
        s.lin(q + ".to_v", dim, kv, bias=False)
        s.lin(q + ".to_out.0", dim, dim)
        if ip:
            s.lin(q + ".to_k_ip", dim, kv, bias=False)
            s.lin(q + ".to_v_ip", dim, kv, bias=False)
        if name == "attn1":
            s.lin(p + ".ff.net.0.proj", 8 * dim, dim)
            s.lin(p + ".ff.net.2", dim, 4 * dim)
    for n in ("norm1", "norm2", "norm3"):
        s.norm(f"{p}.{n}", dim)


def _reorder_like_reference(s: _Spec) -> _Spec:
    return s


def unet_layout(*, in_channels: int, out_channels: int, model_channels: int, num_res_blocks: int,
                attention_resolutions: Sequence[int], channel_mult: Sequence[int] = (1, 2, 4, 8),
                num_head_channels: int = 64, context_dim: int = 1024, temporal_conv: bool = True,
                addition_attention: bool = True, image_cross_attention: bool = True, fs_condition: bool = True,
                class_label_condition: bool = True, **_ignored) -> Dict[str, Shape]:
    mc, ted = model_channels, 4 * model_channels
    s = _Spec()

    def mlp(p):
        s.lin(p + ".0", ted, mc)
        s.lin(p + ".2", ted, ted)

    def res(p, ci, co):
        s.norm(p + ".in_layers.0", ci)
        s.conv(p + ".in_layers.2", co, ci, 3, 3)
        s.lin(p + ".emb_layers.1", co, ted)
        s.norm(p + ".out_layers.0", co)
        s.conv(p + ".out_layers.3", co, co, 3, 3)
        if ci != co:
            s.conv(p + ".skip_connection", co, ci, 1, 1)
        if temporal_conv:
            for j, idx in ((1, 2), (2, 3), (3, 3), (4, 3)):
                q = f"{p}.temopral_conv.conv{j}"                  # sic (openaimodel3d.py:190)
                s.norm(q + ".0", co)
                s.conv(f"{q}.{idx}", co, co, 3, 1, 1)

    def spatial(p, ch):
        s.norm(p + ".norm", ch)
        s.lin(p + ".proj_in", ch, ch)
        _transformer_block(s, p + ".transformer_blocks.0", ch, context_dim, image_cross_attention)
        s.lin(p + ".proj_out", ch, ch)

    def temporal(p, ch, inner=None, conv1d=False):
        inner = inner or ch
        s.norm(p + ".norm", ch)
        if conv1d:
            s.conv(p + ".proj_in", inner, ch, 1)
        else:
            s.lin(p + ".proj_in", inner, ch)
        _transformer_block(s, p + ".transformer_blocks.0", inner, None, False)
        if conv1d:
            s.conv(p + ".proj_out", ch, inner, 1)
        else:
            s.lin(p + ".proj_out", ch, inner)

    mlp("time_embed")
    if class_label_condition:
        mlp("class_embed")
    if fs_condition:
        mlp("fps_embedding")
    s.conv("input_blocks.0.0", mc, in_channels, 3, 3)
    if addition_attention:
        temporal("init_attn.0", mc, inner=8 * num_head_channels, conv1d=True)
    chans = [mc]
    ch, ds, idx = mc, 1, 1
    for level, mult in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            res(f"input_blocks.{idx}.0", ch, mult * mc)
            ch = mult * mc
            if ds in attention_resolutions:
                spatial(f"input_blocks.{idx}.1", ch)
                temporal(f"input_blocks.{idx}.2", ch)
            chans.append(ch)
            idx += 1
        if level != len(channel_mult) - 1:
            s.conv(f"input_blocks.{idx}.0.op", ch, ch, 3, 3)
            chans.append(ch)
            idx += 1
            ds *= 2
    res("middle_block.0", ch, ch)
    spatial("middle_block.1", ch)
    temporal("middle_block.2", ch)
    res("middle_block.3", ch, ch)
    oidx = 0
    for level, mult in list(enumerate(channel_mult))[::-1]:
        for i in range(num_res_blocks + 1):
            ich = chans.pop()
            li = 0
            res(f"output_blocks.{oidx}.{li}", ch + ich, mc * mult)
            li += 1
            ch = mc * mult
            if ds in attention_resolutions:
                spatial(f"output_blocks.{oidx}.{li}", ch)
                temporal(f"output_blocks.{oidx}.{li + 1}", ch)
                li += 2
            if level and i == num_res_blocks:
                s.conv(f"output_blocks.{oidx}.{li}.conv", ch, ch, 3, 3)
                ds //= 2
            oidx += 1
    s.norm("out.0", mc)
    s.conv("out.2", out_channels, mc, 3, 3)
    return s


#: parameters the reference zero-initialises
def unet_zero_init_keys(layout: Iterable[str]):
    out = []
    for k in layout:
        if (".out_layers.3." in k or ".temopral_conv.conv4.3." in k or k.startswith("fps_embedding.2.")
                or k.startswith("out.2.") or (k.endswith((".proj_out.weight", ".proj_out.bias")))):
            out.append(k)
    return out


def vae_layout(*, ch: int, ch_mult: Sequence[int], num_res_blocks: int, z_channels: int, out_ch: int = 3,
               in_channels: int = 3, embed_dim: int = 4, double_z: bool = True, **_ignored) -> Dict[str, Shape]:
    s = _Spec()

    def res(p, ci, co):
        s.norm(p + ".norm1", ci)
        s.conv(p + ".conv1", co, ci, 3, 3)
        s.norm(p + ".norm2", co)
        s.conv(p + ".conv2", co, co, 3, 3)
        if ci != co:
            s.conv(p + ".nin_shortcut", co, ci, 1, 1)

    def attn(p, c):
        s.norm(p + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            s.conv(f"{p}.{n}", c, c, 1, 1)

    n = len(ch_mult)
    s.conv("encoder.conv_in", ch, in_channels, 3, 3)
    in_mult = (1,) + tuple(ch_mult)
    block_in = ch
    for lvl in range(n):
        block_in, block_out = ch * in_mult[lvl], ch * ch_mult[lvl]
        for ib in range(num_res_blocks):
            res(f"encoder.down.{lvl}.block.{ib}", block_in, block_out)
            block_in = block_out
        if lvl != n - 1:
            s.conv(f"encoder.down.{lvl}.downsample.conv", block_in, block_in, 3, 3)
    res("encoder.mid.block_1", block_in, block_in)
    attn("encoder.mid.attn_1", block_in)
    res("encoder.mid.block_2", block_in, block_in)
    s.norm("encoder.norm_out", block_in)
    s.conv("encoder.conv_out", 2 * z_channels if double_z else z_channels, block_in, 3, 3)
    block_in = ch * ch_mult[-1]
    s.conv("decoder.conv_in", block_in, z_channels, 3, 3)
    res("decoder.mid.block_1", block_in, block_in)
    attn("decoder.mid.attn_1", block_in)
    res("decoder.mid.block_2", block_in, block_in)
    for lvl in reversed(range(n)):
        block_out = ch * ch_mult[lvl]
        for ib in range(num_res_blocks + 1):
            res(f"decoder.up.{lvl}.block.{ib}", block_in, block_out)
            block_in = block_out
        if lvl != 0:
            s.conv(f"decoder.up.{lvl}.upsample.conv", block_in, block_in, 3, 3)
    s.norm("decoder.norm_out", block_in)
    s.conv("decoder.conv_out", out_ch, block_in, 3, 3)
    s.conv("quant_conv", 2 * embed_dim, 2 * z_channels, 1, 1)
    s.conv("post_quant_conv", z_channels, embed_dim, 1, 1)
    return s
