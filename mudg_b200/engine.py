"""Python binding of libmudg_sm100.so: one `Engine` per (process, device).

The engine owns packed fp16 weights and a workspace arena inside the library; PyTorch only provides
device memory for inputs/outputs and the CUDA stream.  There is no CPU / PyTorch fallback: constructing
an Engine without the built library or without a Blackwell GPU raises.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Mapping, Optional, Sequence

import torch

from ._lib import MudgError, check, cur_stream, lib, ptr

MUDG_F32, MUDG_F16, MUDG_U8 = 0, 1, 2
MUDG_UNET, MUDG_VAE, MUDG_RESAMPLER, MUDG_CLIP_IMAGE, MUDG_CLIP_TEXT = 0, 1, 2, 3, 4
MUDG_POST_COLOR, MUDG_POST_DEPTH, MUDG_POST_SEMANTIC = 0, 1, 2
# class labels the driver gives the three modalities (virtual_pose_render.py:247-318)
LABEL_TO_POST_MODE = {0: MUDG_POST_COLOR, 500: MUDG_POST_DEPTH, 1: MUDG_POST_SEMANTIC}


def postdecode(frames: torch.Tensor, modes: Sequence[int]):
    """Post-decode frame pipeline on the GPU (mudg_postdecode): frames [B, 3, T, H, W] (fp16 / fp32 in [-1,1], or uint8)
    and one mode per sample -> (rgb uint8 [B,T,3,H,W], depth fp32 [B,T,H,W], cls uint8 [B,T,H,W]).  depth / cls rows of
    samples in another mode are left zero.  Bit-exact against the reference's CPU code (eval_tools.py)."""
    if not frames.is_cuda:
        raise MudgError("postdecode needs a CUDA tensor; there is no CPU fallback")
    if frames.dim() != 5 or frames.shape[1] != 3:
        raise MudgError(f"postdecode expects [B, 3, T, H, W], got {tuple(frames.shape)}")
    dt = {torch.float32: MUDG_F32, torch.float16: MUDG_F16, torch.uint8: MUDG_U8}.get(frames.dtype)
    if dt is None:
        frames, dt = frames.float(), MUDG_F32
    frames = frames.detach().contiguous()
    B, _, T, H, W = frames.shape
    modes = [int(m) for m in modes]
    if len(modes) != B:
        raise MudgError(f"postdecode: {len(modes)} modes for {B} samples")
    rgb = torch.empty((B, T, 3, H, W), device=frames.device, dtype=torch.uint8)
    depth = torch.zeros((B, T, H, W), device=frames.device, dtype=torch.float32) if MUDG_POST_DEPTH in modes else None
    cls = torch.zeros((B, T, H, W), device=frames.device, dtype=torch.uint8) if MUDG_POST_SEMANTIC in modes else None
    marr = (ctypes.c_int * B)(*modes)
    with torch.cuda.device(frames.device):
        check(lib().mudg_postdecode(ptr(frames), dt, B, T, H, W, marr, ptr(rgb), ptr(depth), ptr(cls), cur_stream()))
    return rgb, depth, cls


def colormap_spectral(depth01: torch.Tensor) -> torch.Tensor:
    """[...] fp32 map in [0,1] on the device -> [..., 3] uint8 (the reference's colormap(..., 'Spectral', bytes=True))."""
    if not depth01.is_cuda:
        raise MudgError("colormap_spectral needs a CUDA tensor; there is no CPU fallback")
    m = depth01.detach().float().contiguous()
    out = torch.empty(tuple(m.shape) + (3,), device=m.device, dtype=torch.uint8)
    with torch.cuda.device(m.device):
        check(lib().mudg_colormap_spectral(ptr(m), ctypes.c_int64(m.numel()), ptr(out), cur_stream()))
    return out


class _UNetConfig(ctypes.Structure):
    _fields_ = [("in_channels", ctypes.c_int), ("out_channels", ctypes.c_int), ("model_channels", ctypes.c_int),
                ("num_res_blocks", ctypes.c_int), ("channel_mult", ctypes.c_int * 8), ("n_channel_mult", ctypes.c_int),
                ("attention_resolutions", ctypes.c_int * 8), ("n_attention_resolutions", ctypes.c_int),
                ("num_head_channels", ctypes.c_int), ("context_dim", ctypes.c_int), ("init_attn_heads", ctypes.c_int),
                ("text_context_len", ctypes.c_int)]


class _VaeConfig(ctypes.Structure):
    _fields_ = [("ch", ctypes.c_int), ("ch_mult", ctypes.c_int * 8), ("n_ch_mult", ctypes.c_int),
                ("num_res_blocks", ctypes.c_int), ("z_channels", ctypes.c_int), ("out_ch", ctypes.c_int),
                ("embed_dim", ctypes.c_int)]


def _arr8(values: Sequence[int]):
    vals = list(values)
    if len(vals) > 8:
        raise MudgError("at most 8 levels are supported")
    return (ctypes.c_int * 8)(*(vals + [0] * (8 - len(vals)))), len(vals)


# smallest graph the library accepts: for contexts that only run the Resampler / post-decode kernels
_MINIMAL_UNET = dict(in_channels=8, out_channels=4, model_channels=32, num_res_blocks=1, channel_mult=(1,),
                     attention_resolutions=(1,), num_head_channels=64, context_dim=64)
_DEFAULT_VAE = dict(ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4, out_ch=3, embed_dim=4)


class Engine:
    def __init__(self, unet: Optional[Mapping] = None, vae: Optional[Mapping] = None, device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise MudgError("mudg_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        unet = dict(_MINIMAL_UNET) if unet is None else unet
        self.device = torch.cuda.current_device() if device is None else int(device)
        vae = dict(_DEFAULT_VAE, **(vae or {}))
        uc = _UNetConfig()
        uc.in_channels = int(unet["in_channels"]); uc.out_channels = int(unet["out_channels"])
        uc.model_channels = int(unet["model_channels"]); uc.num_res_blocks = int(unet["num_res_blocks"])
        uc.channel_mult, uc.n_channel_mult = _arr8(unet.get("channel_mult", (1, 2, 4, 8)))
        uc.attention_resolutions, uc.n_attention_resolutions = _arr8(unet["attention_resolutions"])
        uc.num_head_channels = int(unet.get("num_head_channels", 64))
        uc.context_dim = int(unet.get("context_dim", 1024))
        uc.init_attn_heads = 8
        uc.text_context_len = 77
        vc = _VaeConfig()
        vc.ch = int(vae["ch"]); vc.ch_mult, vc.n_ch_mult = _arr8(vae["ch_mult"])
        vc.num_res_blocks = int(vae["num_res_blocks"]); vc.z_channels = int(vae["z_channels"])
        vc.out_ch = int(vae["out_ch"]); vc.embed_dim = int(vae.get("embed_dim", vc.z_channels))
        self.in_channels, self.out_channels = uc.in_channels, uc.out_channels
        self.z_channels, self.vae_out_ch = vc.z_channels, vc.out_ch
        self._h = ctypes.c_void_p()
        L = lib()
        L.mudg_workspace_bytes.restype = ctypes.c_size_t
        L.mudg_launch_count.restype = ctypes.c_int64
        check(L.mudg_create(self.device, ctypes.byref(uc), ctypes.byref(vc), ctypes.byref(self._h)))
        self._ctx_key = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().mudg_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd: Mapping[str, torch.Tensor], which: int = MUDG_UNET, prefix: str = "",
                        chunk_bytes: int = 1 << 30) -> None:
        """Upload `sd` (reference key names, optionally below `prefix`).  Host tensors are staged through the GPU
        a chunk at a time so the fp32 copy of the 1.44 B-parameter UNet never has to be resident."""
        L = lib()
        dev = torch.device("cuda", self.device)
        staged, staged_bytes = [], 0
        for key, t in sd.items():
            if prefix:
                if not key.startswith(prefix):
                    continue
                key = key[len(prefix):]
            if not torch.is_floating_point(t) or t.dim() == 0:      # index buffers; scalars (CLIP logit_scale) are not on any path
                continue
            t = t.detach()
            if t.dtype not in (torch.float32, torch.float16):
                t = t.float()
            g = t.to(dev, non_blocking=True).contiguous()
            if which == MUDG_RESAMPLER and key == "latents":      # [1, nq, dim] parameter -> a [nq, dim] row matrix
                g = g.reshape(-1, g.shape[-1])
            shape = (ctypes.c_int64 * max(1, g.dim()))(*g.shape)
            check(L.mudg_load_weight(self._h, which, key.encode(), ptr(g), MUDG_F32 if g.dtype == torch.float32 else MUDG_F16,
                                     shape, g.dim(), cur_stream()))
            staged.append(g)
            staged_bytes += g.numel() * g.element_size()
            if staged_bytes > chunk_bytes:
                torch.cuda.current_stream().synchronize()
                staged, staged_bytes = [], 0
        torch.cuda.current_stream().synchronize()
        check(L.mudg_finalize_weights(self._h, which, cur_stream()))

    # ------------------------------------------------------------------ hot path
    def set_context(self, context: torch.Tensor, T: int) -> None:
        """context [N, L, context_dim]; precomputes the cross-attention K/V of all spatial transformers."""
        c = context.detach()
        if c.dtype not in (torch.float32, torch.float16):
            c = c.float()
        c = c.contiguous()
        N, Lc, _ = c.shape
        check(lib().mudg_set_context(self._h, ptr(c), MUDG_F32 if c.dtype == torch.float32 else MUDG_F16, N, Lc, int(T),
                                     cur_stream()))

    def unet_forward(self, x: torch.Tensor, t: torch.Tensor, c_label: torch.Tensor, fs: torch.Tensor,
                     out: Optional[torch.Tensor] = None, dup: int = 1) -> torch.Tensor:
        """x [N, Cin, T, h, w] fp32; t/c_label/fs [N] int64; returns [N, Cout, T, h, w] fp16.
        dup > 1 (mudg_unet_forward_shared): the caller promises that x / t / c_label / fs are N / dup distinct samples
        tiled dup times (a classifier-free-guidance batch: only the context rows differ)."""
        N, C, T, h, w = x.shape
        assert C == self.in_channels, (C, self.in_channels)
        x = x.detach().float().contiguous()
        t = t.to(device=x.device, dtype=torch.long).contiguous()
        c_label = c_label.to(device=x.device, dtype=torch.long).contiguous()
        fs = fs.to(device=x.device, dtype=torch.long).contiguous()
        if out is None:
            out = torch.empty((N, self.out_channels, T, h, w), device=x.device, dtype=torch.float16)
        if dup > 1:
            if N % dup:
                raise MudgError(f"unet_forward: batch {N} is not a multiple of dup={dup}")
            check(lib().mudg_unet_forward_shared(self._h, ptr(x), ptr(t), ptr(c_label), ptr(fs), N, int(dup), T, h, w, ptr(out),
                                                 cur_stream()))
        else:
            check(lib().mudg_unet_forward(self._h, ptr(x), ptr(t), ptr(c_label), ptr(fs), N, T, h, w, ptr(out), cur_stream()))
        return out

    def vae_decode(self, z: torch.Tensor) -> torch.Tensor:
        """z [F, zc, h, w] (already divided by scale_factor) -> [F, 3, 8h, 8w] fp16."""
        F_, C, h, w = z.shape
        assert C == self.z_channels
        z = z.detach().float().contiguous()
        out = torch.empty((F_, self.vae_out_ch, 8 * h, 8 * w), device=z.device, dtype=torch.float16)
        check(lib().mudg_vae_decode(self._h, ptr(z), F_, h, w, ptr(out), cur_stream()))
        return out

    def vae_encode_moments(self, x: torch.Tensor) -> torch.Tensor:
        """x [F, 3, H, W] in [-1,1] -> posterior moments [F, 2*zc, H/8, W/8] fp32 (mean | logvar)."""
        F_, C, H, W = x.shape
        assert C == 3
        x = x.detach().float().contiguous()
        out = torch.empty((F_, 2 * self.z_channels, H // 8, W // 8), device=x.device, dtype=torch.float32)
        check(lib().mudg_vae_encode(self._h, ptr(x), F_, H, W, ptr(out), cur_stream()))
        return out

    def resampler_forward(self, x: torch.Tensor, n_out: int, out_dim: int) -> torch.Tensor:
        """x [B, L, embedding_dim] -> [B, n_out, out_dim] fp32 (Resampler.forward)."""
        x = x.detach()
        if x.dtype not in (torch.float32, torch.float16):
            x = x.float()
        x = x.contiguous()
        B, L, _ = x.shape
        out = torch.empty((B, n_out, out_dim), device=x.device, dtype=torch.float32)
        check(lib().mudg_resampler_forward(self._h, ptr(x), MUDG_F32 if x.dtype == torch.float32 else MUDG_F16, B, L, ptr(out),
                                           cur_stream()))
        return out

    def clip_image_forward(self, img: torch.Tensor, tokens: int, width: int, heads: int, resize: bool = True) -> torch.Tensor:
        """img [B, 3, H, W] -> [B, tokens, width] fp32 (FrozenOpenCLIPImageEmbedderV2.forward; resize=True runs the
        reference's preprocess on an image in [-1, 1], resize=False takes the normalised tower input)."""
        img = img.detach()
        if img.dtype not in (torch.float32, torch.float16):
            img = img.float()
        img = img.contiguous()
        B, C, H, W = img.shape
        if C != 3:
            raise MudgError(f"clip_image_forward: {C} channels")
        out = torch.empty((B, tokens, width), device=img.device, dtype=torch.float32)
        check(lib().mudg_clip_image_forward(self._h, ptr(img), MUDG_F32 if img.dtype == torch.float32 else MUDG_F16, B, H, W,
                                            int(bool(resize)), int(heads), ptr(out), cur_stream()))
        return out

    def clip_text_forward(self, tokens: torch.Tensor, width: int, heads: int, skip_last: int = 1) -> torch.Tensor:
        """tokens [B, L] int64 (device) -> [B, L, width] fp32 (FrozenOpenCLIPEmbedder.encode_with_transformer)."""
        tokens = tokens.detach().to(torch.long).contiguous()
        B, L = tokens.shape
        out = torch.empty((B, L, width), device=tokens.device, dtype=torch.float32)
        check(lib().mudg_clip_text_forward(self._h, ptr(tokens), B, L, int(heads), int(skip_last), ptr(out), cur_stream()))
        return out

    def ddim_step(self, x, v_cond, v_uncond, noise, *, cfg_scale, guidance_rescale, sqrt_ac, sqrt_1mac, rescale,
                  a_prev, sigma):
        B = x.shape[0]
        n = x[0].numel()
        x = x.float().contiguous(); noise = noise.float().contiguous()
        v_cond = v_cond.half().contiguous()
        v_uncond = None if v_uncond is None else v_uncond.half().contiguous()
        x_prev = torch.empty_like(x); pred = torch.empty_like(x)
        f = ctypes.c_float
        check(lib().mudg_ddim_step(ptr(x), ptr(v_cond), ptr(v_uncond), ptr(noise), ptr(x_prev), ptr(pred), B,
                                   ctypes.c_int64(n), f(cfg_scale), f(guidance_rescale), f(sqrt_ac), f(sqrt_1mac),
                                   f(rescale), f(a_prev), f(sigma), cur_stream()))
        return x_prev, pred

    def workspace_bytes(self, N: int, T: int, h: int, w: int) -> int:
        return int(lib().mudg_workspace_bytes(self._h, N, T, h, w))

    def launch_count(self) -> int:
        return int(lib().mudg_launch_count(self._h))
