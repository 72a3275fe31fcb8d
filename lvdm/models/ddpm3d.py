"""Inference surface of the reference's Lightning model classes (lvdm/models/ddpm3d.py) without Lightning:
`DDPM` -> `LatentDiffusion` -> `LatentVisualDiffusion` and `DiffusionWrapper`, with the same constructor keys
(both *_infer.yaml configs), buffers, attributes and state-dict key layout, so
virtual_render/virtual_pose_render.py builds, loads and drives the model unchanged.  Training-only methods
(losses, optimisers, logging, EMA) are out of scope for the sampler path and not provided.
"""
from __future__ import annotations

from functools import partial

import numpy as np
import torch
import torch.nn as nn
from einops import rearrange

from lvdm.common import default, extract_into_tensor
from lvdm.distributions import DiagonalGaussianDistribution
from lvdm.models.utils_diffusion import make_beta_schedule, rescale_zero_terminal_snr
from utils.utils import instantiate_from_config


def _frozen(module):
    module.eval()
    module.train = lambda mode=True: module       # "disabled_train" (basics.py:12-15)
    for p in module.parameters():
        p.requires_grad = False
    return module


def _get(cfg, *path):
    for k in path:
        cfg = cfg[k] if isinstance(cfg, dict) or hasattr(cfg, "__getitem__") else getattr(cfg, k)
    return cfg


class DiffusionWrapper(nn.Module):
    """ddpm3d.py:1303-1372 (conditioning-key dispatch; the MuDG configs use 'hybrid')."""

    def __init__(self, diff_model_config, conditioning_key):
        super().__init__()
        self.diffusion_model = instantiate_from_config(diff_model_config)
        self.conditioning_key = conditioning_key

    def _context(self, c_crossattn):
        """torch.cat(c_crossattn, 1) of the reference (ddpm3d.py:1322), but the SAME tensor object for the same inputs:
        the context is constant over the DDIM steps and the UNet keys its cross-attention K/V cache on the tensor, so a
        fresh concatenation per step would recompute all 16 layers' K/V (and synchronise) fifty times per clip."""
        if len(c_crossattn) == 1:
            return c_crossattn[0]
        key = tuple((id(c), c._version) for c in c_crossattn)
        cache = getattr(self, "_ctx_cache", None)
        if cache is None or cache[0] != key:
            cache = (key, torch.cat(c_crossattn, 1), list(c_crossattn))
            self._ctx_cache = cache
        return cache[1]

    def forward(self, x, t, c_label=None, c_concat=None, c_crossattn=None, c_adm=None, s=None, mask=None, **kwargs):
        key = self.conditioning_key
        if key is None:
            return self.diffusion_model(x, t)
        if key == "concat":
            return self.diffusion_model(torch.cat([x] + c_concat, dim=1), t, **kwargs)
        if key == "crossattn":
            return self.diffusion_model(x, t, context=self._context(c_crossattn), **kwargs)
        if key == "hybrid":
            xc = torch.cat([x] + c_concat, dim=1)            # [b, 4+8, t, h, w]
            return self.diffusion_model(xc, t, c_label=c_label, context=self._context(c_crossattn), **kwargs)
        raise NotImplementedError(f"conditioning_key {key!r} is not used by the MuDG sampler path")


class DDPM(nn.Module):
    def __init__(self, unet_config, timesteps=1000, beta_schedule="linear", loss_type="l2", ckpt_path=None,
                 ignore_keys=(), load_only_unet=False, monitor=None, use_ema=True, first_stage_key="image",
                 image_size=256, channels=3, log_every_t=100, clip_denoised=True, linear_start=1e-4, linear_end=2e-2,
                 cosine_s=8e-3, given_betas=None, original_elbo_weight=0.0, v_posterior=0.0, l_simple_weight=1.0,
                 conditioning_key=None, parameterization="eps", scheduler_config=None,
                 use_positional_encodings=False, learn_logvar=False, logvar_init=0.0, rescale_betas_zero_snr=False):
        super().__init__()
        assert parameterization in ("eps", "x0", "v")
        if use_ema:
            raise NotImplementedError("use_ema=True is a training feature; the infer configs set use_ema: False")
        self.parameterization = parameterization
        self.cond_stage_model = None
        self.clip_denoised, self.log_every_t = clip_denoised, log_every_t
        self.first_stage_key, self.channels = first_stage_key, channels
        self.temporal_length = _get(unet_config, "params", "temporal_length")
        self.image_size = [image_size, image_size] if isinstance(image_size, int) else image_size
        self.use_positional_encodings = use_positional_encodings
        self.model = DiffusionWrapper(unet_config, conditioning_key)
        self.use_ema = False
        self.rescale_betas_zero_snr = rescale_betas_zero_snr
        self.v_posterior, self.original_elbo_weight, self.l_simple_weight = v_posterior, original_elbo_weight, l_simple_weight
        if monitor is not None:
            self.monitor = monitor
        self.register_schedule(given_betas=given_betas, beta_schedule=beta_schedule, timesteps=timesteps,
                               linear_start=linear_start, linear_end=linear_end, cosine_s=cosine_s)
        self.given_betas, self.beta_schedule, self.timesteps, self.cosine_s = given_betas, beta_schedule, timesteps, cosine_s
        self.loss_type, self.learn_logvar = loss_type, learn_logvar
        self.logvar = torch.full(fill_value=logvar_init, size=(self.num_timesteps,))
        self._ckpt = (ckpt_path, ignore_keys, load_only_unet)

    @property
    def device(self):
        return self.betas.device

    def register_schedule(self, given_betas=None, beta_schedule="linear", timesteps=1000, linear_start=1e-4,
                          linear_end=2e-2, cosine_s=8e-3):
        """Same persistent buffers as ddpm3d.py:123-186 (fp32 tables computed in f64)."""
        betas = given_betas if given_betas is not None else make_beta_schedule(
            beta_schedule, timesteps, linear_start=linear_start, linear_end=linear_end, cosine_s=cosine_s)
        if self.rescale_betas_zero_snr:
            betas = rescale_zero_terminal_snr(betas)
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        self.num_timesteps = int(betas.shape[0])
        self.linear_start, self.linear_end = linear_start, linear_end
        f32 = partial(torch.tensor, dtype=torch.float32)
        reg = self.register_buffer
        reg("betas", f32(betas))
        reg("alphas_cumprod", f32(ac))
        reg("alphas_cumprod_prev", f32(ac_prev))
        reg("sqrt_alphas_cumprod", f32(np.sqrt(ac)))
        reg("sqrt_one_minus_alphas_cumprod", f32(np.sqrt(1.0 - ac)))
        with np.errstate(divide="ignore", invalid="ignore"):
            reg("log_one_minus_alphas_cumprod", f32(np.log(1.0 - ac)))
            if self.parameterization != "v":
                reg("sqrt_recip_alphas_cumprod", f32(np.sqrt(1.0 / ac)))
                reg("sqrt_recipm1_alphas_cumprod", f32(np.sqrt(1.0 / ac - 1)))
            else:
                reg("sqrt_recip_alphas_cumprod", torch.zeros(self.num_timesteps))
                reg("sqrt_recipm1_alphas_cumprod", torch.zeros(self.num_timesteps))
            post_var = (1 - self.v_posterior) * betas * (1.0 - ac_prev) / (1.0 - ac) + self.v_posterior * betas
            reg("posterior_variance", f32(post_var))
            reg("posterior_log_variance_clipped", f32(np.log(np.maximum(post_var, 1e-20))))
            reg("posterior_mean_coef1", f32(betas * np.sqrt(ac_prev) / (1.0 - ac)))
            reg("posterior_mean_coef2", f32((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)))
        reg("lvlb_weights", torch.ones(self.num_timesteps), persistent=False)

    # v-parameterisation helpers (ddpm3d.py:239-251,305-308)
    def predict_start_from_z_and_v(self, x_t, t, v):
        return (extract_into_tensor(self.sqrt_alphas_cumprod, t, x_t.shape) * x_t
                - extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_t.shape) * v)

    def predict_eps_from_z_and_v(self, x_t, t, v):
        return (extract_into_tensor(self.sqrt_alphas_cumprod, t, x_t.shape) * v
                + extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_t.shape) * x_t)

    def q_sample(self, x_start, t, noise=None):
        noise = default(noise, lambda: torch.randn_like(x_start))
        return (extract_into_tensor(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
                + extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)

    def init_from_ckpt(self, path, ignore_keys=(), only_model=False):
        sd = torch.load(path, map_location="cpu")
        sd = sd.get("state_dict", sd)
        sd = {k: v for k, v in sd.items() if not any(k.startswith(i) for i in ignore_keys)}
        return (self.model if only_model else self).load_state_dict(sd, strict=False)


class LatentDiffusion(DDPM):
    """ddpm3d.py:464-739 (inference surface)."""

    def __init__(self, first_stage_config, cond_stage_config, num_timesteps_cond=None, cond_stage_key="caption",
                 cond_stage_trainable=False, cond_stage_forward=None, conditioning_key=None, uncond_prob=0.2,
                 uncond_type="empty_seq", scale_factor=1.0, scale_by_std=False, encoder_type="2d", only_model=False,
                 noise_strength=0, use_dynamic_rescale=False, base_scale=0.7, turning_step=400, interp_mode=False,
                 fps_condition_type="fs", perframe_ae=False, logdir=None, rand_cond_frame=False,
                 en_and_decode_n_samples_a_time=None, *args, **kwargs):
        self.num_timesteps_cond = default(num_timesteps_cond, 1)
        self.scale_by_std = scale_by_std
        assert self.num_timesteps_cond <= kwargs["timesteps"]
        ckpt_path = kwargs.pop("ckpt_path", None)
        ignore_keys = kwargs.pop("ignore_keys", [])
        super().__init__(conditioning_key=default(conditioning_key, "crossattn"), *args, **kwargs)
        self.cond_stage_trainable, self.cond_stage_key = cond_stage_trainable, cond_stage_key
        self.noise_strength, self.use_dynamic_rescale, self.interp_mode = noise_strength, use_dynamic_rescale, interp_mode
        self.fps_condition_type, self.perframe_ae = fps_condition_type, perframe_ae
        self.logdir, self.rand_cond_frame = logdir, rand_cond_frame
        self.en_and_decode_n_samples_a_time = en_and_decode_n_samples_a_time
        try:
            self.num_downs = len(_get(first_stage_config, "params", "ddconfig", "ch_mult")) - 1
        except Exception:
            self.num_downs = 0
        if not scale_by_std:
            self.scale_factor = scale_factor
        else:
            self.register_buffer("scale_factor", torch.tensor(scale_factor))
        if use_dynamic_rescale:        # ddpm3d.py:522-527
            arr = np.concatenate((np.linspace(1.0, base_scale, turning_step), np.full(self.num_timesteps, base_scale)))
            self.register_buffer("scale_arr", torch.tensor(arr, dtype=torch.float32))
        self.first_stage_model = _frozen(instantiate_from_config(first_stage_config))
        model = instantiate_from_config(cond_stage_config)
        self.cond_stage_model = model if cond_stage_trainable else _frozen(model)
        self.first_stage_config, self.cond_stage_config = first_stage_config, cond_stage_config
        self.clip_denoised = False
        self.cond_stage_forward, self.encoder_type = cond_stage_forward, encoder_type
        assert encoder_type in ("2d", "3d") and uncond_type in ("zero_embed", "empty_seq")
        self.uncond_prob, self.classifier_free_guidance, self.uncond_type = uncond_prob, uncond_prob > 0, uncond_type
        self.restarted_from_ckpt = False
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys, only_model=only_model)
            self.restarted_from_ckpt = True

    def get_learned_conditioning(self, c):
        m = self.cond_stage_model
        if self.cond_stage_forward is not None:
            return getattr(m, self.cond_stage_forward)(c)
        if callable(getattr(m, "encode", None)):
            c = m.encode(c)
            return c.mode() if isinstance(c, DiagonalGaussianDistribution) else c
        return m(c)

    def get_first_stage_encoding(self, encoder_posterior, noise=None):
        z = encoder_posterior.sample(noise=noise) if isinstance(encoder_posterior, DiagonalGaussianDistribution) \
            else encoder_posterior
        return self.scale_factor * z

    @torch.no_grad()
    def encode_first_stage(self, x):
        """ddpm3d.py:620-644 -- per-frame when perframe_ae (posterior noise is drawn per frame, on the CPU generator)."""
        reshape_back = self.encoder_type == "2d" and x.dim() == 5
        if reshape_back:
            b, _, t, _, _ = x.shape
            x = rearrange(x, "b c t h w -> (b t) c h w")
        if not self.perframe_ae:
            z = self.get_first_stage_encoding(self.first_stage_model.encode(x)).detach()
        else:
            z = torch.cat([self.get_first_stage_encoding(self.first_stage_model.encode(x[i:i + 1])).detach()
                           for i in range(x.shape[0])], dim=0)
        return rearrange(z, "(b t) c h w -> b c t h w", b=b, t=t) if reshape_back else z

    def decode_core(self, z, **kwargs):
        """ddpm3d.py:646-667.  All frames go through ONE library call; the library itself decodes frame by frame,
        which is what perframe_ae does and is numerically identical to the batched path (no cross-frame op)."""
        reshape_back = self.encoder_type == "2d" and z.dim() == 5
        if reshape_back:
            b, _, t, _, _ = z.shape
            z = rearrange(z, "b c t h w -> (b t) c h w")
        out = self.first_stage_model.decode((1.0 / self.scale_factor) * z, **kwargs)
        return rearrange(out, "(b t) c h w -> b c t h w", b=b, t=t) if reshape_back else out

    @torch.no_grad()
    def decode_first_stage(self, z, **kwargs):
        return self.decode_core(z, **kwargs)

    def apply_model(self, x_noisy, t, cond, **kwargs):
        """ddpm3d.py:723-739."""
        if not isinstance(cond, dict):
            cond = {"c_concat" if self.model.conditioning_key == "concat" else "c_crossattn":
                    cond if isinstance(cond, list) else [cond]}
        class_label = kwargs.get("class_label", None)[:, 0]
        out = self.model(x_noisy, t, class_label, **cond, **kwargs)
        return out[0] if isinstance(out, tuple) else out

    def apply_model_multi(self, x_noisy, t, conds, **kwargs):
        """The 2 (ddim.py:221-222) or 3 (ddim_multiplecond.py:213-235) UNet evaluations of one guided DDIM step as ONE
        batch of len(conds)*B: the calls differ only in the conditioning.  When they differ only in the cross-attention
        context (same c_concat: the driver's case) the copies also share the layers before the first cross-attention
        (UNetModel.forward(mudg_shared_copies=...)).  Returns one output per conditioning."""
        b, d = x_noisy.shape[0], len(conds)
        tensors = [v for c in conds for k in sorted(c) for v in c[k]]
        key = tuple(id(v) for v in tensors) + tuple(v._version for v in tensors)
        cache = getattr(self, "_multi_cache", None)
        if cache is None or cache[0] != key:
            both = {k: [torch.cat(vs, dim=0) for vs in zip(*(c[k] for c in conds))] for k in conds[0]}
            shared = all(k == "c_crossattn" or all(v is w or torch.equal(v, w) for c in conds[1:] for v, w in zip(c[k], conds[0][k]))
                         for k in conds[0])
            cache = (key, both, tuple(conds), shared)
            self._multi_cache = cache
        both, shared = cache[1], cache[3]
        kw = {k: (torch.cat([v] * d, dim=0) if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == b else v)
              for k, v in kwargs.items()}
        if shared and hasattr(self.model.diffusion_model, "engine"):
            kw["mudg_shared_copies"] = d
        out = self.apply_model(torch.cat([x_noisy] * d, dim=0), torch.cat([t] * d, dim=0), both, **kw)
        return [out[i * b:(i + 1) * b] for i in range(d)]

    def apply_model_cfg(self, x_noisy, t, cond, uncond, **kwargs):
        """cond + uncond evaluated as ONE batch of 2B; returns (e_t_cond, e_t_uncond)."""
        e_c, e_u = self.apply_model_multi(x_noisy, t, [cond, uncond], **kwargs)
        return e_c, e_u


class LatentVisualDiffusion(LatentDiffusion):
    """ddpm3d.py:1033-1054: adds the frozen image embedder and the image-context projector (Resampler)."""

    def __init__(self, img_cond_stage_config, image_proj_stage_config, freeze_embedder=True,
                 image_proj_model_trainable=True, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.image_proj_model_trainable = image_proj_model_trainable
        emb = instantiate_from_config(img_cond_stage_config)
        self.embedder = _frozen(emb) if freeze_embedder else emb
        proj = instantiate_from_config(image_proj_stage_config)
        self.image_proj_model = proj if image_proj_model_trainable else _frozen(proj)
