"""`DDIMSampler` with the reference's public surface (lvdm/models/samplers/ddim.py:10-279): `make_schedule`, `sample`,
`ddim_sampling`, `p_sample_ddim`, same arguments and return values, same RNG draw order (x_T, then one draw per step).

B200-first differences, all parity-neutral:
  * cond and uncond are evaluated as one 2B batch (`model.apply_model_cfg`) when the model offers it;
  * CFG mix + guidance rescale + v->(eps,x0) + dynamic rescale + DDIM update run as ONE fused kernel
    (`mudg_ddim_step`) instead of ~12 elementwise/reduction launches;
  * buffers live on the model's device (the reference hard-codes "cuda", ddim.py:18-22).
"""
from __future__ import annotations

import numpy as np
import torch

from lvdm.common import noise_like
from lvdm.models.utils_diffusion import make_ddim_sampling_parameters, make_ddim_timesteps, rescale_noise_cfg


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", **kwargs):
        super().__init__()
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self.counter = 0

    def register_buffer(self, name, attr):
        if torch.is_tensor(attr) and attr.device != self.model.device:
            attr = attr.to(self.model.device)
        setattr(self, name, attr)

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0.0, verbose=True):
        m = self.model
        self.ddim_timesteps = make_ddim_timesteps(ddim_discretize, ddim_num_steps, self.ddpm_num_timesteps, verbose=verbose)
        ac = m.alphas_cumprod
        assert ac.shape[0] == self.ddpm_num_timesteps, "alphas have to be defined for each timestep"
        if m.use_dynamic_rescale:
            self.ddim_scale_arr = m.scale_arr[self.ddim_timesteps]
            self.ddim_scale_arr_prev = torch.cat([self.ddim_scale_arr[0:1], self.ddim_scale_arr[:-1]])
        f32 = lambda x: x.clone().detach().to(torch.float32).to(m.device)
        acc = ac.detach().cpu()
        self.register_buffer("betas", f32(m.betas))
        self.register_buffer("alphas_cumprod", f32(ac))
        self.register_buffer("alphas_cumprod_prev", f32(m.alphas_cumprod_prev))
        self.register_buffer("sqrt_alphas_cumprod", f32(acc.sqrt()))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", f32((1.0 - acc).sqrt()))
        sig, a, a_prev = make_ddim_sampling_parameters(acc, self.ddim_timesteps, ddim_eta, verbose=verbose)
        self.ddim_sigmas, self.ddim_alphas, self.ddim_alphas_prev = sig, a, a_prev
        self.ddim_sqrt_one_minus_alphas = np.sqrt(1.0 - a)
        # host copies of the per-timestep scalars the fused step needs (no device sync inside the loop)
        # (the model's own buffers, derived in float64 as the reference indexes them: ddpm3d.py:239-251)
        self._h_sqrt_ac = m.sqrt_alphas_cumprod.detach().cpu().numpy()
        self._h_sqrt_1mac = m.sqrt_one_minus_alphas_cumprod.detach().cpu().numpy()
        if m.use_dynamic_rescale:
            self._h_scale = self.ddim_scale_arr.detach().cpu().numpy()
            self._h_scale_prev = self.ddim_scale_arr_prev.detach().cpu().numpy()

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0.0, mask=None, x0=None, temperature=1.0, noise_dropout=0.0,
               score_corrector=None, corrector_kwargs=None, verbose=True, schedule_verbose=False, x_T=None,
               log_every_t=100, unconditional_guidance_scale=1.0, unconditional_conditioning=None, precision=None,
               fs=None, timestep_spacing="uniform", guidance_rescale=0.0, **kwargs):
        if conditioning is not None:
            first = conditioning[next(iter(conditioning))] if isinstance(conditioning, dict) else conditioning
            cbs = (first[0] if isinstance(first, (list, tuple)) else first).shape[0]
            if cbs != batch_size:
                print(f"Warning: Got {cbs} conditionings but batch-size is {batch_size}")
        self.make_schedule(ddim_num_steps=S, ddim_discretize=timestep_spacing, ddim_eta=eta, verbose=schedule_verbose)
        size = (batch_size, *shape)
        return self.ddim_sampling(conditioning, size, callback=callback, img_callback=img_callback,
                                  quantize_denoised=quantize_x0, mask=mask, x0=x0, ddim_use_original_steps=False,
                                  noise_dropout=noise_dropout, temperature=temperature, score_corrector=score_corrector,
                                  corrector_kwargs=corrector_kwargs, x_T=x_T, log_every_t=log_every_t,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning, verbose=verbose,
                                  precision=precision, fs=fs, guidance_rescale=guidance_rescale, **kwargs)

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100, temperature=1.0,
                      noise_dropout=0.0, score_corrector=None, corrector_kwargs=None, unconditional_guidance_scale=1.0,
                      unconditional_conditioning=None, verbose=True, precision=None, fs=None, guidance_rescale=0.0,
                      **kwargs):
        if ddim_use_original_steps:
            raise NotImplementedError("ddim_use_original_steps is not used by the MuDG sampler path")
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T       # RNG draw #0 (ddim.py:145)
        if precision == 16:
            img = img.to(dtype=torch.float16)
        if timesteps is None:
            timesteps = self.ddim_timesteps
        else:
            n = self.ddim_timesteps.shape[0]
            timesteps = self.ddim_timesteps[:int(min(timesteps / n, 1) * n) - 1]
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        total = timesteps.shape[0]
        clean_cond = kwargs.pop("clean_cond", False)
        order = np.flip(timesteps)
        if verbose:
            from tqdm import tqdm
            order = tqdm(order, desc="DDIM Sampler", total=total)
        for i, step in enumerate(order):
            index = total - i - 1
            ts = torch.full((b,), int(step), device=device, dtype=torch.long)
            if mask is not None:
                assert x0 is not None
                img_orig = x0 if clean_cond else self.model.q_sample(x0, ts)
                img = img_orig * mask + (1.0 - mask) * img
            img, pred_x0 = self.p_sample_ddim(img, cond, ts, index=index, quantize_denoised=quantize_denoised,
                                              temperature=temperature, noise_dropout=noise_dropout,
                                              score_corrector=score_corrector, corrector_kwargs=corrector_kwargs,
                                              unconditional_guidance_scale=unconditional_guidance_scale,
                                              unconditional_conditioning=unconditional_conditioning, mask=mask, x0=x0,
                                              fs=fs, guidance_rescale=guidance_rescale, **kwargs)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total - 1:
                intermediates["x_inter"].append(img)
                intermediates["pred_x0"].append(pred_x0)
        return img, intermediates

    def _fused_step_ok(self, x, quantize_denoised, noise_dropout, score_corrector):
        m = self.model
        return (x.is_cuda and m.parameterization == "v" and not quantize_denoised and noise_dropout == 0.0
                and score_corrector is None and hasattr(m.model.diffusion_model, "engine"))

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1.0, noise_dropout=0.0, score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1.0, unconditional_conditioning=None, uc_type=None,
                      conditional_guidance_scale_temporal=None, mask=None, x0=None, guidance_rescale=0.0, **kwargs):
        m = self.model
        guided = unconditional_conditioning is not None and unconditional_guidance_scale != 1.0
        e_cond, e_uncond = None, None
        if not guided:
            e_cond = m.apply_model(x, t, c, **kwargs)
        elif isinstance(c, dict) and isinstance(unconditional_conditioning, dict) and hasattr(m, "apply_model_cfg"):
            e_cond, e_uncond = m.apply_model_cfg(x, t, c, unconditional_conditioning, **kwargs)
        else:
            e_cond = m.apply_model(x, t, c, **kwargs)
            e_uncond = m.apply_model(x, t, unconditional_conditioning, **kwargs)

        a_prev, sigma = float(self.ddim_alphas_prev[index]), float(self.ddim_sigmas[index])
        noise = noise_like(x.shape, x.device, repeat_noise)                   # one draw per step, always (ddim.py:273)
        if self._fused_step_ok(x, quantize_denoised, noise_dropout, score_corrector):
            step = int(self.ddim_timesteps[index])
            rescale = float(self._h_scale_prev[index] / self._h_scale[index]) if m.use_dynamic_rescale else 1.0
            eng = m.model.diffusion_model.engine()
            return eng.ddim_step(x, e_cond, e_uncond, noise * temperature, cfg_scale=float(unconditional_guidance_scale),
                                 guidance_rescale=float(guidance_rescale), sqrt_ac=float(self._h_sqrt_ac[step]),
                                 sqrt_1mac=float(self._h_sqrt_1mac[step]), rescale=rescale, a_prev=a_prev, sigma=sigma)

        # generic path (eps / x0 parameterisations, quantisation, dropout): plain tensor ops, reference order
        out = e_cond
        if guided:
            out = e_uncond + unconditional_guidance_scale * (e_cond - e_uncond)
            if guidance_rescale > 0.0:
                out = rescale_noise_cfg(out, e_cond, guidance_rescale=guidance_rescale)
        e_t = m.predict_eps_from_z_and_v(x, t, out) if m.parameterization == "v" else out
        if score_corrector is not None:
            assert m.parameterization == "eps", "not implemented"
            e_t = score_corrector.modify_score(m, e_t, x, t, c, **corrector_kwargs)
        size = (x.shape[0],) + (1,) * (x.dim() - 1)
        full = lambda v: torch.full(size, float(v), device=x.device)
        if m.parameterization != "v":
            pred_x0 = (x - full(self.ddim_sqrt_one_minus_alphas[index]) * e_t) / full(self.ddim_alphas[index]).sqrt()
        else:
            pred_x0 = m.predict_start_from_z_and_v(x, t, out)
        if m.use_dynamic_rescale:
            pred_x0 = pred_x0 * (full(self.ddim_scale_arr_prev[index]) / full(self.ddim_scale_arr[index]))
        if quantize_denoised:
            pred_x0, _, *_ = m.first_stage_model.quantize(pred_x0)
        dir_xt = (1.0 - full(a_prev) - full(sigma) ** 2).sqrt() * e_t
        noise = full(sigma) * noise * temperature
        if noise_dropout > 0.0:
            noise = torch.nn.functional.dropout(noise, p=noise_dropout)
        return full(a_prev).sqrt() * pred_x0 + dir_xt + noise, pred_x0
