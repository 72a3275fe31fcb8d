"""`DDIMSampler` with separate image / text guidance (lvdm/models/samplers/ddim_multiplecond.py:210-237): three UNet
evaluations per step -- cond, uncond, and "image yes / text empty" -- mixed as

    v = v_u + cfg_img * (v_ui - v_u) + s * (v_c - v_ui)

Everything else (schedule, RNG order, update) is the single-guidance sampler's.  Selected by the driver with
--multiple_cond_cfg (virtual_render/virtual_pose_render.py:65)."""
from __future__ import annotations

import torch

from lvdm.common import noise_like
from lvdm.models.samplers.ddim import DDIMSampler as _Base
from lvdm.models.utils_diffusion import rescale_noise_cfg


class DDIMSampler(_Base):
    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1.0, noise_dropout=0.0, score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1.0, unconditional_conditioning=None, uc_type=None, cfg_img=None,
                      mask=None, x0=None, guidance_rescale=0.0, **kwargs):
        m = self.model
        if cfg_img is None:
            cfg_img = unconditional_guidance_scale
        uc_img = kwargs["unconditional_conditioning_img_nonetext"]
        if unconditional_conditioning is None or unconditional_guidance_scale == 1.0:
            out = m.apply_model(x, t, c, **kwargs)
        else:
            if all(isinstance(d, dict) for d in (c, unconditional_conditioning, uc_img)) and hasattr(m, "apply_model_multi"):
                # one 3B forward (cross-attention K/V of the three contexts cached together, shared prefix computed once)
                v_c, v_u, v_ui = m.apply_model_multi(x, t, [c, unconditional_conditioning, uc_img], **kwargs)
            else:
                v_c = m.apply_model(x, t, c, **kwargs)
                v_u = m.apply_model(x, t, unconditional_conditioning, **kwargs)
                v_ui = m.apply_model(x, t, uc_img, **kwargs)
            out = v_u + cfg_img * (v_ui - v_u) + unconditional_guidance_scale * (v_c - v_ui)
            if guidance_rescale > 0.0:
                out = rescale_noise_cfg(out, v_c, guidance_rescale=guidance_rescale)
        # the mixed prediction goes through the same fused update as an unguided step
        a_prev, sigma = float(self.ddim_alphas_prev[index]), float(self.ddim_sigmas[index])
        noise = noise_like(x.shape, x.device, repeat_noise)
        if self._fused_step_ok(x, quantize_denoised, noise_dropout, score_corrector):
            step = int(self.ddim_timesteps[index])
            rescale = float(self._h_scale_prev[index] / self._h_scale[index]) if m.use_dynamic_rescale else 1.0
            eng = m.model.diffusion_model.engine()
            return eng.ddim_step(x, out, None, noise * temperature, cfg_scale=1.0, guidance_rescale=0.0,
                                 sqrt_ac=float(self._h_sqrt_ac[step]), sqrt_1mac=float(self._h_sqrt_1mac[step]),
                                 rescale=rescale, a_prev=a_prev, sigma=sigma)
        raise NotImplementedError("DDIMSampler_multicond supports the v-parameterised CUDA path only")
