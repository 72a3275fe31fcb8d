"""Per-window synthesis flow of the reference driver (virtual_render/virtual_pose_render.py:62-147, `image_guided_synthesis`),
restated against this package's drop-in classes so the whole window -- conditioning encoders -> VAE encode of the sparse
RGB / depth videos -> 50-step CFG DDIM sampling -> VAE decode -- can be run and tested without the driver's file I/O.
The unchanged driver calls exactly the same model / sampler methods in the same order."""
from __future__ import annotations

import torch
from einops import rearrange

from lvdm.models.samplers.ddim import DDIMSampler
from lvdm.models.samplers.ddim_multiplecond import DDIMSampler as DDIMSamplerMulticond


def get_latent_z(model, videos):
    b, c, t, h, w = videos.shape
    z = model.encode_first_stage(rearrange(videos, "b c t h w -> (b t) c h w"))
    return rearrange(z, "(b t) c h w -> b c t h w", b=b, t=t)


@torch.no_grad()
def image_guided_synthesis(model, prompts, sparse_x, sparse_depth, class_label, noise_shape, n_samples=1, ddim_steps=50,
                           ddim_eta=1.0, unconditional_guidance_scale=1.0, cfg_img=None, fs=None, text_input=False,
                           multiple_cond_cfg=False, timestep_spacing="uniform", guidance_rescale=0.0, **kwargs):
    sampler = DDIMSamplerMulticond(model) if multiple_cond_cfg else DDIMSampler(model)
    batch = sparse_x.shape[0]
    fs = torch.tensor([fs] * batch, dtype=torch.long, device=model.device)
    if not text_input:
        prompts = [""] * batch
    img = sparse_x[:, :, 0]
    img_emb = model.image_proj_model(model.embedder(img))
    cond_emb = model.get_learned_conditioning(prompts)
    cond = {"c_crossattn": [torch.cat([cond_emb, img_emb], dim=1)]}
    img_cat = None
    if model.model.conditioning_key == "hybrid":
        sparse_z = get_latent_z(model, sparse_x)
        depth_z = get_latent_z(model, sparse_depth)
        kwargs.update({"sparse_x": sparse_z, "class_label": class_label})
        img_cat = torch.cat([sparse_z, depth_z], dim=1)
        cond["c_concat"] = [img_cat]
    uc = None
    if unconditional_guidance_scale != 1.0:
        uc_emb = model.get_learned_conditioning(batch * [""]) if model.uncond_type == "empty_seq" else torch.zeros_like(cond_emb)
        uc_img = model.image_proj_model(model.embedder(torch.zeros_like(img)))
        uc = {"c_crossattn": [torch.cat([uc_emb, uc_img], dim=1)]}
        if img_cat is not None:
            uc["c_concat"] = [img_cat]
    if multiple_cond_cfg and cfg_img != 1.0:
        uc2 = {"c_crossattn": [torch.cat([uc_emb, img_emb], dim=1)]}
        if img_cat is not None:
            uc2["c_concat"] = [img_cat]
        kwargs["unconditional_conditioning_img_nonetext"] = uc2
    else:
        kwargs["unconditional_conditioning_img_nonetext"] = None
    variants = []
    for _ in range(n_samples):
        samples, _ = sampler.sample(S=ddim_steps, conditioning=cond, batch_size=batch, shape=noise_shape[1:], verbose=False,
                                    unconditional_guidance_scale=unconditional_guidance_scale, unconditional_conditioning=uc,
                                    eta=ddim_eta, cfg_img=cfg_img, mask=None, x0=None, fs=fs,
                                    timestep_spacing=timestep_spacing, guidance_rescale=guidance_rescale, **kwargs)
        variants.append(model.decode_first_stage(samples))
    return torch.stack(variants).permute(1, 0, 2, 3, 4, 5)
