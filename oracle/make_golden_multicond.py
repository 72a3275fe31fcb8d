"""Generate tests/golden/ddim_multicond_small.npz with the UNCHANGED reference DDIMSampler of
lvdm/models/samplers/ddim_multiplecond.py: 2 steps, text scale 7.5, image scale 3.0,
guidance_rescale 0.7, eta 1 on the small UNet of make_golden.py (same seeded weights, same x / contexts).
Build container only:  python oracle/make_golden_multicond.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
    MG = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(MG)
    spec = importlib.util.spec_from_file_location("mudg_oracle", os.path.join(HERE, "mudg_oracle.py"))
    O = importlib.util.module_from_spec(spec)
    sys.modules["mudg_oracle"] = O
    spec.loader.exec_module(O)
    MG.install_shims()
    sys.path[:] = [REF] + [p for p in sys.path if os.path.abspath(p or os.getcwd()) not in (ROOT, HERE)]
    for m in [m for m in sys.modules if m.split(".")[0] in ("lvdm", "utils")]:
        del sys.modules[m]
    from lvdm.models.samplers.ddim_multiplecond import DDIMSampler
    from lvdm.models.ddpm3d import LatentVisualDiffusion
    import lvdm.models.samplers.ddim_multiplecond as _m
    assert _m.__file__.startswith(REF), _m.__file__
    torch.set_grad_enabled(False)
    small = O.UNetCfg(model_channels=64, temporal_length=4)
    B, T, H, W = 2, 4, 16, 16
    unet_kw = dict(in_channels=small.in_channels, out_channels=small.out_channels, model_channels=small.model_channels,
                   attention_resolutions=list(small.attention_resolutions), num_res_blocks=small.num_res_blocks,
                   channel_mult=list(small.channel_mult), dropout=0.1, num_head_channels=small.num_head_channels,
                   transformer_depth=1, context_dim=small.context_dim, use_linear=True, use_checkpoint=False,
                   temporal_conv=True, temporal_attention=True, temporal_selfatt_only=True,
                   use_relative_position=False, use_causal_attention=False, temporal_length=small.temporal_length,
                   addition_attention=True, image_cross_attention=True, default_fs=24, fs_condition=True,
                   class_label_condition=True)
    dd = dict(double_z=True, z_channels=4, resolution=64, in_channels=3, out_ch=3, ch=64,
              ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    to_attr = MG.to_attr
    ident = to_attr(dict(target="torch.nn.Identity"))
    DDIMSampler.register_buffer = lambda self, n, a: setattr(self, n, a)          # CPU shim (ddim_multiplecond.py:18-22)
    model = LatentVisualDiffusion(
        img_cond_stage_config=ident, image_proj_stage_config=ident,
        first_stage_config=to_attr(dict(target="lvdm.models.autoencoder.AutoencoderKL",
                                        params=dict(embed_dim=4, ddconfig=dd, lossconfig=dict(target="torch.nn.Identity")))),
        cond_stage_config=ident,
        unet_config=to_attr(dict(target="lvdm.modules.networks.openaimodel3d.UNetModel", params=unet_kw)),
        rescale_betas_zero_snr=True, parameterization="v", linear_start=0.00085, linear_end=0.012,
        num_timesteps_cond=1, timesteps=1000, first_stage_key="video", cond_stage_key="caption",
        cond_stage_trainable=False, conditioning_key="hybrid", image_size=[H, W], channels=4,
        scale_by_std=False, scale_factor=0.18215, use_ema=False, uncond_type="empty_seq",
        use_dynamic_rescale=True, base_scale=0.3, fps_condition_type="fps", perframe_ae=True)
    sd = O.seeded_state_dict(O.unet_param_shapes(small), seed=1)
    model.model.diffusion_model.load_state_dict(sd, strict=True)
    model.eval()
    g = torch.Generator().manual_seed(21)
    ctx = torch.randn(B, 77 + 16 * T, small.context_dim, generator=g)
    uc_ctx = torch.randn(B, 77 + 16 * T, small.context_dim, generator=g)
    uc_img_ctx = torch.cat([uc_ctx[:, :77], ctx[:, 77:]], dim=1)       # "image yes / text empty" (virtual_pose_render.py:101-106)
    c_concat = 0.5 * torch.randn(B, 8, T, H, W, generator=g)
    lab = torch.tensor([0, 1], dtype=torch.long)
    fs = torch.tensor([10, 10], dtype=torch.long)
    S = 2
    cond = {"c_crossattn": [ctx], "c_concat": [c_concat]}
    uc = {"c_crossattn": [uc_ctx], "c_concat": [c_concat]}
    uc2 = {"c_crossattn": [uc_img_ctx], "c_concat": [c_concat]}
    torch.manual_seed(321)
    samples, _ = DDIMSampler(model).sample(
        S=S, conditioning=cond, batch_size=B, shape=[4, T, H, W], verbose=False, unconditional_guidance_scale=7.5,
        unconditional_conditioning=uc, eta=1.0, cfg_img=3.0, mask=None, x0=None, fs=fs,
        timestep_spacing="uniform_trailing", guidance_rescale=0.7, sparse_x=None, class_label=lab[:, None],
        unconditional_conditioning_img_nonetext=uc2)
    torch.manual_seed(321)
    noises = [torch.randn(B, 4, T, H, W) for _ in range(S + 1)]
    tab = O.make_tables(base_scale=0.3)
    mine = O.ddim_sample_multicond(sd, small, tab, S=S, shape=(B, 4, T, H, W), c_concat=c_concat, context=ctx,
                                   uc_context=uc_ctx, uc_img_context=uc_img_ctx, class_label=lab, fs=fs, cfg_scale=7.5,
                                   cfg_img=3.0, guidance_rescale=0.7, eta=1.0, noises=noises)
    err = float((samples - mine).abs().max())
    print("multicond 2 steps: ref vs oracle max|d| =", err, "ref absmax", float(samples.abs().max()))
    assert err < 1e-3
    np.savez_compressed(os.path.join(OUT, "ddim_multicond_small.npz"), ctx=ctx.numpy(), uc_ctx=uc_ctx.numpy(),
                        c_concat=c_concat.numpy(), lab=lab.numpy(), fs=fs.numpy(), samples=samples.numpy())
    print("wrote ddim_multicond_small.npz")


if __name__ == "__main__":
    main()
