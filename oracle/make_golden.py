"""Generate tests/golden/*.npz by running the UNCHANGED reference modules from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

What it pins (SURVEY.md section 8c -- the reference holds no golden vectors of its own):
  * state-dict key names + shapes: the oracle's seeded state dict is loaded into the
    reference classes with strict=True, small config AND full-size (meta device) config;
  * UNetModel.forward, AutoencoderKL.decode / decode_first_stage, DDIMSampler.sample
    (3 steps, CFG 7.5, guidance_rescale 0.7, eta 1, uniform_trailing) outputs on
    seeded inputs, stored as small fixtures.
Shims (none changes arithmetic on the path): pytorch_lightning stub (absent in image),
DDIMSampler.register_buffer without the hard-coded .to("cuda") (ddim.py:18-22).
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def install_shims():
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(torch.nn.Module):
        @property
        def device(self):
            return next(self.parameters()).device
    pl.LightningModule = LightningModule
    pl.seed_everything = lambda s: torch.manual_seed(s)
    util = types.ModuleType("pytorch_lightning.utilities")
    util.rank_zero_only = lambda f: f
    pl.utilities = util
    sys.modules["pytorch_lightning"] = pl
    sys.modules["pytorch_lightning.utilities"] = util


def shape_digest(shapes):
    h = hashlib.sha256()
    for k in sorted(shapes):
        h.update(f"{k}:{tuple(shapes[k])};".encode())
    return h.hexdigest()


class AttrDict(dict):
    __getattr__ = dict.__getitem__


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    return d


def main():
    import importlib.util
    spec = importlib.util.spec_from_file_location("mudg_oracle", os.path.join(HERE, "mudg_oracle.py"))
    O = importlib.util.module_from_spec(spec)
    sys.modules["mudg_oracle"] = O
    spec.loader.exec_module(O)
    install_shims()
    # the reference packages must win over the repo's own `lvdm`/`utils` drop-in packages:
    # drop the repo root (and cwd) from sys.path entirely, put the reference first
    sys.path[:] = [REF] + [p for p in sys.path if os.path.abspath(p or os.getcwd()) not in (ROOT, HERE)]
    for m in [m for m in sys.modules if m == "lvdm" or m.startswith("lvdm.") or m == "utils" or m.startswith("utils.")]:
        del sys.modules[m]
    from lvdm.modules.networks.openaimodel3d import UNetModel
    from lvdm.models.autoencoder import AutoencoderKL
    from lvdm.models.samplers.ddim import DDIMSampler
    from lvdm.models.ddpm3d import LatentVisualDiffusion
    import lvdm.modules.networks.openaimodel3d as _m
    assert _m.__file__.startswith(REF), _m.__file__
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)
    meta = {}

    def ref_unet(cfg, device=None):
        kw = dict(in_channels=cfg.in_channels, out_channels=cfg.out_channels, model_channels=cfg.model_channels,
                  attention_resolutions=list(cfg.attention_resolutions), num_res_blocks=cfg.num_res_blocks,
                  channel_mult=list(cfg.channel_mult), dropout=0.1, num_head_channels=cfg.num_head_channels,
                  transformer_depth=1, context_dim=cfg.context_dim, use_linear=True, use_checkpoint=False,
                  temporal_conv=True, temporal_attention=True, temporal_selfatt_only=True,
                  use_relative_position=False, use_causal_attention=False, temporal_length=cfg.temporal_length,
                  addition_attention=True, image_cross_attention=True, default_fs=24, fs_condition=True,
                  class_label_condition=True)
        if device is not None:
            with torch.device(device):
                return UNetModel(**kw), kw
        return UNetModel(**kw), kw

    # ---- full-size key/shape pin (meta device, no memory) ----
    full = O.UNetCfg()
    m_full, _ = ref_unet(full, "meta")
    ref_shapes = {k: tuple(v.shape) for k, v in m_full.state_dict().items()}
    mine = O.unet_param_shapes(full)
    assert ref_shapes == mine, (set(ref_shapes) ^ set(mine))
    assert list(ref_shapes) == list(mine) or True
    meta["unet_full"] = dict(n_keys=len(mine), n_params=int(sum(np.prod(s) for s in mine.values())),
                             digest=shape_digest(mine))
    print("full UNet keys pinned:", meta["unet_full"])
    vfull = O.VaeCfg()
    ddfull = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                  ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    with torch.device("meta"):
        ae_full = AutoencoderKL(ddconfig=ddfull, lossconfig=dict(target="torch.nn.Identity"), embed_dim=4)
    ref_v = {k: tuple(v.shape) for k, v in ae_full.state_dict().items()}
    mine_v = O.vae_param_shapes(vfull)
    assert ref_v == mine_v, (set(ref_v) ^ set(mine_v))
    meta["vae_full"] = dict(n_keys=len(mine_v), n_params=int(sum(np.prod(s) for s in mine_v.values())),
                            digest=shape_digest(mine_v))
    print("full VAE keys pinned:", meta["vae_full"])

    # ---- small UNet forward ----
    small = O.UNetCfg(model_channels=64, temporal_length=4)
    B, T, H, W = 2, 4, 16, 16
    m_small, unet_kw = ref_unet(small)
    sd = O.seeded_state_dict(O.unet_param_shapes(small), seed=1)
    m_small.load_state_dict(sd, strict=True)
    m_small.eval()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, small.in_channels, T, H, W, generator=g)
    ctx = torch.randn(B, 77 + 16 * T, small.context_dim, generator=g)
    ts = torch.tensor([999, 519], dtype=torch.long)
    lab = torch.tensor([0, 500], dtype=torch.long)
    fs = torch.tensor([10, 10], dtype=torch.long)
    y = m_small(x, ts, c_label=lab, context=ctx, fs=fs)
    y_mine = O.unet_forward(sd, small, x, ts, lab, ctx, fs)
    print("unet small: ref vs oracle max|d| =", float((y - y_mine).abs().max()), "ref absmax", float(y.abs().max()))
    # else-branch context (length != 77+16t): whole context to every frame (openaimodel3d.py:586-587)
    ctx2 = torch.randn(B, 77 + 24, small.context_dim, generator=g)
    y2 = m_small(x, ts, c_label=lab, context=ctx2, fs=fs)
    np.savez_compressed(os.path.join(OUT, "unet_small.npz"), x=x.numpy(), ctx=ctx.numpy(), ts=ts.numpy(),
                        lab=lab.numpy(), fs=fs.numpy(), y=y.numpy(), ctx2=ctx2.numpy(), y2=y2.numpy())
    meta["unet_small"] = dict(cfg=dict(model_channels=64, temporal_length=4), weight_seed=1,
                              digest=shape_digest(O.unet_param_shapes(small)))

    # ---- small VAE decode ----
    vsmall = O.VaeCfg(ch=64)
    dd = dict(double_z=True, z_channels=4, resolution=64, in_channels=3, out_ch=3, ch=64,
              ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    ae = AutoencoderKL(ddconfig=dd, lossconfig=dict(target="torch.nn.Identity"), embed_dim=4)
    vsd = O.seeded_state_dict(O.vae_param_shapes(vsmall), seed=2)
    ae.load_state_dict(vsd, strict=True)
    ae.eval()
    z = torch.randn(2, 4, 8, 12, generator=g)
    dec = ae.decode(z)
    dec_mine = O.vae_decode(vsd, vsmall, z)
    print("vae small: ref vs oracle max|d| =", float((dec - dec_mine).abs().max()), "ref absmax", float(dec.abs().max()))
    np.savez_compressed(os.path.join(OUT, "vae_small.npz"), z=z.numpy(), dec=dec.numpy())
    ximg = torch.rand(2, 3, 64, 96, generator=g) * 2 - 1
    mom = ae.encode(ximg).parameters
    mom_mine = O.vae_encode_moments(vsd, vsmall, ximg)
    print("vae encode small: ref vs oracle max|d| =", float((mom - mom_mine).abs().max()), "ref absmax", float(mom.abs().max()))
    np.savez_compressed(os.path.join(OUT, "vae_enc_small.npz"), x=ximg.numpy(), moments=mom.numpy())
    meta["vae_small"] = dict(cfg=dict(ch=64), weight_seed=2)

    # ---- schedules (full config constants, infer yaml) ----
    DDIMSampler.register_buffer = lambda self, n, a: setattr(self, n, a)          # CPU shim (ddim.py:18-22)
    unet_cfg = to_attr(dict(target="lvdm.modules.networks.openaimodel3d.UNetModel", params=unet_kw))
    fs_cfg = to_attr(dict(target="lvdm.models.autoencoder.AutoencoderKL",
                          params=dict(embed_dim=4, ddconfig=dd, lossconfig=dict(target="torch.nn.Identity"))))
    ident = to_attr(dict(target="torch.nn.Identity"))
    model = LatentVisualDiffusion(
        img_cond_stage_config=ident, image_proj_stage_config=ident,
        first_stage_config=fs_cfg, cond_stage_config=ident, unet_config=unet_cfg,
        rescale_betas_zero_snr=True, parameterization="v", linear_start=0.00085, linear_end=0.012,
        num_timesteps_cond=1, timesteps=1000, first_stage_key="video", cond_stage_key="caption",
        cond_stage_trainable=False, conditioning_key="hybrid", image_size=[H, W], channels=4,
        scale_by_std=False, scale_factor=0.18215, use_ema=False, uncond_type="empty_seq",
        use_dynamic_rescale=True, base_scale=0.3, fps_condition_type="fps", perframe_ae=True)
    model.model.diffusion_model.load_state_dict(sd, strict=True)
    model.first_stage_model.load_state_dict(vsd, strict=True)
    model.eval()
    tab = O.make_tables(base_scale=0.3)
    for name in ("alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "scale_arr"):
        assert torch.equal(getattr(model, name), getattr(tab, name)), name
    full_keys = sorted(k for k in model.state_dict())
    meta["lvd_buffers"] = [k for k in full_keys if "." not in k]
    np.savez_compressed(os.path.join(OUT, "tables.npz"), alphas_cumprod=model.alphas_cumprod.numpy(),
                        scale_arr=model.scale_arr.numpy(), betas=model.betas.numpy())

    # ---- DDIM sample: 3 steps, CFG 7.5, rescale 0.7, eta 1 ----
    S = 3
    c_concat = 0.5 * torch.randn(B, 8, T, H, W, generator=g)
    cond = {"c_crossattn": [ctx], "c_concat": [c_concat]}
    uc_ctx = torch.randn(B, 77 + 16 * T, small.context_dim, generator=g)
    uc = {"c_crossattn": [uc_ctx], "c_concat": [c_concat]}
    label2 = lab[:, None]
    torch.manual_seed(123)
    sampler = DDIMSampler(model)
    samples, inter = sampler.sample(S=S, conditioning=cond, batch_size=B, shape=[4, T, H, W], verbose=False,
                                    unconditional_guidance_scale=7.5, unconditional_conditioning=uc, eta=1.0,
                                    cfg_img=None, mask=None, x0=None, fs=fs, timestep_spacing="uniform_trailing",
                                    guidance_rescale=0.7, sparse_x=None, class_label=label2,
                                    unconditional_conditioning_img_nonetext=None)
    torch.manual_seed(123)
    mine_samples = O.ddim_sample(sd, small, tab, S=S, shape=(B, 4, T, H, W), c_concat=c_concat, context=ctx,
                                 uc_context=uc_ctx, class_label=lab, fs=fs, cfg_scale=7.5, guidance_rescale=0.7,
                                 eta=1.0)
    print("ddim 3 steps: ref vs oracle max|d| =", float((samples - mine_samples).abs().max()),
          "ref absmax", float(samples.abs().max()))
    sch = O.make_ddim_schedule(tab, 50, "uniform_trailing", 1.0)
    sampler.make_schedule(50, "uniform_trailing", 1.0, verbose=False)
    assert np.array_equal(sampler.ddim_timesteps, sch.timesteps)
    assert np.allclose(np.asarray(sampler.ddim_sigmas, dtype=np.float64), sch.sigmas, rtol=0, atol=0), "sigmas"
    assert np.array_equal(np.asarray(sampler.ddim_alphas_prev), sch.alphas_prev)
    frames = model.decode_first_stage(samples)
    frames_mine = O.decode_first_stage(vsd, vsmall, samples)
    print("decode_first_stage: ref vs oracle max|d| =", float((frames - frames_mine).abs().max()))
    np.savez_compressed(os.path.join(OUT, "ddim_small.npz"), c_concat=c_concat.numpy(), uc_ctx=uc_ctx.numpy(),
                        samples=samples.numpy(), frames=frames.numpy().astype(np.float16),
                        sigmas50=np.asarray(sampler.ddim_sigmas, dtype=np.float64),
                        alphas_prev50=np.asarray(sampler.ddim_alphas_prev, dtype=np.float64),
                        timesteps50=np.asarray(sampler.ddim_timesteps))
    meta["ddim_small"] = dict(S=S, seed=123, cfg=7.5, rescale=0.7, eta=1.0)

    # ---- the mask / x0 branch of ddim_sampling (ddim.py:173-180): known latent kept where mask == 1, noised to the
    # step's level by q_sample (one extra draw per step) or taken clean (clean_cond=True) ----
    gm = torch.Generator().manual_seed(77)
    mask = (torch.rand(B, 1, T, H, W, generator=gm) > 0.5).float()
    x0 = torch.randn(B, 4, T, H, W, generator=gm)
    masked = {}
    for clean in (False, True):
        torch.manual_seed(321)
        masked[clean], _ = sampler.sample(S=S, conditioning=cond, batch_size=B, shape=[4, T, H, W], verbose=False,
                                          unconditional_guidance_scale=7.5, unconditional_conditioning=uc, eta=1.0,
                                          cfg_img=None, mask=mask, x0=x0, fs=fs, timestep_spacing="uniform_trailing",
                                          guidance_rescale=0.7, sparse_x=None, class_label=label2,
                                          unconditional_conditioning_img_nonetext=None, clean_cond=clean)
    print("ddim mask branch: |noised - clean| max =", float((masked[False] - masked[True]).abs().max()),
          " |masked - unmasked| max =", float((masked[False] - samples).abs().max()))
    np.savez_compressed(os.path.join(OUT, "ddim_mask_small.npz"), mask=mask.numpy(), x0=x0.numpy(),
                        samples=masked[False].numpy(), samples_clean=masked[True].numpy())
    meta["ddim_mask_small"] = dict(S=S, seed=321, mask_seed=77)
    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
