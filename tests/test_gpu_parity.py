"""GPU parity tests (pytest -m gpu): every CUDA kernel and the full hot path, called through the C-ABI, against the
oracle / the golden vectors produced by the unchanged reference.  Tolerances are fp16 tolerances: operands and
outputs are fp16 (as in the reference's autocast path), accumulation is fp32."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
pytestmark = pytest.mark.gpu

SMALL_UNET = dict(in_channels=12, out_channels=4, model_channels=64, num_res_blocks=2, channel_mult=(1, 2, 4, 4),
                  attention_resolutions=(4, 2, 1), num_head_channels=64, context_dim=1024)
SMALL_VAE = dict(ch=64, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4, out_ch=3, embed_dim=4)
# Stated fp16 tolerance at BASELINE full size: 2 x the values measured on B200 in round 2
# (profiles/r2_parity_full.json), and never looser than the reference's own autocast-fp16 path against fp32.
# Measured: UNet forward max |d| 0.0060-0.0080 (output range +-3), mean |d| 0.0010-0.0011, PSNR 65.6-68.7 dB -- the unchanged
# reference under autocast-fp16 on the same B200 is at 0.0073-0.0082 / 0.00125 / 64.5-65.7 dB; 50-step CFG clip: latent PSNR
# 62.5 dB (reference fp16: 61.3), decoded frames max |d| 0.018 on [-1,1] (reference fp16: 0.019).
FULL_MAX_ABS, FULL_MEAN_ABS, FULL_PSNR_DB = 0.016, 0.0022, 60.0
SAMPLER_PSNR_DB, SAMPLER_FRAMES_MAX_ABS = 56.0, 0.036


@pytest.fixture(scope="module")
def engine():
    from mudg_b200.engine import Engine, MUDG_UNET, MUDG_VAE
    from oracle import mudg_oracle as O
    eng = Engine(SMALL_UNET, SMALL_VAE)
    eng.load_state_dict(O.seeded_state_dict(O.unet_param_shapes(O.UNetCfg(model_channels=64, temporal_length=4)), seed=1), MUDG_UNET)
    eng.load_state_dict(O.seeded_state_dict(O.vae_param_shapes(O.VaeCfg(ch=64)), seed=2), MUDG_VAE)
    return eng


def _gemm_cases():
    import gpu_probe_gemm as P
    return list(P.CASES)


@pytest.mark.parametrize("case", _gemm_cases())
def test_tapgemm(case):
    """Every tap-GEMM variant (single-CTA tc2, CTA-pair tc3 with each compile-time epilogue, GEGLU, folded LayerNorm)
    against fp32 torch, at the BASELINE level-0 shapes for the pair kernel; the launch must really take the expected
    kernel (mudg_test_last_gemm_path), so a dispatch change cannot silently leave the pair path untested."""
    import gpu_probe_gemm as P
    out = P.run_case(case)
    path, want = out.pop("_path")
    gn = out.pop("_gn", None)
    rm = out.pop("_resmma", None)
    lo = out.pop("_ln_out", None)
    if lo is not None:          # per-chunk LayerNorm partial sums of the output rows, stored by the pair kernel's epilogue
        taken, rel, nan_p, dfin = lo
        assert taken and nan_p == 0 and rel < 2e-3 and dfin < 1e-3, (case, "LayerNorm partials of the output", lo)
    if rm is not None:          # residual added by the tensor core (identity k-steps) or by the epilogue, as the dispatch rule says
        assert rm[0] == rm[1], (case, "residual through the MMA", rm)
    if gn is not None:          # GroupNorm statistics of the output, accumulated by the pair kernel's epilogue
        taken, rel = gn
        assert taken == ((path & 255) == 4), (case, taken, path)      # the pair kernel fuses them (knob gn_fuse = 2: any tap count)
        if taken:
            assert rel < 1e-5, (case, "fused GroupNorm statistics", rel)
    if want is not None:
        assert (path & 255) == (want & 255), (case, "kernel", path & 255, "expected", want & 255)
        if want >> 8:
            assert (path >> 8) == (want >> 8), (case, "epilogue variant", (path >> 8) - 1, "expected", (want >> 8) - 1)
    for label, (err, nans) in out.items():
        assert nans == 0 and err < 0.02, (case, label, err, nans)


def test_silu_tanh_accuracy():
    """The GroupNorm apply's SiLU goes through MUFU.TANH (h + h tanh h): PTX bounds tanh.approx at 2^-11 relative only, so
    the accuracy is pinned by measurement -- on a level-0 sized tensor the fp16 output must be as close to the float64 SiLU
    as the ex2 + rcp form is (both dominated by the final fp16 rounding), including the inputs below -2 where 1 + tanh cancels."""
    import gpu_probe_silu as P
    out = P.main()
    exact, fast = out[1], out[3]
    assert fast["nan"] == 0 and exact["nan"] == 0
    assert fast["mean_abs"] < 1.01 * exact["mean_abs"] and fast["mean_abs"] < 1.01 * fast["round_mean"], out
    assert fast["max_abs"] <= 1.05 * exact["max_abs"] and fast["neg_max_abs"] < 1.5e-4, out


@pytest.mark.parametrize("case", ["flash_self_small", "flash_self_l1", "flash_self_ragged", "flash_cross", "flash_cross_else",
                                  "flash_self_split_ragged", "flash_self_split_tail", "flash_self_small_s2", "flash_self_ragged_s2",
                                  "flash_self_l1_s1", "flash_self_l0",
                                  "xattn_small", "xattn_ragged", "xattn_l0", "xattn_l3", "xattn_one_tile",
                                  "tattn16", "tattn4", "tattn64", "tattn32", "tattn64g", "tattn48", "gns_frame_c320", "gns_frame_c1280",
                                  "gns_time_c640", "gns_frame_c64", "gns_big_c1280", "gns_frame_c320_fma", "gn_frame_fma", "gn_time_fma", "gn_frame", "gn_time", "ln320", "ln1280", "ln512"])
def test_ops(case):
    import gpu_probe_ops as P
    P.RESULTS.clear()
    P.run_case(case)
    assert P.RESULTS
    for key, (err, nans) in P.RESULTS.items():
        assert nans == 0 and err < 0.01, (key, err, nans)


def test_unet_forward_vs_reference_golden(engine, golden_dir):
    g = np.load(os.path.join(golden_dir, "unet_small.npz"))
    t = lambda k: torch.from_numpy(g[k]).cuda()
    for ck, yk in (("ctx", "y"), ("ctx2", "y2")):      # per-frame image tokens, and the else-branch context
        engine.set_context(t(ck), T=4)
        y = engine.unet_forward(t("x"), t("ts"), t("lab"), t("fs"))
        torch.cuda.synchronize()
        err = (y.float().cpu() - torch.from_numpy(g[yk])).abs()
        print(f"MEASURED unet small {ck}: max {float(err.max()):.5f} mean {float(err.mean()):.6f}")
        assert float(err.max()) < 0.014 and float(err.mean()) < 0.0025, (ck, float(err.max()), float(err.mean()))     # 2x measured (0.0066 / 0.0012)


def test_unet_linearity_in_batch(engine, golden_dir):
    """Size-independent property: samples of a batch are independent (N=2 result == two N=1 results, bitwise)."""
    g = np.load(os.path.join(golden_dir, "unet_small.npz"))
    t = lambda k: torch.from_numpy(g[k]).cuda()
    engine.set_context(t("ctx"), T=4)
    both = engine.unet_forward(t("x"), t("ts"), t("lab"), t("fs")).clone()
    for i in range(2):
        engine.set_context(t("ctx")[i:i + 1].contiguous(), T=4)
        one = engine.unet_forward(t("x")[i:i + 1].contiguous(), t("ts")[i:i + 1], t("lab")[i:i + 1], t("fs")[i:i + 1])
        torch.cuda.synchronize()
        d = float((one[0].float() - both[i].float()).abs().max())
        assert d < 1.5e-2, d      # not bitwise: GroupNorm sums use atomics and tile shapes differ; fp16 re-rounding amplifies


def test_vae_decode_vs_reference_golden(engine, golden_dir):
    g = np.load(os.path.join(golden_dir, "vae_small.npz"))
    dec = engine.vae_decode(torch.from_numpy(g["z"]).cuda())
    torch.cuda.synchronize()
    err = (dec.float().cpu() - torch.from_numpy(g["dec"])).abs()
    print(f"MEASURED vae decode small: max {float(err.max()):.5f}")
    assert float(err.max()) < 0.014, float(err.max())          # 2x measured (0.0066)


def test_vae_encode_vs_reference_golden(engine, golden_dir):
    g = np.load(os.path.join(golden_dir, "vae_enc_small.npz"))
    mom = engine.vae_encode_moments(torch.from_numpy(g["x"]).cuda())
    torch.cuda.synchronize()
    err = (mom.cpu() - torch.from_numpy(g["moments"])).abs()
    print(f"MEASURED vae encode small: max {float(err.max()):.5f}")
    assert float(err.max()) < 0.008, float(err.max())          # 2x measured (0.0039)


def test_ddim_step_vs_oracle(engine):
    from oracle import mudg_oracle as O
    tab = O.make_tables(base_scale=0.3)
    sch = O.make_ddim_schedule(tab, 50, "uniform_trailing", 1.0)
    gen = torch.Generator().manual_seed(3)
    x, vc, vu, nz = (torch.randn(2, 4, 4, 16, 16, generator=gen) for _ in range(4))
    vc, vu = vc.half().float(), vu.half().float()
    for index in (49, 20, 0):
        ref_prev, ref_x0 = O.ddim_step(tab, sch, index, x, vc, vu, nz, 7.5, 0.7)
        tt = int(sch.timesteps[index])
        xp, x0 = engine.ddim_step(x.cuda(), vc.cuda(), vu.cuda(), nz.cuda(), cfg_scale=7.5, guidance_rescale=0.7,
                                  sqrt_ac=float(tab.sqrt_alphas_cumprod[tt]), sqrt_1mac=float(tab.sqrt_one_minus_alphas_cumprod[tt]),
                                  rescale=float(sch.scale_arr_prev[index] / sch.scale_arr[index]),
                                  a_prev=float(sch.alphas_prev[index]), sigma=float(sch.sigmas[index]))
        torch.cuda.synchronize()
        assert float((xp.cpu() - ref_prev).abs().max()) < 1e-4
        assert float((x0.cpu() - ref_x0).abs().max()) < 1e-4


def test_sampler_public_api_vs_reference_golden(golden_dir):
    """DDIMSampler.sample + decode_first_stage through the drop-in classes (3 steps, CFG 7.5, rescale 0.7, eta 1) against
    the samples the reference produced with the same seed.  CUDA and CPU generators differ, so the reference's noise
    draws are replayed by seeding the CPU generator and moving the draws to the GPU."""
    from mudg_b200 import compat
    compat.install()
    from omegaconf import OmegaConf
    from utils.utils import instantiate_from_config
    from lvdm.models.samplers.ddim import DDIMSampler
    import lvdm.models.samplers.ddim as ddim_mod
    from oracle import mudg_oracle as O
    cfg = OmegaConf.load(os.path.join(ROOT, "configs", "stage2-1024_mdm_waymo_infer_synthetic.yaml")).model
    p = cfg.params
    p.unet_config.params.model_channels = 64
    p.unet_config.params.temporal_length = 4
    p.first_stage_config.params.ddconfig.ch = 64
    p.image_size = [16, 16]
    model = instantiate_from_config(cfg)
    sd = O.seeded_state_dict(O.unet_param_shapes(O.UNetCfg(model_channels=64, temporal_length=4)), seed=1)
    vsd = O.seeded_state_dict(O.vae_param_shapes(O.VaeCfg(ch=64)), seed=2)
    model.model.diffusion_model.load_state_dict(sd, strict=True)
    model.first_stage_model.load_state_dict(vsd, strict=True)
    model = model.cuda().eval()
    g = np.load(os.path.join(golden_dir, "unet_small.npz"))
    d = np.load(os.path.join(golden_dir, "ddim_small.npz"))
    t = lambda a: torch.from_numpy(a).cuda()
    cond = {"c_crossattn": [t(g["ctx"])], "c_concat": [t(d["c_concat"])]}
    uc = {"c_crossattn": [t(d["uc_ctx"])], "c_concat": [t(d["c_concat"])]}
    # replay the reference's CPU-generator draws (x_T, then one per step)
    torch.manual_seed(123)
    draws = [torch.randn(2, 4, 4, 16, 16) for _ in range(4)]
    it = iter(draws[1:])
    orig = ddim_mod.noise_like
    ddim_mod.noise_like = lambda shape, device, repeat=False: next(it).to(device)
    try:
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            z, inter = DDIMSampler(model).sample(
                S=3, conditioning=cond, batch_size=2, shape=[4, 4, 16, 16], verbose=False,
                unconditional_guidance_scale=7.5, unconditional_conditioning=uc, eta=1.0, cfg_img=None, mask=None, x0=None,
                fs=t(g["fs"]), timestep_spacing="uniform_trailing", guidance_rescale=0.7, sparse_x=None,
                class_label=t(g["lab"])[:, None], unconditional_conditioning_img_nonetext=None, x_T=draws[0].cuda())
            frames = model.decode_first_stage(z)
    finally:
        ddim_mod.noise_like = orig
    torch.cuda.synchronize()
    ref = torch.from_numpy(d["samples"])
    err = (z.float().cpu() - ref).abs()
    psnr = 10 * torch.log10(ref.abs().max() ** 2 / ((z.float().cpu() - ref) ** 2).mean())
    print(f"3-step CFG sample: max|d|={float(err.max()):.4f} latent PSNR={float(psnr):.1f} dB")
    assert float(err.max()) < 0.11 and float(psnr) > 46.0           # measured 0.054 / 52.4 dB (small config, 3 steps)
    assert set(inter) == {"x_inter", "pred_x0"}
    ferr = (frames.float().cpu() - torch.from_numpy(d["frames"]).float()).abs()
    print(f"MEASURED decoded frames of the 3-step sample: max {float(ferr.max()):.4f} mean {float(ferr.mean()):.5f}")
    assert frames.shape == (2, 3, 4, 128, 128) and float(ferr.max()) < 0.09 and float(ferr.mean()) < 0.01, float(ferr.max())   # 2x measured (0.041 / 0.0045)
    # the mask / x0 branch (ddim.py:173-180): q_sample-noised (one extra draw per step, before the step's own) and clean_cond
    m = np.load(os.path.join(golden_dir, "ddim_mask_small.npz"))
    for clean, key in ((False, "samples"), (True, "samples_clean")):
        torch.manual_seed(321)
        n_draws = 1 + 3 * (1 if clean else 2)
        seq = iter([torch.randn(2, 4, 4, 16, 16) for _ in range(n_draws)])
        x_T = next(seq).cuda()
        ddim_mod.noise_like = lambda shape, device, repeat=False: next(seq).to(device)
        q_orig = model.q_sample
        model.q_sample = lambda x_start, ts, noise=None: q_orig(x_start, ts, noise=next(seq).to(x_start.device))
        try:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                zm, _ = DDIMSampler(model).sample(
                    S=3, conditioning=cond, batch_size=2, shape=[4, 4, 16, 16], verbose=False,
                    unconditional_guidance_scale=7.5, unconditional_conditioning=uc, eta=1.0, cfg_img=None, mask=t(m["mask"]),
                    x0=t(m["x0"]), fs=t(g["fs"]), timestep_spacing="uniform_trailing", guidance_rescale=0.7, sparse_x=None,
                    class_label=t(g["lab"])[:, None], unconditional_conditioning_img_nonetext=None, x_T=x_T, clean_cond=clean)
        finally:
            ddim_mod.noise_like = orig
            del model.q_sample
        refm = torch.from_numpy(m[key])
        em = (zm.float().cpu() - refm).abs()
        pm = 10 * torch.log10(refm.abs().max() ** 2 / ((zm.float().cpu() - refm) ** 2).mean())
        print(f"3-step CFG sample with mask (clean_cond={clean}): max|d|={float(em.max()):.4f} latent PSNR={float(pm):.1f} dB")
        assert float(em.max()) < 0.11 and float(pm) > 46.0, (clean, float(em.max()), float(pm))     # measured 0.043-0.051 / 53.5-54.9 dB


def test_window_pipeline_three_modalities():
    """The driver's per-window flow (embedders -> VAE encode -> CFG DDIM -> VAE decode) with the three modalities
    stacked on the batch dim (colour 0, depth 500, semantic 1), as virtual_pose_render.py:206-233 does."""
    from mudg_b200 import compat
    compat.install()
    from omegaconf import OmegaConf
    from utils.utils import instantiate_from_config
    from mudg_b200.pipeline import image_guided_synthesis
    cfg = OmegaConf.load(os.path.join(ROOT, "configs", "stage2-1024_mdm_waymo_infer_synthetic.yaml")).model
    p = cfg.params
    p.unet_config.params.model_channels = 64
    p.unet_config.params.temporal_length = 4
    p.first_stage_config.params.ddconfig.ch = 64
    p.image_proj_stage_config.params.video_length = 4
    p.image_size = [8, 16]
    torch.manual_seed(0)
    model = instantiate_from_config(cfg)
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for name, q in model.named_parameters():        # undo the zero-inits so the output is non-trivial
            if q.dim() > 1 and float(q.abs().sum()) == 0.0:
                q.copy_(torch.randn(q.shape, generator=g) / q[0].numel() ** 0.5)
    model = model.cuda().eval()
    T, H, W = 4, 64, 128
    sparse_x = (torch.rand(3, 3, T, H, W, generator=g) * 2 - 1).cuda()
    sparse_d = (torch.rand(3, 3, T, H, W, generator=g) * 2 - 1).cuda()
    labels = torch.tensor([[0], [500], [1]], dtype=torch.long).cuda()
    torch.manual_seed(123)
    with torch.autocast("cuda", dtype=torch.float16):
        out = image_guided_synthesis(model, ["a street"] * 3, sparse_x, sparse_d, labels, [1, 4, T, H // 8, W // 8],
                                     ddim_steps=4, ddim_eta=1.0, unconditional_guidance_scale=7.5, fs=10, text_input=True,
                                     timestep_spacing="uniform_trailing", guidance_rescale=0.7)
    torch.cuda.synchronize()
    assert out.shape == (3, 1, 3, T, H, W)
    assert bool(torch.isfinite(out.float()).all()) and float(out.float().abs().max()) > 1e-3
    # modalities differ (different class labels / contexts)
    assert float((out[0].float() - out[1].float()).abs().max()) > 1e-3


# ---------------------------------------------------------------- post-decode frame pipeline (section 8f row 2): bit-exact
def _post(frames, modes):
    from mudg_b200.engine import postdecode
    rgb, depth, cls = postdecode(frames, modes)
    torch.cuda.synchronize()
    return rgb.cpu().numpy(), None if depth is None else depth.cpu().numpy(), None if cls is None else cls.cpu().numpy()


def test_postdecode_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "post_small.npz"))
    for dt in (torch.float16, torch.float32):
        rgb, depth, cls = _post(torch.from_numpy(g["frames"]).cuda().to(dt), [0, 1, 2])
        assert np.array_equal(rgb, np.stack([g["u8"][0], g["depth_vis"], g["sem_vis"]]))
        assert depth[1].tobytes() == g["depth_pred"].tobytes()
        assert np.array_equal(cls[2].astype(np.int64), g["sem_cls"])
        assert not depth[0].any() and not depth[2].any() and not cls[0].any() and not cls[1].any()


@pytest.mark.parametrize("shape", [(1, 3, 1, 1, 1), (2, 3, 2, 5, 7), (3, 3, 2, 16, 24), (1, 3, 1, 3, 8)])
def test_postdecode_ragged_vs_oracle(shape):
    """Odd / tiny frame sizes (scalar path when H*W is not a multiple of 8) and every mode for every sample."""
    from oracle import post_oracle as P
    g = torch.Generator().manual_seed(sum(shape))
    frames = (0.9 * torch.randn(shape, generator=g)).half()
    for modes in ([0] * shape[0], [1] * shape[0], [2] * shape[0]):
        ref_rgb, ref_d, ref_c = P.postdecode(frames.numpy(), modes)
        rgb, depth, cls = _post(frames.cuda(), modes)
        assert np.array_equal(rgb, ref_rgb)
        if modes[0] == 1:
            assert depth.tobytes() == ref_d.tobytes()
        if modes[0] == 2:
            assert np.array_equal(cls, ref_c)


def test_postdecode_full_size_properties():
    """BASELINE full size (3 modalities x 16 frames x 576 x 1024): two frames against the oracle bit for bit, plus
    size-independent properties on the whole clip: the colour output fed back as uint8 is a fixed point, the semantic
    visualisation is idempotent, class indices and palette colours agree, depth in [0,1] is the RGB mean."""
    from oracle import post_oracle as P
    g = torch.Generator(device="cuda").manual_seed(7)
    frames = (0.7 * torch.randn(3, 3, 16, 576, 1024, generator=g, device="cuda")).half()
    from mudg_b200.engine import postdecode
    rgb, depth, cls = postdecode(frames, [0, 1, 2])
    torch.cuda.synchronize()
    for t in (0, 15):
        sub = frames[:, :, t:t + 1].cpu().numpy()
        r_rgb, r_d, r_c = P.postdecode(sub, [0, 1, 2])
        assert np.array_equal(rgb[:, t].cpu().numpy(), r_rgb[:, 0])
        assert depth[1, t].cpu().numpy().tobytes() == r_d[1, 0].tobytes()
        assert np.array_equal(cls[2, t].cpu().numpy(), r_c[2, 0])
    u8 = rgb.permute(0, 2, 1, 3, 4).contiguous()                      # [B, 3, T, H, W] uint8
    rgb2, depth2, cls2 = postdecode(u8, [0, 0, 2])
    assert torch.equal(rgb2[0], rgb[0])                               # colour: uint8 in == uint8 out
    assert torch.equal(rgb2[2], rgb[2]) and torch.equal(cls2[2], cls[2])   # semantic: idempotent
    pal = torch.tensor(P.PALETTE, dtype=torch.uint8, device="cuda")
    assert torch.equal(pal[cls[2].long()].permute(0, 3, 1, 2), rgb[2])
    # depth == mean of the colour conversion / 255, whole clip (host fp32 division: torch's CUDA `/ scalar` multiplies by
    # the reciprocal and is not the reference's arithmetic)
    col_u8 = postdecode(frames[1:2], [0])[0][0].cpu().numpy().astype(np.float32)
    want = (col_u8.sum(1) / np.float32(3.0)) / np.float32(255.0)
    assert depth[1].cpu().numpy().tobytes() == want.tobytes()
    assert float(depth[1].min()) >= 0.0 and float(depth[1].max()) <= 1.0


def test_eval_tools_dropin(tmp_path):
    """virtual_render.eval_tools keeps the reference's function surface and writes the same files."""
    from virtual_render import eval_tools as E
    from oracle import post_oracle as P
    g = torch.Generator().manual_seed(1)
    samples = (0.8 * torch.randn(1, 3, 3, 16, 24, generator=g)).half().cuda()
    gts = torch.rand(1, 3, 3, 16, 24, generator=g) * 2 - 1
    sparses = torch.rand(1, 3, 3, 16, 24, generator=g) * 2 - 1
    fakedir = str(tmp_path / "samples")
    E.save_virtual_color_results("p", samples, 0, fakedir, gts, sparses, base_index=5, dir_name="virtual_color")
    E.save_virtual_depth_results("p", samples, 0, fakedir, gts, sparses, base_index=5, is_virtual=True, dir_name="virtual_depth")
    E.save_virtual_semantic_results("p", samples, 0, fakedir, gts, sparses, base_index=5, dir_name="virtual_semantic")
    assert sorted(os.listdir(tmp_path / "virtual_color")) == sorted(
        f"color_{k}_{i}.png" for k in ("re", "gt", "sp", "all") for i in (6, 7))
    ref_rgb, ref_d, ref_c = P.postdecode(samples.cpu().numpy().repeat(3, 0), [0, 1, 2])
    d = np.load(tmp_path / "depth" / "depth_re_6.npy")
    assert d.shape == (1, 16, 24) and d.tobytes() == ref_d[1, 1].tobytes()
    s = np.load(tmp_path / "semantic" / "semantic_re_7.npy")
    assert s.dtype == np.int64 and np.array_equal(s, ref_c[2, 2].astype(np.int64))
    import torchvision
    png = torchvision.io.read_image(str(tmp_path / "virtual_color" / "color_re_6.png")).numpy()
    assert np.array_equal(png, ref_rgb[0, 1])
    vis, idx = E.visualize_semantic(torch.from_numpy(ref_rgb[0, 0]), return_pt=True)
    rv, rc = P.semantic_from_uint8(ref_rgb[0, 0])
    assert np.array_equal(vis.numpy(), rv) and np.array_equal(idx.numpy(), rc)
    dv = np.array(E.visualize_depth(ref_d[1, 0][None])[0])
    assert np.array_equal(dv.transpose(2, 0, 1), P.spectral_u8(ref_d[1, 0]))


def test_multicond_sampler_vs_reference_golden(golden_dir):
    """Next row f.4: lvdm.models.samplers.ddim_multiplecond.DDIMSampler (separate image / text guidance, three UNet
    evaluations per step) through the drop-in classes against the reference's own sampler output (2 steps, text scale 7.5,
    image scale 3.0, guidance_rescale 0.7, eta 1; oracle/make_golden_multicond.py)."""
    from mudg_b200 import compat
    compat.install()
    from omegaconf import OmegaConf
    from utils.utils import instantiate_from_config
    from lvdm.models.samplers.ddim_multiplecond import DDIMSampler
    import lvdm.models.samplers.ddim_multiplecond as mc_mod
    from oracle import mudg_oracle as O
    cfg = OmegaConf.load(os.path.join(ROOT, "configs", "stage2-1024_mdm_waymo_infer_synthetic.yaml")).model
    p = cfg.params
    p.unet_config.params.model_channels = 64
    p.unet_config.params.temporal_length = 4
    p.first_stage_config.params.ddconfig.ch = 64
    p.image_size = [16, 16]
    model = instantiate_from_config(cfg)
    sd = O.seeded_state_dict(O.unet_param_shapes(O.UNetCfg(model_channels=64, temporal_length=4)), seed=1)
    model.model.diffusion_model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    d = np.load(os.path.join(golden_dir, "ddim_multicond_small.npz"))
    t = lambda a: torch.from_numpy(a).cuda()
    ctx, uc_ctx = t(d["ctx"]), t(d["uc_ctx"])
    uc_img_ctx = torch.cat([uc_ctx[:, :77], ctx[:, 77:]], dim=1)
    cc = t(d["c_concat"])
    cond = {"c_crossattn": [ctx], "c_concat": [cc]}
    uc = {"c_crossattn": [uc_ctx], "c_concat": [cc]}
    uc2 = {"c_crossattn": [uc_img_ctx], "c_concat": [cc]}
    torch.manual_seed(321)
    draws = [torch.randn(2, 4, 4, 16, 16) for _ in range(3)]
    it = iter(draws[1:])
    orig = mc_mod.noise_like
    mc_mod.noise_like = lambda shape, device, repeat=False: next(it).to(device)
    try:
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            z, _ = DDIMSampler(model).sample(
                S=2, conditioning=cond, batch_size=2, shape=[4, 4, 16, 16], verbose=False,
                unconditional_guidance_scale=7.5, unconditional_conditioning=uc, eta=1.0, cfg_img=3.0, mask=None, x0=None,
                fs=t(d["fs"]), timestep_spacing="uniform_trailing", guidance_rescale=0.7, sparse_x=None,
                class_label=t(d["lab"])[:, None], unconditional_conditioning_img_nonetext=uc2, x_T=draws[0].cuda())
    finally:
        mc_mod.noise_like = orig
    torch.cuda.synchronize()
    ref = torch.from_numpy(d["samples"])
    err = (z.float().cpu() - ref).abs()
    psnr = 10 * torch.log10(ref.abs().max() ** 2 / ((z.float().cpu() - ref) ** 2).mean())
    print(f"2-step multicond sample: max|d|={float(err.max()):.4f} latent PSNR={float(psnr):.1f} dB")
    assert float(err.max()) < 0.08 and float(psnr) > 46.0           # measured 0.037 / 52.6 dB


def test_resampler_vs_reference_golden(golden_dir):
    """Next row f.3: lvdm.modules.encoders.resampler.Resampler (drop-in, one C-ABI call) against the reference module's
    output on the same seeded weights; then the shipped full-size configuration against the oracle."""
    from lvdm.modules.encoders.resampler import Resampler
    from oracle import mudg_oracle as O
    g = np.load(os.path.join(golden_dir, "resampler_small.npz"))
    cfg = dict(dim=128, depth=2, dim_head=64, heads=2, num_queries=4, embedding_dim=96, output_dim=128, ff_mult=4, video_length=4)
    sd = O.seeded_state_dict(O.resampler_param_shapes(**cfg), seed=5)
    m = Resampler(**cfg)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    y = m(torch.from_numpy(g["x"]).cuda())
    torch.cuda.synchronize()
    err = (y.float().cpu() - torch.from_numpy(g["y"])).abs()
    print(f"MEASURED resampler small: max {float(err.max()):.5f} mean {float(err.mean()):.6f}")
    assert y.shape == (3, 16, 128) and float(err.max()) < 0.012 and float(err.mean()) < 0.002, (float(err.max()), float(err.mean()))
    # full size (infer yaml): [2, 257, 1280] -> [2, 256, 1024]
    full = dict(dim=1024, depth=4, dim_head=64, heads=12, num_queries=16, embedding_dim=1280, output_dim=1024, ff_mult=4,
                video_length=16)
    fsd = O.seeded_state_dict(O.resampler_param_shapes(**full), seed=6)
    mf = Resampler(**full)
    mf.load_state_dict(fsd, strict=True)
    mf = mf.cuda().eval()
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(2, 257, 1280, generator=gen)
    ref = O.resampler_forward(fsd, x, heads=12)
    out = mf(x.cuda())
    torch.cuda.synchronize()
    ferr = (out.float().cpu() - ref).abs()
    print(f"MEASURED resampler full: max {float(ferr.max()):.5f} mean {float(ferr.mean()):.6f}")
    assert out.shape == (2, 256, 1024) and float(ferr.max()) < 0.016 and float(ferr.mean()) < 0.002, (float(ferr.max()), float(ferr.mean()))
    # batch independence (size-independent property): sample 1 alone == sample 1 of the batch
    one = mf(x[1:2].cuda())
    assert float((one[0].float() - out[1].float()).abs().max()) < 2e-2


def test_unet_full_size_properties():
    """BASELINE full size (MDM1024: N=2 CFG batch, T=16, 72x128 latent, 1.44 B-parameter graph, random weights generated
    on the GPU): size-independent properties of the forward through the C-ABI -- finite output of the right shape,
    samples of a batch are independent (N=2 result == two N=1 results within fp16 re-rounding), and a replay of the
    captured CUDA graph reproduces the eager result."""
    import gpu_probe_full as PF
    from mudg_b200.engine import Engine, MUDG_UNET
    from mudg_b200.layout import unet_layout
    eng = Engine(PF.UNET, PF.VAE)
    eng.load_state_dict(PF.gpu_weights(unet_layout(**PF.UNET), 0), MUDG_UNET)
    g = torch.Generator(device="cuda").manual_seed(3)
    N, T, h, w = 2, 16, 72, 128
    x = torch.randn(N, 12, T, h, w, device="cuda", generator=g)
    ctx = torch.randn(N, 77 + 16 * T, 1024, device="cuda", generator=g)
    ts = torch.tensor([999, 499], device="cuda")
    lab = torch.tensor([0, 500], device="cuda")
    fs = torch.full((N,), 10, device="cuda", dtype=torch.long)
    eng.set_context(ctx, T)
    y_eager = eng.unet_forward(x, ts, lab, fs).clone()          # first call: eager
    y_cap = eng.unet_forward(x, ts, lab, fs).clone()            # second: capture + launch
    y_rep = eng.unet_forward(x, ts, lab, fs).clone()            # third: graph replay
    torch.cuda.synchronize()
    assert y_eager.shape == (N, 4, T, h, w) and bool(torch.isfinite(y_eager.float()).all())
    assert float(y_eager.float().abs().max()) > 0.1
    # measured 0.0 for both; the fp64 GroupNorm atomics are the only order-dependent arithmetic, far below fp16 resolution
    assert float((y_cap.float() - y_eager.float()).abs().max()) < 1e-3
    assert float((y_rep.float() - y_cap.float()).abs().max()) < 1e-3
    for i in range(N):
        eng.set_context(ctx[i:i + 1].contiguous(), T)
        one = eng.unet_forward(x[i:i + 1].contiguous(), ts[i:i + 1], lab[i:i + 1], fs[i:i + 1])
        torch.cuda.synchronize()
        d = (one[0].float() - y_eager[i].float()).abs()
        print(f"MEASURED full-size N=2 vs N=1 sample {i}: max {float(d.max()):.5f} mean {float(d.mean()):.6f}; eager/capture {float((y_cap.float() - y_eager.float()).abs().max()):.5f} replay {float((y_rep.float() - y_cap.float()).abs().max()):.5f}")
        assert float(d.max()) < 1.4e-2 and float(d.mean()) < 2.2e-3, (i, float(d.max()), float(d.mean()))     # 2x measured (0.0068 / 0.0011)
    del eng
    torch.cuda.empty_cache()


def test_unchanged_reference_driver_function(golden_dir):
    """B1 end to end: the reference driver's OWN `image_guided_synthesis` (virtual_render/virtual_pose_render.py:62-147,
    file copied verbatim into the git-ignored baseline/_ref/ so it travels to the GPU box) drives this repo's model and
    sampler, and gives the same clip as this repo's restatement of that flow (mudg_b200/pipeline.py) under the same seed."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(ref_root, "virtual_render", "virtual_pose_render.py")):
        pytest.skip("baseline/_ref/virtual_render/virtual_pose_render.py not present (copy it from the reference checkout)")
    from mudg_b200 import compat
    compat.install()
    if ref_root not in sys.path:
        sys.path.append(ref_root)                      # after the repo root: drop-in modules win
    import importlib
    import virtual_render
    virtual_render.__path__ = __import__("mudg_b200.compat.pkgpath", fromlist=["extended"]).extended(
        virtual_render.__path__, "virtual_render")
    V = importlib.import_module("virtual_render.virtual_pose_render")
    assert V.__file__.startswith(ref_root) and V.DDIMSampler.__module__ == "lvdm.models.samplers.ddim"
    from omegaconf import OmegaConf
    from utils.utils import instantiate_from_config
    from mudg_b200.pipeline import image_guided_synthesis
    cfg = OmegaConf.load(os.path.join(ROOT, "configs", "stage2-1024_mdm_waymo_infer_synthetic.yaml")).model
    p = cfg.params
    p.unet_config.params.model_channels = 64
    p.unet_config.params.temporal_length = 4
    p.first_stage_config.params.ddconfig.ch = 64
    p.image_proj_stage_config.params.video_length = 4
    p.image_size = [8, 16]
    torch.manual_seed(0)
    model = instantiate_from_config(cfg)
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for name, q in model.named_parameters():
            if q.dim() > 1 and float(q.abs().sum()) == 0.0:
                q.copy_(torch.randn(q.shape, generator=g) / q[0].numel() ** 0.5)
    model = model.cuda().eval()
    T, H, W = 4, 64, 128
    sparse_x = (torch.rand(3, 3, T, H, W, generator=g) * 2 - 1).cuda()
    sparse_d = (torch.rand(3, 3, T, H, W, generator=g) * 2 - 1).cuda()
    labels = torch.tensor([[0], [500], [1]], dtype=torch.long).cuda()
    args = (["a street"] * 3, sparse_x, sparse_d, labels, [1, 4, T, H // 8, W // 8])
    kw = dict(n_samples=1, ddim_steps=3, ddim_eta=1.0, unconditional_guidance_scale=7.5, cfg_img=None, fs=10, text_input=True,
              multiple_cond_cfg=False, timestep_spacing="uniform_trailing", guidance_rescale=0.7)
    outs = []
    for fn in (V.image_guided_synthesis, V.image_guided_synthesis, image_guided_synthesis):
        torch.manual_seed(123)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            outs.append(fn(model, *args, **kw).float().cpu())
    torch.cuda.synchronize()
    assert outs[0].shape == (3, 1, 3, T, H, W) and bool(torch.isfinite(outs[0]).all())
    assert float(outs[0].abs().max()) > 1e-3
    # Same calls, same seed -- but not bitwise: GroupNorm sums use atomics, and three CFG-7.5 steps plus the decoder amplify
    # the fp16 re-rounding of a random-weight network.  The yardstick is therefore the driver function against ITSELF.
    self_d = float((outs[0] - outs[1]).abs().mean())
    cross_d = float((outs[0] - outs[2]).abs().mean())
    print(f"driver vs driver mean|d| {self_d:.5f}; driver vs mudg_b200.pipeline mean|d| {cross_d:.5f}")
    assert cross_d <= 3.0 * self_d + 2e-3, (self_d, cross_d)


def test_bitwise_repeatability(engine, golden_dir):
    """Two identical calls give identical bits -- eager, graph capture and graph replay of the UNet forward, VAE decode /
    encode, post-decode (every reduction has a fixed order; the GroupNorm atomics accumulate in fp64)."""
    g = np.load(os.path.join(golden_dir, "unet_small.npz"))
    t = lambda k: torch.from_numpy(g[k]).cuda()
    engine.set_context(t("ctx"), T=4)
    ys = [engine.unet_forward(t("x"), t("ts"), t("lab"), t("fs")).clone() for _ in range(4)]
    z = torch.from_numpy(np.load(os.path.join(golden_dir, "vae_small.npz"))["z"]).cuda()
    ds = [engine.vae_decode(z).clone() for _ in range(2)]
    torch.cuda.synchronize()
    for y in ys[1:]:
        assert torch.equal(y, ys[0])
    assert torch.equal(ds[0], ds[1])


def test_weight_reload_replaces_the_whole_set(golden_dir):
    """loading a second state dict into a live engine (after forwards, graph capture and a cached context) must
    give exactly what a fresh engine gives -- no stale fused QKV / K|V / GEGLU / LayerNorm-fold tensors, no replay of a
    graph that points at freed weights, no K/V cache projected with the old weights."""
    from mudg_b200.engine import Engine, MUDG_UNET, MUDG_VAE
    from oracle import mudg_oracle as O
    g = np.load(os.path.join(golden_dir, "unet_small.npz"))
    t = lambda k: torch.from_numpy(g[k]).cuda()
    shapes = O.unet_param_shapes(O.UNetCfg(model_channels=64, temporal_length=4))
    sd1, sd2 = O.seeded_state_dict(shapes, seed=1), O.seeded_state_dict(shapes, seed=7)
    vshapes = O.vae_param_shapes(O.VaeCfg(ch=64))
    v1, v2 = O.seeded_state_dict(vshapes, seed=2), O.seeded_state_dict(vshapes, seed=8)
    z = torch.from_numpy(np.load(os.path.join(golden_dir, "vae_small.npz"))["z"]).cuda()

    def run(eng):
        eng.set_context(t("ctx"), T=4)
        ys = [eng.unet_forward(t("x"), t("ts"), t("lab"), t("fs")).clone() for _ in range(3)]   # eager, capture, replay
        d = eng.vae_decode(z).clone()
        torch.cuda.synchronize()
        return ys, d

    eng = Engine(SMALL_UNET, SMALL_VAE)
    eng.load_state_dict(sd1, MUDG_UNET); eng.load_state_dict(v1, MUDG_VAE)
    y_old, d_old = run(eng)
    eng.load_state_dict(sd2, MUDG_UNET); eng.load_state_dict(v2, MUDG_VAE)
    with pytest.raises(Exception):                       # the old K/V cache is gone: a forward without set_context must refuse
        eng.unet_forward(t("x"), t("ts"), t("lab"), t("fs"))
    y_new, d_new = run(eng)
    fresh = Engine(SMALL_UNET, SMALL_VAE)
    fresh.load_state_dict(sd2, MUDG_UNET); fresh.load_state_dict(v2, MUDG_VAE)
    y_ref, d_ref = run(fresh)
    assert float((y_new[0].float() - y_old[0].float()).abs().max()) > 0.05      # the weights really changed
    for a, b in zip(y_new, y_ref):
        assert torch.equal(a, b)
    assert torch.equal(d_new, d_ref) and not torch.equal(d_new, d_old)
    ref = O.unet_forward(sd2, O.UNetCfg(model_channels=64, temporal_length=4), t("x").cpu(), t("ts").cpu(), t("lab").cpu(),
                         t("ctx").cpu(), t("fs").cpu())
    assert float((y_new[2].float().cpu() - ref).abs().max()) < 0.03


def test_shared_cfg_prefix_small(engine, golden_dir):
    """mudg_unet_forward_shared (dup copies of one latent, different contexts): same result as the plain forward on the
    tiled batch, eager and replayed; dup = 2 (CFG) and dup = 3 (multi-cond guidance)."""
    g = np.load(os.path.join(golden_dir, "unet_small.npz"))
    t = lambda k: torch.from_numpy(g[k]).cuda()
    for dup in (2, 3):
        x = t("x")[:1].repeat(dup, 1, 1, 1, 1).contiguous()
        ts, lab, fs = t("ts")[:1].repeat(dup), t("lab")[:1].repeat(dup), t("fs")[:1].repeat(dup)
        gen = torch.Generator(device="cuda").manual_seed(dup)
        ctx = torch.randn(dup, t("ctx").shape[1], 1024, device="cuda", generator=gen)
        engine.set_context(ctx, T=4)
        plain = engine.unet_forward(x, ts, lab, fs).clone()
        shared = [engine.unet_forward(x, ts, lab, fs, dup=dup).clone() for _ in range(3)]
        torch.cuda.synchronize()
        assert float((plain[0].float() - plain[1].float()).abs().max()) > 1e-2          # the contexts matter
        for s_ in shared:
            assert float((s_.float() - plain.float()).abs().max()) < 1.5e-2
        assert torch.equal(shared[1], shared[2])


def test_full_size_parity_unet():
    """BASELINE full size against the reference / oracle in fp32 on the GPU, with the reference's own autocast-fp16 gap
    as the yardstick (tests/gpu_parity_full.py; numbers in profiles/r2_parity_full.json and DESIGN.md section 2)."""
    import gpu_parity_full as PFU
    res = PFU.run(["unet40", "unet72", "unet_t64"], out=os.path.join(ROOT, "gpurun_out", "r2_parity_unet_pytest.json"))
    u40, u72, t64 = res["unet40"], res["unet72"], res["unet_t64"]
    if "oracle_vs_reference_fp32" in u40:
        assert u40["oracle_vs_reference_fp32"]["max_abs"] < 2e-3, u40["oracle_vs_reference_fp32"]
    for r in (u40["ours_vs_truth"], u40["ours_vs_truth_n3_labels"], u72["ours_vs_truth"], u72["ours_shared_prefix_vs_truth"], t64["ours_vs_truth"]):
        assert r["max_abs"] < FULL_MAX_ABS and r["mean_abs"] < FULL_MEAN_ABS and r["psnr_db"] > FULL_PSNR_DB, r
    assert u72["ours_shared_prefix_vs_ours_plain"]["max_abs"] < FULL_MAX_ABS
    for r in (u40, u72):
        g16 = r.get("reference_fp16_vs_truth", {})
        if "mean_abs" in g16:        # ours must be no further from fp32 than twice the reference's own fp16 path
            assert r["ours_vs_truth"]["mean_abs"] < 1.25 * g16["mean_abs"], (r["ours_vs_truth"], g16)


def test_full_size_parity_sampler():
    """A whole 50-step CFG-7.5 MDM512 clip + decode against the fp32 oracle with replayed noise."""
    import gpu_parity_full as PFU
    res = PFU.run(["sample512"], out=os.path.join(ROOT, "gpurun_out", "r2_parity_sampler_pytest.json"))
    r = res["sample512"]
    assert r["ours_latent_vs_truth"]["psnr_db"] > SAMPLER_PSNR_DB, r["ours_latent_vs_truth"]
    assert r["ours_frames_vs_truth"]["max_abs"] < SAMPLER_FRAMES_MAX_ABS, r["ours_frames_vs_truth"]
    assert r["ours_decoder_only_vs_oracle"]["max_abs"] < 0.03, r["ours_decoder_only_vs_oracle"]
    g16 = r.get("reference_fp16_latent_vs_truth")
    if g16:
        assert r["ours_latent_vs_truth"]["mean_abs"] < 1.25 * g16["mean_abs"], (r["ours_latent_vs_truth"], g16)


def test_engine_on_second_device_matches_first(golden_dir):
    """A context created on cuda:1 while cuda:0 is the current device (every C entry point switches to the context's
    device; kernel attributes are set per device): same forward, VAE decode and Resampler results as on cuda:0, bitwise."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    from mudg_b200.engine import Engine, MUDG_UNET, MUDG_VAE
    from oracle import mudg_oracle as O
    g = np.load(os.path.join(golden_dir, "unet_small.npz"))
    v = np.load(os.path.join(golden_dir, "vae_small.npz"))
    sd = O.seeded_state_dict(O.unet_param_shapes(O.UNetCfg(model_channels=64, temporal_length=4)), seed=1)
    vsd = O.seeded_state_dict(O.vae_param_shapes(O.VaeCfg(ch=64)), seed=2)
    outs = []
    torch.cuda.set_device(0)
    for dev in (0, 1):
        eng = Engine(SMALL_UNET, SMALL_VAE, device=dev)
        assert torch.cuda.current_device() == 0
        eng.load_state_dict(sd, MUDG_UNET)
        eng.load_state_dict(vsd, MUDG_VAE)
        t = lambda k: torch.from_numpy(g[k]).to(f"cuda:{dev}")
        with torch.cuda.device(dev):                      # torch allocations and the stream handed to the library
            eng.set_context(t("ctx"), T=4)
            y = eng.unet_forward(t("x"), t("ts"), t("lab"), t("fs"))
            y2 = eng.unet_forward(t("x"), t("ts"), t("lab"), t("fs"))      # captured graph
            dec = eng.vae_decode(torch.from_numpy(v["z"]).to(f"cuda:{dev}"))
            torch.cuda.synchronize(dev)
        assert torch.cuda.current_device() == 0
        assert torch.equal(y, y2)
        outs.append((y.cpu(), dec.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    err = (outs[1][0].float() - torch.from_numpy(g["y"])).abs()
    assert float(err.max()) < 0.014
