"""CPU tests of the host side: drop-in class surface, state-dict layout, schedules, C-ABI symbol table,
loud failure without a GPU.  No compute call is made into the CUDA library here."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from mudg_b200 import compat

compat.install()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def small_model_config():
    from omegaconf import OmegaConf
    cfg = OmegaConf.load(os.path.join(ROOT, "configs", "stage2-1024_mdm_waymo_infer_synthetic.yaml"))
    p = cfg.model.params
    p.unet_config.params.model_channels = 64
    p.unet_config.params.temporal_length = 4
    p.first_stage_config.params.ddconfig.ch = 64
    p.image_size = [16, 16]
    p.image_proj_stage_config.params.video_length = 4
    return cfg.model


def test_yaml_configs_cover_reference_keys():
    from omegaconf import OmegaConf
    for name, size, base in (("stage1-512_mdm_waymo_infer", [40, 64], 0.7), ("stage2-1024_mdm_waymo_infer", [72, 128], 0.3)):
        cfg = OmegaConf.load(os.path.join(ROOT, "configs", name + ".yaml")).model
        assert cfg.target == "lvdm.models.ddpm3d.LatentVisualDiffusion"
        assert cfg.params.image_size == size and cfg.params.base_scale == base
        assert cfg.params.unet_config.params.temporal_length == 16
        assert cfg.params.cond_stage_config.target.endswith("FrozenOpenCLIPEmbedder")


def test_instantiate_from_config_builds_reference_layout(golden_dir):
    from utils.utils import instantiate_from_config
    from oracle import mudg_oracle as O
    model = instantiate_from_config(small_model_config())
    assert type(model).__name__ == "LatentVisualDiffusion"
    sd = model.state_dict()
    want = O.unet_param_shapes(O.UNetCfg(model_channels=64, temporal_length=4))
    got = {k[len("model.diffusion_model."):]: tuple(v.shape) for k, v in sd.items() if k.startswith("model.diffusion_model.")}
    assert got == want
    vae = {k[len("first_stage_model."):]: tuple(v.shape) for k, v in sd.items() if k.startswith("first_stage_model.")}
    assert vae == O.vae_param_shapes(O.VaeCfg(ch=64))
    with open(os.path.join(golden_dir, "meta.json")) as f:
        meta = json.load(f)
    buffers = sorted(k for k in sd if "." not in k)
    assert buffers == sorted(meta["lvd_buffers"])          # same persistent buffers as the reference module
    # attributes the driver / sampler read
    assert model.model.conditioning_key == "hybrid" and model.model.diffusion_model.out_channels == 4
    assert model.uncond_type == "empty_seq" and model.parameterization == "v" and model.use_dynamic_rescale
    assert model.num_timesteps == 1000 and model.perframe_ae is True
    tab = np.load(os.path.join(golden_dir, "tables.npz"))
    assert np.array_equal(model.alphas_cumprod.numpy(), tab["alphas_cumprod"])
    assert np.array_equal(model.scale_arr.numpy(), tab["scale_arr"])
    assert np.array_equal(model.betas.numpy(), tab["betas"])
    # strict load of a reference-layout state dict works
    model.load_state_dict(sd, strict=True)


def test_sampler_schedule_matches_reference(golden_dir):
    from utils.utils import instantiate_from_config
    from lvdm.models.samplers.ddim import DDIMSampler
    model = instantiate_from_config(small_model_config())
    s = DDIMSampler(model)
    s.make_schedule(50, "uniform_trailing", 1.0, verbose=False)
    d = np.load(os.path.join(golden_dir, "ddim_small.npz"))
    assert np.array_equal(s.ddim_timesteps, d["timesteps50"])
    assert np.array_equal(np.asarray(s.ddim_sigmas, dtype=np.float64), d["sigmas50"])
    assert np.array_equal(np.asarray(s.ddim_alphas_prev, dtype=np.float64), d["alphas_prev50"])
    assert s.ddim_scale_arr.shape == (50,)


def test_product_path_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from utils.utils import instantiate_from_config
    from mudg_b200._lib import MudgError
    model = instantiate_from_config(small_model_config())
    x = torch.zeros(1, 12, 4, 16, 16)
    with pytest.raises(MudgError):
        model.model.diffusion_model(x, torch.zeros(1, dtype=torch.long), c_label=torch.zeros(1, dtype=torch.long),
                                    context=torch.zeros(1, 77 + 64, 1024))
    with pytest.raises(MudgError):
        model.decode_first_stage(torch.zeros(1, 4, 4, 16, 16))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under mudg_b200/, lvdm/, utils/ may reference it."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|mudg_oracle", re.M)
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|mudg_oracle|post_oracle", re.M)
    for top in ("mudg_b200", "lvdm", "utils", "virtual_render"):
        for dp, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".h", ".cuh")):
                    assert not pat.search(open(os.path.join(dp, f)).read()), os.path.join(dp, f)


def test_next_rows_fail_loudly_without_gpu():
    """Resampler (f.3) and the post-decode helpers (f.2): reference surface present, no CPU fallback behind it."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mudg_b200._lib import MudgError
    from lvdm.modules.encoders.resampler import Resampler
    from virtual_render import eval_tools as E
    m = Resampler(dim=128, depth=1, dim_head=64, heads=2, num_queries=4, embedding_dim=96, output_dim=128, video_length=4)
    assert set(m.state_dict()) >= {"latents", "proj_in.weight", "layers.0.0.to_kv.weight", "layers.0.1.3.weight", "norm_out.bias"}
    with pytest.raises(MudgError):
        m(torch.zeros(1, 9, 96))
    for name in ("save_virtual_color_results", "save_virtual_depth_results", "save_virtual_semantic_results", "colormap",
                 "visualize_depth", "visualize_semantic"):
        assert callable(getattr(E, name))
    with pytest.raises(MudgError):
        E.visualize_semantic(torch.zeros(3, 4, 8, dtype=torch.uint8))
    with pytest.raises(MudgError):
        E.convert_clip(torch.zeros(1, 3, 2, 4, 8), 0)


def test_clip_holders_have_open_clip_layout_and_fail_loudly():
    """OpenCLIP towers (f.3): the drop-in modules expose the key set of the reference's embedders (an open_clip CLIP under
    `model.` with `visual` / `transformer` deleted) at ViT-H/14 size, and have no CPU path."""
    from mudg_b200._lib import MudgError
    from lvdm.modules.encoders.condition import FrozenOpenCLIPEmbedder, FrozenOpenCLIPImageEmbedderV2
    from oracle import clip_oracle as C
    with torch.device("meta"):
        t, v = FrozenOpenCLIPEmbedder(layer="penultimate"), FrozenOpenCLIPImageEmbedderV2()
    text_side = {"model." + k: s for k, s in C.clip_text_param_shapes().items()}
    text_side["model.logit_scale"] = ()
    assert {k: tuple(x.shape) for k, x in t.state_dict().items()} == text_side and t.layer_idx == 1
    want_v = {"model.visual." + k: s for k, s in C.clip_vision_param_shapes().items()}
    want_v.update({k: s for k, s in text_side.items() if "resblocks" not in k})
    assert {k: tuple(x.shape) for k, x in v.state_dict().items()} == want_v
    assert "mean" not in v.state_dict() and tuple(v.mean.shape) == (3,)            # non-persistent buffers, as in the reference
    with pytest.raises(NotImplementedError):
        FrozenOpenCLIPImageEmbedderV2(layer="penultimate")
    if torch.cuda.is_available():
        return
    small = dict(embed_dim=64, vision=dict(width=128, layers=1, heads=2, mlp=512, image_size=28, patch=14),
                 text=dict(width=128, layers=2, heads=2, mlp=512, vocab=100, ctx=77))
    with pytest.raises(MudgError):
        FrozenOpenCLIPImageEmbedderV2(arch=small)(torch.zeros(1, 3, 32, 32))
    tm = FrozenOpenCLIPEmbedder(arch=small, layer="penultimate")
    with pytest.raises(MudgError):
        tm.encode_with_transformer(torch.zeros(1, 77, dtype=torch.long))
    with pytest.raises(MudgError):
        tm(["a prompt"])                                                            # no tokenizer in this image


def test_context_tensor_identity_is_kept():
    """DiffusionWrapper hands the UNet the SAME context tensor object every DDIM step (the K/V cache is keyed on it)."""
    from lvdm.models.ddpm3d import DiffusionWrapper
    w = DiffusionWrapper.__new__(DiffusionWrapper)
    a, b = torch.zeros(1, 77, 8), torch.ones(1, 16, 8)
    assert DiffusionWrapper._context(w, [a]) is a
    c1 = DiffusionWrapper._context(w, [a, b])
    assert DiffusionWrapper._context(w, [a, b]) is c1 and torch.equal(c1, torch.cat([a, b], 1))
    b.add_(1)                                   # in-place change -> new version -> re-concatenated
    c2 = DiffusionWrapper._context(w, [a, b])
    assert c2 is not c1 and torch.equal(c2, torch.cat([a, b], 1))


def test_c_abi_exports_every_declared_symbol():
    from mudg_b200._lib import LIB_PATH
    if not os.path.exists(LIB_PATH):
        from mudg_b200.build import build
        build()
    header = open(os.path.join(ROOT, "include", "mudg.h")).read()
    declared = set(re.findall(r"MUDG_EXPORT[^;(]*?\b(mudg_\w+)\s*\(", header))
    assert {"mudg_create", "mudg_unet_forward", "mudg_vae_decode", "mudg_ddim_step", "mudg_set_context",
            "mudg_load_weight", "mudg_finalize_weights", "mudg_destroy", "mudg_last_error"} <= declared
    lib = ctypes.CDLL(LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib.mudg_last_error.restype = ctypes.c_char_p
    assert lib.mudg_last_error() is not None


def test_test_hooks_live_only_in_the_test_library():
    """libmudg_sm100_test.so (include/mudg_test.h) carries the single-kernel hooks, checkers and knobs; the product
    library exports none of them and reads no tuning switch from the environment."""
    from mudg_b200._lib import LIB_PATH, TEST_LIB_PATH
    if not (os.path.exists(LIB_PATH) and os.path.exists(TEST_LIB_PATH)):
        from mudg_b200.build import build
        build()
    header = open(os.path.join(ROOT, "include", "mudg_test.h")).read()
    declared = set(re.findall(r"MUDG_EXPORT[^;(]*?\b(mudg_\w+)\s*\(", header))
    assert {"mudg_test_tapgemm", "mudg_test_flash", "mudg_test_set_knob", "mudg_test_last_gemm_path"} <= declared
    prod, test = ctypes.CDLL(LIB_PATH), ctypes.CDLL(TEST_LIB_PATH)
    for name in declared:
        assert hasattr(test, name), name
        assert not hasattr(prod, name), name
    assert hasattr(test, "mudg_unet_forward")        # the test library is a superset of the product ABI
    blob = open(LIB_PATH, "rb").read()
    for env in (b"MUDG_GEMM_PAIR", b"MUDG_FORCE_SIMT", b"MUDG_GEMM_V1", b"MUDG_GEMM_DBG", b"MUDG_FLASH_V1", b"MUDG_FLASH_POLY"):
        assert env not in blob, env


def test_unchanged_reference_driver_imports_against_the_dropin_packages():
    """B1: with this repo first and a reference checkout later on sys.path, the UNCHANGED driver module
    (virtual_render/virtual_pose_render.py) imports, and its sampler / config / post-decode symbols are this repo's,
    while modules not replaced here (data_tools) come from the reference.  Build container only (needs /root/reference)."""
    import subprocess
    import sys
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "virtual_render")):
        pytest.skip("reference checkout not present")
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from mudg_b200 import compat; compat.install()\n"
        "import inspect, virtual_render.virtual_pose_render as V, virtual_render.data_tools as D\n"
        "print(inspect.getsourcefile(V)); print(inspect.getsourcefile(V.DDIMSampler)); print(inspect.getsourcefile(V.DDIMSampler_multicond))\n"
        "print(inspect.getsourcefile(V.instantiate_from_config)); print(inspect.getsourcefile(V.save_virtual_depth_results))\n"
        "print(inspect.getsourcefile(D))\n" % (ref, ROOT))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()[-6:]
    assert lines[0].startswith(ref) and lines[5].startswith(ref)
    for ln in lines[1:5]:
        assert ln.startswith(ROOT), lines


def test_schedule_helpers_match_reference_golden(golden_dir):
    """lvdm.models.utils_diffusion (host side of a1 / a3 / a6): every beta schedule, DDIM spacing, eta, the timestep
    embedding and rescale_noise_cfg against the reference's own outputs (oracle/make_golden_schedules.py)."""
    from lvdm.models import utils_diffusion as U
    g = np.load(os.path.join(golden_dir, "schedules.npz"))
    for sched in ("linear", "cosine", "sqrt_linear", "sqrt"):
        b = np.asarray(U.make_beta_schedule(sched, 1000, linear_start=0.00085, linear_end=0.012), dtype=np.float64)
        assert np.array_equal(b, g[f"betas_{sched}"]), sched
    betas = U.make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012)
    zs = np.asarray(U.rescale_zero_terminal_snr(betas), dtype=np.float64)
    assert np.array_equal(zs, g["betas_zero_snr"])
    ac = torch.tensor(np.cumprod(1.0 - zs, axis=0), dtype=torch.float32)
    for method in ("uniform", "quad", "uniform_trailing"):
        for S in (50, 25, 7):
            ts = U.make_ddim_timesteps(method, S, 1000, verbose=False)
            assert np.array_equal(np.asarray(ts), g[f"ts_{method}_{S}"]), (method, S)
            for eta in (0.0, 1.0):
                sig, a, ap = U.make_ddim_sampling_parameters(ac.cpu(), ts, eta, verbose=False)
                assert np.array_equal(np.asarray(sig, dtype=np.float64), g[f"sig_{method}_{S}_{eta}"]), (method, S, eta)
                assert np.array_equal(np.asarray(a, dtype=np.float64), g[f"a_{method}_{S}_{eta}"])
                assert np.array_equal(np.asarray(ap, dtype=np.float64), g[f"ap_{method}_{S}_{eta}"])
    t = torch.tensor([0, 1, 19, 500, 999], dtype=torch.long)
    for dim in (320, 64, 7):
        assert np.array_equal(U.timestep_embedding(t, dim).numpy(), g[f"temb_{dim}"]), dim
    assert np.array_equal(U.timestep_embedding(t, 8, repeat_only=True).numpy(), g["temb_repeat"])
    out = U.rescale_noise_cfg(torch.from_numpy(g["rescale_in_cfg"]), torch.from_numpy(g["rescale_in_txt"]), guidance_rescale=0.7)
    assert np.array_equal(out.numpy(), g["rescale_out"])


def test_host_tensor_helpers_match_reference_golden(golden_dir):
    """q_sample / predict_start_from_z_and_v / predict_eps_from_z_and_v of the drop-in model and the VAE posterior
    (CPU-generator sampling, logvar clamp) against the reference's outputs (oracle/make_golden_host.py), bit for bit."""
    from utils.utils import instantiate_from_config
    from lvdm.distributions import DiagonalGaussianDistribution
    g = np.load(os.path.join(golden_dir, "host_small.npz"))
    model = instantiate_from_config(small_model_config())
    t = lambda k: torch.from_numpy(g[k])
    assert np.array_equal(model.q_sample(t("x0"), t("t"), t("noise")).numpy(), g["q_sample"])
    assert np.array_equal(model.predict_start_from_z_and_v(t("x0"), t("t"), t("v")).numpy(), g["pred_start"])
    assert np.array_equal(model.predict_eps_from_z_and_v(t("x0"), t("t"), t("v")).numpy(), g["pred_eps"])
    d = DiagonalGaussianDistribution(t("moments"))
    torch.manual_seed(5)
    assert np.array_equal(d.sample().numpy(), g["sample"]) and np.array_equal(d.mode().numpy(), g["mode"])
