"""Debug probe: is the small-config UNet forward / VAE decode / DDIM step bitwise repeatable?"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mudg_oracle as O                          # noqa: E402
from mudg_b200.engine import Engine, MUDG_UNET, MUDG_VAE     # noqa: E402

SMALL_UNET = dict(in_channels=12, out_channels=4, model_channels=64, num_res_blocks=2, channel_mult=(1, 2, 4, 4),
                  attention_resolutions=(4, 2, 1), num_head_channels=64, context_dim=1024)
SMALL_VAE = dict(ch=64, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4, out_ch=3, embed_dim=4)


def main():
    eng = Engine(SMALL_UNET, SMALL_VAE)
    eng.load_state_dict(O.seeded_state_dict(O.unet_param_shapes(O.UNetCfg(model_channels=64, temporal_length=4)), seed=1), MUDG_UNET)
    eng.load_state_dict(O.seeded_state_dict(O.vae_param_shapes(O.VaeCfg(ch=64)), seed=2), MUDG_VAE)
    g = np.load(os.path.join(ROOT, "tests", "golden", "unet_small.npz"))
    t = lambda k: torch.from_numpy(g[k]).cuda()
    for shape_name, x in (("golden 16x16", t("x")), ("8x16", t("x")[:, :, :, :8, :].contiguous())):
        eng.set_context(t("ctx"), T=4)
        ys = [eng.unet_forward(x, t("ts"), t("lab"), t("fs")).clone() for _ in range(4)]
        torch.cuda.synchronize()
        print(shape_name, "unet repeat max|d| (eager/capture/replay/replay):", [float((y.float() - ys[0].float()).abs().max()) for y in ys[1:]])
        os.environ["X"] = "1"
    z = torch.randn(2, 4, 8, 16, device="cuda")
    d = [eng.vae_decode(z).clone() for _ in range(3)]
    print("vae decode repeat max|d|:", [float((a.float() - d[0].float()).abs().max()) for a in d[1:]])
    xi = torch.rand(2, 3, 64, 128, device="cuda") * 2 - 1
    m = [eng.vae_encode_moments(xi).clone() for _ in range(3)]
    print("vae encode repeat max|d|:", [float((a - m[0]).abs().max()) for a in m[1:]])
    x = torch.randn(2, 4, 4, 8, 16, device="cuda"); vc = torch.randn_like(x).half(); vu = torch.randn_like(x).half(); nz = torch.randn_like(x)
    s = [eng.ddim_step(x, vc, vu, nz, cfg_scale=7.5, guidance_rescale=0.7, sqrt_ac=0.5, sqrt_1mac=0.8, rescale=1.0, a_prev=0.5, sigma=0.1)[0].clone() for _ in range(3)]
    print("ddim step repeat max|d|:", [float((a - s[0]).abs().max()) for a in s[1:]])


if __name__ == "__main__":
    main()
