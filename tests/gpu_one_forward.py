"""One eager pass over every kernel of the path, for ncu (profiles/): `unet` = one MDM1024 CFG forward (N = 2, shared prefix)
through the C-ABI; `unet512` = one MDM512 forward (N = 1, 40x64); `tail` = one VAE decode frame at 72x128, one fused DDIM step
and one post-decode call; `clip` = the ViT-H/14 image tower on one 576x1024 frame and the text tower on one prompt.
Run with MUDG_GRAPH=0 so the forward launches its kernels one by one instead of replaying a CUDA graph."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import gpu_probe_full as PF
    from mudg_b200.engine import Engine, MUDG_UNET, MUDG_VAE, postdecode
    from mudg_b200.layout import unet_layout, vae_layout
    mode = sys.argv[1] if len(sys.argv) > 1 else "unet"
    eng = Engine(PF.UNET, PF.VAE)
    g = torch.Generator(device="cuda").manual_seed(3)
    if mode == "clip":
        from mudg_b200.engine import MUDG_CLIP_IMAGE, MUDG_CLIP_TEXT
        from oracle import clip_oracle as C
        eng.load_state_dict(C.seeded_clip_state_dict(C.clip_vision_param_shapes(), 31), MUDG_CLIP_IMAGE)
        eng.load_state_dict(C.seeded_clip_state_dict(C.clip_text_param_shapes(), 32), MUDG_CLIP_TEXT)
        img = torch.rand(1, 3, 576, 1024, device="cuda", generator=g) * 2 - 1
        tok = torch.randint(1, 49000, (1, 77), device="cuda", generator=g)
        a = eng.clip_image_forward(img, 257, 1280, 16, resize=True)
        b = eng.clip_text_forward(tok, 1024, 16, 1)
        torch.cuda.synchronize()
        print("clip towers done", float(a.abs().max()), float(b.abs().max()))
    elif mode in ("unet", "unet512"):
        eng.load_state_dict(PF.gpu_weights(unet_layout(**PF.UNET), 0), MUDG_UNET)
        N, T, h, w = (2, 16, 72, 128) if mode == "unet" else (1, 16, 40, 64)
        x = torch.randn(1, 12, T, h, w, device="cuda", generator=g).repeat(N, 1, 1, 1, 1).contiguous()
        ctx = torch.randn(N, 77 + 16 * T, 1024, device="cuda", generator=g)
        ts = torch.full((N,), 499, device="cuda", dtype=torch.long)
        lab = torch.zeros(N, device="cuda", dtype=torch.long)
        fs = torch.full((N,), 10, device="cuda", dtype=torch.long)
        eng.set_context(ctx, T)
        y = eng.unet_forward(x, ts, lab, fs, dup=N)
        torch.cuda.synchronize()
        print("unet forward done", float(y.float().abs().max()), "launches", eng.launch_count())
    else:
        eng.load_state_dict(PF.gpu_weights(vae_layout(**PF.VAE), 1), MUDG_VAE)
        z = torch.randn(1, 4, 72, 128, device="cuda", generator=g)
        fr = eng.vae_decode(z)
        x = torch.randn(1, 4, 16, 72, 128, device="cuda", generator=g)
        v = torch.randn(1, 4, 16, 72, 128, device="cuda", generator=g).half()
        eng.ddim_step(x, v, v * 0.5, torch.randn_like(x), cfg_scale=7.5, guidance_rescale=0.7, sqrt_ac=0.6, sqrt_1mac=0.8,
                      rescale=1.0, a_prev=0.5, sigma=0.1)
        clip = torch.randn(1, 3, 16, 576, 1024, device="cuda", generator=g).half()
        postdecode(clip, [0])
        torch.cuda.synchronize()
        print("tail done", float(fr.float().abs().max()))


if __name__ == "__main__":
    main()
