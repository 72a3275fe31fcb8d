"""Pin oracle/clip_oracle.py and write tests/golden/clip_small.npz.

open_clip (the package the reference calls, condition.py:186,303) is not in this image and not vendored under
/root/reference, so the pin is the OTHER published implementation of the same model that is installed here:
transformers' CLIPVisionModel / CLIPTextModel (hidden_act "gelu", as the HF port of laion/CLIP-ViT-H-14-laion2B-s32B-b79K
configures them).  Seeded open_clip-named weights are renamed to transformers' layout (clip_oracle.open_clip_to_hf) and
both implementations run on the same inputs: small towers (golden vectors, committed) and the full ViT-H/14 towers
(asserted here, too large to commit).  Build container only:  python oracle/make_golden_clip.py [--full]
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import clip_oracle as C  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "clip_small.npz")

SMALL_V = dict(width=128, layers=3, mlp=512, image_size=56, patch=14, embed_dim=64)     # heads 2 -> head dim 64
SMALL_V80 = dict(width=320, layers=2, mlp=640, image_size=42, patch=14, embed_dim=64)   # heads 4 -> head dim 80 (ViT-H's)
SMALL_T = dict(width=128, layers=4, mlp=512, vocab=1000, ctx=77, embed_dim=64)          # heads 2


def hf_vision(shapes_kw, heads, sd):
    from transformers import CLIPVisionConfig, CLIPVisionModel
    cfg = CLIPVisionConfig(hidden_size=shapes_kw["width"], intermediate_size=shapes_kw["mlp"],
                           num_hidden_layers=shapes_kw["layers"], num_attention_heads=heads, image_size=shapes_kw["image_size"],
                           patch_size=shapes_kw["patch"], hidden_act="gelu", layer_norm_eps=1e-5)
    m = CLIPVisionModel(cfg).eval()
    missing = m.load_state_dict(C.open_clip_to_hf(sd, "vision"), strict=False)
    assert not missing.unexpected_keys and all("position_ids" in k for k in missing.missing_keys), missing
    return m


def hf_text(shapes_kw, heads, sd):
    from transformers import CLIPTextConfig, CLIPTextModel
    cfg = CLIPTextConfig(vocab_size=shapes_kw["vocab"], hidden_size=shapes_kw["width"], intermediate_size=shapes_kw["mlp"],
                         num_hidden_layers=shapes_kw["layers"], num_attention_heads=heads,
                         max_position_embeddings=shapes_kw["ctx"], hidden_act="gelu", layer_norm_eps=1e-5,
                         eos_token_id=shapes_kw["vocab"] - 1, bos_token_id=shapes_kw["vocab"] - 2, pad_token_id=0)
    m = CLIPTextModel(cfg).eval()
    missing = m.load_state_dict(C.open_clip_to_hf(sd, "text"), strict=False)
    assert not missing.unexpected_keys and all("position_ids" in k for k in missing.missing_keys), missing
    return m


def vision_case(kw, heads, seed, B):
    sd = C.seeded_clip_state_dict(C.clip_vision_param_shapes(**kw), seed)
    g = torch.Generator().manual_seed(seed + 100)
    img = torch.randn(B, 3, kw["image_size"], kw["image_size"], generator=g)
    ref = hf_vision(kw, heads, sd)(pixel_values=img).last_hidden_state        # encoder output, before post_layernorm
    mine = C.clip_image_tokens(sd, img, heads)
    err = float((ref - mine).abs().max())
    print(f"vision {kw['width']}x{kw['layers']} heads {heads}: transformers vs oracle max|d| = {err:.3g}  (absmax {float(ref.abs().max()):.3g})")
    assert err < 2e-4 * max(1.0, float(ref.abs().max()))
    return img, ref


def text_case(kw, heads, seed, B):
    sd = C.seeded_clip_state_dict(C.clip_text_param_shapes(**kw), seed)
    g = torch.Generator().manual_seed(seed + 100)
    tok = torch.randint(1, kw["vocab"] - 2, (B, kw["ctx"]), generator=g)
    tok[:, 0] = kw["vocab"] - 2
    tok[:, 20:] = 0
    tok[:, 20] = kw["vocab"] - 1                                              # <start> ... <end> then padding, as open_clip.tokenize
    m = hf_text(kw, heads, sd)
    hs = m(input_ids=tok, output_hidden_states=True).hidden_states
    ref = m.text_model.final_layer_norm(hs[-2])                               # "penultimate" + ln_final (condition.py:220-226)
    mine = C.clip_text_encode(sd, tok, heads, layer_idx=1)
    err = float((ref - mine).abs().max())
    last = float((m.text_model.final_layer_norm(hs[-1]) - C.clip_text_encode(sd, tok, heads, layer_idx=0)).abs().max())
    print(f"text {kw['width']}x{kw['layers']} heads {heads}: transformers vs oracle max|d| = {err:.3g} (penultimate), {last:.3g} (last)")
    assert err < 2e-4 and last < 2e-4
    return tok, ref


def main():
    torch.set_grad_enabled(False)
    torch.manual_seed(0)
    img, vis = vision_case(SMALL_V, 2, 21, 2)
    img80, vis80 = vision_case(SMALL_V80, 4, 22, 1)
    tok, txt = text_case(SMALL_T, 2, 23, 2)
    np.savez_compressed(OUT, img=img.numpy(), vis=vis.numpy(), img80=img80.numpy(), vis80=vis80.numpy(), tok=tok.numpy(),
                        txt=txt.numpy())
    print("wrote", OUT)
    if "--full" in sys.argv:
        # the real ViT-H/14 towers (630 M + 350 M parameters): key / shape inventory and one forward each
        vision_case(dict(width=1280, layers=32, mlp=5120, image_size=224, patch=14, embed_dim=1024), 16, 31, 1)
        text_case(dict(width=1024, layers=24, mlp=4096, vocab=49408, ctx=77, embed_dim=1024), 16, 32, 1)


if __name__ == "__main__":
    main()
