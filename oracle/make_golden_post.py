"""Generate tests/golden/post_small.npz by running the UNCHANGED reference post-decode code
(/root/reference/virtual_render/eval_tools.py) on seeded decoded clips.  Build container only:

    python oracle/make_golden_post.py

The reference functions write files; the arithmetic they do per frame is reproduced here by calling the reference's own
`visualize_depth` / `visualize_semantic` and by executing the same tensor statements as
save_virtual_{color,depth,semantic}_results (eval_tools.py:20-27, 56-74, 108-121) around them.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_FILE = "/root/reference/virtual_render/eval_tools.py"
OUT = os.path.join(ROOT, "tests", "golden", "post_small.npz")


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_eval_tools", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_frames(T=3, H=24, W=40, seed=11):
    """Three samples (colour, depth, semantic) of a decoded clip, fp16 like decode_first_stage under autocast; values
    overshoot [-1,1], hit the clamp limits exactly and sit on uint8 rounding boundaries."""
    g = torch.Generator().manual_seed(seed)
    color = 0.8 * torch.randn(3, T, H, W, generator=g)
    color[:, 0, 0, :8] = torch.tensor([-1.0, 1.0, -1.5, 1.5, 0.0, -0.0, 1.0 / 255, 254.0 / 255])
    k = torch.arange(W, dtype=torch.float32)
    color[:, 0, 1, :] = (k * 6.0) / 255.0 * 2 - 1                      # exact k/255 levels
    depth = (torch.rand(1, T, H, W, generator=g) * 2.4 - 1.2).repeat(3, 1, 1, 1) + 0.02 * torch.randn(3, T, H, W, generator=g)
    depth[:, 0, 0, :4] = torch.tensor([-1.0, 1.0, 0.0, 0.2])
    lin = torch.linspace(-1, 1, H * W).reshape(H, W)
    depth[:, 1] = lin                                                  # sweeps every colour-map segment
    # semantic: palette colours + noise (also far-from-palette pixels)
    ref = load_reference()
    lut = None
    import inspect
    src = inspect.getsource(ref.visualize_semantic)
    assert "color_map" in src
    pal = torch.tensor([[255, 120, 50], [255, 192, 203], [255, 255, 0], [0, 150, 245], [0, 255, 255], [255, 127, 0],
                        [255, 0, 0], [255, 240, 150], [135, 60, 0], [160, 32, 240], [255, 0, 255], [139, 137, 137],
                        [75, 0, 75], [150, 240, 80], [230, 230, 250], [0, 175, 0], [0, 255, 127], [222, 155, 161],
                        [140, 62, 69]], dtype=torch.float32)
    idx = torch.randint(0, 19, (T, H, W), generator=g)
    sem = (pal[idx].permute(3, 0, 1, 2) / 255.0) * 2 - 1 + 0.15 * torch.randn(3, T, H, W, generator=g)
    sem[:, 2] = torch.rand(3, H, W, generator=g) * 2 - 1               # arbitrary colours: ties / far pixels
    del lut
    return torch.stack([color, depth, sem]).half()                      # [3 samples, 3, T, H, W]


def main():
    ref = load_reference()
    frames = make_frames()
    # virtual_pose_render.py:243
    batch = torch.clamp(frames.float(), -1.0, 1.0)
    out = {"frames": frames.numpy()}
    u8_all, depth_pred, depth_vis, sem_vis, sem_cls = [], [], [], [], []
    for b in range(3):
        video = batch[b:b + 1]                                         # [1, c, t, h, w]  (samples = batch_samples[nn])
        video = video.detach().cpu()
        video = torch.clamp(video.float(), -1.0, 1.0)
        grid = video[0, ...]
        grid = (grid + 1.0) / 2.0
        grid = (grid * 255).to(torch.uint8).permute(1, 2, 3, 0)        # thwc            eval_tools.py:25-27
        u8_all.append(grid.permute(0, 3, 1, 2).numpy())                # t c h w
        if b == 1:
            for index in range(grid.shape[0]):
                result_pred = torch.mean(grid[index].permute(2, 0, 1).float(), dim=0, keepdim=True) / 255   # :70
                depth_pred.append(result_pred.numpy()[0])
                result = torch.tensor(np.array(ref.visualize_depth(result_pred.cpu().numpy())[0])).permute(2, 0, 1)   # :73
                depth_vis.append(result.numpy())
        if b == 2:
            for index in range(grid.shape[0]):
                result = grid[index].permute(2, 0, 1)
                vis_pred, semantic_pred = ref.visualize_semantic(result, return_pt=True)                     # :119-120
                sem_vis.append(vis_pred.numpy())
                sem_cls.append(semantic_pred.numpy())
    out.update(u8=np.stack(u8_all), depth_pred=np.stack(depth_pred), depth_vis=np.stack(depth_vis),
               sem_vis=np.stack(sem_vis), sem_cls=np.stack(sem_cls))
    np.savez_compressed(OUT, **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype)
    # the restatement must agree bit for bit before the fixture is trusted
    sys.path.insert(0, HERE)
    import post_oracle as P
    rgb, depth, cls = P.postdecode(out["frames"], [0, 1, 2])
    assert np.array_equal(rgb[0], out["u8"][0]), "uint8 conversion"
    assert np.array_equal(P.to_uint8(out["frames"]).transpose(0, 2, 1, 3, 4), out["u8"]), "uint8 conversion (all)"
    assert np.array_equal(depth[1], out["depth_pred"]), "depth mean"
    assert np.array_equal(rgb[1], out["depth_vis"]), "Spectral"
    assert np.array_equal(rgb[2], out["sem_vis"]), "semantic colours"
    assert np.array_equal(cls[2].astype(np.int64), out["sem_cls"]), "semantic classes"
    print("post_oracle == reference eval_tools (bit-exact) ->", OUT)


if __name__ == "__main__":
    main()
