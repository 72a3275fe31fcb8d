"""CPU oracle of the post-decode frame pipeline -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(virtual_render/eval_tools.py -> mudg_postdecode in libmudg_sm100.so) never does.

numpy restatement of what the reference driver does on the CPU with the decoded clip:
  * to_uint8          virtual_render/virtual_pose_render.py:243 (clamp) + eval_tools.py:22-27 (uint8 conversion)
  * depth_from_uint8  eval_tools.py:70 (mean over RGB of the uint8 frame, / 255)
  * spectral_u8       eval_tools.py:205-236 (colormap, method_custom, "Spectral", bytes=True) via visualize_depth :282-289
  * semantic_from_uint8  eval_tools.py:309-347 (visualize_semantic: nearest of the 19 palette colours)
Byte / integer / index outputs: the parity bar is BIT-EXACT.  Pinned against the reference functions themselves by
oracle/make_golden_post.py -> tests/golden/post_small.npz.
"""
from __future__ import annotations

import numpy as np

MODE_COLOR, MODE_DEPTH, MODE_SEMANTIC = 0, 1, 2
LABEL_TO_MODE = {0: MODE_COLOR, 500: MODE_DEPTH, 1: MODE_SEMANTIC}     # class labels of the driver (virtual_pose_render.py:247-318)

SPECTRAL = np.array([   # eval_tools.py:170-182 (matplotlib/_cm.py), rounded to fp32 like torch.tensor(..., dtype=torch.float)
    (0.61960784313725492, 0.003921568627450980, 0.25882352941176473),
    (0.83529411764705885, 0.24313725490196078, 0.30980392156862746),
    (0.95686274509803926, 0.42745098039215684, 0.2627450980392157),
    (0.99215686274509807, 0.68235294117647061, 0.38039215686274508),
    (0.99607843137254903, 0.8784313725490196, 0.54509803921568623),
    (1.0, 1.0, 0.74901960784313726),
    (0.90196078431372551, 0.96078431372549022, 0.59607843137254901),
    (0.6705882352941176, 0.8666666666666667, 0.64313725490196083),
    (0.4, 0.76078431372549016, 0.6470588235294118),
    (0.19607843137254902, 0.53333333333333333, 0.74117647058823533),
    (0.36862745098039218, 0.30980392156862746, 0.63529411764705879),
], dtype=np.float32)

PALETTE = np.array([    # eval_tools.py:312-332
    [255, 120, 50], [255, 192, 203], [255, 255, 0], [0, 150, 245], [0, 255, 255], [255, 127, 0], [255, 0, 0],
    [255, 240, 150], [135, 60, 0], [160, 32, 240], [255, 0, 255], [139, 137, 137], [75, 0, 75], [150, 240, 80],
    [230, 230, 250], [0, 175, 0], [0, 255, 127], [222, 155, 161], [140, 62, 69]], dtype=np.int64)

_F = np.float32


def to_uint8(video: np.ndarray) -> np.ndarray:
    """clamp(x.float(), -1, 1); (x + 1.0) / 2.0; (x * 255).to(uint8)  -- any shape, fp32 arithmetic, truncation."""
    v = np.clip(video.astype(np.float32), _F(-1.0), _F(1.0))
    g = (v + _F(1.0)) / _F(2.0)
    return (g * _F(255.0)).astype(np.uint8)


def depth_from_uint8(frame_u8: np.ndarray) -> np.ndarray:
    """frame [3, H, W] uint8 -> [H, W] fp32: torch.mean(frame.float(), dim=0) / 255 (sum, then / 3, then / 255)."""
    s = frame_u8[0].astype(np.float32) + frame_u8[1].astype(np.float32) + frame_u8[2].astype(np.float32)
    return (s / _F(3.0)) / _F(255.0)


def spectral_u8(depth01: np.ndarray) -> np.ndarray:
    """[H, W] fp32 in [0,1] -> [3, H, W] uint8 (method_custom: K = 11 anchors, linear interpolation, * 255, truncation)."""
    pos = np.clip(depth01.astype(np.float32), _F(0.0), _F(1.0)) * _F(10.0)
    left = pos.astype(np.int64)
    right = np.minimum(left + 1, 10)
    d = (pos - left.astype(np.float32))[..., None]
    out = (_F(1.0) - d) * SPECTRAL[left] + d * SPECTRAL[right]
    return np.ascontiguousarray((out * _F(255.0)).astype(np.uint8).transpose(2, 0, 1))


def semantic_from_uint8(frame_u8: np.ndarray):
    """frame [3, H, W] uint8 -> (vis [3, H, W] uint8, cls [H, W] int64): argmin_k ||rgb - palette[k]||_2, first minimum."""
    px = frame_u8.astype(np.int64).transpose(1, 2, 0)[:, :, None, :]            # H, W, 1, 3
    dist = np.linalg.norm(px - PALETTE[None, None], axis=3)                     # fp64 like the reference
    cls = np.argmin(dist, axis=2)
    vis = PALETTE.astype(np.uint8)[cls].transpose(2, 0, 1)
    return np.ascontiguousarray(vis), cls


def postdecode(frames: np.ndarray, modes):
    """frames [B, 3, T, H, W] float; modes [B] in {0,1,2}.  Returns rgb [B,T,3,H,W] uint8, depth [B,T,H,W] fp32 (zeros for
    non-depth samples), cls [B,T,H,W] uint8 (zeros for non-semantic samples)."""
    B, C, T, H, W = frames.shape
    assert C == 3
    u8 = to_uint8(frames).transpose(0, 2, 1, 3, 4)                               # B, T, 3, H, W
    rgb = np.empty((B, T, 3, H, W), np.uint8)
    depth = np.zeros((B, T, H, W), np.float32)
    cls = np.zeros((B, T, H, W), np.uint8)
    for b in range(B):
        for t in range(T):
            if modes[b] == MODE_COLOR:
                rgb[b, t] = u8[b, t]
            elif modes[b] == MODE_DEPTH:
                depth[b, t] = depth_from_uint8(u8[b, t])
                rgb[b, t] = spectral_u8(depth[b, t])
            else:
                vis, c = semantic_from_uint8(u8[b, t])
                rgb[b, t] = vis
                cls[b, t] = c.astype(np.uint8)
    return rgb, depth, cls
