"""Drop-in for the reference's virtual_render/eval_tools.py (same function names, arguments and files written), with the
per-pixel work -- clamp -> uint8, depth mean + Spectral colour map, nearest-of-19-palette semantic classes -- done in ONE
pass on the GPU by `mudg_postdecode` (libmudg_sm100.so) instead of per-frame CPU NumPy.  Results are bit-exact against
the reference functions (tests/golden/post_small.npz); only uint8 / fp32-depth / class-index planes cross PCIe.

Reference: virtual_render/eval_tools.py:13-134 (save_virtual_*_results), :137-250 (colormap), :253-295 (visualize_depth),
:298-347 (visualize_semantic).  There is no CPU fallback for model outputs: the conversion kernels need the library.
"""
from __future__ import annotations

import os
from typing import List, Optional, Union

import numpy as np
import torch

from mudg_b200.engine import (MUDG_POST_COLOR, MUDG_POST_DEPTH, MUDG_POST_SEMANTIC, MudgError, colormap_spectral,
                              postdecode)


def _write_png(img_u8_chw: torch.Tensor, path: str) -> None:
    try:
        import torchvision
        torchvision.io.write_png(img_u8_chw.cpu().contiguous(), path, 0)
    except ImportError:                                   # torchvision is an I/O convenience only
        import PIL.Image
        PIL.Image.fromarray(img_u8_chw.permute(1, 2, 0).cpu().numpy()).save(path, compress_level=0)


def _to_u8_host(x: torch.Tensor) -> torch.Tensor:
    """((x + 1) / 2 * 255).to(uint8) of a dataset frame (ground truth / sparse input; eval_tools.py:35-36) -- not a model
    output, stays on the host exactly as the reference computes it."""
    return ((x.cpu() + 1) / 2 * 255).to(torch.uint8)


def _device_clip(samples: torch.Tensor) -> torch.Tensor:
    if not torch.is_tensor(samples) or samples.dim() != 5:
        raise MudgError("samples must be a [b, c, t, h, w] tensor")
    if not samples.is_cuda:
        if not torch.cuda.is_available():
            raise MudgError("the post-decode kernels need a CUDA device (B200); there is no CPU fallback")
        samples = samples.cuda()
    return samples


def convert_clip(samples: torch.Tensor, mode: int):
    """samples [b, 3, t, h, w] in [-1,1] (any float dtype) -> (rgb uint8 [b,t,3,h,w], depth fp32 [b,t,h,w] | None,
    cls uint8 [b,t,h,w] | None), all on the device.  The clamp of virtual_pose_render.py:243 is part of the kernel."""
    samples = _device_clip(samples)
    return postdecode(samples, [mode] * samples.shape[0])


def _sample_path(fakedir: str, dir_name: str) -> str:
    p = os.path.join(fakedir.replace("samples", dir_name))
    os.makedirs(p, exist_ok=True)
    return p


def save_virtual_color_results(prompt, samples, filename, fakedir, gts, sparses, base_index, fps=10,
                               dir_name="virtual_samples_separate"):
    rgb, _, _ = convert_clip(samples, MUDG_POST_COLOR)
    rgb = rgb.cpu()
    sample_path = _sample_path(fakedir, dir_name)
    for i in range(rgb.shape[0]):
        for index in range(1, rgb.shape[1]):              # the reference skips frame 0 (eval_tools.py:33)
            result = rgb[i, index]
            dense, sparse = _to_u8_host(gts[i][:, index]), _to_u8_host(sparses[i][:, index])
            _write_png(result, os.path.join(sample_path, f"color_re_{base_index + index}.png"))
            _write_png(dense, os.path.join(sample_path, f"color_gt_{base_index + index}.png"))
            _write_png(sparse, os.path.join(sample_path, f"color_sp_{base_index + index}.png"))
            all_result = torch.stack([dense, result, sparse], dim=2).view(3, result.shape[1], result.shape[2] * 3)
            _write_png(all_result, os.path.join(sample_path, f"color_all_{base_index + index}.png"))


def save_virtual_depth_results(prompt, samples, filename, fakedir, gts, sparses, base_index, fps=10, is_virtual=False,
                               dir_name="virtual_samples_separate"):
    rgb, depth, _ = convert_clip(samples, MUDG_POST_DEPTH)
    rgb, depth = rgb.cpu(), depth.cpu()
    sample_path = _sample_path(fakedir, dir_name)
    depth_path = _sample_path(fakedir, "depth")
    for i in range(rgb.shape[0]):
        for index in range(1, rgb.shape[1]):
            np.save(os.path.join(depth_path, f"depth_re_{base_index + index}.npy"), depth[i, index][None].numpy())
            result = rgb[i, index]
            gt = (torch.mean(gts[i][:, index], dim=0, keepdim=True) + 1) / 2
            np.save(os.path.join(depth_path, f"depth_gt_{base_index + index}.npy"), gt.cpu().numpy())
            if is_virtual:
                dense = _to_u8_host(gts[i][:, index])
            else:
                dense = torch.tensor(np.array(visualize_depth(gt.cpu().numpy())[0])).permute(2, 0, 1)
            sparse = _to_u8_host(sparses[i][:, index])
            _write_png(result, os.path.join(sample_path, f"color_re_{base_index + index}.png"))
            _write_png(dense, os.path.join(sample_path, f"color_gt_{base_index + index}.png"))
            _write_png(sparse, os.path.join(sample_path, f"color_sp_{base_index + index}.png"))
            all_result = torch.stack([dense, result, sparse], dim=2).view(3, result.shape[1], result.shape[2] * 3)
            _write_png(all_result, os.path.join(sample_path, f"color_all_{base_index + index}.png"))


def save_virtual_semantic_results(prompt, samples, filename, fakedir, gts, sparses, base_index, fps=10,
                                  dir_name="virtual_samples_separate"):
    rgb, _, cls = convert_clip(samples, MUDG_POST_SEMANTIC)
    rgb, cls = rgb.cpu(), cls.cpu()
    sample_path = _sample_path(fakedir, dir_name)
    semantic_path = _sample_path(fakedir, "semantic")
    for i in range(rgb.shape[0]):
        for index in range(1, rgb.shape[1]):
            vis_pred = rgb[i, index]
            np.save(os.path.join(semantic_path, f"semantic_re_{base_index + index}.npy"), cls[i, index].long().numpy())
            dense = _to_u8_host(gts[i][:, index])
            _, semantic_gt = visualize_semantic(dense, return_pt=True)
            np.save(os.path.join(semantic_path, f"semantic_gt_{base_index + index}.npy"), semantic_gt.numpy())
            sparse = _to_u8_host(sparses[i][:, index])
            _write_png(vis_pred, os.path.join(sample_path, f"color_re_{base_index + index}.png"))
            _write_png(dense, os.path.join(sample_path, f"color_gt_{base_index + index}.png"))
            _write_png(sparse, os.path.join(sample_path, f"color_sp_{base_index + index}.png"))
            all_result = torch.stack([dense, vis_pred, sparse], dim=2).view(3, vis_pred.shape[1], vis_pred.shape[2] * 3)
            _write_png(all_result, os.path.join(sample_path, f"color_all_{base_index + index}.png"))


def colormap(image: Union[np.ndarray, torch.Tensor], cmap: str = "Spectral", bytes: bool = False,
             _force_method: Optional[str] = None):
    """The reference's `colormap` restricted to what the path uses: cmap="Spectral", bytes=True on a [H, W] map of
    values in [0,1] (eval_tools.py:288).  Returns [H, W, 3] uint8 of the input's kind (ndarray / tensor)."""
    if not (torch.is_tensor(image) or isinstance(image, np.ndarray)):
        raise ValueError("Argument must be a numpy array or torch tensor.")
    if cmap != "Spectral" or not bytes:
        raise MudgError("only colormap(..., cmap='Spectral', bytes=True) is implemented (the one the driver uses)")
    is_np = isinstance(image, np.ndarray)
    t = torch.as_tensor(image)
    if t.dtype == torch.uint8:
        t = t.float() / 255
    t = t.float()
    if t.dim() != 2:
        raise MudgError("colormap expects a 2-D map")
    if not torch.cuda.is_available():
        raise MudgError("the post-decode kernels need a CUDA device (B200); there is no CPU fallback")
    out = colormap_spectral(t.cuda()).cpu()
    return out.numpy() if is_np else out


def visualize_depth(depth, val_min: float = 0.0, val_max: float = 1.0, color_map: str = "Spectral") -> List:
    """Ground-truth depth visualisation (eval_tools.py:253-295): list of PIL images, one per map in `depth` [n, H, W]."""
    import PIL.Image
    if depth is None:
        raise ValueError("Input depth is `None`")
    depth = np.asarray(depth)
    if depth.ndim == 2:
        depth = depth[None, ...]
    if val_max <= val_min:
        raise ValueError(f"Invalid values range: [{val_min}, {val_max}].")
    out = []
    for img in depth:
        img = torch.from_numpy(np.ascontiguousarray(img))
        if val_min != 0.0 or val_max != 1.0:
            img = (img - val_min) / (val_max - val_min)
        out.append(PIL.Image.fromarray(colormap(img, cmap=color_map, bytes=True).cpu().numpy()))
    return out


def visualize_semantic(semantic, return_pt: bool = False):
    """semantic: [3, H, W] uint8 frame -> (palette-coloured frame, class-index map) as the reference returns them
    (eval_tools.py:298-347): ([H, W, 3] uint8 ndarray, [H, W] int64 ndarray), or ([3, H, W] uint8, [H, W] int64) tensors
    with return_pt.  Runs the semantic branch of `mudg_postdecode` on the uint8 frame."""
    t = torch.as_tensor(semantic)
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[0] != 3:
        raise MudgError("visualize_semantic expects a [3, H, W] uint8 frame")
    if not torch.cuda.is_available():
        raise MudgError("the post-decode kernels need a CUDA device (B200); there is no CPU fallback")
    rgb, _, cls = postdecode(t.cuda()[None, :, None], [MUDG_POST_SEMANTIC])
    vis, idx = rgb[0, 0].cpu(), cls[0, 0].cpu().long()
    if return_pt:
        return vis, idx
    return vis.permute(1, 2, 0).numpy(), idx.numpy()
