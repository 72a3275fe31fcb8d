"""nn.Module scaffolding that exposes a flat {key: shape} layout as a tree of sub-modules, so that
state_dict()/load_state_dict(strict=True) use exactly the reference's key names."""
from __future__ import annotations

import math
from typing import Iterable, Mapping, Tuple

import torch
import torch.nn as nn


class ParamNode(nn.Module):
    """A bare container; children are ParamNodes (named by key segments) or leaf Parameters."""


def build_param_tree(root: nn.Module, layout: Mapping[str, Tuple[int, ...]], zero_keys: Iterable[str] = (),
                     seed: int = 0) -> None:
    zero = set(zero_keys)
    g = torch.Generator().manual_seed(seed)
    for key, shape in layout.items():
        parts = key.split(".")
        node = root
        for seg in parts[:-1]:
            child = node._modules.get(seg)
            if child is None:
                child = ParamNode()
                node.add_module(seg, child)
            node = child
        if key in zero:
            val = torch.zeros(shape)
        elif len(shape) == 1:
            val = torch.ones(shape) if parts[-1] == "weight" else torch.zeros(shape)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = 1.0 / math.sqrt(fan_in)
            val = (torch.rand(shape, generator=g) * 2 - 1) * bound
        node.register_parameter(parts[-1], nn.Parameter(val, requires_grad=False))


def mark_dirty_on_load(module: nn.Module, attr: str = "_engine_dirty") -> None:
    """Any load_state_dict / .to() / .cuda() invalidates the packed copy inside the engine."""
    setattr(module, attr, True)

    def _post(mod, incompatible):
        setattr(module, attr, True)
    module.register_load_state_dict_post_hook(_post)
