"""Minimal stand-ins for three packages the reference driver imports but this image does not ship
(virtual_render/virtual_pose_render.py:3,10,22): omegaconf, pytorch_lightning.seed_everything, megfile.smart_open.
`install()` registers them in sys.modules ONLY when the real package is missing, so the unchanged driver imports."""
from __future__ import annotations

import importlib
import random
import sys
import types


class AttrDict(dict):
    """dict with attribute access, enough for `config.model`, `cfg.params.temporal_length`, `.pop`, `.get`."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(obj):
    if isinstance(obj, dict):
        return AttrDict({k: to_attr(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [to_attr(v) for v in obj]
    return obj


def _omegaconf():
    import yaml
    mod = types.ModuleType("omegaconf")

    class OmegaConf:
        @staticmethod
        def load(path):
            with open(path) as f:
                return to_attr(yaml.safe_load(f))

        @staticmethod
        def create(obj=None):
            return to_attr(obj or {})

        @staticmethod
        def to_container(cfg, resolve=True):
            return cfg
    mod.OmegaConf = OmegaConf
    return mod


def _pytorch_lightning():
    mod = types.ModuleType("pytorch_lightning")

    def seed_everything(seed, workers=False):
        import numpy as np
        import torch
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        return seed
    mod.seed_everything = seed_everything
    return mod


def _megfile():
    mod = types.ModuleType("megfile")
    mod.smart_open = open
    import os
    mod.smart_exists = os.path.exists
    mod.smart_makedirs = lambda p, exist_ok=True: os.makedirs(p, exist_ok=exist_ok)
    mod.smart_path_join = os.path.join           # virtual_render/data_tools.py:5 (local paths only)
    mod.smart_listdir = os.listdir
    mod.smart_isdir = os.path.isdir
    return mod


def install():
    for name, factory in (("omegaconf", _omegaconf), ("pytorch_lightning", _pytorch_lightning), ("megfile", _megfile)):
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except ImportError:
            sys.modules[name] = factory()
