"""Generate tests/golden/resampler_small.npz with the UNCHANGED reference Resampler
(/root/reference/lvdm/modules/encoders/resampler.py) and pin the oracle's key layout at full size.
Build container only:  python oracle/make_golden_resampler.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden", "resampler_small.npz")


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    R = load("/root/reference/lvdm/modules/encoders/resampler.py", "ref_resampler")
    O = load(os.path.join(HERE, "mudg_oracle.py"), "mudg_oracle")
    torch.set_grad_enabled(False)
    # full-size key / shape pin (shipped infer yaml: dim 1024, depth 4, heads 12, 16 queries x 16 frames, 1280 -> 1024)
    with torch.device("meta"):
        full = R.Resampler(dim=1024, depth=4, dim_head=64, heads=12, num_queries=16, embedding_dim=1280, output_dim=1024,
                           ff_mult=4, video_length=16)
    ref_shapes = {k: tuple(v.shape) for k, v in full.state_dict().items()}
    assert ref_shapes == O.resampler_param_shapes(), set(ref_shapes) ^ set(O.resampler_param_shapes())
    cfg = dict(dim=128, depth=2, dim_head=64, heads=2, num_queries=4, embedding_dim=96, output_dim=128, ff_mult=4, video_length=4)
    m = R.Resampler(**cfg).eval()
    sd = O.seeded_state_dict(O.resampler_param_shapes(**cfg), seed=5)
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(3, 9, 96, generator=g)
    y = m(x)
    mine = O.resampler_forward(sd, x, heads=2)
    err = float((y - mine).abs().max())
    print("resampler small: ref vs oracle max|d| =", err, "ref absmax", float(y.abs().max()), "keys", len(ref_shapes))
    assert err < 1e-4
    np.savez_compressed(OUT, x=x.numpy(), y=y.numpy())
    print("wrote", OUT)


if __name__ == "__main__":
    main()
