"""Post-decode frame pipeline: the numpy oracle replayed against the golden vectors that
oracle/make_golden_post.py produced with the UNCHANGED reference functions (virtual_render/eval_tools.py).  Byte /
index work: every comparison is bit-exact."""
import os

import numpy as np

from oracle import post_oracle as P


def _g(golden_dir):
    return np.load(os.path.join(golden_dir, "post_small.npz"))


def test_uint8_conversion_matches_reference(golden_dir):
    g = _g(golden_dir)
    assert np.array_equal(P.to_uint8(g["frames"]).transpose(0, 2, 1, 3, 4), g["u8"])


def test_depth_and_spectral_match_reference(golden_dir):
    g = _g(golden_dir)
    rgb, depth, _ = P.postdecode(g["frames"], [P.MODE_COLOR, P.MODE_DEPTH, P.MODE_SEMANTIC])
    assert np.array_equal(depth[1], g["depth_pred"])           # fp32 bit patterns
    assert depth[1].tobytes() == g["depth_pred"].tobytes()
    assert np.array_equal(rgb[1], g["depth_vis"])


def test_semantic_matches_reference(golden_dir):
    g = _g(golden_dir)
    rgb, _, cls = P.postdecode(g["frames"], [0, 1, 2])
    assert np.array_equal(rgb[2], g["sem_vis"])
    assert np.array_equal(cls[2].astype(np.int64), g["sem_cls"])
    assert len(np.unique(g["sem_cls"])) == 19                   # the fixture exercises every class


def test_edge_values():
    # clamp limits, signed zero, exact k/255 levels
    x = np.array([-2.0, -1.0, -0.0, 0.0, 1.0, 2.0, 1.0 / 255 * 2 - 1, 0.5], np.float32)
    # truncation, not rounding: fp32((1/255*2-1 + 1) / 2 * 255) = 0.99999994 -> 0
    assert P.to_uint8(x).tolist() == [0, 0, 127, 127, 255, 255, 0, 191]
    # Spectral end points: pos == 10 -> left == right == last anchor
    ends = P.spectral_u8(np.array([[0.0, 1.0, -3.0, 7.0]], np.float32))
    assert ends[:, 0, 0].tolist() == [158, 1, 66] and ends[:, 0, 1].tolist() == [94, 79, 162]
    assert np.array_equal(ends[:, 0, 0], ends[:, 0, 2]) and np.array_equal(ends[:, 0, 1], ends[:, 0, 3])
    # palette colours map to themselves; ties resolve to the first class
    pal = P.PALETTE.astype(np.uint8).T[:, None, :]              # [3, 1, 19]
    vis, cls = P.semantic_from_uint8(pal)
    assert cls[0].tolist() == list(range(19)) and np.array_equal(vis, pal)


def test_semantic_idempotent():
    rng = np.random.default_rng(0)
    f = rng.integers(0, 256, size=(3, 16, 32), dtype=np.uint8)
    vis, cls = P.semantic_from_uint8(f)
    vis2, cls2 = P.semantic_from_uint8(vis)
    assert np.array_equal(vis, vis2) and np.array_equal(cls, cls2)
