"""Prompt simulators (pvpuformer_b200/inference/prompts.py <- isegm/engine/trainer.py:703-768,1061-1243) under fixed seeds."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import ref_harness as rh
from pvpuformer_b200.inference import prompts as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prompt_simulators.npz")


def make_case(seed, B=3, H=96, W=128):
    """Ground truth of a few blobs (one of them dominant), a prediction that misses part of it and adds a false region."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[:H, :W]
    gt = np.zeros((B, 1, H, W), np.float32)
    pred = np.zeros((B, 1, H, W), np.float32)
    for b in range(B):
        for k in range(int(rng.integers(1, 4))):
            cy, cx, ry, rx = rng.uniform(20, H - 20), rng.uniform(20, W - 20), rng.uniform(6, 30), rng.uniform(6, 40)
            gt[b, 0] = np.maximum(gt[b, 0], ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1)
        shift = np.roll(gt[b, 0], (int(rng.integers(-12, 12)), int(rng.integers(-12, 12))), axis=(0, 1))
        pred[b, 0] = np.clip(0.8 * shift + 0.15 * rng.random((H, W)), 0, 1)
    if seed % 3 == 0:
        gt[B - 1] = 0                                              # an empty object: zero scribble, zero box
    n = 4
    pts = -torch.ones(B, 2 * n, 3)
    for b in range(B):
        pts[b, 0] = torch.tensor([float(rng.integers(0, H)), float(rng.integers(0, W)), 0.0])
        if b % 2:
            pts[b, n] = torch.tensor([float(rng.integers(0, H)), float(rng.integers(0, W)), 1.0])
    return torch.from_numpy(pred), torch.from_numpy(gt), pts


def run(fn, seed, **kw):
    pred, gt, pts = make_case(seed)
    random.seed(seed)
    np.random.seed(seed)
    p, boxes, (scr, rect) = fn(pred, gt, pts, **kw)
    return p.numpy(), boxes.numpy(), np.asarray(scr), np.asarray(rect)


SETTINGS = [dict(as_allmask=False, jitter_box=False), dict(as_allmask=True, jitter_box=True), dict(as_allmask=False, jitter_box=True)]


def test_simulators_match_the_committed_golden_vectors():
    g = np.load(GOLD)
    for seed in range(6):
        for k, kw in enumerate(SETTINGS):
            out = run(P.get_next_promts, seed, **kw)
            for name, a in zip(("points", "boxes", "scribbles", "rects"), out):
                want = g["s%d_k%d_%s" % (seed, k, name)]
                assert a.dtype == want.dtype and a.shape == want.shape and np.array_equal(a, want), (seed, kw, name)


def test_largest_region_rule():
    m = np.zeros((20, 40), bool)
    m[1:5, 1:5] = True        # 16 px, visited first: the running maximum
    m[10:18, 10:30] = True    # 160 px: becomes the maximum, the first component is NOT merged into it
    m[1:6, 30:38] = True      # 40 px > 10 % of 216: visited second in raster order (row 1), merged into component 1 ...
    out = P.max_connected_regions(m)
    assert out.dtype == np.int8 and out[12, 15] == 1 and out[2, 2] == 0 and out[2, 32] == 0      # ... and dropped with it
    assert P.max_connected_regions(np.zeros((4, 4), bool)).sum() == 0


@pytest.mark.reference
def test_simulators_match_the_unmodified_reference():
    rh.import_reference()
    from isegm.engine.trainer import get_next_promts as ref_fn
    for seed in range(12):
        for kw in SETTINGS:
            a, b = run(P.get_next_promts, seed, **kw), run(ref_fn, seed, **kw)
            for name, x, y in zip(("points", "boxes", "scribbles", "rects"), a, b):
                assert x.dtype == y.dtype and x.shape == y.shape and np.array_equal(x, y), (seed, kw, name)
