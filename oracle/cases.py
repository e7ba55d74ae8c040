"""Synthetic parity cases shared by oracle/make_golden.py and tests/ (TEST INFRASTRUCTURE ONLY).

Inputs are pure functions of seeds (torch CPU generators / numpy RandomState), so the GPU box
regenerates exactly what the build container fed to the reference when the golden vectors in
tests/golden/ were made.  Prompts for types 1/2 came from the reference's own prompt simulator
(get_next_promts, reference isegm/engine/trainer.py:703-768) and are stored in the fixture.
"""
import numpy as np
import torch


from pvpuformer_b200.synthetic import images, random_clicks  # noqa: F401  (one definition, shared with bench.py)


# clicks exercising: centre, fractional coords, corner drop quirk (5,440), image corner, (0,0),
# padding rows, float64 dtype (what the NoBRS predictor produces)
CLICKS_A = torch.tensor(
    [[[224., 224, 0], [100, 50.5, 2], [-1, -1, -1], [300, 310, 1], [-1, -1, -1], [-1, -1, -1]],
     [[5, 440, 0], [-1, -1, -1], [-1, -1, -1], [447, 447, 1], [0, 0, 2], [200.7, 13.2, 3]]],
    dtype=torch.float64)


def ellipse_masks(B, seed, size=448):
    """random ellipses, centre U[100,348], semi-axes U[40,140]."""
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:size, 0:size]
    out = []
    for _ in range(B):
        cy, cx = rs.uniform(100, 348, 2)
        a, b = rs.uniform(40, 140, 2)
        out.append((((yy - cy) / a) ** 2 + ((xx - cx) / b) ** 2 <= 1).astype(np.float32))
    return np.stack(out)


def first_clicks_in_masks(masks, seed, n=2):
    rs = np.random.RandomState(seed)
    B = masks.shape[0]
    pts = torch.full((B, 2 * n, 3), -1.0, dtype=torch.float64)
    for b in range(B):
        ys, xs = np.nonzero(masks[b])
        i = rs.randint(len(ys))
        pts[b, 0] = torch.tensor([float(ys[i]), float(xs[i]), 0.])
    return pts


def train12_inputs():
    """ViT-B training shape -- batch 12, points [12, 48, 3] (n = 24 rows per half) with 1-3 valid
    clicks per half, ellipse ground truth; -> (image4, points, gt [12,1,448,448] float)."""
    B, n = 12, 24
    masks = ellipse_masks(B, seed=9)
    rs = np.random.RandomState(10)
    pts = torch.full((B, 2 * n, 3), -1.0, dtype=torch.float32)
    for b in range(B):
        order = 0
        for half, inside in ((0, True), (1, False)):
            ys, xs = np.nonzero(masks[b] > 0.5 if inside else masks[b] < 0.5)
            for j in range(rs.randint(1, 4)):
                i = rs.randint(len(ys))
                pts[b, half * n + j] = torch.tensor([float(ys[i]), float(xs[i]), float(order)])
                order += 1
    return images(B, seed=8), pts, torch.from_numpy(masks)[:, None]
