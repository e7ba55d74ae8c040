"""Synthetic workloads of the per-click forward (no datasets or checkpoints exist offline).

Pure functions of seeds (torch CPU generators / numpy RandomState) so every box regenerates the same
inputs: SURVEY.md 8(d) config 2 -- images U[0,1) fp32, previous-mask channel = sigmoid of N(0,1)
noise (or zeros), per image k in U{1..20} clicks packed by the NoBRS predictor's get_points_nd rule
(reference isegm/inference/predictors/base.py:195-213: positives first, negatives second, each half
padded with (-1,-1,-1) to the batch maximum).
"""
import numpy as np
import torch


def images(B, seed, prev="sigmoid", size=448):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, 3, size, size, generator=g)
    if prev == "sigmoid":
        pm = torch.sigmoid(torch.randn(B, 1, size, size, generator=g))
    else:
        pm = torch.zeros(B, 1, size, size)
    return torch.cat([img, pm], 1)


def random_clicks(B, seed, max_clicks=20, size=448, dtype=torch.float32):
    rs = np.random.RandomState(seed)
    per = []
    for b in range(B):
        k = rs.randint(1, max_clicks + 1)
        pos, neg = [], []
        for i in range(k):
            rc = (float(rs.randint(0, size)), float(rs.randint(0, size)), float(i))
            (pos if (i == 0 or rs.rand() < 0.5) else neg).append(rc)
        per.append((pos, neg))
    n = max(1, max(max(len(p), len(q)) for p, q in per))
    out = []
    for pos, neg in per:
        pos = pos + [(-1., -1., -1.)] * (n - len(pos))
        neg = neg + [(-1., -1., -1.)] * (n - len(neg))
        out.append(pos + neg)
    return torch.tensor(out, dtype=dtype)
