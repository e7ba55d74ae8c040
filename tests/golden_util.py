"""Helpers shared by the golden-vector tests (oracle port and CUDA path use the same checks)."""
import os
import random

import numpy as np
import torch

from oracle import cases

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def case_inputs(name):
    """-> (image4, points, prompts, as_prompt_type) exactly as oracle/make_golden.py built them."""
    if name.endswith("_clicks"):
        return cases.images(2, seed=1), cases.CLICKS_A.clone(), None, 0
    if name.endswith("_train12"):
        image4, pts, _ = cases.train12_inputs()
        return image4, pts, None, 0
    if name.endswith("_manyclicks"):
        return cases.images(3, seed=5), cases.random_clicks(3, seed=6), None, 0
    g = load(name)
    t = 1 if name.endswith("_box") else 2
    prompts = (torch.from_numpy(g["prompt_points"]), torch.from_numpy(g["boxes"]),
               [g["scribbles"], g["rects"]])
    return cases.images(3, seed=2, prev="zeros"), torch.from_numpy(g["points"]), prompts, t


def unpack_disks(g, B, size=448):
    return np.unpackbits(g["disks_packed"])[:B * 2 * size * size].reshape(B, 2, size, size)


def seed_scribble():
    random.seed(7)   # the seed make_golden.py used right before each reference forward
