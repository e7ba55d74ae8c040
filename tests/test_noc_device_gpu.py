"""Device-side oracle clicker + IoU (csrc/noc.cu) against the host clicker that mirrors the reference
(pvpuformer_b200/inference/clicker.py <- isegm/inference/clicker.py:29-69, cv2.distanceTransform) -- bit-exact clicks."""
import numpy as np
import pytest
import torch

from pvpuformer_b200.inference.clicker import Clicker
from pvpuformer_b200.inference.evaluation import get_iou

pytestmark = pytest.mark.gpu


def _blobs(rng, H, W, n):
    m = np.zeros((H, W), dtype=bool)
    yy, xx = np.mgrid[:H, :W]
    for _ in range(n):
        cy, cx = rng.integers(0, H), rng.integers(0, W)
        ry, rx = rng.integers(3, H // 3), rng.integers(3, W // 3)
        m |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
    return m


def _cases(H, W):
    rng = np.random.default_rng(7)
    cases = []
    for i in range(10):
        gt = _blobs(rng, H, W, 2).astype(np.int32)
        pred = _blobs(rng, H, W, 2) if i % 3 else np.zeros((H, W), dtype=bool)
        if i == 4:
            gt[rng.random((H, W)) < 0.05] = -1                          # ignore label
        cases.append((gt, pred))
    cases.append((np.zeros((H, W), np.int32), np.zeros((H, W), bool)))    # nothing wrong: click (0, 0), negative
    cases.append((np.ones((H, W), np.int32), np.zeros((H, W), bool)))     # whole image false negative: centre-most pixel, ties
    cases.append((np.zeros((H, W), np.int32), np.ones((H, W), bool)))     # whole image false positive
    g = np.zeros((H, W), np.int32); g[5:9, 5:9] = 1; g[20:24, 30:34] = 1  # two equal squares: first maximum in row-major order
    cases.append((g, np.zeros((H, W), bool)))
    g = np.zeros((H, W), np.int32); g[0:6, 0:6] = 1                        # region touching the border: the zero padding matters
    cases.append((g, np.zeros((H, W), bool)))
    rnd = rng.random((H, W)) < 0.5
    cases.append((rnd.astype(np.int32), rng.random((H, W)) < 0.5))         # salt and pepper: many ties at distance 1
    return cases


# sizes >= 2e4 pixels: below that cv2 4.13 itself stops returning the correctly rounded sqrt of the exact squared distance
# (+-1 ulp, pixel dependent -- measured map in csrc/noc.cu), so exact ties have no well-defined reference order there
@pytest.mark.parametrize("H,W", [(448, 448), (233, 301)])
def test_device_clicker_matches_host_clicker_over_several_clicks(H, W):
    from pvpuformer_b200 import ops
    cases = _cases(H, W)
    S = len(cases)
    dev = torch.device("cuda:0")
    gt_d = torch.from_numpy(np.stack([c[0] for c in cases]).astype(np.int8)).to(dev)
    not_clicked = torch.ones(S, H, W, dtype=torch.uint8, device=dev)
    hosts = [Clicker(gt_mask=c[0]) for c in cases]
    rng = np.random.default_rng(11)
    preds = [c[1] for c in cases]
    for step in range(5):
        pred_d = torch.from_numpy(np.stack(preds).astype(np.uint8)).to(dev)
        clicks, counts = ops.noc_next_clicks(gt_d, pred_d, not_clicked)
        clicks, counts = clicks.cpu().numpy(), counts.cpu().numpy()
        for s, (h, (gt, _)) in enumerate(zip(hosts, cases)):
            c = h._get_next_click(preds[s])
            h.add_click(c)
            assert (int(clicks[s, 0]) == int(c.is_positive) and int(clicks[s, 1]) == int(c.coords[0]) and
                    int(clicks[s, 2]) == int(c.coords[1])), (step, s, clicks[s], c.is_positive, c.coords)
            keep, obj = gt != -1, gt == 1
            assert counts[s, 0] == np.logical_and(np.logical_and(preds[s], obj), keep).sum()
            assert counts[s, 1] == np.logical_and(np.logical_or(preds[s], obj), keep).sum()
            if counts[s, 1] > 0:
                assert np.float32(counts[s, 0] / counts[s, 1]) == np.float32(get_iou(gt, preds[s]))
        assert np.array_equal(not_clicked.cpu().numpy().astype(bool), np.stack([h.not_clicked_map for h in hosts]))
        # next round: perturb the predictions (the clicked pixels stay excluded through not_clicked)
        preds = [np.logical_xor(p, _blobs(rng, H, W, 1)) for p in preds]


def test_lockstep_with_device_clicker_matches_host_clicker_loop():
    """The NoC loop with the clicker and the IoU tally on the device gives the IoU table of the host-clicker loop bit for bit."""
    from pvpuformer_b200.config import make_config
    from pvpuformer_b200.inference import evaluate_lockstep
    from pvpuformer_b200.inference.datasets import SyntheticEllipseDataset
    from pvpuformer_b200.model import build_model
    from pvpuformer_b200.weights import synthetic_state_dict
    dev = torch.device("cuda:0")
    cfg = make_config("vit_base")
    m = build_model("vit_base", state_dict=synthetic_state_dict(cfg, 0), device=dev)
    m.want_aux = False
    ds = SyntheticEllipseDataset(5, seed0=70)
    samples = [(ds.get_sample(i).image, ds.get_sample(i).gt_mask(1)) for i in range(5)]
    host = evaluate_lockstep(samples, m, dev, 1.01, max_clicks=4, micro_batch=3)
    devc = evaluate_lockstep(samples, m, dev, 1.01, max_clicks=4, micro_batch=3, device_clicker=True)
    for a, b in zip(host, devc):
        assert a.dtype == b.dtype and np.array_equal(a, b), (a, b)
