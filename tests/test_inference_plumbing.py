"""CPU: the NoBRS plumbing (clicker, transforms, predictor, NoC loop, metrics) against the reference's own
implementation run here with a deterministic stand-in network, lock-step batching against the serial loop, and the
sharded loop + all_gather with world_size 2 on gloo."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pvpuformer_b200.inference import (Clicker, compute_noc_metric, evaluate_lockstep, evaluate_sample, gather_iou_tables,
                                       shard_range)
from pvpuformer_b200.inference.datasets import SyntheticEllipseDataset
from pvpuformer_b200.inference.evaluation import evaluate_sharded, iou_table
from pvpuformer_b200.inference.predictor import vpu_eval_predictor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeNet(torch.nn.Module):
    """Deterministic, batch-independent stand-in with the model's call surface: Gaussian bumps at the clicks
    (+ for positive, - for negative) plus a previous-mask and image term, so ZoomIn / flip / prev-mask plumbing matter."""
    with_prev_mask = True

    def forward(self, image, points, prompts=None, as_prompt_type=0):
        B, _, H, W = image.shape
        yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
        out = torch.zeros(B, 1, H, W, dtype=torch.float64)
        n = points.shape[1] // 2
        for b in range(B):
            for i in range(2 * n):
                if points[b, i, 2] >= 0:
                    sgn = 1.0 if i < n else -1.0
                    out[b, 0] += sgn * 4 * torch.exp(-((yy - points[b, i, 0]) ** 2 + (xx - points[b, i, 1]) ** 2) / (2 * 60.0 ** 2))
        out = out - 1.0 + 0.5 * image[:, 3:4].double() + 0.3 * (image[:, 0:1].double() - 0.5)
        return {"instances": out.float(), "instances_aux": None}


def _samples(n, seed0=0):
    ds = SyntheticEllipseDataset(n, seed0=seed0)
    return [(ds.get_sample(i).image, ds.get_sample(i).gt_mask(1)) for i in range(n)]


def test_shard_range_covers_everything():
    for n in (1, 7, 1024):
        for w in (1, 2, 3, 8):
            got = []
            for r in range(w):
                a, b = shard_range(n, r, w)
                got += list(range(a, b))
            assert got == list(range(n))


def test_lockstep_equals_serial_and_early_stop():
    net, samples = FakeNet(), _samples(5)
    serial = []
    for image, gt in samples:
        p = vpu_eval_predictor(net, "cpu")
        serial.append(evaluate_sample(image, gt, p, 0.55, max_clicks=6)[1])
    stats = {}
    lock = evaluate_lockstep(samples, net, "cpu", 0.55, max_clicks=6, micro_batch=3, stats=stats)
    assert len({len(s) for s in serial}) > 1                      # some sessions stop early, others do not
    for a, b in zip(serial, lock):
        assert np.array_equal(a, b)
    assert stats["click_forwards"] == 2 * sum(len(s) for s in serial)      # flip TTA: 2 click-forwards per click
    assert stats["network_calls"] < sum(len(s) for s in serial)


def test_noc_metric_and_table():
    ious = [np.array([0.5, 0.86, 0.91]), np.array([0.2, 0.3]), np.array([0.95])]
    noc, std, over = compute_noc_metric(ious, [0.85, 0.9], max_clicks=20)
    assert noc == [np.mean([2, 20, 1]), np.mean([3, 20, 1])] and over == [1, 1]
    t = iou_table(ious, 4)
    assert t.shape == (3, 4) and np.isnan(t[1, 2]) and t[0, 2] == np.float32(0.91)


@pytest.mark.reference
def test_clicker_matches_reference():
    from oracle import ref_harness as rh
    rh.import_reference()
    from isegm.inference.clicker import Clicker as RefClicker
    rs = np.random.RandomState(0)
    for trial in range(5):
        gt = _samples(1, seed0=trial)[0][1]
        mine, ref = Clicker(gt_mask=gt), RefClicker(gt_mask=gt)
        pred = np.zeros_like(gt, dtype=bool)
        for k in range(6):
            mine.make_next_click(pred)
            ref.make_next_click(pred)
            a, b = mine.clicks_list[-1], ref.clicks_list[-1]
            assert (a.is_positive, tuple(int(v) for v in a.coords), a.indx) == (bool(b.is_positive), tuple(int(v) for v in b.coords), b.indx)
            yy, xx = np.mgrid[0:448, 0:448]          # grow / shrink a blob around the click to create new errors
            blob = (yy - a.coords[0]) ** 2 + (xx - a.coords[1]) ** 2 < rs.randint(20, 90) ** 2
            pred = (pred | blob) if a.is_positive else (pred & ~blob)


@pytest.mark.reference
def test_noc_loop_matches_reference_plumbing():
    """The reference's BasePredictor / ZoomIn / AddHorizontalFlip / evaluate_sample (imported read-only) and this
    repo's, around the same stand-in network: identical clicks, IoU sequences and final probability maps."""
    import random
    from oracle import ref_harness as rh
    rh.import_reference()
    from isegm.inference.predictors import get_predictor as ref_get_predictor
    from isegm.inference.vpu_evaluation import evaluate_sample as ref_evaluate_sample
    from isegm.inference import utils as ref_utils
    net = FakeNet()
    for image, gt in _samples(3, seed0=10):
        random.seed(0)
        np.random.seed(0)
        rp = ref_get_predictor(net, "NoBRS", "cpu", with_flip=True, zoom_in_params={"skip_clicks": -1, "target_size": (448, 448)},
                               predictor_params={"cascade_step": 1, "cascade_adaptive": False, "cascade_clicks": 1})
        rclicks, rious, rprobs = ref_evaluate_sample(image, gt, rp, 0.9, max_clicks=5)
        mclicks, mious, mprobs = evaluate_sample(image, gt, vpu_eval_predictor(net, "cpu"), 0.9, max_clicks=5)
        assert [(bool(c.is_positive), int(c.coords[0]), int(c.coords[1])) for c in rclicks] == \
               [(bool(c.is_positive), int(c.coords[0]), int(c.coords[1])) for c in mclicks]
        assert np.array_equal(rious, mious)
        assert np.array_equal(rprobs, mprobs)
    ious = [np.array([0.5, 0.86, 0.91]), np.array([0.2, 0.3]), np.array([0.95])]
    assert ref_utils.compute_noc_metric(ious, [0.8, 0.85, 0.9], max_clicks=20)[0] == compute_noc_metric(ious, [0.8, 0.85, 0.9], 20)[0]


class PromptNet(FakeNet):
    """FakeNet + a term that depends on the box / scribble prompt, so a wrong prompt changes the masks and the clicks."""

    def forward(self, image, points, prompts=None, as_prompt_type=0):
        out = super().forward(image, points)["instances"]
        if as_prompt_type == 1:
            boxes = prompts[1].double()
            out = out + (0.4 * (boxes[:, :4].sum(dim=1) % 97) / 97).view(-1, 1, 1, 1).float()
        elif as_prompt_type == 2:
            scr = torch.as_tensor(np.asarray(prompts[2][0], dtype=np.float64))
            out = out + (0.4 * (scr.sum(dim=(1, 2, 3)) % 97) / 97).view(-1, 1, 1, 1).float()
        return {"instances": out, "instances_aux": None}


@pytest.mark.reference
@pytest.mark.parametrize("as_prompt_type", [1, 2])
def test_noc_loop_with_simulated_box_and_scribble_prompts_matches_reference(as_prompt_type):
    """Box / scribble prompts rebuilt at every click by the restated simulators (inference/prompts.py) inside this repo's
    predictor against the reference's predictor calling its own get_next_promts: same random seeds, same clicks and IoUs."""
    import random
    from oracle import ref_harness as rh
    rh.import_reference()
    from isegm.inference.predictors import get_predictor as ref_get_predictor
    from isegm.inference.clicker import Clicker as RefClicker
    from isegm.inference.utils import get_iou as ref_get_iou

    def ref_evaluate_sample(image, gt_mask, predictor, max_iou_thr, max_clicks, as_prompt_type):
        # isegm/inference/vpu_evaluation.py:35-98 with its hard-coded `as_prompt_type = 0` (:48) turned into the argument
        clicker, pred_mask, ious = RefClicker(gt_mask=gt_mask), np.zeros_like(gt_mask), []
        with torch.no_grad():
            predictor.set_input_image(image)
            for click_indx in range(max_clicks):
                clicker.make_next_click(pred_mask)
                pred_probs, _ = predictor.get_vqu_prediction(clicker, gt_mask=gt_mask, as_prompt_type=as_prompt_type,
                                                             click_indx=click_indx, as_multi_prompts=True)
                pred_mask = pred_probs > 0.49
                ious.append(ref_get_iou(gt_mask, pred_mask))
                if ious[-1] >= max_iou_thr:
                    break
        return clicker.clicks_list, np.array(ious, dtype=np.float32), pred_probs

    net = PromptNet()
    for image, gt in _samples(2, seed0=20):
        out = []
        for impl in ("ref", "here"):
            random.seed(5)
            np.random.seed(5)
            if impl == "ref":
                pred = ref_get_predictor(net, "NoBRS", "cpu", with_flip=True, zoom_in_params={"skip_clicks": -1, "target_size": (448, 448)},
                                         predictor_params={"cascade_step": 1, "cascade_adaptive": False, "cascade_clicks": 1})
                clicks, ious, probs = ref_evaluate_sample(image, gt, pred, 0.95, max_clicks=4, as_prompt_type=as_prompt_type)
            else:
                clicks, ious, probs = evaluate_sample(image, gt, vpu_eval_predictor(net, "cpu"), 0.95, max_clicks=4,
                                                      as_prompt_type=as_prompt_type)
            out.append(([(bool(c.is_positive), int(c.coords[0]), int(c.coords[1])) for c in clicks], ious, probs))
        assert out[0][0] == out[1][0]
        assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])


def _worker(rank, world, port, n_images, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    table, _, stats = evaluate_sharded(SyntheticEllipseDataset(n_images), FakeNet(), "cpu", rank, world, 0.55, max_clicks=4,
                                       micro_batch=2)
    out_q.put((rank, table, stats["click_forwards"]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_loop_world_size_2_gloo():
    n_images, world = 5, 2                                   # uneven shards: 3 + 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    serial = iou_table(evaluate_lockstep(_samples(n_images), FakeNet(), "cpu", 0.55, max_clicks=4, micro_batch=8), 4)
    for rank, table, _ in got:
        assert table.shape == (n_images, 4)
        assert np.array_equal(np.nan_to_num(table, nan=-1), np.nan_to_num(serial, nan=-1))      # every rank holds the full table
    assert sum(f for _, _, f in got) == 2 * int(np.isfinite(serial).sum())


class _MultiObjectDataset:
    """Image i holds 1 + (i % 3) objects (labels 1..k): rows of the IoU table are (image, object) pairs, not images."""

    def __init__(self, n):
        self.base = SyntheticEllipseDataset(n)

    def __len__(self):
        return len(self.base)

    def get_sample(self, index):
        s = self.base.get_sample(index)
        k = 1 + index % 3
        m = np.zeros_like(s._mask)
        H = m.shape[0]
        for o in range(k):                                 # k horizontal bands of the ellipse -> k objects
            band = slice(o * H // k, (o + 1) * H // k)
            m[band][s._mask[band] == 1] = o + 1
        s._mask = m
        s.objects_ids = [o + 1 for o in range(k) if (m == o + 1).sum() > 0]
        return s


def _worker_multi(rank, world, port, n_images, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    table, _, _ = evaluate_sharded(_MultiObjectDataset(n_images), FakeNet(), "cpu", rank, world, 0.55, max_clicks=3, micro_batch=2)
    out_q.put((rank, table))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_loop_multi_object_rows_world_size_2_gloo():
    """Ranks hold different numbers of (image, object) rows (here 3 + 2 images -> 6 + 3 rows... whatever the dataset gives): the
    gathered table must have one row per pair, in dataset order, on every rank."""
    n_images, world = 5, 2
    ds = _MultiObjectDataset(n_images)
    samples = []
    for i in range(n_images):
        s = ds.get_sample(i)
        samples += [(s.image, s.gt_mask(o)) for o in s.objects_ids]
    assert len(samples) > n_images
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_multi, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    serial = iou_table(evaluate_lockstep(samples, FakeNet(), "cpu", 0.55, max_clicks=3, micro_batch=8), 3)
    for _, table in got:
        assert table.shape == serial.shape
        assert np.array_equal(np.nan_to_num(table, nan=-1), np.nan_to_num(serial, nan=-1))


def test_gt_labels_for_device_clicker():
    from pvpuformer_b200.inference.evaluation import gt_labels_int8
    g = np.array([[0, 1, 255, -1, 2]], dtype=np.int32)
    assert gt_labels_int8([g]).tolist() == [[[0, 1, 0, -1, 0]]]


def test_lockstep_refuses_recalculating_zoom_in():
    from pvpuformer_b200.inference.predictor import get_predictor
    factory = lambda net, dev: get_predictor(net, "NoBRS", dev, zoom_in_params={"skip_clicks": 1, "target_size": (448, 448)})
    with pytest.raises(NotImplementedError):
        evaluate_lockstep(_samples(1), FakeNet(), "cpu", 0.9, max_clicks=2, predictor_factory=factory)


def test_gather_without_process_group_is_identity():
    t = np.zeros((3, 4), np.float32)
    assert gather_iou_tables(t, 3) is t
