"""NoC evaluation loop (reference isegm/inference/vpu_evaluation.py:18-98, metrics of inference/utils.py:80-110)
plus what the reference does not have: lock-step batched evaluation and sharding over ranks.

The click loop is sequential only WITHIN an image (the next click depends on the last mask); images and click
sessions are independent (nothing in the forward mixes batch elements).  `evaluate_lockstep` therefore advances
many sessions one click at a time and feeds all of them to ONE network call (model batch = 2 x sessions with flip
TTA), and `evaluate_sharded` gives every rank a contiguous block of images, with no collective in the loop and a
single all_gather of the per-image IoU table at the end (NCCL on GPUs, gloo in the CPU tests).
"""
from time import time

import numpy as np
import torch

from .clicker import Clicker
from .predictor import vpu_eval_predictor


def get_iou(gt_mask, pred_mask, ignore_label=-1):
    keep = gt_mask != ignore_label
    obj = gt_mask == 1
    inter = np.logical_and(np.logical_and(pred_mask, obj), keep).sum()
    union = np.logical_and(np.logical_or(pred_mask, obj), keep).sum()
    return inter / union


def compute_noc_metric(all_ious, iou_thrs, max_clicks=20):
    """-> (mean NoC, std NoC, number of objects that needed max_clicks) per IoU threshold."""
    noc_list, noc_std, over_max = [], [], []
    for thr in iou_thrs:
        scores = []
        for ious in all_ious:
            hit = np.asarray(ious) >= thr
            scores.append(int(np.argmax(hit)) + 1 if hit.any() else max_clicks)
        scores = np.array(scores, dtype=np.int64)
        noc_list.append(scores.mean())
        noc_std.append(scores.std())
        over_max.append(int((scores == max_clicks).sum()))
    return noc_list, noc_std, over_max


def evaluate_sample(image, gt_mask, predictor, max_iou_thr, pred_thr=0.49, min_clicks=1, max_clicks=20, sample_id=None,
                    callback=None, as_prompt_type=0):
    clicker = Clicker(gt_mask=gt_mask)
    pred_mask = np.zeros_like(gt_mask)
    ious = []
    pred_probs = None
    with torch.no_grad():
        predictor.set_input_image(image)
        for click_indx in range(max_clicks):
            clicker.make_next_click(pred_mask)
            pred_probs, prompts = predictor.get_vqu_prediction(clicker, gt_mask=gt_mask, as_prompt_type=as_prompt_type,
                                                               click_indx=click_indx, as_multi_prompts=True)
            pred_mask = pred_probs > pred_thr
            iou = get_iou(gt_mask, pred_mask)
            ious.append(iou)
            done = iou >= max_iou_thr and click_indx + 1 >= min_clicks
            if callback is not None:
                callback(image, gt_mask, pred_probs, iou, sample_id, click_indx, clicker.clicks_list, done,
                         predictor.zoom_in, prompts, as_prompt_type)
            if done:
                break
    return clicker.clicks_list, np.array(ious, dtype=np.float32), pred_probs


def evaluate_dataset(dataset, predictor, **kwargs):
    all_ious = []
    t0 = time()
    for index in range(len(dataset)):
        sample = dataset.get_sample(index)
        for object_id in sample.objects_ids:
            _, ious, _ = evaluate_sample(sample.image, sample.gt_mask(object_id), predictor, sample_id=index, **kwargs)
            all_ious.append(ious)
    return all_ious, time() - t0


# ---------------------------------------------------------------------------------------------------------
# lock-step batched evaluation
# ---------------------------------------------------------------------------------------------------------
def _repack_points(points_nd, n):
    """[r, 2k, 3] -> [r, 2n, 3]: each half padded with (-1,-1,-1) rows (the network pads to num_max_points anyway)."""
    k = points_nd.shape[1] // 2
    if k == n:
        return points_nd
    pad = points_nd.new_full((points_nd.shape[0], n - k, 3), -1)
    return torch.cat([points_nd[:, :k], pad, points_nd[:, k:], pad], dim=1)


def evaluate_lockstep(samples, net, device, max_iou_thr, pred_thr=0.49, min_clicks=1, max_clicks=20, micro_batch=32,
                      predictor_factory=vpu_eval_predictor, stats=None, device_clicker=False, device_session=False):
    """samples: list of (image HWC, gt_mask HW).  Returns the list of per-sample IoU arrays, identical to running
    evaluate_sample on each (the forward is batch-independent), but with ONE network call per click per micro-batch.
    `stats` (dict, optional) receives the number of network calls and click-forwards executed.

    device_clicker=True keeps the ground truth, the thresholded predictions and the not-clicked maps of the micro-batch on
    the device and gets the next clicks and the IoU counts of ALL its sessions from one call of the CUDA clicker
    (csrc/noc.cu, bit-exact with the cv2 distance-transform clicker): per click the host reads back 16 + 16 bytes per
    session instead of a full-resolution probability map and runs no distance transform.  All images of a micro-batch
    must then have the same size.

    device_session=True additionally keeps the images, the previous probabilities, the click lists and the zoom-in regions
    on the device and runs the predictor transforms either side of the network as CUDA kernels (csrc/session.cu,
    inference/device_session.py): no per-session host work and no read-back inside the click loop."""
    if device_session:
        if predictor_factory is not vpu_eval_predictor:
            raise ValueError("device sessions implement the vpu_eval_predictor configuration only")
        from .device_session import evaluate_device_sessions
        return evaluate_device_sessions(samples, net, device, max_iou_thr, pred_thr=pred_thr, min_clicks=min_clicks,
                                        max_clicks=max_clicks, micro_batch=micro_batch, stats=stats)
    results = [None] * len(samples)
    n_calls = n_fwd = 0
    with torch.no_grad():
        for lo in range(0, len(samples), micro_batch):
            chunk = list(range(lo, min(lo + micro_batch, len(samples))))
            sess = {}
            for i in chunk:
                image, gt = samples[i]
                p = predictor_factory(net, device)
                if p.cascade_step > 1 or p.cascade_adaptive:
                    raise NotImplementedError("lock-step evaluation supports cascade_step <= 1, cascade_adaptive=False")
                if p.zoom_in is not None and p.zoom_in.skip_clicks >= 0:
                    # the serial path re-runs a click when ZoomIn.check_possible_recalculation() fires (base.py:185-186); that
                    # only happens while _object_roi is unset, i.e. with skip_clicks >= 0 -- not reproduced in lock-step
                    raise NotImplementedError("lock-step evaluation needs ZoomIn(skip_clicks=-1) (the VPU evaluation's setting) "
                                              "or no zoom-in; got skip_clicks=%d" % p.zoom_in.skip_clicks)
                p.set_input_image(image)
                sess[i] = dict(pred=p, clicker=Clicker(gt_mask=gt), gt=gt, mask=np.zeros_like(gt), ious=[])
            dc = _DeviceClickerBatch([sess[i]["gt"] for i in chunk], device) if device_clicker else None
            slot = {i: k for k, i in enumerate(chunk)}
            if dc is not None:
                dc.step(None)                                   # first clicks: prediction = all background
            active = list(chunk)
            for click_indx in range(max_clicks):
                if not active:
                    break
                prepared = []
                for i in active:
                    s = sess[i]
                    if dc is None:
                        s["clicker"].make_next_click(s["mask"])
                    else:
                        s["clicker"].add_click(dc.click(slot[i]))
                    image_nd, points_nd, _ = s["pred"].prepare_inputs(s["clicker"], None, s["gt"], 0)
                    prepared.append((image_nd, points_nd))
                n = max(p.shape[1] // 2 for _, p in prepared)
                images = torch.cat([im for im, _ in prepared], dim=0)
                points = torch.cat([_repack_points(p.to(torch.float64), n) for _, p in prepared], dim=0)
                logits = net(images, points)["instances"]
                n_calls += 1
                n_fwd += images.shape[0]
                still, row = [], 0
                preds = {}
                for i, (image_nd, _) in zip(active, prepared):
                    s = sess[i]
                    r = image_nd.shape[0]
                    pred = s["pred"].finish_prediction(logits[row:row + r], image_nd.shape[2:])
                    row += r
                    s["pred"].prev_prediction = pred
                    if dc is None:
                        probs = pred.cpu().numpy()[0, 0]
                        s["mask"] = probs > pred_thr
                    else:
                        preds[slot[i]] = pred[0, 0] > pred_thr
                if dc is not None:
                    dc.step(preds)                              # IoU of this click + the next click, all sessions at once
                for i in active:
                    s = sess[i]
                    iou = dc.iou(slot[i]) if dc is not None else get_iou(s["gt"], s["mask"])
                    s["ious"].append(iou)
                    if not (iou >= max_iou_thr and click_indx + 1 >= min_clicks):
                        still.append(i)
                active = still
            for i in chunk:
                results[i] = np.array(sess[i]["ious"], dtype=np.float32)
    if stats is not None:
        stats["network_calls"] = stats.get("network_calls", 0) + n_calls
        stats["click_forwards"] = stats.get("click_forwards", 0) + n_fwd
    return results


def gt_labels_int8(gts):
    """Ground-truth masks -> int8 [S,H,W] for the device clicker, whose contract is 1 = object, -1 = ignore, anything else =
    background (clicker.py:31-32, utils.py:80-87).  A plain astype(int8) would wrap 255 to -1 and silently turn an ordinary
    label into `ignore`, so labels outside {-1, 0, 1} are mapped to 0 explicitly."""
    a = np.stack([np.asarray(g) for g in gts])
    out = np.zeros(a.shape, dtype=np.int8)
    out[a == 1] = 1
    out[a == -1] = -1
    return out


class _DeviceClickerBatch:
    """Ground truth / prediction / not-clicked maps of one micro-batch on the device + one clicker call per click."""

    def __init__(self, gts, device):
        from .. import ops
        self.ops = ops
        shapes = {g.shape for g in gts}
        if len(shapes) != 1:
            raise ValueError("device_clicker needs equally sized images inside a micro-batch, got %s" % sorted(shapes))
        if gts[0].shape[0] * gts[0].shape[1] < 20000:
            raise ValueError("device_clicker is bit-exact with the cv2 clicker only for masks of >= 2e4 pixels (csrc/noc.cu); "
                             "use device_clicker=False for %s" % (gts[0].shape,))
        self.gt = torch.from_numpy(gt_labels_int8(gts)).to(device)
        self.pred = torch.zeros(self.gt.shape, dtype=torch.uint8, device=device)
        self.not_clicked = torch.ones(self.gt.shape, dtype=torch.uint8, device=device)
        self.workspace = None
        self.clicks = self.counts = None

    def step(self, preds):
        """preds: {slot: bool [H,W] device tensor} of the sessions that ran this click (others keep their last mask)."""
        if preds:
            for k, p in preds.items():
                self.pred[k] = p
        clicks, counts = self.ops.noc_next_clicks(self.gt, self.pred, self.not_clicked)
        self.clicks, self.counts = clicks.cpu().numpy(), counts.cpu().numpy()       # 32 bytes per session

    def click(self, k):
        from .clicker import Click
        c = self.clicks[k]
        # numpy int64 coordinates, as np.where gives them to the reference's clicker (clicker.py:55-69): the ZoomIn-rescaled
        # coordinates then are numpy float64 scalars and get_points_nd builds a float64 tensor, not a float32 one
        return Click(is_positive=bool(c[0]), coords=(np.int64(c[1]), np.int64(c[2])))

    def iou(self, k):
        return self.counts[k, 0] / self.counts[k, 1]           # numpy int64 / int64 -> float64, as get_iou


# ---------------------------------------------------------------------------------------------------------
# sharding over ranks
# ---------------------------------------------------------------------------------------------------------
def shard_range(n, rank, world):
    """Contiguous block of rank `rank`: [start, stop); sizes differ by at most one."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def iou_table(all_ious, max_clicks):
    """list of variable-length IoU arrays -> [n, max_clicks] fp32, NaN after an early stop."""
    t = np.full((len(all_ious), max_clicks), np.nan, dtype=np.float32)
    for i, ious in enumerate(all_ious):
        t[i, :len(ious)] = ious
    return t


def gather_iou_tables(local_table, n_total=None, group=None, device=None):
    """all_gather of the per-rank [rows_local, max_clicks] tables -> [sum of rows, max_clicks] on every rank (rank order ==
    image order because shards are contiguous).  A row is one (image, object) pair, so ranks first exchange their row
    counts (images with several objects make them differ from the image shard sizes) and the tables are padded to the
    largest one.  The only data collective of the sharded NoC loop.  `n_total`, if given, is checked against the result."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local_table
    world = dist.get_world_size(group)
    max_clicks = local_table.shape[1]
    cnt = torch.tensor([local_table.shape[0]], dtype=torch.int64)
    if device is not None:
        cnt = cnt.to(device)
    counts = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    buf = torch.full((cap, max_clicks), float("nan"), dtype=torch.float32)
    buf[:local_table.shape[0]] = torch.from_numpy(local_table)
    if device is not None:
        buf = buf.to(device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    table = np.concatenate([out[r][:counts[r]].cpu().numpy() for r in range(world)], axis=0)
    if n_total is not None and table.shape[0] != n_total:
        raise RuntimeError("gathered %d IoU rows, expected %d" % (table.shape[0], n_total))
    return table


def evaluate_sharded(dataset, net, device, rank, world, max_iou_thr, max_clicks=20, micro_batch=32, group=None,
                     gather_device=None, **kwargs):
    """Rank `rank` evaluates images [start, stop) in lock-step micro-batches; returns the full [objects, max_clicks] IoU
    table (one row per (image, object) pair in dataset order, after the all_gather) and the local wall-clock seconds of
    the loop."""
    start, stop = shard_range(len(dataset), rank, world)
    samples = []
    for index in range(start, stop):
        s = dataset.get_sample(index)
        for object_id in s.objects_ids:
            samples.append((s.image, s.gt_mask(object_id)))
    t0 = time()
    stats = {}
    ious = evaluate_lockstep(samples, net, device, max_iou_thr, max_clicks=max_clicks, micro_batch=micro_batch, stats=stats,
                             **kwargs)
    elapsed = time() - t0
    table = gather_iou_tables(iou_table(ious, max_clicks), group=group, device=gather_device)
    return table, elapsed, stats
