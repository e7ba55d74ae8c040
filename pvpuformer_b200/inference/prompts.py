"""Prompt simulators of the NoC evaluation: next click, largest-error-region box and scribble from
the previous prediction and the ground truth.

Host-side restatement of the reference's `get_next_promts`, `cal_box`, `max_connected_regions`, `cal_scribble` and
`bezier_curve` (isegm/engine/trainer.py:703-768, 1061-1243), which `BasePredictor._get_vqu_prediction_prompts` calls before
every forward (isegm/inference/predictors/base.py:166-177).  The reference runs them even for click-only evaluation; here
`BasePredictor.prepare_inputs` calls `prompt_fn` only when `as_prompt_type != 0`.

Random streams are consumed in the reference's order -- Python's `random` (box jitter, scribble control points), then
`np.random.rand` (curve kind), then `np.random.randint` (click position) -- so that under the same seeds the outputs are
identical (tests/test_prompts_cpu.py, pinned against the unmodified reference and against committed golden vectors).

Third-party pieces the reference imports and this image lacks: `skimage.measure.label(connectivity=2)` is
`scipy.ndimage.label` with the 8-neighbour structure (both number components in raster order of their first pixel);
`bezier.Curve.evaluate_multi` (bezier 2021.2.12 in the reference's requirements) is the Bernstein form
sum_i C(n, i) s^i (1 - s)^(n - i) P_i, evaluated here directly.
"""
import random
from math import comb

import cv2
import numpy as np
import torch
from scipy import ndimage
from scipy.interpolate import make_interp_spline

_EIGHT = ndimage.generate_binary_structure(2, 2)


def max_connected_regions(mask):
    """trainer.py:1186-1201.  Keeps the component that is largest when the scan ends, plus every component holding more than
    10 % of the foreground that was met while that one already was the running maximum ... exactly as the reference's
    in-place relabelling does: component j joins the running maximum of the moment it is visited."""
    labels, num = ndimage.label(np.asarray(mask) != 0, structure=_EIGHT)
    if num == 0:
        return labels
    counts = np.bincount(labels.ravel(), minlength=num + 1)
    total = counts[1:].sum()
    target = np.arange(num + 1)
    best, best_count = 0, 0
    for j in range(1, num + 1):
        if counts[j] > best_count:
            best_count, best = counts[j], j
        if counts[j] > 0.1 * total:
            target[j] = best
    keep = target == best
    keep[0] = False
    return keep[labels].astype(np.int8)


def _first_free(points_row, lo, hi):
    """Index (into the full row) of the first slot of points_row[lo:hi] whose order field is < 0, or None."""
    free = torch.argwhere(points_row[lo:hi, 2] < 0)
    return (free[0, 0] + lo) if len(free) > 0 else None


def cal_box(gt_mask, fn_mask, fp_mask, points, as_allmask=True, jitter_box=True, set_offset=10):
    """trainer.py:1061-1133 -> int32 [B, 5] = (x_center, y_center, width, height, prompt slot); zeros when there is no region."""
    height, width = gt_mask.shape[1], gt_mask.shape[2]
    n = points.size(1) // 2
    boxes = np.zeros([len(fn_mask), 5], np.int32)
    for b in range(len(fn_mask)):
        if as_allmask:
            region = gt_mask[b]
            slot = _first_free(points[b], 0, n)
            slot = n - 1 if slot is None else slot
        elif np.sum(fn_mask[b]) > np.sum(fp_mask[b]):
            region = max_connected_regions(fn_mask[b])
            slot = n - 1
        else:
            region = max_connected_regions(fp_mask[b])
            slot = _first_free(points[b], n, 2 * n)
            slot = 2 * n - 1 if slot is None else slot
        ind = np.argwhere(region == True)  # noqa: E712  (int8 regions: == 1)
        if len(ind) == 0:
            continue
        y0, y1, x0, x1 = ind[:, 0].min(), ind[:, 0].max(), ind[:, 1].min(), ind[:, 1].max()
        if jitter_box:
            bx = np.minimum(np.maximum(x0 + random.randint(-set_offset, 0), 0), width - set_offset)
            ex = np.maximum(np.minimum(x1 + random.randint(0, set_offset), width), bx + set_offset)
            by = np.minimum(np.maximum(y0 + random.randint(-set_offset, 0), 0), height - set_offset)
            ey = np.maximum(np.minimum(y1 + random.randint(0, set_offset), height), by + set_offset)
            y0, y1, x0, x1 = by, ey, bx, ex
        xc, yc, bw, bh = int(0.5 * (x0 + x1)), int(0.5 * (y0 + y1)), int(x1 - x0), int(y1 - y0)
        if min(xc, yc, bw, bh) >= 1:
            boxes[b] = [xc, yc, bw, bh, slot.cpu().numpy() if isinstance(slot, torch.Tensor) else slot]
    return boxes


def _bernstein(nodes, s):
    """nodes [2, P] -> [2, len(s)]: the Bezier curve of degree P - 1 through the control points."""
    n = nodes.shape[1] - 1
    out = np.zeros((nodes.shape[0], s.shape[0]))
    for i in range(n + 1):
        out += np.outer(nodes[:, i], comb(n, i) * (s ** i) * ((1 - s) ** (n - i)))
    return out


def bezier_curve(points, bbox=None, num_samples=100, as_inline=False):
    """trainer.py:1137-1184: a Bezier curve through the control points, or (as_inline False) a cubic interpolating spline of
    column over row when scipy accepts the points (strictly increasing rows, >= 4 of them), else the Bezier curve again."""
    def clipped(a, b):
        return np.column_stack((np.clip(a, bbox[0], bbox[2]).astype(int), np.clip(b, bbox[1], bbox[3]).astype(int)))

    def bez():
        data = _bernstein(np.asarray(points, dtype=np.float64).transpose((1, 0)), np.linspace(0.0, 1.0, num_samples))
        return clipped(data[0], data[1])

    if as_inline:
        return bez()
    try:
        x, y = points[:, 0], points[:, 1]
        spline = make_interp_spline(x, y)
        x_new = np.linspace(x.min(), x.max(), num_samples)
        return clipped(x_new, spline(x_new))
    except Exception:
        return bez()


def cal_scribble(gt_mask, min_p=3, max_p=10, num_samples=1000):
    """trainer.py:1203-1257 -> [scribbles [B,1,num_samples,2] (col, row), rectangles [B,1,4] (col centre, row centre, col
    extent, row extent)]: control points are drawn column-band by column-band inside the object's main region."""
    all_s, all_r = [], []
    for i in range(len(gt_mask)):
        scribble, rect = np.zeros([num_samples, 2]), np.array([[0, 0, 0, 0]])
        if np.sum(gt_mask[i]) > 0:
            ind = np.argwhere(max_connected_regions(gt_mask[i]) == True)  # noqa: E712
            num_p = random.randint(min_p, max_p)
            r0, r1, c0, c1 = ind[:, 0].min(), ind[:, 0].max(), ind[:, 1].min(), ind[:, 1].max()
            ext_r, ext_c = int(r1 - r0), int(c1 - c0)
            value, gap = r0, ext_r // num_p
            ctrl = []
            for _ in range(num_p):
                row = random.randint(value, value + gap - 1) if gap > 0 else random.randint(value, value + gap)
                cand = ind[ind[:, 0] == row]
                if cand.shape[0] > 0:
                    ctrl.append(cand[random.randint(0, cand.shape[0] - 1)])
                value += gap
            ctrl = np.array(ctrl)
            if len(ctrl) > 0:
                as_inline = np.random.rand() > 0.5
                scribble = bezier_curve(ctrl, [r0, c0, r1, c1], num_samples, as_inline=as_inline)[:, ::-1]
                rect = np.array([[int(0.5 * (c0 + c1)), int(0.5 * (r0 + r1)), ext_c, ext_r]])
        all_s.append(np.expand_dims(scribble, 0))
        all_r.append(rect)
    return [np.expand_dims(np.concatenate(all_s, 0), 1), np.array(all_r)]


def get_next_promts(pred, gt, points, pred_thresh=0.49, as_allmask=False, jitter_box=True):
    """trainer.py:703-768 (without the training-only ed_mask_label branch) -> (points with one simulated click added,
    boxes int32 tensor [B,5] on points.device, [scribbles, rectangles])."""
    if isinstance(gt, torch.Tensor):
        gt = gt.cpu().numpy()[:, 0, :, :] > 0.5
    elif len(gt) != len(pred):
        gt = np.expand_dims(gt, axis=0) > 0.5
    else:
        gt = gt[:, 0, :, :] > 0.5
    pred = pred.detach().cpu().numpy()[:, 0, :, :]
    fn = np.logical_and(gt, pred < pred_thresh)
    fp = np.logical_and(np.logical_not(gt), pred > pred_thresh)
    boxes = torch.from_numpy(cal_box(gt, fn, fp, points, as_allmask=as_allmask, jitter_box=jitter_box)).to(points.device)
    scribbles = cal_scribble(gt, min_p=3, max_p=10, num_samples=1000)

    fn = np.pad(fn, ((0, 0), (1, 1), (1, 1)), "constant").astype(np.uint8)
    fp = np.pad(fp, ((0, 0), (1, 1), (1, 1)), "constant").astype(np.uint8)
    n = points.size(1) // 2
    points = points.clone()
    for b in range(fn.shape[0]):
        fn_dt = cv2.distanceTransform(fn[b], cv2.DIST_L2, 5)[1:-1, 1:-1]
        fp_dt = cv2.distanceTransform(fp[b], cv2.DIST_L2, 5)[1:-1, 1:-1]
        fn_max, fp_max = np.max(fn_dt), np.max(fp_dt)
        positive = fn_max > fp_max
        dt = fn_dt if positive else fp_dt
        inside = np.argwhere(dt > max(fn_max, fp_max) / 2.0)
        if len(inside) == 0:
            continue
        coords = inside[np.random.randint(0, len(inside))]
        order = max(points[b, :, 2].max(), 0) + 1
        slot = _first_free(points[b], 0, n) if positive else _first_free(points[b], n, 2 * n)
        if slot is None:
            slot = n - 1 if positive else 2 * n - 1
        points[b, slot, 0] = float(coords[0])
        points[b, slot, 1] = float(coords[1])
        points[b, slot, 2] = float(order)
    return points, boxes, scribbles


def eval_prompt_fn(prev_mask, gt_mask, points_nd):
    """What BasePredictor._get_vqu_prediction_prompts passes to the network (base.py:176): prompts for as_prompt_type 1 / 2."""
    return get_next_promts(prev_mask, gt_mask, points_nd, as_allmask=False, jitter_box=False)
