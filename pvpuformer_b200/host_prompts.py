"""Host-side halves of the prompt encoders that the reference itself runs on the host.

* scribble column/row selection: the reference draws one scribble sample per column and per row
  with Python's global `random` (reference isegm/model/ops.py:271-295), so reproducing its output
  for a given `random.seed` means consuming `random.randint` in exactly its order on the host.
  Only the integer selection happens here; the Gaussian values are evaluated on the device.
* `raster_planes`: the reference's `cv2.rectangle` / `cv2.polylines` calls (isegm/model/is_model.py:97-146), kept as the
  checker of the device rasteriser (csrc/raster.cu) in tests; the forward does not call it.
"""
import random

import numpy as np

INT_MIN = np.iinfo(np.int32).min


def scribble_select(scribble, rect, size=448, rng=random):
    """-> sel[2, size] int32: offsets (coordinate - box origin) or INT_MIN where nothing is written."""
    sel = np.full((2, size), INT_MIN, np.int32)
    pts = np.asarray(scribble).astype(np.int32)
    rect = np.asarray(rect)
    if int(pts.sum()) + int(rect.sum()) == 0:                      # ops.py:249
        return sel
    x0, y0, w0, h0 = (min(int(v), size) for v in rect)               # ops.py:261-265
    origin_w, origin_h = x0 - w0 // 2, y0 - h0 // 2
    for xi in range(w0):                                             # ops.py:271-285
        hits = np.count_nonzero(pts[:, 0] == xi)
        if hits:
            px, py = (int(v) for v in pts[rng.randint(0, hits - 1)])  # index into the full array (ref. quirk)
            sel[0, xi] = py - origin_h
            pts = pts[~((pts[:, 0] == px) & (pts[:, 1] == py))]
    for yj in range(h0):                                             # ops.py:287-294
        hits = np.count_nonzero(pts[:, 1] == yj)
        if hits:
            px = int(pts[rng.randint(0, hits - 1)][0])
            sel[1, yj] = px - origin_w
    return sel


def scribble_slots(ppue_points_cpu, n):
    """Row that receives the scribble vector: the last positive slot whose order != -1
    (reference is_vpu_model.py:329,336-338), or -1."""
    labels = np.asarray(ppue_points_cpu)[:, :n, 2]
    out = np.full(labels.shape[0], -1, np.int32)
    for b in range(labels.shape[0]):
        valid = np.nonzero(labels[b] != -1)[0]
        if len(valid):
            out[b] = valid[-1]
    return out


def raster_planes(as_prompt_type, boxes_cpu, scribbles, n, B, size=448):
    """-> uint8 [B, 2, size, size] planes (1 where the box outline / scribble is drawn)."""
    import cv2
    planes = np.zeros((B, 2, size, size), np.uint8)
    for b in range(B):
        if as_prompt_type == 1:
            xc, yc, w, h, slot = (int(v) for v in boxes_cpu[b])
            ch = 0 if slot < n else 1                                # is_model.py:101-104
            x0, x1, y0, y1 = xc - w // 2, xc + w // 2, yc - h // 2, yc + h // 2
            cv2.rectangle(planes[b, ch], (x0, y0), (x1, y1), 1, 3)
        else:
            s = np.asarray(scribbles[b][0])
            curve = np.column_stack((s[:, 0].astype(np.int32), s[:, 1].astype(np.int32)))
            cv2.polylines(planes[b, 0], [curve], False, 1, 3)
    return planes
