"""Test-time transforms of the NoBRS predictor (reference isegm/inference/transforms/{base,zoom_in,flip}.py,
bbox helpers of isegm/utils/misc.py:36-79).  Each transform maps (image_nd, clicks_lists) forward and the
probability map backward; ZoomIn keeps the region of interest and the last full-size probabilities between clicks."""
import numpy as np
import torch
import torch.nn.functional as F


# ---- bounding-box helpers (rmin, rmax, cmin, cmax; inclusive) ---------------------------------------------
def get_bbox_from_mask(mask):
    rows = np.flatnonzero(mask.any(axis=1))
    cols = np.flatnonzero(mask.any(axis=0))
    return rows[0], rows[-1], cols[0], cols[-1]


def expand_bbox(bbox, expand_ratio, min_crop_size=None):
    rmin, rmax, cmin, cmax = bbox
    rc, cc = 0.5 * (rmin + rmax), 0.5 * (cmin + cmax)
    height, width = expand_ratio * (rmax - rmin + 1), expand_ratio * (cmax - cmin + 1)
    if min_crop_size is not None:
        height, width = max(height, min_crop_size), max(width, min_crop_size)
    return (int(round(rc - 0.5 * height)), int(round(rc + 0.5 * height)),
            int(round(cc - 0.5 * width)), int(round(cc + 0.5 * width)))


def clamp_bbox(bbox, rmin, rmax, cmin, cmax):
    return max(rmin, bbox[0]), min(rmax, bbox[1]), max(cmin, bbox[2]), min(cmax, bbox[3])


def _segment_iou(s1, s2):
    (a, b), (c, d) = s1, s2
    return max(0, min(b, d) - max(a, c) + 1) / max(1e-6, max(b, d) - min(a, c) + 1)


def get_bbox_iou(b1, b2):
    return _segment_iou(b1[:2], b2[:2]) * _segment_iou(b1[2:4], b2[2:4])


def get_object_roi(pred_mask, clicks_list, expansion_ratio, min_crop_size):
    pred_mask = pred_mask.copy()
    for click in clicks_list:
        if click.is_positive:
            pred_mask[int(click.coords[0]), int(click.coords[1])] = 1
    bbox = expand_bbox(get_bbox_from_mask(pred_mask), expansion_ratio, min_crop_size)
    return clamp_bbox(bbox, 0, pred_mask.shape[0] - 1, 0, pred_mask.shape[1] - 1)


def get_roi_image_nd(image_nd, object_roi, target_size):
    rmin, rmax, cmin, cmax = object_roi
    height, width = rmax - rmin + 1, cmax - cmin + 1
    if isinstance(target_size, tuple):
        new_height, new_width = target_size
    else:
        scale = target_size / max(height, width)
        new_height, new_width = int(round(height * scale)), int(round(width * scale))
    with torch.no_grad():
        roi = image_nd[:, :, rmin:rmax + 1, cmin:cmax + 1]
        return F.interpolate(roi, size=(new_height, new_width), mode="bilinear", align_corners=True)


def check_object_roi(object_roi, clicks_list):
    for click in clicks_list:
        if click.is_positive:
            if not (object_roi[0] <= click.coords[0] < object_roi[1]):
                return False
            if not (object_roi[2] <= click.coords[1] < object_roi[3]):
                return False
    return True


# ---- transforms ---------------------------------------------------------------------------------------
class BaseTransform:
    def __init__(self):
        self.image_changed = False

    def transform(self, image_nd, clicks_lists):
        raise NotImplementedError

    def inv_transform(self, prob_map):
        raise NotImplementedError

    def reset(self):
        pass

    def get_state(self):
        return None

    def set_state(self, state):
        pass


class SigmoidForPred(BaseTransform):
    def transform(self, image_nd, clicks_lists):
        return image_nd, clicks_lists

    def inv_transform(self, prob_map):
        return torch.sigmoid(prob_map)


class AddHorizontalFlip(BaseTransform):
    """batch 1 -> 2 (image + mirrored image, clicks mirrored); probabilities are averaged back."""

    def transform(self, image_nd, clicks_lists):
        assert image_nd.dim() == 4
        image_nd = torch.cat([image_nd, torch.flip(image_nd, dims=[3])], dim=0)
        width = image_nd.shape[3]
        flipped = [[c.copy(coords=(c.coords[0], width - c.coords[1] - 1)) for c in cl] for cl in clicks_lists]
        return image_nd, clicks_lists + flipped

    def inv_transform(self, prob_map):
        assert prob_map.dim() == 4 and prob_map.shape[0] % 2 == 0
        n = prob_map.shape[0] // 2
        return 0.5 * (prob_map[:n] + torch.flip(prob_map[n:], dims=[3]))


class ZoomIn(BaseTransform):
    def __init__(self, target_size=400, skip_clicks=1, expansion_ratio=1.4, min_crop_size=200, recompute_thresh_iou=0.5,
                 prob_thresh=0.50):
        super().__init__()
        self.target_size = target_size
        self.min_crop_size = min_crop_size
        self.skip_clicks = skip_clicks
        self.expansion_ratio = expansion_ratio
        self.recompute_thresh_iou = recompute_thresh_iou
        self.prob_thresh = prob_thresh
        self.reset()

    def reset(self):
        self._input_image_shape = None
        self._object_roi = None
        self._prev_probs = None
        self._roi_image = None
        self.image_changed = False

    def transform(self, image_nd, clicks_lists):
        assert image_nd.shape[0] == 1 and len(clicks_lists) == 1
        self.image_changed = False
        clicks_list = clicks_lists[0]
        if len(clicks_list) <= self.skip_clicks:
            return image_nd, clicks_lists
        self._input_image_shape = image_nd.shape

        roi = None
        if self._prev_probs is not None:
            pred_mask = (self._prev_probs > self.prob_thresh)[0, 0]
            if pred_mask.sum() > 0:
                roi = get_object_roi(pred_mask, clicks_list, self.expansion_ratio, self.min_crop_size)
        if roi is None:
            if self.skip_clicks >= 0:
                return image_nd, clicks_lists
            roi = 0, image_nd.shape[2] - 1, 0, image_nd.shape[3] - 1

        if (self._object_roi is None or not check_object_roi(self._object_roi, clicks_list)
                or get_bbox_iou(roi, self._object_roi) < self.recompute_thresh_iou):
            self._object_roi = roi
            self.image_changed = True
        self._roi_image = get_roi_image_nd(image_nd, self._object_roi, self.target_size)
        return self._roi_image.to(image_nd.device), [self._transform_clicks(clicks_list)]

    def inv_transform(self, prob_map):
        if self._object_roi is None:
            self._prev_probs = prob_map.cpu().numpy()
            return prob_map
        assert prob_map.shape[0] == 1
        rmin, rmax, cmin, cmax = self._object_roi
        prob_map = F.interpolate(prob_map, size=(rmax - rmin + 1, cmax - cmin + 1), mode="bilinear", align_corners=True)
        if self._prev_probs is not None:
            full = torch.zeros(*self._prev_probs.shape, device=prob_map.device, dtype=prob_map.dtype)
            full[:, :, rmin:rmax + 1, cmin:cmax + 1] = prob_map
        else:
            full = prob_map
        self._prev_probs = full.cpu().numpy()
        return full

    def check_possible_recalculation(self):
        if self._prev_probs is None or self._object_roi is not None or self.skip_clicks > 0:
            return False
        pred_mask = (self._prev_probs > self.prob_thresh)[0, 0]
        if pred_mask.sum() > 0:
            roi = get_object_roi(pred_mask, [], self.expansion_ratio, self.min_crop_size)
            image_roi = (0, self._input_image_shape[2] - 1, 0, self._input_image_shape[3] - 1)
            if get_bbox_iou(roi, image_roi) < 0.50:
                return True
        return False

    def get_state(self):
        roi_image = self._roi_image.cpu() if self._roi_image is not None else None
        return self._input_image_shape, self._object_roi, self._prev_probs, roi_image, self.image_changed

    def set_state(self, state):
        self._input_image_shape, self._object_roi, self._prev_probs, self._roi_image, self.image_changed = state

    def _transform_clicks(self, clicks_list):
        if self._object_roi is None:
            return clicks_list
        rmin, rmax, cmin, cmax = self._object_roi
        crop_h, crop_w = self._roi_image.shape[2:]
        return [c.copy(coords=(crop_h * (c.coords[0] - rmin) / (rmax - rmin + 1),
                               crop_w * (c.coords[1] - cmin) / (cmax - cmin + 1))) for c in clicks_list]
