"""Oracle clicker of the NoC protocol (reference isegm/inference/clicker.py:6-118).

The next click is the interior-most pixel of the larger error region: false-negative and false-positive masks are
zero-padded by one pixel, distance-transformed (cv2 DIST_L2, mask size 0 = exact), already-clicked pixels are
excluded, the region with the larger maximum distance wins (positive click iff FN strictly larger) and ties inside
a region resolve to the first maximum in row-major order (what np.where(...)[0] gives the reference).
"""
import copy

import cv2
import numpy as np


class Click:
    def __init__(self, is_positive, coords, indx=None):
        self.is_positive = is_positive
        self.coords = coords          # (row, col)
        self.indx = indx

    @property
    def coords_and_indx(self):
        return (*self.coords, self.indx)

    def copy(self, **kwargs):
        c = copy.deepcopy(self)
        for k, v in kwargs.items():
            setattr(c, k, v)
        return c


class Clicker:
    def __init__(self, gt_mask=None, init_clicks=None, ignore_label=-1, click_indx_offset=0):
        self.click_indx_offset = click_indx_offset
        if gt_mask is not None:
            self.gt_mask = gt_mask == 1
            self.not_ignore_mask = gt_mask != ignore_label
        else:
            self.gt_mask = None
        self.reset_clicks()
        for click in (init_clicks or []):
            self.add_click(click)

    def __len__(self):
        return len(self.clicks_list)

    def reset_clicks(self):
        if self.gt_mask is not None:
            self.not_clicked_map = np.ones(self.gt_mask.shape, dtype=bool)
        self.num_pos_clicks = 0
        self.num_neg_clicks = 0
        self.clicks_list = []

    def get_clicks(self, clicks_limit=None):
        return self.clicks_list[:clicks_limit]

    def make_next_click(self, pred_mask):
        assert self.gt_mask is not None
        self.add_click(self._get_next_click(pred_mask))

    def _error_distance(self, err_mask, padding):
        m = err_mask.astype(np.uint8)
        if padding:
            m = np.pad(m, 1)
        dt = cv2.distanceTransform(m, cv2.DIST_L2, 0)
        if padding:
            dt = dt[1:-1, 1:-1]
        return dt * self.not_clicked_map

    def _get_next_click(self, pred_mask, padding=True):
        pred_mask = np.asarray(pred_mask, dtype=bool)
        fn = self.gt_mask & ~pred_mask & self.not_ignore_mask
        fp = ~self.gt_mask & pred_mask & self.not_ignore_mask
        fn_dt, fp_dt = self._error_distance(fn, padding), self._error_distance(fp, padding)
        fn_max, fp_max = fn_dt.max(), fp_dt.max()
        is_positive = bool(fn_max > fp_max)
        dt = fn_dt if is_positive else fp_dt
        y, x = np.unravel_index(np.argmax(dt), dt.shape)       # first maximum in row-major order
        return Click(is_positive=is_positive, coords=(y, x))

    def add_click(self, click):
        click.indx = self.click_indx_offset + self.num_pos_clicks + self.num_neg_clicks
        if click.is_positive:
            self.num_pos_clicks += 1
        else:
            self.num_neg_clicks += 1
        self.clicks_list.append(click)
        if self.gt_mask is not None:
            self.not_clicked_map[click.coords[0], click.coords[1]] = False

    def _remove_last_click(self):
        click = self.clicks_list.pop()
        if click.is_positive:
            self.num_pos_clicks -= 1
        else:
            self.num_neg_clicks -= 1
        if self.gt_mask is not None:
            self.not_clicked_map[click.coords[0], click.coords[1]] = True

    def get_state(self):
        return copy.deepcopy(self.clicks_list)

    def set_state(self, state):
        self.reset_clicks()
        for click in state:
            self.add_click(click)
