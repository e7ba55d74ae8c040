"""Device-resident click sessions of the NoC evaluation loop.

What the reference does on the host for every click of every image -- oracle clicker (two cv2 distance transforms),
`BasePredictor` input assembly, `ZoomIn` crop / resize / click rescaling, flip TTA, `get_points_nd`, and after the
network: flip average, sigmoid, resize back, paste, `.cpu().numpy()` of the probability map, threshold, IoU
(isegm/inference/predictors/base.py:106-213, transforms/zoom_in.py:30-112, transforms/flip.py:9-28,
inference/clicker.py:29-69, inference/utils.py:80-87) -- runs here as CUDA kernels over ALL sessions of a micro-batch
(csrc/session.cu, csrc/noc.cu) on state that never leaves the device.  Per click the host launches
clicker -> prepare -> net(image, points) -> finish and reads nothing back unless sessions can stop early
(max_iou_thr <= 1: 16 bytes per session per click).

The network is any callable `net(image [2A,4,T,T] fp32 cuda, points [2A,2n,3] float64 cuda) -> {'instances': [2A,1,T,T]}`:
the drop-in module of pvpuformer_b200.model, unchanged.  There is no host fallback: every step is a C-ABI call.
"""
import ctypes

import numpy as np
import torch

from .. import lib as L

INT32_MAX = 2 ** 31 - 1


class DeviceClickSessions:
    """State of S sessions (equal image size) + the three per-click stages."""

    def __init__(self, images, gt_masks, device, target_size=448, max_clicks=20, pred_thr=0.49, num_max_points=24,
                 expansion_ratio=1.4, min_crop_size=200, recompute_thresh_iou=0.5, zoom_prob_thresh=0.5):
        gts = [np.asarray(g) for g in gt_masks]
        shapes = {g.shape for g in gts} | {tuple(np.asarray(im).shape[:2]) for im in images}
        if len(shapes) != 1:
            raise ValueError("device sessions need equally sized images inside a micro-batch, got %s" % sorted(shapes))
        H, W = gts[0].shape
        if H * W < 20000:
            raise ValueError("the device clicker is bit-exact with the cv2 clicker only for masks of >= 2e4 pixels (csrc/noc.cu)")
        if max_clicks > num_max_points:
            raise ValueError("max_clicks %d > num_max_points %d: the reference cannot hold that many clicks of one kind "
                             "(is_vpu_model.py:218-228)" % (max_clicks, num_max_points))
        if isinstance(target_size, (tuple, list)):
            if target_size[0] != target_size[1]:
                raise NotImplementedError("square zoom-in targets only (the VPU models take 448 x 448)")
            target_size = target_size[0]
        self.lib = L.load()
        self.device = device
        self.S, self.H, self.W, self.T = len(gts), H, W, int(target_size)
        self.max_clicks = int(max_clicks)
        self.n_half = int(max_clicks)
        S = self.S
        im = np.stack([np.asarray(i) for i in images])                       # [S,H,W,3]
        t = torch.from_numpy(np.ascontiguousarray(im)).to(device)
        if im.dtype == np.uint8:
            # predictor._to_tensor (ToTensor): x / 255 as a true IEEE division, on the device (vpu_image_from_u8); torch's CUDA
            # division by a Python scalar would multiply by the reciprocal (1 ulp off the host result for some x)
            from .. import ops
            self.images = ops.image_from_u8(t)[:, :3].contiguous()
        else:
            self.images = t.permute(0, 3, 1, 2).float().contiguous()
        from .evaluation import gt_labels_int8
        self.gt = torch.from_numpy(gt_labels_int8(gts)).to(device)
        self.prev_probs = torch.zeros(S, H, W, dtype=torch.float32, device=device)
        self.pred = torch.zeros(S, H, W, dtype=torch.uint8, device=device)
        self.not_clicked = torch.ones(S, H, W, dtype=torch.uint8, device=device)
        self.clicks = torch.zeros(S, self.max_clicks, 3, dtype=torch.int32, device=device)
        self.nclicks = torch.zeros(S, dtype=torch.int32, device=device)
        self.roi = torch.full((S, 4), -1, dtype=torch.int32, device=device)
        self.fgbox = torch.tensor([[INT32_MAX, -1, INT32_MAX, -1, -1]] * S, dtype=torch.int32, device=device)
        self.next_click = torch.zeros(S, 4, dtype=torch.int32, device=device)
        self.counts = torch.zeros(self.max_clicks + 1, S, 2, dtype=torch.int64, device=device)
        self.noc_ws = torch.empty(self.lib.vpu_noc_workspace_bytes(S, H, W) + 256, dtype=torch.uint8, device=device)
        self.noc_ws = self.noc_ws[(-self.noc_ws.data_ptr()) % 256:]
        self.state = L.VpuSessionState(S, H, W, self.T, self.max_clicks, self.n_half, self.images.data_ptr(),
                                       self.prev_probs.data_ptr(), self.pred.data_ptr(), self.clicks.data_ptr(),
                                       self.nclicks.data_ptr(), self.roi.data_ptr(), self.fgbox.data_ptr(), float(pred_thr),
                                       float(zoom_prob_thresh), float(expansion_ratio), float(recompute_thresh_iou),
                                       -1 if min_crop_size is None else int(min_crop_size))
        self._have_click = False
        self.set_active(list(range(S)))

    def set_active(self, active):
        self.active = list(active)
        self.active_dev = torch.tensor(self.active, dtype=torch.int32, device=self.device)

    # ---- the stages of one click ----------------------------------------------------------------------
    def clicker_step(self, slot):
        """IoU counts of the current masks -> counts[slot]; next oracle click of every session -> next_click."""
        L.check(self.lib.vpu_noc_next_clicks(L.ptr(self.gt), L.ptr(self.pred), L.ptr(self.not_clicked), self.S, self.H, self.W,
                                             L.ptr(self.next_click), L.ptr(self.counts[slot]), L.ptr(self.noc_ws),
                                             self.noc_ws.numel(), L.current_stream()))
        self._have_click = True

    def prepare(self):
        """Append the pending clicks of the active sessions, update their zoom-in regions -> (net_image, net_points)."""
        A = len(self.active)
        net_image = torch.empty(2 * A, 4, self.T, self.T, dtype=torch.float32, device=self.device)
        net_points = torch.empty(2 * A, 2 * self.n_half, 3, dtype=torch.float64, device=self.device)
        L.check(self.lib.vpu_session_prepare(ctypes.byref(self.state), L.ptr(self.active_dev), A,
                                             L.ptr(self.next_click) if self._have_click else None, L.ptr(net_image),
                                             L.ptr(net_points), L.current_stream()))
        self._have_click = False
        return net_image, net_points

    def finish(self, logits):
        A = len(self.active)
        if tuple(logits.shape) != (2 * A, 1, self.T, self.T) or logits.dtype != torch.float32 or not logits.is_cuda:
            raise ValueError("logits must be a cuda fp32 [%d,1,%d,%d] tensor, got %s" % (2 * A, self.T, self.T, tuple(logits.shape)))
        logits = logits.contiguous()
        L.check(self.lib.vpu_session_finish(ctypes.byref(self.state), L.ptr(self.active_dev), A, L.ptr(logits), L.current_stream()))
        self._keep = logits

    # ---- read-back --------------------------------------------------------------------------------------
    def ious(self, slot, sessions=None):
        c = self.counts[slot].cpu().numpy()
        if sessions is not None:
            c = c[sessions]
        with np.errstate(invalid="ignore", divide="ignore"):
            return c[:, 0] / c[:, 1]                    # int64 / int64 -> float64, as inference/utils.py:80-87

    def clicks_list(self, s):
        from .clicker import Click
        n = int(self.nclicks[s].item())
        rows = self.clicks[s, :n].cpu().numpy()
        return [Click(is_positive=bool(r[0]), coords=(int(r[1]), int(r[2])), indx=i) for i, r in enumerate(rows)]


def evaluate_device_sessions(samples, net, device, max_iou_thr, pred_thr=0.49, min_clicks=1, max_clicks=20, micro_batch=32,
                             target_size=448, stats=None, on_click=None):
    """The NoC loop of evaluate_sample for all `samples` [(image HWC, gt HW)], micro_batch sessions at a time, entirely on
    device state.  Returns per-sample IoU arrays (float32, one entry per executed click) like evaluate_lockstep.
    on_click(engine, click_indx) is a test hook called after every finish."""
    results = [None] * len(samples)
    n_calls = n_fwd = 0
    can_stop = max_iou_thr <= 1.0
    nmp = getattr(net, "num_max_points", 24)
    with torch.no_grad():
        for lo in range(0, len(samples), micro_batch):
            chunk = samples[lo:lo + micro_batch]
            eng = DeviceClickSessions([c[0] for c in chunk], [c[1] for c in chunk], device, target_size=target_size,
                                      max_clicks=max_clicks, pred_thr=pred_thr, num_max_points=nmp)
            executed = np.zeros(eng.S, dtype=np.int64)
            stopped_iou = {}
            eng.clicker_step(0)                                        # masks are all background: the first clicks
            for k in range(max_clicks):
                if not eng.active:
                    break
                image, points = eng.prepare()
                logits = net(image, points)["instances"]
                eng.finish(logits)
                n_calls += 1
                n_fwd += image.shape[0]
                executed[eng.active] += 1
                if on_click is not None:
                    on_click(eng, k)
                eng.clicker_step(k + 1)                                # IoU of click k (+ click k+1 of every session)
                if can_stop:
                    iou = eng.ious(k + 1, eng.active)
                    still = [s for s, v in zip(eng.active, iou) if not (v >= max_iou_thr and k + 1 >= min_clicks)]
                    if len(still) != len(eng.active):
                        eng.set_active(still)
            table = eng.counts.cpu().numpy()                            # [max_clicks+1, S, 2]
            with np.errstate(invalid="ignore", divide="ignore"):
                iou_all = table[1:, :, 0] / table[1:, :, 1]
            for s in range(eng.S):
                results[lo + s] = iou_all[:executed[s], s].astype(np.float32)
    if stats is not None:
        stats["network_calls"] = stats.get("network_calls", 0) + n_calls
        stats["click_forwards"] = stats.get("click_forwards", 0) + n_fwd
    return results
