"""NoBRS predictor (reference isegm/inference/predictors/base.py:10-223, get_predictor of predictors/__init__.py:9-99).

Same public surface: set_input_image, get_prediction, get_vqu_prediction, get_points_nd, get_states / set_states.
The per-click work is split into `prepare_inputs` (clicks + previous mask -> transformed network inputs) and
`finish_prediction` (network logits -> full-size probabilities, state update) so that the lock-step batched
evaluator (evaluation.evaluate_lockstep) can put many click sessions into ONE network call; the serial methods are
exactly prepare -> net -> finish.

Prompt simulation: the reference rebuilds box and scribble prompts on the host at every click
(engine/trainer.py:703-768) even though `as_prompt_type == 0` never reads them.  Here click-only prediction passes
prompts=None -- which also means the host `random` / numpy RNG draws of those unused simulations are NOT consumed for
type 0: a seeded run that mixes prompt types per click sees a different RNG stream than the reference from the first
type-0 click on (runs of a single type, what evaluate_vpumodel.py does, are unaffected); for prompt types 1/2 `prompt_fn(prev_mask_roi, gt_mask_roi, points_nd) -> prompts` is called, by default
the restated simulators of inference/prompts.py (`eval_prompt_fn`).
"""
import numpy as np
import torch
import torch.nn.functional as F

from .transforms import AddHorizontalFlip, SigmoidForPred, ZoomIn, get_roi_image_nd


def _to_tensor(image):
    """HWC uint8 / float array -> CHW float32 in [0, 1] (what torchvision's ToTensor does for ndarrays)."""
    if isinstance(image, torch.Tensor):
        return image
    a = np.asarray(image)
    if a.ndim == 2:
        a = a[:, :, None]
    t = torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1)))
    return t.float().div(255) if a.dtype == np.uint8 else t.float()


class BasePredictor:
    def __init__(self, model, device, net_clicks_limit=None, with_flip=False, with_sigmoid=True, zoom_in=None, max_size=None,
                 cascade_step=0, cascade_adaptive=False, cascade_clicks=1, prompt_fn=None, **kwargs):
        if max_size is not None:
            raise NotImplementedError("LimitLongestSide is not used by the VPU evaluation (scripts/evaluate_vpumodel.py:187-192)")
        if isinstance(model, tuple):
            raise NotImplementedError("multi-stage click models are outside this path")
        self.net = model
        self.device = device
        self.with_flip = with_flip
        self.with_sigmoid = with_sigmoid
        self.net_clicks_limit = net_clicks_limit
        self.zoom_in = zoom_in
        self.cascade_step = cascade_step
        self.cascade_adaptive = cascade_adaptive
        self.cascade_clicks = cascade_clicks
        self.prompt_fn = prompt_fn
        self.original_image = None
        self.prev_prediction = None
        self.transforms = [zoom_in] if zoom_in is not None else []
        if with_sigmoid:
            self.transforms.append(SigmoidForPred())
        if with_flip:
            self.transforms.append(AddHorizontalFlip())

    # ---- state ------------------------------------------------------------------------------------
    def set_input_image(self, image):
        image_nd = _to_tensor(image)
        for t in self.transforms:
            t.reset()
        self.original_image = image_nd.to(self.device)
        if self.original_image.dim() == 3:
            self.original_image = self.original_image.unsqueeze(0)
        self.prev_prediction = torch.zeros_like(self.original_image[:, :1, :, :])

    def get_states(self):
        return {"transform_states": [t.get_state() for t in self.transforms], "prev_prediction": self.prev_prediction.clone()}

    def set_states(self, states):
        assert len(states["transform_states"]) == len(self.transforms)
        for s, t in zip(states["transform_states"], self.transforms):
            t.set_state(s)
        self.prev_prediction = states["prev_prediction"]

    # ---- the two halves of one click ---------------------------------------------------------------
    def apply_transforms(self, image_nd, clicks_lists):
        changed = False
        for t in self.transforms:
            image_nd, clicks_lists = t.transform(image_nd, clicks_lists)
            changed |= t.image_changed
        return image_nd, clicks_lists, changed

    def get_points_nd(self, clicks_lists):
        """[len(clicks_lists), 2n, 3] (row, col, click order): positives first, negatives second, each half padded
        with (-1,-1,-1) to n = the largest per-list count of either kind (reference base.py:195-213)."""
        n_pos = [sum(c.is_positive for c in cl) for cl in clicks_lists]
        n_neg = [len(cl) - p for cl, p in zip(clicks_lists, n_pos)]
        n = max(n_pos + n_neg)
        if self.net_clicks_limit is not None:
            n = min(self.net_clicks_limit, n)
        n = max(1, n)
        rows = []
        for cl in clicks_lists:
            cl = cl[:self.net_clicks_limit]
            pos = [c.coords_and_indx for c in cl if c.is_positive]
            neg = [c.coords_and_indx for c in cl if not c.is_positive]
            rows.append(pos + (n - len(pos)) * [(-1, -1, -1)] + neg + (n - len(neg)) * [(-1, -1, -1)])
        return torch.tensor(rows, device=self.device)

    def prepare_inputs(self, clicker, prev_mask=None, gt_mask=None, as_prompt_type=0):
        """-> (image_nd [1 or 2, C, h, w], points_nd, prompts_nd or None) for one network call."""
        clicks_list = clicker.get_clicks()
        if prev_mask is None:
            prev_mask = self.prev_prediction
        input_image = self.original_image
        if getattr(self.net, "with_prev_mask", False):
            input_image = torch.cat((input_image, prev_mask), dim=1)
        image_nd, clicks_lists, _ = self.apply_transforms(input_image, [clicks_list])
        points_nd = self.get_points_nd(clicks_lists)
        prompts_nd = None
        if as_prompt_type != 0:
            if self.prompt_fn is None:
                raise ValueError("as_prompt_type %d needs a prompt_fn (box / scribble simulator)" % as_prompt_type)
            gt = torch.as_tensor(np.asarray(gt_mask, dtype=np.float32))[None, None]
            pm = prev_mask
            if self.with_flip:
                gt = torch.cat([gt, torch.flip(gt, dims=[3])], dim=0)
                pm = torch.cat([pm, torch.flip(pm, dims=[3])], dim=0)
            gt = gt.to(pm.device)
            if self.zoom_in is not None and self.zoom_in._object_roi is not None:
                gt = get_roi_image_nd(gt, self.zoom_in._object_roi, self.zoom_in.target_size)
                pm = get_roi_image_nd(pm, self.zoom_in._object_roi, self.zoom_in.target_size)
            prompts_nd = self.prompt_fn(pm, gt, points_nd)
        return image_nd, points_nd, prompts_nd

    def finish_prediction(self, pred_logits, image_size):
        """logits of this session's rows -> full-size probability map; updates prev_prediction / ZoomIn state."""
        prediction = F.interpolate(pred_logits, mode="bilinear", align_corners=True, size=tuple(image_size))
        for t in reversed(self.transforms):
            prediction = t.inv_transform(prediction)
        return prediction

    # ---- serial API (reference semantics) ------------------------------------------------------------
    def _cascade(self, clicker, on_cascade):
        return len(clicker.get_clicks()) <= self.cascade_clicks and self.cascade_step > 0 and not on_cascade

    def get_vqu_prediction(self, clicker, prev_mask=None, on_cascade=False, gt_mask=None, as_prompt_type=0, click_indx=0,
                           as_multi_prompts=True):
        if self._cascade(clicker, on_cascade):
            for _ in range(self.cascade_step):
                prediction, prompts_nd = self.get_vqu_prediction(clicker, None, True, gt_mask, as_prompt_type, click_indx,
                                                                 as_multi_prompts)
                if self.cascade_adaptive and prev_mask is not None:
                    if ((prediction > 0.49) != (prev_mask > 0.49)).sum() <= 20:
                        return prediction, prompts_nd
                prev_mask = prediction
            return prediction, prompts_nd
        if not as_multi_prompts:
            # the reference's as_multi_prompts=False branch (base.py:153-163, trainer.get_next_promts_inference) rewrites points_nd
            # from simulated prompts and calls net(image, points) without them: a different prediction, not part of this path
            raise NotImplementedError("as_multi_prompts=False (_get_vqu_prediction_points, base.py:153-163) is outside the B200 path")
        image_nd, points_nd, prompts_nd = self.prepare_inputs(clicker, prev_mask, gt_mask, as_prompt_type)
        if prompts_nd is None:
            logits = self.net(image_nd, points_nd)["instances"]
        else:
            logits = self.net(image_nd, points_nd, prompts_nd, as_prompt_type)["instances"]
        prediction = self.finish_prediction(logits, image_nd.shape[2:])
        if self.zoom_in is not None and self.zoom_in.check_possible_recalculation():
            return self.get_prediction(clicker), prompts_nd
        self.prev_prediction = prediction
        return prediction.cpu().numpy()[0, 0], prompts_nd

    def get_prediction(self, clicker, prev_mask=None, on_cascade=False):
        if self._cascade(clicker, on_cascade):
            for _ in range(self.cascade_step):
                prediction = self.get_prediction(clicker, None, True)
                if self.cascade_adaptive and prev_mask is not None:
                    if ((prediction > 0.49) != (prev_mask > 0.49)).sum() <= 20:
                        return prediction
                prev_mask = prediction
            return prediction
        image_nd, points_nd, _ = self.prepare_inputs(clicker, prev_mask)
        logits = self.net(image_nd, points_nd)["instances"]
        prediction = self.finish_prediction(logits, image_nd.shape[2:])
        if self.zoom_in is not None and self.zoom_in.check_possible_recalculation():
            return self.get_prediction(clicker)
        self.prev_prediction = prediction
        return prediction.cpu().numpy()[0, 0]


def get_predictor(net, brs_mode, device, prob_thresh=0.49, with_flip=True, zoom_in_params=dict(), predictor_params=None,
                  brs_opt_func_params=None, lbfgs_params=None):
    """NoBRS only (the BRS predictors are outside this path)."""
    if brs_mode != "NoBRS":
        raise NotImplementedError("only the NoBRS predictor is part of the B200 path (got %r)" % (brs_mode,))
    zoom_in = ZoomIn(**zoom_in_params) if zoom_in_params is not None else None
    params = {"optimize_after_n_clicks": 1}
    if predictor_params is not None:
        params.update(predictor_params)
    return BasePredictor(net, device, zoom_in=zoom_in, with_flip=with_flip, **params)


def vpu_eval_predictor(net, device, prompt_fn=None):
    """The configuration scripts/evaluate_vpumodel.py builds for 448-px VPU models (:161-165,187-192): NoBRS, flip TTA,
    fixed 448x448 zoom-in with skip_clicks=-1, one cascade step on the first click."""
    p = get_predictor(net, "NoBRS", device, with_flip=True, zoom_in_params={"skip_clicks": -1, "target_size": (448, 448)},
                      predictor_params={"cascade_step": 1, "cascade_adaptive": False, "cascade_clicks": 1})
    if prompt_fn is None:
        from .prompts import eval_prompt_fn as prompt_fn
    p.prompt_fn = prompt_fn
    return p
