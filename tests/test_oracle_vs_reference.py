"""Build container only: the oracle restatement against the UNMODIFIED reference, live."""
import random

import numpy as np
import pytest
import torch

from oracle import cases, ref_harness as rh, vpu_oracle as vo
from pvpuformer_b200.config import make_config
from pvpuformer_b200.weights import param_spec, synthetic_state_dict

pytestmark = pytest.mark.reference


def test_state_dict_layout_matches_reference():
    for arch in ("vit_base",):
        m = rh.build_reference_model(arch)
        sd = m.state_dict()
        spec = param_spec(make_config(arch))
        assert list(sd.keys()) == list(spec.keys())
        for k in sd:
            assert tuple(sd[k].shape) == tuple(spec[k][0]), k


def test_forward_bitexact_vs_reference_clicks():
    cfg = make_config("vit_base")
    m = rh.build_reference_model("vit_base")
    sd = synthetic_state_dict(cfg, 0)
    m.load_state_dict(sd, strict=True)
    image4 = cases.images(2, seed=11)
    pts = cases.random_clicks(2, seed=12, dtype=torch.float64)
    with torch.no_grad():
        ref = m(image4, pts)
        out = vo.forward(sd, cfg, image4, pts)
    assert torch.equal(ref["instances"], out["instances"])
    assert torch.equal(ref["instances_aux"], out["instances_aux"])


def test_ppue_fuzz_vs_reference():
    m = rh.build_reference_model("vit_base")
    rs = np.random.RandomState(0)
    for it in range(6):
        n = int(rs.randint(1, 25))
        pts = torch.tensor(rs.uniform(-30, 480, (3, 2 * n, 3)))
        pts[:, :, 2] = torch.tensor(rs.randint(-1, 3, (3, 2 * n))).double()
        a = m._guassinvector_click(pts)
        b = vo.ppue(pts)
        assert torch.equal(a, b)
        boxes = torch.tensor(np.column_stack([rs.randint(0, 448, 3), rs.randint(0, 448, 3), rs.randint(0, 300, 3),
                                              rs.randint(0, 300, 3), rs.randint(0, 2 * n, 3)]).astype(np.int32))
        a = m._guassinvector_box(pts, boxes)
        b = vo.ppue(pts, (pts, boxes, None), 1)
        assert torch.equal(a, b)
        disks_ref = m.dist_maps(torch.zeros(3, 3, 448, 448), pts.clone())
        assert torch.equal(disks_ref, vo.disk_maps(pts, 448, 448))


def test_training_losses_bitexact_vs_reference_classes():
    """oracle/losses.py against the reference's own loss modules (configuration of vpu_base448_cocolvis.py:72-80) on random
    logits / probabilities with an ignore region."""
    rh.import_reference()
    from isegm.model.losses import DiceLoss, NormalizedFocalLossSigmoid, SigmoidBinaryCrossEntropyLoss
    from oracle import losses as ol
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(3, 1, 64, 64, generator=g) * 2
    gt = (torch.rand(3, 1, 64, 64, generator=g) > 0.6).float()
    gt_ign = gt.clone()
    gt_ign[:, :, :5] = -1
    aux = torch.rand(3, 48, 64, 64, generator=g)
    assert torch.equal(NormalizedFocalLossSigmoid(alpha=0.5, gamma=2, penalty_loss=False)(logits, gt_ign),
                       ol.normalized_focal_loss_sigmoid(logits, gt_ign))
    assert torch.equal(DiceLoss(use_sigmoid=True, activate=True, naive_dice=True, loss_weight=1.0)(logits, gt),
                       ol.dice_loss_sigmoid_naive(logits, gt))
    lab = ol.ed_mask_label(gt)
    assert torch.equal(SigmoidBinaryCrossEntropyLoss(from_sigmoid=True)(aux, lab), ol.sigmoid_bce_from_sigmoid(aux, lab))
