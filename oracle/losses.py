"""CPU restatement of the three training losses of the reference's VPU config (TEST INFRASTRUCTURE ONLY).

Used for the config-5 parity case (ViT-B, batch 12, training-shape forward): the CUDA forward's outputs are
pushed through these losses and compared with the loss values the UNMODIFIED reference produced on its own outputs
(tests/golden/vit_base_train12.npz, oracle/make_golden.py).  Pinned against the reference classes in
tests/test_oracle_vs_reference.py.  Loss configuration: reference models/iSegNet/vpu_base448_cocolvis.py:72-80.
"""
import torch


def normalized_focal_loss_sigmoid(pred, label, alpha=0.5, gamma=2, eps=1e-12, ignore_label=-1):
    """isegm/model/losses.py:11-83, NormalizedFocalLossSigmoid(alpha=0.5, gamma=2, penalty_loss=False), from_sigmoid=False,
    detach_delimeter=True, max_mult=-1, size_average=True, weight=1 -> [B]."""
    one_hot = label > 0.5
    sample_weight = label != ignore_label
    pred = torch.sigmoid(pred)
    alpha_t = torch.where(one_hot, alpha * sample_weight, (1 - alpha) * sample_weight)
    pt = torch.where(sample_weight, 1.0 - torch.abs(label - pred), torch.ones_like(pred))
    beta = (1 - pt) ** gamma
    sw_sum = torch.sum(sample_weight, dim=(-2, -1), keepdim=True)
    beta_sum = torch.sum(beta, dim=(-2, -1), keepdim=True)
    beta = beta * (sw_sum / (beta_sum + eps))
    loss = -alpha_t * beta * torch.log(torch.min(pt + eps, torch.ones(1, dtype=torch.float)))
    loss = loss * sample_weight
    dims = tuple(range(1, loss.dim()))
    bsum = torch.sum(sample_weight, dim=dims)
    return torch.sum(loss, dim=dims) / (bsum + eps)


def dice_loss_sigmoid_naive(pred, target, eps=1e-3):
    """isegm/model/losses.py:225-369, DiceLoss(use_sigmoid=True, activate=True, naive_dice=True, loss_weight=1.0),
    reduction 'mean' -> scalar."""
    p = pred.sigmoid().flatten(1)
    t = target.flatten(1).float()
    a = torch.sum(p * t, 1)
    d = (2 * a + eps) / (torch.sum(p, 1) + torch.sum(t, 1) + eps)
    return (1 - d).mean()


def sigmoid_bce_from_sigmoid(pred, label, ignore_label=-1):
    """isegm/model/losses.py:155-176, SigmoidBinaryCrossEntropyLoss(from_sigmoid=True) on probabilities -> [B]."""
    label = label.view(pred.size())
    sample_weight = label != ignore_label
    label = torch.where(sample_weight, label, torch.zeros_like(label))
    eps = 1e-12
    loss = -(torch.log(pred + eps) * label + torch.log(1. - pred + eps) * (1. - label))
    loss = loss * sample_weight
    return torch.mean(loss, dim=tuple(range(1, loss.dim())))


def ed_mask_label(gt, num_max_points=24):
    """isegm/engine/trainer.py:329-331: positive rows carry the object mask, negative rows its complement -> [B, 2n, H, W]."""
    pos = gt.repeat(1, num_max_points, 1, 1)
    neg = torch.logical_not(gt).repeat(1, num_max_points, 1, 1)
    return torch.cat([pos, neg], dim=1).float()


def training_losses(out, gt, num_max_points=24):
    """-> dict of the three loss terms of one training step's first forward (trainer.py:399-419)."""
    return {"nfl": normalized_focal_loss_sigmoid(out["instances"], gt),
            "dice": dice_loss_sigmoid_naive(out["instances"], gt),
            "bce_aux": sigmoid_bce_from_sigmoid(out["instances_aux"], ed_mask_label(gt, num_max_points))}
