"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Run in the build container only (needs /root/reference):   python -m oracle.make_golden
The reference is imported read-only through oracle/ref_harness.py; weights are
pvpuformer_b200.weights.synthetic_state_dict (seed 0) loaded with strict=True.  Stored tensors
are small slices/low-res taps so the fixtures stay a few hundred KB each.
"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cases, ref_harness as rh  # noqa: E402
from pvpuformer_b200.config import make_config  # noqa: E402
from pvpuformer_b200.weights import synthetic_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def run_reference(model, image4, points, prompts=None, t=0, scribble_seed=7):
    """Forward with taps on the head's low-res outputs."""
    taps = {}
    orig = model.head.forward_feat

    def ff(inputs, inputs2, pclout=False):
        out, logits = orig(inputs, inputs2, pclout=pclout)
        taps["seg_lowres"], taps["aux_lowres"] = out, logits
        taps["q_out"] = inputs2
        return out, logits
    model.head.forward_feat = ff
    orig_bb = model.backbone.forward_backbone

    def fb(*a, **k):
        r = orig_bb(*a, **k)
        taps["backbone_features"] = r
        return r
    model.backbone.forward_backbone = fb
    try:
        with torch.no_grad():
            random.seed(scribble_seed)
            out = model(image4, points, prompts, t)
            img, prev = model.prepare_input(image4)
            taps["coord_features"] = model.get_coord_features_with_prompt(img, prev, points, prompts, t)
            random.seed(scribble_seed)
            if t == 0:
                taps["ppue"] = model._guassinvector_click(points)
            elif t == 1:
                taps["ppue"] = model._guassinvector_box(prompts[0], prompts[1])
            else:
                taps["ppue"] = model._guassinvector_scribble(prompts[0], prompts[2])
    finally:
        model.head.forward_feat = orig
        model.backbone.forward_backbone = orig_bb
    return out, taps


def pack(out, taps):
    d = {
        "ppue": taps["ppue"].float().numpy(),
        "disks_packed": np.packbits(taps["coord_features"][:, 1:].numpy().astype(np.uint8)),
        "seg_lowres": taps["seg_lowres"].numpy(),
        "aux_lowres_sel": taps["aux_lowres"][:, [0, 1, 23, 24, 25, 47]].numpy().astype(np.float32),
        "instances_s4": out["instances"][:, :, ::4, ::4].numpy(),
        "instances_row100": out["instances"][:, 0, 100, :].numpy(),
        "aux_s8_sel": out["instances_aux"][:, [0, 24], ::8, ::8].numpy(),
        "backbone_slice": taps["backbone_features"][:, ::49, ::16].numpy(),
        "q_out_slice": taps["q_out"][:, :, ::16].numpy(),
    }
    return d


def train12(model):
    from isegm.model.losses import DiceLoss, NormalizedFocalLossSigmoid, SigmoidBinaryCrossEntropyLoss
    from oracle import losses as ol
    image4, pts, gt = cases.train12_inputs()
    out, taps = run_reference(model, image4, pts)
    nfl = NormalizedFocalLossSigmoid(alpha=0.5, gamma=2, penalty_loss=False)(out["instances"], gt)
    dice = DiceLoss(use_sigmoid=True, activate=True, naive_dice=True, loss_weight=1.0)(out["instances"], gt)
    bce = SigmoidBinaryCrossEntropyLoss(from_sigmoid=True)(out["instances_aux"], ol.ed_mask_label(gt))
    keep = {"ppue_support_packed": np.packbits(taps["ppue"].numpy() != 0),          # all 48 rows in use (n = 24 per half)
            "seg_lowres_s4": taps["seg_lowres"][:, :, ::4, ::4].numpy(),
            "instances_row100": out["instances"][:, 0, 100, :].numpy(),
            "aux_s16_sel": out["instances_aux"][:, [0, 24], ::16, ::16].numpy()}
    np.savez_compressed(os.path.join(OUT, "vit_base_train12.npz"), loss_nfl=nfl.numpy(), loss_dice=dice.numpy(),
                        loss_bce_aux=bce.numpy(), **keep)


def main():
    os.makedirs(OUT, exist_ok=True)
    rh.import_reference()
    if "--only-large-huge" not in sys.argv:
        base_cases()
    large_huge_cases()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def base_cases():
    from isegm.engine.trainer import get_next_promts

    cfg = make_config("vit_base")
    model = rh.build_reference_model("vit_base")
    model.load_state_dict(synthetic_state_dict(cfg, 0), strict=True)

    # case 1: clicks (type 0), B=2, float64 points, sigmoid-noise prev mask
    image4 = cases.images(2, seed=1)
    out, taps = run_reference(model, image4, cases.CLICKS_A)
    np.savez_compressed(os.path.join(OUT, "vit_base_clicks.npz"), **pack(out, taps))

    # case 2/3: box and scribble prompts from the reference's simulator, B=3
    masks = cases.ellipse_masks(3, seed=3)
    pts = cases.first_clicks_in_masks(masks, seed=4)
    image4 = cases.images(3, seed=2, prev="zeros")
    random.seed(3)
    np.random.seed(3)
    prompts = get_next_promts(image4[:, 3:], torch.tensor(masks)[:, None], pts,
                              as_allmask=False, jitter_box=False)
    pr = dict(prompt_points=prompts[0].numpy(), boxes=prompts[1].numpy(),
              scribbles=np.asarray(prompts[2][0]).astype(np.int32), rects=np.asarray(prompts[2][1]).astype(np.int32),
              points=pts.numpy())
    for t, name in ((1, "box"), (2, "scribble")):
        out, taps = run_reference(model, image4, pts, prompts, t)
        np.savez_compressed(os.path.join(OUT, "vit_base_%s.npz" % name), **pack(out, taps), **pr)

    # case 4: random 1..20 clicks per image (config-2 style), B=3, float32 points
    image4 = cases.images(3, seed=5)
    pts = cases.random_clicks(3, seed=6)
    out, taps = run_reference(model, image4, pts)
    np.savez_compressed(os.path.join(OUT, "vit_base_manyclicks.npz"), **pack(out, taps))

    # case 5: training shape, B=12, points [12,48,3]; the reference's own loss classes on its outputs
    train12(model)
    del model



def large_huge_cases():
    """ViT-L / ViT-H clicks (B=2): the full fixture set of the ViT-B cases (PPuE rows, disks, low-res and full-res logits, aux
    selection), so that the L/H parity tests carry the same gates as the ViT-B ones."""
    for arch in ("vit_large", "vit_huge"):
        cfg = make_config(arch)
        model = rh.build_reference_model(arch)
        model.load_state_dict(synthetic_state_dict(cfg, 0), strict=True)
        image4 = cases.images(2, seed=1)
        out, taps = run_reference(model, image4, cases.CLICKS_A)
        np.savez_compressed(os.path.join(OUT, "%s_clicks.npz" % arch), **pack(out, taps))
        del model


if __name__ == "__main__":
    main()
