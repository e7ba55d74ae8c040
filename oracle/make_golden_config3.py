"""Golden vectors for BASELINE.json configs[2] / SURVEY.md 8(d) config 3 AT ITS STATED SIZE, from the UNMODIFIED reference
(build container only):   python -m oracle.make_golden_config3  ->  tests/golden/vit_large_config3.npz

ViT-Large, batch 32 split 11 / 11 / 10 over click / box / scribble prompts (the reference takes one `as_prompt_type` per
call, is_vpu_model.py:233-352, so the batch is three sub-batches, none of size 4 -- the reference's crashing size).  Ground
truth = random ellipses, clicks = one positive click inside each, boxes / scribbles from the reference's own simulator
get_next_promts(prev = 0, gt, points, as_allmask=False, jitter_box=False) (engine/trainer.py:703-768) under
random.seed / np.random.seed = 3, images = synthetic.images(32, seed=12, prev='zeros').  Stored: the prompts (so that the GPU
box feeds exactly these), the PPuE support, and stride-8 logits of all 32 samples + stride-16 aux of two queries.
"""
import os
import random

import numpy as np
import torch

from oracle import cases, ref_harness as rh
from oracle.make_golden import OUT, run_reference
from pvpuformer_b200.config import make_config
from pvpuformer_b200.weights import synthetic_state_dict

SPLIT = ((0, 0, 11), (1, 11, 22), (2, 22, 32))     # (as_prompt_type, start, stop)


def inputs():
    masks = cases.ellipse_masks(32, seed=13)
    pts = cases.first_clicks_in_masks(masks, seed=14)
    image4 = cases.images(32, seed=12, prev="zeros")
    return image4, pts, masks


def main():
    rh.import_reference()
    from isegm.engine.trainer import get_next_promts
    arch = "vit_large"
    model = rh.build_reference_model(arch)
    model.load_state_dict(synthetic_state_dict(make_config(arch), 0), strict=True)
    image4, pts, masks = inputs()
    keep = {}
    inst, aux, sup = [], [], []
    for t, a, b in SPLIT:
        prompts = None
        if t != 0:
            random.seed(3)
            np.random.seed(3)
            prompts = get_next_promts(image4[a:b, 3:], torch.tensor(masks[a:b])[:, None], pts[a:b], as_allmask=False, jitter_box=False)
            keep["t%d_prompt_points" % t] = prompts[0].numpy()
            keep["t%d_boxes" % t] = prompts[1].numpy()
            keep["t%d_scribbles" % t] = np.asarray(prompts[2][0]).astype(np.int32)
            keep["t%d_rects" % t] = np.asarray(prompts[2][1]).astype(np.int32)
        out, taps = run_reference(model, image4[a:b], pts[a:b], prompts, t)
        inst.append(out["instances"][:, :, ::8, ::8].numpy())
        aux.append(out["instances_aux"][:, [0, 24], ::16, ::16].numpy())
        sup.append(np.packbits(taps["ppue"].float().numpy() != 0, axis=None))   # fp32, as the network consumes the rows
        print("type", t, "rows", a, b, "done", flush=True)
    path = os.path.join(OUT, "vit_large_config3.npz")
    np.savez_compressed(path, instances_s8=np.concatenate(inst), aux_s16_sel=np.concatenate(aux),
                        ppue_support_t0=sup[0], ppue_support_t1=sup[1], ppue_support_t2=sup[2], **keep)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
