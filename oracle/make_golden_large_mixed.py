"""Golden vectors for SURVEY.md 8(d) config 3 (ViT-Large, box and scribble prompts through PPuE) from the UNMODIFIED reference
(build container only):   python -m oracle.make_golden_large_mixed  ->  tests/golden/vit_large_{box,scribble}.npz
Same inputs and prompt simulation as the ViT-B box / scribble cases of oracle/make_golden.py (the prompts do not depend on
the backbone); the head's hard-coded d_model = 768 FFN is replaced by FFNBlock(C, 2C, 256) in the harness (ref_harness.py)."""
import os
import random

import numpy as np
import torch

from oracle import cases, ref_harness as rh
from oracle.make_golden import OUT, pack, run_reference
from pvpuformer_b200.config import make_config
from pvpuformer_b200.weights import synthetic_state_dict


def main():
    rh.import_reference()
    from isegm.engine.trainer import get_next_promts
    arch = "vit_large"
    model = rh.build_reference_model(arch)
    model.load_state_dict(synthetic_state_dict(make_config(arch), 0), strict=True)
    masks = cases.ellipse_masks(3, seed=3)
    pts = cases.first_clicks_in_masks(masks, seed=4)
    image4 = cases.images(3, seed=2, prev="zeros")
    random.seed(3)
    np.random.seed(3)
    prompts = get_next_promts(image4[:, 3:], torch.tensor(masks)[:, None], pts, as_allmask=False, jitter_box=False)
    pr = dict(prompt_points=prompts[0].numpy(), boxes=prompts[1].numpy(), scribbles=np.asarray(prompts[2][0]).astype(np.int32),
              rects=np.asarray(prompts[2][1]).astype(np.int32), points=pts.numpy())
    for t, name in ((1, "box"), (2, "scribble")):
        out, taps = run_reference(model, image4, pts, prompts, t)
        path = os.path.join(OUT, "%s_%s.npz" % (arch, name))
        np.savez_compressed(path, **pack(out, taps), **pr)
        print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
