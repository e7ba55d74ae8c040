"""Where does the config-3 scribble PPuE support differ from the fixture?  (GPU box)"""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
import golden_util as gu
from oracle import vpu_oracle as vo
from oracle.make_golden_config3 import SPLIT, inputs
from pvpuformer_b200.model import build_model
from pvpuformer_b200.config import make_config
from pvpuformer_b200.weights import synthetic_state_dict

g = gu.load("vit_large_config3")
image4, pts, _ = inputs()
m = build_model("vit_large", state_dict=synthetic_state_dict(make_config("vit_large"), 0), device="cuda")
for t, a, b in SPLIT[1:]:
    prompts = (torch.from_numpy(g["t%d_prompt_points" % t]), torch.from_numpy(g["t%d_boxes" % t]), [g["t%d_scribbles" % t], g["t%d_rects" % t]])
    gu.seed_scribble()
    ref = vo.ppue(prompts[0], prompts, t).float().numpy()
    gu.seed_scribble()
    rows = m.ppue(pts[a:b].cuda(), (prompts[0].cuda(), prompts[1].cuda(), prompts[2]), t).cpu().numpy()
    gold = np.unpackbits(g["ppue_support_t%d" % t])[:ref.size].reshape(ref.shape).astype(bool)
    print("type", t, "oracle==gold", np.array_equal(ref != 0, gold), "gpu==gold", np.array_equal(rows != 0, gold),
          "max|gpu-oracle|", np.abs(rows - ref).max())
    d = np.argwhere((rows != 0) != gold)
    print(" differing entries", len(d), d[:12].tolist())
    for bb, j, c in d[:12]:
        print("  ", bb, j, c, "gpu", rows[bb, j, c], "oracle", ref[bb, j, c])
