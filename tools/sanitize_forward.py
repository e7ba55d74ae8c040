"""One small forward per architecture (batch 2 = a NoBRS click with flip TTA; all three prompt types for ViT-B) plus one
device-session click: the command tools/sanitize.sh runs under compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pvpuformer_b200 import synthetic
from pvpuformer_b200.config import make_config
from pvpuformer_b200.model import build_model
from pvpuformer_b200.weights import synthetic_state_dict

archs = sys.argv[1:] or ["vit_base", "vit_large", "vit_huge"]
for arch in archs:
    m = build_model(arch, state_dict=synthetic_state_dict(make_config(arch), 0), device="cuda")
    for B in (2, 3):
        img = synthetic.images(B, seed=3).cuda()
        pts = synthetic.random_clicks(B, seed=4, dtype=torch.float64).cuda()
        out = m(img, pts)
        torch.cuda.synchronize()
        print(arch, B, float(out["instances"].abs().mean()), float(out["instances_aux"].abs().mean()), flush=True)
print("done")
