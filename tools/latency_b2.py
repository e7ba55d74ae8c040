"""Batch-2 forward latency (one NoBRS click with flip TTA, instances only) through the module call.  GPU box: python tools/latency_b2.py [arch]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvpuformer_b200 import synthetic
from pvpuformer_b200.config import make_config
from pvpuformer_b200.model import build_model
from pvpuformer_b200.weights import synthetic_state_dict

arch = sys.argv[1] if len(sys.argv) > 1 else "vit_base"
m = build_model(arch, state_dict=synthetic_state_dict(make_config(arch), 0), device="cuda")
m.want_aux = False
for B in (2, 8):
    img = synthetic.images(B, seed=3).cuda()
    pts = synthetic.random_clicks(B, seed=4, dtype=torch.float64).cuda()
    for graphs in (True, False):
        m.graph_max_batch = 8 if graphs else 0
        for _ in range(5):
            m(img, pts)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            m(img, pts)
        e1.record()
        torch.cuda.synchronize()
        print("%s batch %d graph=%s: %.3f ms" % (arch, B, graphs, e0.elapsed_time(e1) / 50), flush=True)
