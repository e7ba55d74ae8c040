"""Device-time split of one NoC click over a micro-batch of device-resident sessions (clicker / prepare / forward / finish).
Usage on the GPU box: python tools/noc_profile.py [--arch vit_base] [--sessions 32] [--clicks 6]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvpuformer_b200.config import make_config  # noqa: E402
from pvpuformer_b200.inference.datasets import SyntheticEllipseDataset  # noqa: E402
from pvpuformer_b200.inference.device_session import DeviceClickSessions  # noqa: E402
from pvpuformer_b200.model import build_model  # noqa: E402
from pvpuformer_b200.weights import synthetic_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="vit_base")
    ap.add_argument("--sessions", type=int, default=32)
    ap.add_argument("--clicks", type=int, default=6)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = make_config(args.arch)
    m = build_model(args.arch, state_dict=synthetic_state_dict(cfg, 0), device=dev)
    m.want_aux = False
    ds = SyntheticEllipseDataset(args.sessions)
    samples = [(ds.get_sample(i).image, ds.get_sample(i).gt_mask(1)) for i in range(args.sessions)]
    eng = DeviceClickSessions([s[0] for s in samples], [s[1] for s in samples], dev, max_clicks=args.clicks + 2)
    names = ["clicker", "prepare", "forward", "finish"]
    tot = {n: [] for n in names}
    eng.clicker_step(0)
    with torch.no_grad():
        for k in range(args.clicks + 1):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            ev[0].record()
            if k:
                eng.clicker_step(k)
            ev[1].record()
            image, points = eng.prepare()
            ev[2].record()
            logits = m(image, points)["instances"]
            ev[3].record()
            eng.finish(logits)
            ev[4].record()
            torch.cuda.synchronize()
            if k:                                       # click 0 warms up (weights packed, workspace allocated)
                for i, n in enumerate(names):
                    tot[n].append(ev[i].elapsed_time(ev[i + 1]))
    print({n: "%.3f ms" % float(np.median(v)) for n, v in tot.items()}, "sessions", args.sessions, "model batch", 2 * args.sessions)


if __name__ == "__main__":
    main()
