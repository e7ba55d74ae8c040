"""Per-launch-class profile of one forward through the C ABI's event profiler in detail mode (scalar debug.profile_detail = 1:
every GEMM is its own class, labelled by weight key and M x N x K; ViT blocks pooled).  GPU box:
    python tools/profile_detail.py [arch] [batch]      -> table sorted by time, TFLOP/s and GB/s per class"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvpuformer_b200 import lib as L, synthetic  # noqa: E402
from pvpuformer_b200.config import make_config  # noqa: E402
from pvpuformer_b200.model import build_model  # noqa: E402
from pvpuformer_b200.weights import synthetic_state_dict  # noqa: E402


def main():
    arch = sys.argv[1] if len(sys.argv) > 1 else "vit_base"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    steps = 3
    m = build_model(arch, state_dict=synthetic_state_dict(make_config(arch), 0), device="cuda")
    m.graph_max_batch = 0        # launch by launch: the event brackets need the individual launches
    img = synthetic.images(B, seed=3).cuda()
    pts = synthetic.random_clicks(B, seed=4, dtype=torch.float64).cuda()
    for _ in range(3):
        m(img, pts)
    torch.cuda.synchronize()
    lib = L.load()
    L.check(lib.vpu_set_scalar(m._handle, b"debug.profile_detail", 1.0))
    L.check(lib.vpu_profile_begin(m._handle))
    for _ in range(steps):
        m(img, pts)
    torch.cuda.synchronize()
    arr = (L.VpuProfileEntry * 512)()
    n = ctypes.c_int(0)
    L.check(lib.vpu_profile_end(m._handle, arr, 512, ctypes.byref(n)))
    rows = [(arr[i].name.decode(), arr[i].ms / steps, arr[i].flops / steps, arr[i].bytes / steps, arr[i].launches // steps)
            for i in range(n.value)]
    tot = sum(r[1] for r in rows)
    print("%s batch %d: %.3f ms summed over %d classes" % (arch, B, tot, len(rows)))
    for name, ms, fl, by, ln in sorted(rows, key=lambda r: -r[1]):
        print("%-47s %3d x %7.1f us = %7.3f ms %5.1f%%  %7.1f TF/s %7.1f GB/s" %
              (name, ln, ms / max(ln, 1) * 1e3, ms, 100 * ms / tot, fl / ms / 1e9 if fl else 0.0, by / ms / 1e6))


if __name__ == "__main__":
    main()
