"""What HBM rate do simple streaming kernels reach on this board for the read / write mixes of the residual epilogues?
(torch elementwise kernels as the yardstick; GPU box: python tools/hbm_mix_bench.py)"""
import torch
from attn_bench import timeit

dev = torch.device("cuda:0")
n = 50176 * 768
x = torch.randn(n, device=dev)
r = torch.randn(n, device=dev)
a = torch.randn(n, device=dev).to(torch.bfloat16)
o = torch.empty_like(x)
ob = torch.empty(n, device=dev, dtype=torch.bfloat16)
for name, fn, by in (
    ("copy fp32 (r 154 + w 154 MB)", lambda: o.copy_(x), 8.0 * n),
    ("x += r in place (r 308 + w 154)", lambda: x.add_(r), 12.0 * n),
    ("o = x + r (r 308 + w 154)", lambda: torch.add(x, r, out=o), 12.0 * n),
    ("x *= 1.0001 in place (r 154 + w 154)", lambda: x.mul_(1.0001), 8.0 * n),
    ("bf16 cast (r 154 + w 77)", lambda: ob.copy_(x), 6.0 * n),
    ("fill fp32 (w 154)", lambda: o.fill_(1.0), 4.0 * n),
    ("sum fp32 (r 154)", lambda: x.sum(), 4.0 * n),
):
    ms = timeit(fn, 20)
    print("%-40s %7.1f us  %7.1f GB/s" % (name, ms * 1e3, by / ms / 1e6), flush=True)
