"""GEMM micro-benchmark through the C ABI (vpu_gemm): the forward's dominant shapes, 2-CTA vs 1-CTA kernel.
Usage on the GPU box: python tools/gemm_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvpuformer_b200 import ops  # noqa: E402

SHAPES = [  # name, M, N, K, act, residual(fp32 in place), out dtype
    ("qkv", 50176, 2304, 768, None, False, torch.bfloat16),
    ("proj+res", 50176, 768, 768, None, True, torch.float32),
    ("fc1+gelu", 50176, 3072, 768, "gelu", False, torch.bfloat16),
    ("fc2+res", 50176, 768, 3072, None, True, torch.float32),
    ("dma.img", 50176, 1152, 768, None, False, torch.bfloat16),
    ("dma.small", 3072, 768, 768, None, False, torch.bfloat16),
    ("neck.d4b", 200704, 768, 384, None, False, torch.bfloat16),
    ("head.c0", 802816, 256, 128, "relu", False, torch.bfloat16),
]


def main():
    dev = torch.device("cuda:0")
    for name, M, N, K, act, res, odt in SHAPES:
        A = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
        W = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        out = torch.zeros(M, N, device=dev, dtype=odt)
        line = "%-10s M=%6d N=%4d K=%4d " % (name, M, N, K)
        for impl in (0, 2):
            def run():
                ops.gemm(A, W, bias=bias, residual=out if res else None, act=act, out_dtype=odt, impl=impl, out=out)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            line += " | impl%d %7.1f us %7.1f TF/s" % (impl, ms * 1e3, 2.0 * M * N * K / (ms * 1e-3) / 1e12)
        print(line, flush=True)


if __name__ == "__main__":
    main()
