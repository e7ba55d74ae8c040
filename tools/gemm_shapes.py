"""A/B micro-benchmark of the GEMM shapes outside the ViT blocks (DMA, neck, head) through the C ABI (vpu_gemm).
Usage on the GPU box: python tools/gemm_shapes.py   (VPU_GEMM_RAGGED256=0 for the 128-wide tiles on N = 1152)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvpuformer_b200 import ops  # noqa: E402

CASES = [  # name, M, N, K, act, residual dtype or None, out dtype, table rows
    ("dma.img+tab", 50176, 1152, 768, None, None, torch.bfloat16, 784),
    ("dma.img", 50176, 1152, 768, None, None, torch.bfloat16, 0),
    ("dmaf.img+tab", 50176, 768, 768, None, None, torch.bfloat16, 784),
    ("i2t.o resbf16", 50176, 768, 384, None, torch.bfloat16, torch.float32, 0),
    ("i2t.o resf32", 50176, 768, 384, None, torch.float32, torch.float32, 0),
    ("proj resf32", 50176, 768, 768, None, torch.float32, torch.float32, 0),
    ("fc2 resf32", 50176, 768, 3072, None, torch.float32, torch.float32, 0),
    ("qkv", 50176, 2304, 768, None, None, torch.bfloat16, 0),
    ("head.c0 relu", 802816, 256, 128, "relu", None, torch.bfloat16, 0),
    ("head.f0", 802816, 256, 256, None, None, torch.bfloat16, 0),
    ("dma.small", 3072, 768, 768, None, None, torch.bfloat16, 0),
    ("dma.sa.qk", 3072, 1536, 768, None, None, torch.bfloat16, 0),
    ("dma.t2i.q", 3072, 384, 768, None, None, torch.bfloat16, 0),
]


def main():
    dev = torch.device("cuda:0")
    for name, M, N, K, act, rdt, odt, tab in CASES:
        A = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
        W = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        out = torch.zeros(M, N, device=dev, dtype=odt)
        res = torch.randn(M, N, device=dev).to(rdt) if rdt is not None else None
        table = torch.randn(tab, N, device=dev) if tab else None

        def run():
            ops.gemm(A, W, bias=None if tab else bias, bias2d=table, residual=res, act=act, out_dtype=odt, out=out)
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
        by = 2.0 * (M * K + N * K) + M * N * (out.element_size() + (res.element_size() if res is not None else 0))
        print("%-14s M=%6d N=%4d K=%4d  %7.1f us %7.1f TF/s %7.1f GB/s" % (name, M, N, K, ms * 1e3, 2.0 * M * N * K / ms / 1e9, by / ms / 1e6),
              flush=True)


if __name__ == "__main__":
    main()
