"""DMA attention micro-benchmark through the C ABI (vpu_attention): the three shapes of a DMA layer (prompt self-attention,
tokens -> image, image -> tokens; reference transformer.py:499-521) for ViT-B / L at batch 64 and ViT-H at batch 32.
GPU box: python tools/dma_attn_bench.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvpuformer_b200 import ops  # noqa: E402
from tools.attn_bench import timeit  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for C, B, N in ((768, 64, 784), (1024, 64, 784), (1280, 32, 1024)):
        Q, H, Ci = 48, 8, C // 2
        ds, dc = C // H, Ci // H
        qk = torch.randn(B * Q, 2 * C, device=dev).to(torch.bfloat16)
        v = torch.randn(B * Q, C, device=dev).to(torch.bfloat16)
        tq = torch.randn(B * Q, Ci, device=dev).to(torch.bfloat16)
        kvq = torch.randn(B * N, 3 * Ci, device=dev).to(torch.bfloat16)
        ik = torch.randn(B * Q, Ci, device=dev).to(torch.bfloat16)
        iv = torch.randn(B * Q, Ci, device=dev).to(torch.bfloat16)
        for name, fn, fl, by in (
            ("self 48x48", lambda: ops.attention(qk, qk, v, Q, Q, H, ds, B, 1 / math.sqrt(ds), 0, C, 0), 4.0 * Q * Q * C * B, 8.0 * B * Q * C),
            ("t2i  48xN ", lambda: ops.attention(tq, kvq, kvq, Q, N, H, dc, B, 1 / math.sqrt(dc), 0, 0, Ci), 4.0 * Q * N * Ci * B,
             4.0 * B * N * Ci + 4.0 * B * Q * Ci),
            ("i2t  Nx48 ", lambda: ops.attention(kvq, ik, iv, N, Q, H, dc, B, 1 / math.sqrt(dc), 2 * Ci, 0, 0, out_cols=Ci), 4.0 * Q * N * Ci * B,
             4.0 * B * N * Ci + 4.0 * B * Q * Ci),
        ):
            def cold():
                flush.zero_()
                fn()
            ms = timeit(fn, 20)
            ms_cold = timeit(cold, 10) - timeit(lambda: flush.zero_(), 10)
            print("C=%4d B=%2d %s %7.1f us (L2-cold %7.1f us)  %6.1f TF/s  %7.1f GB/s algorithmic" %
                  (C, B, name, ms * 1e3, ms_cold * 1e3, fl / ms / 1e9, by / ms / 1e6), flush=True)


if __name__ == "__main__":
    main()
