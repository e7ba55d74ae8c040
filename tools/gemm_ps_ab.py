"""Pixel-shuffle GEMM epilogue against the plain bf16 epilogue on the neck's ConvTranspose shapes (no statistics in either)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvpuformer_b200 import ops  # noqa: E402


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 * 1e3


def main():
    dev = torch.device("cuda:0")
    for name, B, g, K, cout in (("d4.a", 64, 28, 768, 384), ("d4.b", 64, 56, 384, 192), ("d8.a", 64, 28, 768, 384)):
        M, N = B * g * g, 4 * cout
        A = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
        W = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        t_plain = timeit(lambda: ops.gemm(A, W, bias=bias, out=out))
        t_ps = timeit(lambda: ops.gemm_pixel_shuffle(A, W, bias, g, cout))
        fl = 2.0 * M * N * K
        print("%-5s M=%6d N=%4d K=%3d  plain %6.1f us (%6.1f TF/s)   pixel-shuffle %6.1f us (%6.1f TF/s)" %
              (name, M, N, K, t_plain, fl / t_plain / 1e6, t_ps, fl / t_ps / 1e6), flush=True)


if __name__ == "__main__":
    main()
