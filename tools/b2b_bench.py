"""The head's back-to-back GEMM pair (vpu_gemm_b2b) at the batch-64 ViT-B shapes.  GPU box: python tools/b2b_bench.py [level]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvpuformer_b200 import ops  # noqa: E402
from tools.attn_bench import timeit  # noqa: E402

LEVELS = [(802816, 128), (200704, 256), (50176, 512), (12544, 1024)]


def main():
    dev = torch.device("cuda:0")
    which = [int(sys.argv[1])] if len(sys.argv) > 1 else range(4)
    for i in which:
        M, K1 = LEVELS[i]
        A = (torch.randn(M, K1, device=dev)).to(torch.bfloat16)
        W1 = (torch.randn(256, K1, device=dev) * K1 ** -0.5).to(torch.bfloat16)
        W2 = (torch.randn(256, 256, device=dev) / 16).to(torch.bfloat16)
        b1 = torch.randn(256, device=dev)
        ms = timeit(lambda: ops.gemm_b2b(A, W1, b1, W2), 10)
        fl = 2.0 * M * 256 * (K1 + 256)
        by = 2.0 * M * (K1 + 256)
        print("level %d M=%7d K1=%4d  %7.1f us  %7.1f TF/s  %7.1f GB/s" % (i, M, K1, ms * 1e3, fl / ms / 1e9, by / ms / 1e6), flush=True)


if __name__ == "__main__":
    main()
