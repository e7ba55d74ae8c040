"""Attention micro-benchmark through the C ABI (vpu_attention): ViT-B window / global and DMA shapes at batch 64.
Usage on the GPU box: python tools/attn_bench.py   (VPU_ATTN_TC=0 selects the mma.sync kernel for window attention)"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvpuformer_b200 import ops  # noqa: E402


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda:0")
    B, heads, hd, grid, win = 64, 12, 64, 28, 14
    N, C = grid * grid, heads * hd
    qkv = (torch.randn(B * N, 3 * C, device=dev)).to(torch.bfloat16)
    scale = hd ** -0.5
    ms = timeit(lambda: ops.attention(qkv, qkv, qkv, win * win, win * win, heads, hd, B * 4, scale, 0, C, 2 * C, window=win, grid=grid))
    fl = 4.0 * 196 * 196 * hd * heads * B * 4
    by = B * N * C * 2 * 4.0
    print("vit window  %8.1f us  %7.1f TF/s  %7.1f GB/s (algorithmic q,k,v,o)" % (ms * 1e3, fl / ms / 1e9, by / ms / 1e6))
    ms = timeit(lambda: ops.attention(qkv, qkv, qkv, N, N, heads, hd, B, scale, 0, C, 2 * C))
    fl = 4.0 * N * N * hd * heads * B
    print("vit global  %8.1f us  %7.1f TF/s  %7.1f GB/s" % (ms * 1e3, fl / ms / 1e9, by / ms / 1e6))
    if os.environ.get("ATTN_BENCH_VITH", "1") != "0":
        Bh, hh, dh, gh, wh = 32, 16, 80, 32, 16
        Nh, Ch = gh * gh, hh * dh
        qh = (torch.randn(Bh * Nh, 3 * Ch, device=dev)).to(torch.bfloat16)
        byh = Bh * Nh * Ch * 2 * 4.0
        ms = timeit(lambda: ops.attention(qh, qh, qh, wh * wh, wh * wh, hh, dh, Bh * 4, dh ** -0.5, 0, Ch, 2 * Ch, window=wh, grid=gh))
        fl = 4.0 * 256 * 256 * dh * hh * Bh * 4
        print("vit-h window %7.1f us  %7.1f TF/s  %7.1f GB/s" % (ms * 1e3, fl / ms / 1e9, byh / ms / 1e6))
        ms = timeit(lambda: ops.attention(qh, qh, qh, Nh, Nh, hh, dh, Bh, dh ** -0.5, 0, Ch, 2 * Ch))
        fl = 4.0 * Nh * Nh * dh * hh * Bh
        print("vit-h global %7.1f us  %7.1f TF/s  %7.1f GB/s" % (ms * 1e3, fl / ms / 1e9, byh / ms / 1e6))
        del qh
    Ci, Q, H = C // 2, 48, 8
    tq = torch.randn(B * Q, Ci, device=dev).to(torch.bfloat16)
    kvq = torch.randn(B * N, 3 * Ci, device=dev).to(torch.bfloat16)
    ms = timeit(lambda: ops.attention(tq, kvq, kvq, Q, N, H, Ci // H, B, 1 / math.sqrt(Ci // H), 0, 0, Ci))
    print("dma t2i     %8.1f us  %7.1f TF/s" % (ms * 1e3, 4.0 * Q * N * Ci * B / ms / 1e9))
    ik = torch.randn(B * Q, Ci, device=dev).to(torch.bfloat16)
    ms = timeit(lambda: ops.attention(kvq, ik, ik, N, Q, H, Ci // H, B, 1 / math.sqrt(Ci // H), 2 * Ci, 0, 0, out_cols=Ci))
    print("dma i2t     %8.1f us  %7.1f TF/s" % (ms * 1e3, 4.0 * Q * N * Ci * B / ms / 1e9))


if __name__ == "__main__":
    main()
