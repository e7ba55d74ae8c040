"""A/B of the image-side table projection (gemm_res.cu MODE_TAB against the generic epilogue) through the C ABI; the bf16-residual
out-projection is timed beside it.
Usage on the GPU box (debug library): VPU_LIB_PATH=pvpuformer_b200/libvpuformer_b200_debug.so VPU_GEMM_RES_MODES={0,1} python tools/res_modes_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvpuformer_b200 import ops  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    dev = torch.device("cuda:0")
    M = 50176
    g = torch.Generator(device=dev).manual_seed(0)
    for N, K in ((1152, 768), (768, 768)):
        A = (torch.randn(M, K, device=dev, generator=g) * 0.5).to(torch.bfloat16)
        W = (torch.randn(N, K, device=dev, generator=g) * 0.05).to(torch.bfloat16)
        tab = torch.randn(784 + 128, N, device=dev, generator=g)
        print("table  %dx%dx%d: %.1f us" % (M, N, K, timeit(lambda: ops.gemm_table(A, W, tab, 784))), flush=True)
    N, K = 768, 384
    A = (torch.randn(M, K, device=dev, generator=g) * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev, generator=g) * 0.05).to(torch.bfloat16)
    res = torch.randn(M, N, device=dev, generator=g).to(torch.bfloat16)
    bias = torch.randn(N, device=dev, generator=g)
    out = torch.empty(M, N, device=dev)
    print("res16  %dx%dx%d: %.1f us" % (M, N, K, timeit(lambda: ops.gemm(A, W, bias=bias, residual=res, out_dtype=torch.float32, out=out))), flush=True)


if __name__ == "__main__":
    main()
