"""GPU: each hand-written kernel, called through the C ABI, against the same op in plain torch fp32
(floating-point kernels) or the CPU oracle (integer / rasterisation kernels, bit-exact)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _rand_bf16(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).to(_dev())


GEMM_SHAPES = [
    # M, N, K            tile config exercised
    (128, 256, 64),      # one tile, one k-block, BN=256
    (256, 64, 128),      # BN=64
    (384, 768, 768),     # BN=256, 12 k-blocks (ring wraps)
    (1000, 384, 768),    # BN=192, ragged M
    (3072, 1152, 768),   # BN=192 (DMA image-side projection width)
    (777, 2304, 768),    # BN=256 ragged M, qkv width
    (96, 2048, 904),     # PPuE FFN: K=904 (ragged k-block), M < 128
    (512, 1280, 1176),   # ViT-H patch embed K, BN=256
    (640, 640, 3072),    # BN=128, long K
    (20000, 768, 768),   # > 148 tiles: persistent loop + TMEM double buffering
]


@pytest.mark.parametrize("impl", [0, 2])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_plain(M, N, K, impl):
    from pvpuformer_b200 import ops
    A, W = _rand_bf16((M, K), 1), _rand_bf16((N, K), 2, 0.05)
    bias = torch.randn(N, device=_dev())
    out = ops.gemm(A, W, bias=bias, out_dtype=torch.float32, impl=impl)
    ref = A.double() @ W.double().t() + bias.double()
    err = (out.double() - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("impl", [0, 2])
def test_gemm_epilogues(impl):
    from pvpuformer_b200 import ops
    M, N, K = 1568, 768, 256
    A, W = _rand_bf16((M, K), 3), _rand_bf16((N, K), 4, 0.1)
    bias = torch.randn(N, device=_dev())
    tab = torch.randn(784, N, device=_dev())
    base = A.double() @ W.double().t()
    # bias2d table (positional tables), fp32 out
    out = ops.gemm(A, W, bias2d=tab, out_dtype=torch.float32, impl=impl)
    ref = base + tab.double().repeat(2, 1)
    assert (out.double() - ref).abs().max().item() < 2e-3 * ref.abs().max().item()
    # GELU, bf16 out
    out = ops.gemm(A, W, bias=bias, act="gelu", impl=impl)
    ref = F.gelu((base + bias.double()).float())
    assert (out.float() - ref).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())
    # ReLU
    out = ops.gemm(A, W, bias=bias, act="relu", impl=impl)
    ref = F.relu((base + bias.double()).float())
    assert (out.float() - ref).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())
    # fp32 residual, in place (the ViT residual stream)
    x = torch.randn(M, N, device=_dev())
    ref = x.double() + base + bias.double()
    out = ops.gemm(A, W, bias=bias, residual=x, out_dtype=torch.float32, impl=impl, out=x)
    assert out.data_ptr() == x.data_ptr()
    assert (out.double() - ref).abs().max().item() < 2e-3 * ref.abs().max().item()
    # bf16 residual, fp32 out (DMA image-side residual)
    r = _rand_bf16((M, N), 5)
    out = ops.gemm(A, W, bias=bias, residual=r, out_dtype=torch.float32, impl=impl)
    ref = r.double() + base + bias.double()
    assert (out.double() - ref).abs().max().item() < 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("M,N,K,R", [(784 * 5 + 300, 1152, 768, 784), (1024 * 3, 1280, 1280, 1024), (700, 768, 768, 784)])
def test_gemm_table_tma_epilogue_matches_generic(M, N, K, R):
    """Image-side K|V|Q projection (reference transformer.py:444-449): the TMA-staged epilogue (gemm_res.cu MODE_TAB, padded
    periodic table, ragged last column tile) against fp64 and, bit for bit, against the generic 1-CTA kernel."""
    from pvpuformer_b200 import ops
    A, W = _rand_bf16((M, K), 31), _rand_bf16((N, K), 32, 0.05)
    tab = torch.randn(R, N, device=_dev())
    padded = torch.cat([tab, tab[:128]], 0).contiguous()
    out = ops.gemm_table(A, W, padded, R, impl=0)
    gen = ops.gemm_table(A, W, padded, R, impl=2)
    idx = torch.arange(M, device=_dev()) % R
    ref = A.double() @ W.double().t() + tab.double()[idx]
    assert (out.double() - ref).abs().max().item() < 1e-2 * ref.abs().max().item()
    assert torch.equal(out, gen)


@pytest.mark.parametrize("M", [50176 // 8, 1000])
def test_gemm_bf16_residual_2cta_matches_1cta(M):
    """Image -> tokens out-projection (reference transformer.py:459-461): fp32 out = A W^T + bias + bf16 residual, 2-CTA pairs
    and 1-CTA kernel bit-identical."""
    from pvpuformer_b200 import ops
    N, K = 768, 384
    A, W = _rand_bf16((M, K), 33), _rand_bf16((N, K), 34, 0.05)
    bias = torch.randn(N, device=_dev())
    res = _rand_bf16((M, N), 35)
    out = ops.gemm(A, W, bias=bias, residual=res, out_dtype=torch.float32, impl=0)
    gen = ops.gemm(A, W, bias=bias, residual=res, out_dtype=torch.float32, impl=2)
    ref = A.double() @ W.double().t() + bias.double() + res.double()
    assert (out.double() - ref).abs().max().item() < 2e-3 * ref.abs().max().item()
    assert torch.equal(out, gen)


@pytest.mark.parametrize("M,C,K", [(784 * 3, 768, 384), (1024 * 2 + 77, 1280, 640), (50176, 768, 384), (900, 1024, 512)])
def test_gemm_layernorm_cluster_matches_torch(M, C, K):
    """Image -> tokens out-projection + residual + norm4 (reference transformer.py:459-463) as one kernel whose CTAs exchange the
    row moments through distributed shared memory (csrc/gemm_ln.cu): against the same chain in torch fp32, in place over the
    residual like the forward uses it, and run to run."""
    from pvpuformer_b200 import ops
    A, W = _rand_bf16((M, K), 51), _rand_bf16((C, K), 52, K ** -0.5)
    g = torch.Generator().manual_seed(53)
    bias, gamma, beta = (torch.randn(C, generator=g).to(_dev()) for _ in range(3))
    res = _rand_bf16((M, C), 54, 2.0)
    out, parts = ops.gemm_layernorm(A, W, bias, res, gamma, beta)
    x = A.float() @ W.float().t() + bias + res.float()
    ref = F.layer_norm(x, (C,), gamma, beta, 1e-5)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())
    assert (parts.max(0).values - ref.max(1).values).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())
    inplace = res.clone()
    out2, parts2 = ops.gemm_layernorm(A, W, bias, inplace, gamma, beta, out=inplace)
    assert torch.equal(out2, out) and torch.equal(parts2, parts)


@pytest.mark.parametrize("M,K1", [(256, 128), (12544, 1024), (3136 * 3, 256), (50176, 512), (802816 // 8, 128)])
def test_gemm_b2b_head_pair(M, K1):
    """Back-to-back head GEMM (csrc/gemm_b2b.cu): conv 1x1 + ReLU + fusion-conv slice of one pyramid level with the
    intermediate in shared memory, against the same two products in fp32 with the intermediate rounded to bf16 (what
    the two separate GEMMs store).  M covers one tile, a ragged last tile (9408 = 36.75 x 256) and many tiles per CTA pair."""
    from pvpuformer_b200 import ops
    A = _rand_bf16((M, K1), 31)
    W1, W2 = _rand_bf16((256, K1), 32, K1 ** -0.5), _rand_bf16((256, 256), 33, 1 / 16)
    b1 = torch.randn(256, generator=torch.Generator().manual_seed(34)).to(_dev())
    out = ops.gemm_b2b(A, W1, b1, W2)
    h = torch.relu(A.float() @ W1.float().t() + b1).to(torch.bfloat16).float()
    ref = h @ W2.float().t()
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), err
    assert torch.equal(out, ops.gemm_b2b(A, W1, b1, W2))     # run to run


@pytest.mark.parametrize("B,R,aux", [(1, 112, True), (3, 112, True), (5, 112, False), (2, 128, True), (151, 112, True)])
def test_head_tail_resize_sum_seg_and_cosine_logits(B, R, aux):
    """Fused head tail (csrc/head_tail.cu): relu(b + y0 + sum_l resize(y_l)) -> conv_seg and P2CL cosine logits, with the
    bilinear resize (align_corners=False, swin_transformer.py:727-737) done as a tensor-core product with precomputed tap
    matrices, against F.interpolate in fp32.  B = 3 / 151 put image boundaries inside a CTA's tile range (query reload),
    R = 128 is the ViT-H geometry, more tiles than 2 x SMs exercise the accumulator double buffering."""
    from pvpuformer_b200 import ops
    ys = [_rand_bf16((B, R >> l, R >> l, 256), 40 + l) for l in range(4)]
    g = torch.Generator().manual_seed(45)
    bias, wseg = torch.randn(256, generator=g).to(_dev()), (torch.randn(256, generator=g) / 16).to(_dev())
    qn = torch.zeros(B, 64, 256)
    qn[:, :48] = F.normalize(torch.randn(B, 48, 256, generator=g), dim=2)
    qn = qn.to(torch.bfloat16).to(_dev())
    seg, auxo = ops.head_tail(ys, bias, wseg, 0.25, qn if aux else None)
    f = bias.view(1, 256, 1, 1) + ys[0].float().permute(0, 3, 1, 2)
    for l in range(1, 4):
        f = f + F.interpolate(ys[l].float().permute(0, 3, 1, 2), size=(R, R), mode="bilinear", align_corners=False)
    f = torch.relu(f)
    seg_ref = (f * wseg.view(1, 256, 1, 1)).sum(1) + 0.25
    assert torch.isfinite(seg).all()
    assert (seg - seg_ref).abs().max().item() < 1e-3 * max(1.0, seg_ref.abs().max().item())
    if aux:
        fn = F.normalize(f, dim=1).to(torch.bfloat16).float()      # the kernel rounds f to bf16 for the product, like the stored F did
        ref = (torch.einsum("bnc,bchw->bnhw", qn[:, :48].float(), F.normalize(f, dim=1)) + 1) / 2
        assert torch.isfinite(auxo).all()
        assert (auxo - ref).abs().max().item() < 4e-3
        del fn
    else:
        assert auxo is None
    seg2, aux2 = ops.head_tail(ys, bias, wseg, 0.25, qn if aux else None)
    assert torch.equal(seg, seg2) and (not aux or torch.equal(auxo, aux2))


@pytest.mark.parametrize("impl", [0, 2])
def test_gemm_pixel_shuffle_matches_conv_transpose(impl):
    from pvpuformer_b200 import ops
    B, g, cin, cout = 2, 28, 256, 192
    x = _rand_bf16((B, g, g, cin), 6)                       # NHWC
    w = (torch.randn(cin, cout, 2, 2, generator=torch.Generator().manual_seed(7)) * 0.05).to(_dev())
    b = torch.randn(cout, device=_dev())
    wp = w.permute(2, 3, 1, 0).reshape(4 * cout, cin).to(torch.bfloat16).contiguous()
    out = ops.gemm_pixel_shuffle(x.view(-1, cin), wp, b.repeat(4).contiguous(), g, cout, impl=impl)
    ref = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), b, stride=2).permute(0, 2, 3, 1)
    assert out.shape == ref.shape
    assert (out.float() - ref).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("B,g,cin,cout", [(3, 28, 768, 384), (2, 56, 384, 192), (2, 32, 1280, 640), (1, 64, 640, 320)])
def test_gemm_pixel_shuffle_tma_store_matches_generic(B, g, cin, cout):
    """Neck ConvTranspose2d(2, 2) shapes (is_vpu_model.py:57-75; ViT-B g = 28 / 56, ViT-H g = 32 / 64): the 5-D TMA pixel-shuffle
    store of gemm_gn.cu (whole grid rows per CTA: 112 of 128 accumulator rows at g = 28 / 56) bit for bit against the 1-CTA kernel."""
    from pvpuformer_b200 import ops
    x = _rand_bf16((B, g, g, cin), 36)
    wp = _rand_bf16((4 * cout, cin), 37, 0.05)
    b4 = torch.randn(cout, device=_dev()).repeat(4).contiguous()
    out = ops.gemm_pixel_shuffle(x.view(-1, cin), wp, b4, g, cout, impl=0)
    gen = ops.gemm_pixel_shuffle(x.view(-1, cin), wp, b4, g, cout, impl=2)
    y = (x.view(-1, cin).double() @ wp.double().t() + b4.double()).view(B, g, g, 2, 2, cout)
    ref = y.permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * g, 2 * g, cout)
    assert (out.double() - ref).abs().max().item() < 1e-2 * ref.abs().max().item()
    assert torch.equal(out, gen)


def _attn_ref(q, k, v, scale):
    a = torch.softmax((q.float() @ k.float().transpose(-1, -2)) * scale, dim=-1)
    return a @ v.float()


@pytest.mark.parametrize("heads,hd,grid,win", [(12, 64, 28, 14), (16, 80, 32, 16)])
def test_attention_vit_window_and_global(heads, hd, grid, win):
    from pvpuformer_b200 import ops
    B, N, C = 2, grid * grid, heads * hd
    qkv = _rand_bf16((B * N, 3 * C), 8)
    scale = hd ** -0.5
    t = qkv.view(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)         # [3,B,h,N,d]
    # global
    o = ops.attention(qkv, qkv, qkv, N, N, heads, hd, B, scale, 0, C, 2 * C)
    ref = _attn_ref(t[0], t[1], t[2], scale).transpose(1, 2).reshape(B * N, C)
    assert (o.float() - ref).abs().max().item() < 2e-2
    # windowed: reference patchify -> attention per window -> unpatchify (models_vit.py:225-255)
    nw = grid // win

    def part(x):      # [B,h,N,d] -> [B*nw*nw, h, win*win, d]
        x = x.reshape(B, heads, nw, win, nw, win, hd).permute(0, 2, 4, 1, 3, 5, 6)
        return x.reshape(B * nw * nw, heads, win * win, hd)
    refw = _attn_ref(part(t[0]), part(t[1]), part(t[2]), scale)      # [B*nw2, h, S, d]
    refw = refw.reshape(B, nw, nw, heads, win, win, hd).permute(0, 1, 4, 2, 5, 3, 6).reshape(B * N, C)
    o = ops.attention(qkv, qkv, qkv, win * win, win * win, heads, hd, B * nw * nw, scale, 0, C, 2 * C, window=win, grid=grid)
    assert (o.float() - refw).abs().max().item() < 2e-2


def test_window_attention_tcgen05_many_problems():
    """ViT-B window attention (14x14 windows, d=64) on the tcgen05/TMEM kernel with more problems than SMs, so
    every CTA runs its software-pipelined multi-problem loop (both smem stages, both TMEM row tiles reused)."""
    from pvpuformer_b200 import ops
    heads, hd, grid, win, B = 12, 64, 28, 14, 24
    N, C = grid * grid, heads * hd
    qkv = _rand_bf16((B * N, 3 * C), 21, 2.0)                         # larger logits: exercises the max subtraction
    scale = hd ** -0.5
    t = qkv.view(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    nw = grid // win

    def part(x):
        x = x.reshape(B, heads, nw, win, nw, win, hd).permute(0, 2, 4, 1, 3, 5, 6)
        return x.reshape(B * nw * nw, heads, win * win, hd)
    refw = _attn_ref(part(t[0]), part(t[1]), part(t[2]), scale)
    refw = refw.reshape(B, nw, nw, heads, win, win, hd).permute(0, 1, 4, 2, 5, 3, 6).reshape(B * N, C)
    o = ops.attention(qkv, qkv, qkv, win * win, win * win, heads, hd, B * nw * nw, scale, 0, C, 2 * C, window=win, grid=grid)
    torch.cuda.synchronize()
    assert not torch.isnan(o.float()).any()
    # outputs reach |3|: one bf16 ulp there is 1.6e-2, so the bound scales with the output magnitude
    assert (o.float() - refw).abs().max().item() < 1e-2 * max(1.0, refw.abs().max().item())


@pytest.mark.parametrize("B,amp", [(1, 1.0), (12, 2.0)])
def test_window_attention_tcgen05_vit_huge(B, amp):
    """ViT-H window attention (16x16 windows, d=80 = a 128-byte-swizzled 64-column part + a 32-byte-swizzled 16-column
    part per operand) on the tcgen05/TMEM kernel: fewer problems than SMs (B=1) and the multi-problem loop with the
    split Q/K and V rings and the shared output staging tile (B=12: 768 problems).  On failure the assertion message
    carries the per-head-dim-column error, which separates the 64-column from the 16-column part."""
    from pvpuformer_b200 import ops
    heads, hd, grid, win = 16, 80, 32, 16
    N, C = grid * grid, heads * hd
    qkv = _rand_bf16((B * N, 3 * C), 41 + B, amp)
    scale = hd ** -0.5
    t = qkv.view(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    nw = grid // win

    def part(x):
        x = x.reshape(B, heads, nw, win, nw, win, hd).permute(0, 2, 4, 1, 3, 5, 6)
        return x.reshape(B * nw * nw, heads, win * win, hd)
    refw = _attn_ref(part(t[0]), part(t[1]), part(t[2]), scale)
    refw = refw.reshape(B, nw, nw, heads, win, win, hd).permute(0, 1, 4, 2, 5, 3, 6).reshape(B * N, C)
    o = ops.attention(qkv, qkv, qkv, win * win, win * win, heads, hd, B * nw * nw, scale, 0, C, 2 * C, window=win, grid=grid)
    torch.cuda.synchronize()
    assert not torch.isnan(o.float()).any()
    err = (o.float() - refw).abs()
    assert err.max().item() < 1e-2 * max(1.0, refw.abs().max().item()), (err.max().item(), err.view(B * N, heads, hd).amax((0, 1)))


@pytest.mark.parametrize("B,amp", [(3, 1.0), (5, 6.0)])
def test_global_attention_tcgen05_vit_huge(B, amp):
    """ViT-H global attention (S=1024, d=80) on the tcgen05 flash kernel with the 64 + 16 column operand split and 16 key
    blocks of 64; more work units than SMs, and (amp=6) block maxima far enough apart for the lazy O / l rescale."""
    from pvpuformer_b200 import ops
    heads, hd, N = 16, 80, 1024
    C = heads * hd
    qkv = _rand_bf16((B * N, 3 * C), 51 + B, amp)
    if amp > 1.0:
        qkv.view(B, N, 3 * C)[:, : N // 2, C:2 * C] *= 0.05
    scale = hd ** -0.5
    t = qkv.view(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    ref = _attn_ref(t[0], t[1], t[2], scale).transpose(1, 2).reshape(B * N, C)
    o = ops.attention(qkv, qkv, qkv, N, N, heads, hd, B, scale, 0, C, 2 * C)
    torch.cuda.synchronize()
    assert not torch.isnan(o.float()).any()
    err = (o.float() - ref).abs()
    assert err.max().item() < 1e-2 * max(1.0, ref.abs().max().item()), (err.max().item(), err.view(B * N, heads, hd).amax((0, 1)))


@pytest.mark.parametrize("heads,B,amp", [(12, 20, 1.0), (16, 5, 6.0)])
def test_global_attention_tcgen05_many_problems(heads, B, amp):
    """ViT-B / ViT-L global attention (S=784, d=64) on the tcgen05 flash kernel with more work units than SMs (every
    CTA loops over units: Q double buffer, K/V ring wrap, single-tile last pairs between two-tile units).  amp=6
    gives score blocks whose maxima differ by far more than 2^8, so the lazy O / l rescale path runs."""
    from pvpuformer_b200 import ops
    hd, N = 64, 784
    C = heads * hd
    qkv = _rand_bf16((B * N, 3 * C), 33, amp)
    if amp > 1.0:     # the late keys carry the large scores: the running maximum has to be raised mid-stream
        qkv.view(B, N, 3 * C)[:, : N // 2, C:2 * C] *= 0.05
    scale = hd ** -0.5
    t = qkv.view(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    ref = _attn_ref(t[0], t[1], t[2], scale).transpose(1, 2).reshape(B * N, C)
    o = ops.attention(qkv, qkv, qkv, N, N, heads, hd, B, scale, 0, C, 2 * C)
    torch.cuda.synchronize()
    assert not torch.isnan(o.float()).any()
    assert (o.float() - ref).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("C", [768, 1024, 1280])
def test_attention_dma_shapes(C):
    from pvpuformer_b200 import ops
    B, N, Q, H = 3, 784 if C != 1280 else 1024, 48, 8
    Ci = C // 2
    d_self, d_cross = C // H, Ci // H
    # prompt self-attention 48 x 48
    qk, v = _rand_bf16((B * Q, 2 * C), 9), _rand_bf16((B * Q, C), 10)
    o = ops.attention(qk, qk, v, Q, Q, H, d_self, B, 1 / math.sqrt(d_self), 0, C, 0)
    r = lambda t, n, d: t.reshape(B, n, H, d).transpose(1, 2)
    ref = _attn_ref(r(qk[:, :C], Q, d_self), r(qk[:, C:], Q, d_self), r(v, Q, d_self), 1 / math.sqrt(d_self))
    assert (o.float() - ref.transpose(1, 2).reshape(B * Q, C)).abs().max().item() < 2e-2
    # tokens -> image: 48 queries x N keys, K|V|Q fused buffer of width 3*Ci
    tq, kvq = _rand_bf16((B * Q, Ci), 11), _rand_bf16((B * N, 3 * Ci), 12)
    o = ops.attention(tq, kvq, kvq, Q, N, H, d_cross, B, 1 / math.sqrt(d_cross), 0, 0, Ci)
    ref = _attn_ref(r(tq, Q, d_cross), r(kvq[:, :Ci], N, d_cross), r(kvq[:, Ci:2 * Ci], N, d_cross), 1 / math.sqrt(d_cross))
    assert (o.float() - ref.transpose(1, 2).reshape(B * Q, Ci)).abs().max().item() < 2e-2
    # image -> tokens: N queries x 48 keys
    ik, iv = _rand_bf16((B * Q, Ci), 13), _rand_bf16((B * Q, Ci), 14)
    o = ops.attention(kvq, ik, iv, N, Q, H, d_cross, B, 1 / math.sqrt(d_cross), 2 * Ci, 0, 0, out_cols=Ci)
    ref = _attn_ref(r(kvq[:, 2 * Ci:], N, d_cross), r(ik, Q, d_cross), r(iv, Q, d_cross), 1 / math.sqrt(d_cross))
    assert (o.float() - ref.transpose(1, 2).reshape(B * N, Ci)).abs().max().item() < 2e-2


@pytest.mark.parametrize("C,B,amp", [(768, 70, 1.0), (768, 5, 6.0), (1024, 41, 3.0), (1280, 33, 3.0), (1280, 1, 6.0)])
def test_attention_dma_tcgen05_many_problems(C, B, amp):
    """The three DMA attention shapes (transformer.py:499-521) on the tcgen05 kernel (csrc/attention_dma.cu) with more
    (image, head) units than SMs, odd batch sizes and logits of a few hundred (amp): the persistent loops wrap, the lazy
    online-softmax rescale of the tokens -> image shape is taken, and the zero-filled query rows past 48 / past the image
    end must not leak into a neighbouring image."""
    from pvpuformer_b200 import ops
    N, Q, H = 784 if C != 1280 else 1024, 48, 8
    Ci = C // 2
    d_self, d_cross = C // H, Ci // H
    r = lambda t, n, d: t.reshape(B, n, H, d).transpose(1, 2)

    def check(o, ref, rows, cols):
        ref = ref.transpose(1, 2).reshape(rows, cols)
        assert torch.isfinite(o.float()).all()
        assert (o.float() - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())
    qk, v = _rand_bf16((B * Q, 2 * C), 21, amp), _rand_bf16((B * Q, C), 22)
    o = ops.attention(qk, qk, v, Q, Q, H, d_self, B, 1 / math.sqrt(d_self), 0, C, 0)
    check(o, _attn_ref(r(qk[:, :C], Q, d_self), r(qk[:, C:], Q, d_self), r(v, Q, d_self), 1 / math.sqrt(d_self)), B * Q, C)
    tq, kvq = _rand_bf16((B * Q, Ci), 23, amp), _rand_bf16((B * N, 3 * Ci), 24)
    kvq[:, :Ci] *= amp
    kvq[:, 2 * Ci:] *= amp
    o = ops.attention(tq, kvq, kvq, Q, N, H, d_cross, B, 1 / math.sqrt(d_cross), 0, 0, Ci)
    check(o, _attn_ref(r(tq, Q, d_cross), r(kvq[:, :Ci], N, d_cross), r(kvq[:, Ci:2 * Ci], N, d_cross), 1 / math.sqrt(d_cross)),
          B * Q, Ci)
    ik, iv = _rand_bf16((B * Q, Ci), 25, amp), _rand_bf16((B * Q, Ci), 26)
    o = ops.attention(kvq, ik, iv, N, Q, H, d_cross, B, 1 / math.sqrt(d_cross), 2 * Ci, 0, 0, out_cols=Ci)
    check(o, _attn_ref(r(kvq[:, 2 * Ci:], N, d_cross), r(ik, Q, d_cross), r(iv, Q, d_cross), 1 / math.sqrt(d_cross)), B * N, Ci)


@pytest.mark.parametrize("C", [768, 1024, 1280])
def test_layernorm(C):
    from pvpuformer_b200 import ops
    rows = 1003
    x = torch.randn(rows, C, device=_dev()) * 3 + 0.5
    g, b, pe = torch.randn(C, device=_dev()), torch.randn(C, device=_dev()), torch.randn(rows, C, device=_dev())
    of, ob, ope, rm = ops.layernorm(x, g, b, 1e-6, pe=pe, want_rowmax=True)
    ref = F.layer_norm(x, (C,), g, b, 1e-6)
    assert (of - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    assert torch.equal(ob, of.to(torch.bfloat16))
    assert (ope.float() - (ref + pe)).abs().max().item() < 1e-2 * (ref + pe).abs().max().item()
    assert (rm - ref.max(dim=1).values).abs().max().item() < 1e-4 * ref.abs().max().item()


@pytest.mark.parametrize("gelu", [0, 1])
def test_groupnorm_nhwc(gelu):
    from pvpuformer_b200 import ops
    B, H, C = 3, 56, 384
    x = _rand_bf16((B, H, H, C), 15, 2.0) + 0.25
    g, b = torch.randn(C, device=_dev()), torch.randn(C, device=_dev())
    ref = F.group_norm(x.float().permute(0, 3, 1, 2), 1, g, b, 1e-5)
    if gelu:
        ref = F.gelu(ref)
    ref = ref.permute(0, 2, 3, 1)
    out = ops.groupnorm_nhwc_(x.clone(), g, b, gelu)
    assert (out.float() - ref).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())


def test_upsample_align_corners():
    from pvpuformer_b200 import ops
    x = torch.randn(2, 5, 112, 112, device=_dev())
    out = ops.upsample_align_corners(x, 448, 448)
    ref = F.interpolate(x, size=(448, 448), mode="bilinear", align_corners=True)
    assert (out - ref).abs().max().item() < 1e-5
    x = torch.randn(1, 3, 128, 128, device=_dev())
    out = ops.upsample_align_corners(x, 448, 448)
    ref = F.interpolate(x, size=(448, 448), mode="bilinear", align_corners=True)
    assert (out - ref).abs().max().item() < 1e-5


def test_prompt_rasteriser_bitexact_vs_cv2():
    """csrc/raster.cu against cv2.rectangle / cv2.polylines (thickness 3) -- the calls of reference is_model.py:109,129 -- for
    vertices inside the image (long and short segments, repeated points, corners, borders) and outside it."""
    import cv2
    from pvpuformer_b200 import host_prompts, ops
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(5)
    for size in (448, 97):
        # boxes: random corners, degenerate (zero) boxes, boxes touching the border
        B = 64
        xc, yc = rng.integers(0, size, B), rng.integers(0, size, B)
        w = np.minimum(rng.integers(0, size, B), 2 * np.minimum(xc, size - 1 - xc))
        h = np.minimum(rng.integers(0, size, B), 2 * np.minimum(yc, size - 1 - yc))
        boxes = np.stack([xc, yc, w, h, rng.integers(0, 8, B)], axis=1).astype(np.int32)
        boxes[0] = 0
        boxes[1] = [size // 2, size // 2, size - 1 - (size - 1) % 2, size - 1 - (size - 1) % 2, 5]
        n = 4
        got = ops.raster_prompts(1, torch.from_numpy(boxes).to(dev), None, n, B, size).cpu().numpy()
        want = host_prompts.raster_planes(1, boxes, None, n, B, size)            # the reference's cv2.rectangle calls
        for b in range(B):
            assert np.array_equal(got[b], want[b]), (size, b, boxes[b], int((got[b] != want[b]).sum()))
        # boxes and curves that leave the image (cv2 clips each segment to the image grown by the thickness first)
        far = np.stack([rng.integers(-60, size + 60, B), rng.integers(-60, size + 60, B), rng.integers(0, 2 * size, B),
                        rng.integers(0, 2 * size, B), rng.integers(0, 8, B)], axis=1).astype(np.int32)
        got = ops.raster_prompts(1, torch.from_numpy(far).to(dev), None, n, B, size).cpu().numpy()
        want = host_prompts.raster_planes(1, far, None, n, B, size)
        assert np.array_equal(got, want), (size, int((got != want).sum()))
        wild = rng.integers(-60, size + 60, (B, 40, 2)).astype(np.int32)
        got = ops.raster_prompts(2, None, torch.from_numpy(wild).to(dev), n, B, size).cpu().numpy()
        for b in range(B):
            ref = np.zeros((2, size, size), np.uint8)
            cv2.polylines(ref[0], [wild[b]], False, 1, 3)
            assert np.array_equal(got[b], ref), (size, b, int((got[b] != ref).sum()))
        # polylines: smooth curves sampled densely (the reference's 1000-point scribbles), random walks, far-apart points
        B, S = 24, 200
        curves = np.zeros((B, S, 2), np.int32)
        t = np.linspace(0, 1, S)
        for b in range(B):
            if b % 3 == 0:
                c = rng.uniform(0, size - 1, (4, 2))
                pts = ((1 - t)[:, None] ** 3 * c[0] + 3 * ((1 - t) ** 2 * t)[:, None] * c[1] + 3 * ((1 - t) * t ** 2)[:, None] * c[2] + (t ** 3)[:, None] * c[3])
            elif b % 3 == 1:
                pts = np.cumsum(rng.integers(-3, 4, (S, 2)), axis=0) + size // 2
            else:
                pts = rng.uniform(0, size - 1, (S, 2))
            curves[b] = np.clip(pts, 0, size - 1).astype(np.int32)
        curves[3, :, :] = curves[3, :1, :]                       # one point repeated: circles only
        got = ops.raster_prompts(2, None, torch.from_numpy(curves).to(dev), 4, B, size).cpu().numpy()
        for b in range(B):
            ref = np.zeros((2, size, size), np.uint8)
            cv2.polylines(ref[0], [curves[b]], False, 1, 3)
            assert np.array_equal(got[b], ref), (size, b, int((got[b] != ref).sum()))
