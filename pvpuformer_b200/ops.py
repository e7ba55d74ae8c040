"""Thin torch-tensor wrappers over the kernel-level C-ABI entry points (used by tests and tools).
Every function launches the hand-written CUDA kernel on the current stream; none has a fallback."""
import ctypes

import torch

from . import lib as L

_ACT = {None: 0, "none": 0, "gelu": 1, "relu": 2}


def _dt(t):
    return L.VPU_BF16 if t.dtype == torch.bfloat16 else L.VPU_F32


def gemm(A, W, bias=None, bias2d=None, residual=None, act=None, out_dtype=torch.bfloat16, impl=0, out=None):
    """out = act(A @ W.T + bias + bias2d[m % rows] + residual); A [M,K] bf16, W [N,K] bf16."""
    assert A.is_cuda and A.dtype == torch.bfloat16 and W.dtype == torch.bfloat16
    M, K = A.shape
    N = W.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=A.device)
    L.check(L.load().vpu_gemm(L.ptr(A), A.stride(0), L.ptr(W), W.stride(0), M, N, K, L.ptr(bias), L.ptr(bias2d),
                              0 if bias2d is None else bias2d.shape[0], L.ptr(residual),
                              0 if residual is None else _dt(residual), 0 if residual is None else residual.stride(0),
                              _ACT[act], L.ptr(out), _dt(out), out.stride(0), impl, L.current_stream()))
    return out


def gemm_table(A, W, table, table_rows, impl=0):
    """bf16(A @ W.T + table[m % table_rows]); table: fp32 [table_rows + pad, N] with its first `pad` rows repeated at the end."""
    M, K = A.shape
    N = W.shape[0]
    out = torch.empty(M, N, dtype=torch.bfloat16, device=A.device)
    L.check(L.load().vpu_gemm_table(L.ptr(A), A.stride(0), L.ptr(W), W.stride(0), M, N, K, L.ptr(table), table_rows,
                                    table.shape[0] - table_rows, L.ptr(out), out.stride(0), impl, L.current_stream()))
    return out


def gemm_layernorm(A, W, bias, res, gamma, beta, eps=1e-5, want_rowmax=True, out=None):
    """bf16(LayerNorm(A @ W.T + bias + res) * gamma + beta) and the per-256-column-block row maxima [C // 256, M]."""
    M, K = A.shape
    C = W.shape[0]
    if out is None:
        out = torch.empty(M, C, dtype=torch.bfloat16, device=A.device)
    parts = torch.empty(C // 256, M, dtype=torch.float32, device=A.device) if want_rowmax else None
    L.check(L.load().vpu_gemm_layernorm(L.ptr(A), A.stride(0), L.ptr(W), W.stride(0), L.ptr(bias), L.ptr(res), res.stride(0),
                                        L.ptr(gamma), L.ptr(beta), eps, M, K, C, L.ptr(out), out.stride(0), L.ptr(parts),
                                        L.current_stream()))
    return out, parts


def gemm_b2b(A, W1, bias1, W2):
    """bf16( bf16(relu(A @ W1.T + bias1)) @ W2.T ): the head's conv + ReLU + fusion-conv slice of one pyramid level."""
    M, K1 = A.shape
    out = torch.empty(M, W2.shape[0], dtype=torch.bfloat16, device=A.device)
    L.check(L.load().vpu_gemm_b2b(L.ptr(A), A.stride(0), L.ptr(W1), L.ptr(bias1), L.ptr(W2), M, K1, L.ptr(out), out.stride(0),
                                  L.current_stream()))
    return out


def head_tail(ys, bias, wseg, seg_bias, qn=None, nq=48):
    """ys: four NHWC bf16 maps [B, R >> l, R >> l, 256]; -> (seg [B, R, R] fp32, aux [B, nq, R, R] fp32 or None)."""
    B, R = ys[0].shape[0], ys[0].shape[1]
    seg = torch.empty(B, R, R, dtype=torch.float32, device=ys[0].device)
    aux = torch.empty(B, nq, R, R, dtype=torch.float32, device=ys[0].device) if qn is not None else None
    L.check(L.load().vpu_head_tail(L.ptr(ys[0]), L.ptr(ys[1]), L.ptr(ys[2]), L.ptr(ys[3]), B, R, L.ptr(bias), L.ptr(wseg),
                                   float(seg_bias), L.ptr(qn), nq, L.ptr(seg), L.ptr(aux), L.current_stream()))
    return seg, aux


def gemm_pixel_shuffle(A, W, bias4, g, cout, impl=0):
    """ConvTranspose2d(k=2,s=2): A [B*g*g, K] bf16, W [4*cout, K] -> NHWC [B, 2g, 2g, cout] bf16."""
    M, K = A.shape
    B = M // (g * g)
    out = torch.empty(B, 2 * g, 2 * g, cout, dtype=torch.bfloat16, device=A.device)
    L.check(L.load().vpu_gemm_pixel_shuffle(L.ptr(A), L.ptr(W), M, cout, K, L.ptr(bias4), g, L.ptr(out), impl,
                                            L.current_stream()))
    return out


def attention(q, k, v, Sq, Sk, heads, head_dim, nprob, scale, qoff=0, koff=0, voff=0, window=0, grid=0, out_cols=None):
    """q/k/v: 2-D bf16 token-major buffers (rows, ld); returns o [rows_q, heads*head_dim] bf16."""
    cols = out_cols or heads * head_dim
    o = torch.zeros(q.shape[0], cols, dtype=torch.bfloat16, device=q.device)
    L.check(L.load().vpu_attention(L.ptr(q), q.stride(0), qoff, L.ptr(k), k.stride(0), koff, L.ptr(v), v.stride(0), voff,
                                   L.ptr(o), o.stride(0), Sq, Sk, heads, head_dim, nprob, float(scale), window, grid,
                                   L.current_stream()))
    return o


def layernorm(x, gamma, beta, eps, pe=None, want_rowmax=False):
    rows, C = x.shape
    of = torch.empty_like(x)
    ob = torch.empty(rows, C, dtype=torch.bfloat16, device=x.device)
    ope = torch.empty(rows, C, dtype=torch.bfloat16, device=x.device) if pe is not None else None
    rm = torch.empty(rows, dtype=torch.float32, device=x.device) if want_rowmax else None
    L.check(L.load().vpu_layernorm(L.ptr(x), L.ptr(gamma), L.ptr(beta), float(eps), rows, C, L.ptr(of), L.ptr(ob),
                                   L.ptr(pe), L.ptr(ope), L.ptr(rm), L.current_stream()))
    return of, ob, ope, rm


def groupnorm_nhwc_(x, gamma, beta, gelu):
    """In-place GroupNorm(1, C) (+GELU) on NHWC bf16 [B, ..., C]."""
    B, C = x.shape[0], x.shape[-1]
    per = x[0].numel()
    scratch = torch.empty(B * 8200, dtype=torch.uint8, device=x.device)
    L.check(L.load().vpu_groupnorm_nhwc(L.ptr(x), B, per, C, L.ptr(gamma), L.ptr(beta), int(gelu), L.ptr(scratch),
                                        L.current_stream()))
    return x


def upsample_align_corners(x, H, W):
    planes = x.numel() // (x.shape[-1] * x.shape[-2])
    out = torch.empty(*x.shape[:-2], H, W, dtype=torch.float32, device=x.device)
    L.check(L.load().vpu_upsample_align_corners(L.ptr(x), L.ptr(out), x.shape[-2], x.shape[-1], H, W, planes,
                                                L.current_stream()))
    return out


def noc_next_clicks(gt, pred, not_clicked, workspace=None):
    """Device-side oracle clicker + IoU counts (csrc/noc.cu; reference inference/clicker.py:29-69, inference/utils.py:80-87).
    gt int8 [S,H,W] (1 object / 0 background / -1 ignore), pred uint8 [S,H,W], not_clicked uint8 [S,H,W] (updated in place)
    -> (clicks int32 [S,4] = is_positive,row,col,d2 ; counts int64 [S,2] = intersection, union), both on the device."""
    S, H, W = gt.shape
    assert gt.dtype == torch.int8 and pred.dtype == torch.uint8 and not_clicked.dtype == torch.uint8
    assert pred.shape == gt.shape and not_clicked.shape == gt.shape and gt.is_contiguous() and pred.is_contiguous() and not_clicked.is_contiguous()
    lib = L.load()
    need = lib.vpu_noc_workspace_bytes(S, H, W)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=gt.device)
    clicks = torch.empty(S, 4, dtype=torch.int32, device=gt.device)
    counts = torch.empty(S, 2, dtype=torch.int64, device=gt.device)
    L.check(lib.vpu_noc_next_clicks(L.ptr(gt), L.ptr(pred), L.ptr(not_clicked), S, H, W, L.ptr(clicks), L.ptr(counts), L.ptr(workspace),
                                    workspace.numel(), L.current_stream()))
    return clicks, counts


def raster_prompts(as_prompt_type, boxes, scribbles, n, B, size=448, device=None):
    """Box / scribble outline planes uint8 [B,2,size,size] (csrc/raster.cu; reference is_model.py:97-146 through cv2).
    boxes: int32 cuda [B,5]; scribbles: int32 cuda [B,S,2] (x, y); vertices may lie outside the image (clipped as cv2 does)."""
    dev = device or (boxes.device if as_prompt_type == 1 else scribbles.device)
    planes = torch.empty(B, 2, size, size, dtype=torch.uint8, device=dev)
    S = 0 if as_prompt_type == 1 else scribbles.shape[1]
    L.check(L.load().vpu_raster_prompts(int(as_prompt_type), L.ptr(boxes) if as_prompt_type == 1 else None,
                                        L.ptr(scribbles) if as_prompt_type == 2 else None, S, n, B, size, L.ptr(planes),
                                        L.current_stream()))
    return planes


def image_from_u8(rgb_nhwc, prev_mask=None, out=None):
    """ToTensor + cat(prev mask) of the predictor for a batch on the device (csrc/session.cu; reference
    inference/predictors/base.py:30,45,113-115): uint8 cuda [B,H,W,3] (+ fp32 [B,H,W] or [B,1,H,W]) -> fp32 [B,4,H,W]."""
    assert rgb_nhwc.is_cuda and rgb_nhwc.dtype == torch.uint8 and rgb_nhwc.dim() == 4 and rgb_nhwc.shape[3] == 3
    rgb_nhwc = rgb_nhwc.contiguous()
    B, H, W, _ = rgb_nhwc.shape
    if prev_mask is not None:
        assert prev_mask.is_cuda and prev_mask.dtype == torch.float32 and prev_mask.numel() == B * H * W
        prev_mask = prev_mask.contiguous()
    if out is None:
        out = torch.empty(B, 4, H, W, dtype=torch.float32, device=rgb_nhwc.device)
    L.check(L.load().vpu_image_from_u8(L.ptr(rgb_nhwc), L.ptr(prev_mask), L.ptr(out), B, H, W, L.current_stream()))
    return out
