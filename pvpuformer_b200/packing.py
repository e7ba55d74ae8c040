"""One-time weight packing: reference state_dict (fp32) -> GEMM-ready device buffers.

Runs once at load time with torch ops (plumbing, not the hot path).  Layouts (DESIGN.md
"packed weights"): every GEMM weight is bf16 [N, K] row-major (the nn.Linear layout), biases /
norm affine / additive tables stay fp32.

Algebraic folds, each exact in real arithmetic:
  * image normalisation (reference ops.py:398-407) folded into the image patch-embed weights; both
    patch embeds (models_vit.py:94-104, is_vpu_model.py:165-170,385) concatenated along K;
    pos_embed[:,1:] + both conv biases - the mean fold become one additive [N, C] table;
  * DMA key_pe (transformer.py:290-318, a constant of (C, grid)): k_proj(keys + pe) =
    k_proj(keys) + pe W_k^T, so pe W^T + b is a precomputed [N, internal] table and the three
    image-side projections of a layer (t2i K, t2i V, i2t Q) share one GEMM;
  * ConvTranspose2d(k=2,s=2) / Conv2d(k=2,s=2) (is_vpu_model.py:56-86) reshaped to plain GEMM
    weights with (kh, kw) folded into N resp. K;
  * the head's fusion 1x1 conv split into four [256,256] slices (it commutes with the bilinear
    resize, swin_transformer.py:727-737);
  * norm1 / norm2 of every ViT block (models_vit.py:72-75) folded into the qkv / fc1 weights: W' = W diag(gamma),
    b' = W beta + b, and the per-row mean / rstd are applied in the GEMM epilogue from statistics the preceding residual
    GEMM wrote (VPU_LN_FOLD=0 keeps the separate LayerNorm pass for A/B measurements).
"""
import math
import os

import torch


def key_pe_table(C, g, device):
    """transformer.py:290-318 pos2d -> [g*g, C] fp32."""
    pe = torch.zeros(C, g, g, device=device)
    d = C // 2
    div = torch.exp(torch.arange(0., d, 2, device=device) * -(math.log(10000.0) / d))
    pw = torch.arange(0., g, device=device).unsqueeze(1)
    ph = torch.arange(0., g, device=device).unsqueeze(1)
    pe[0:d:2] = torch.sin(pw * div).transpose(0, 1).unsqueeze(1).repeat(1, g, 1)
    pe[1:d:2] = torch.cos(pw * div).transpose(0, 1).unsqueeze(1).repeat(1, g, 1)
    pe[d::2] = torch.sin(ph * div).transpose(0, 1).unsqueeze(2).repeat(1, 1, g)
    pe[d + 1::2] = torch.cos(ph * div).transpose(0, 1).unsqueeze(2).repeat(1, 1, g)
    return pe.reshape(C, g * g).t().contiguous()


TAB_PAD = 128   # api.cu TAB_PAD: the positional tables of the image-side projections repeat their first rows after the last one,
                # so that a 128-row TMA box starting at row (m mod N) never wraps (gemm_res.cu MODE_TAB)


def _pad_cols(w, mult=64):
    k = w.shape[1]
    kp = (k + mult - 1) // mult * mult
    return w if kp == k else torch.cat([w, w.new_zeros(w.shape[0], kp - k)], 1)


def _pad_table(t):
    reps = (TAB_PAD + t.shape[0] - 1) // t.shape[0] + 1
    return torch.cat([t] * reps, 0)[: t.shape[0] + TAB_PAD].contiguous()


def pack_weights(sd, cfg, device):
    """-> (dict key -> device tensor (bf16 or fp32, contiguous), dict scalar key -> float)."""
    f = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in sd.items()}
    out, scalars = {}, {}
    bf = torch.bfloat16
    C, p, g, N = cfg.embed_dim, cfg.patch, cfg.grid, cfg.num_tokens

    def W(key, t):
        out[key] = t.to(bf).contiguous()

    def F32(key, t):
        out[key] = t.to(torch.float32).contiguous()

    def lin(dst, src):
        W(dst + ".w", f[src + ".weight"])
        F32(dst + ".b", f[src + ".bias"])

    def norm(dst, src):
        F32(dst + ".g", f[src + ".weight"])
        F32(dst + ".b", f[src + ".bias"])

    # ---- fused patch embed ----
    mean = torch.tensor(cfg.norm_mean, device=device).view(1, 3, 1, 1)
    std = torch.tensor(cfg.norm_std, device=device).view(1, 3, 1, 1)
    w_img = f["backbone.patch_embed.proj.weight"]                # [C,3,p,p]
    w_crd = f["patch_embed_coords.proj.weight"]
    w_pe = torch.cat([(w_img / std).reshape(C, -1), w_crd.reshape(C, -1)], dim=1)     # [C, 6pp] fp32
    w_hi = w_pe.to(bf)
    pp = p * p
    # split-bf16 (prompt.cu patch_operand_kernel): A = [hi planes (6pp) | lo planes of R,G,B,prev (4pp)];
    # GEMM 1: A x [W_hi | W_hi[:, :4pp]], GEMM 2: A[:, :6pp] x W_lo
    # rows padded with zero columns to a multiple of 64 elements (api.cu K0s_ld / K0_ld): 128-byte aligned rows for the TMA loads
    W("pe.w", _pad_cols(torch.cat([w_hi, w_hi[:, :4 * pp]], dim=1)))
    W("pe.w_lo", _pad_cols(w_pe - w_hi.float()))
    fold = (w_img * (mean / std)).sum(dim=(1, 2, 3))             # [C]
    tab = f["backbone.pos_embed"][0, 1:] + (f["backbone.patch_embed.proj.bias"] + f["patch_embed_coords.proj.bias"] - fold)
    F32("pe.tab", tab)

    def lin_ln(dst, src, ln):     # Linear applied to LayerNorm(x) with x stored un-normalised (gemm.cuh Epi::ln_in):
        w = f[src + ".weight"]    #   W (gamma (x - mean) rstd + beta) + b = rstd (W' x) - mean rstd rowsum(W') + (W beta + b)
        wq = (w * f[ln + ".weight"].view(1, -1)).to(bf)
        out[dst + ".w"] = wq.contiguous()
        F32(dst + ".s", wq.float().sum(dim=1))                                # row sums of the ROUNDED W': consistent with the MMA
        F32(dst + ".b", w @ f[ln + ".bias"] + f[src + ".bias"])

    # development knob, honoured only together with the -DVPU_DEBUG library (build.py --debug): the shipped path always folds
    ln_fold = not (os.environ.get("VPU_LN_FOLD") == "0" and os.environ.get("VPU_LIB_PATH", "").endswith("_debug.so"))
    scalars["vit.ln_fold"] = 1.0 if ln_fold else 0.0
    if ln_fold:
        F32("pe.zero_b", torch.zeros(C, device=device))
    for i in range(cfg.depth):
        s, d = "backbone.blocks.%d" % i, "blk%d" % i
        if ln_fold:
            lin_ln(d + ".qkv", s + ".attn.qkv", s + ".norm1")
            lin_ln(d + ".fc1", s + ".mlp.fc1", s + ".norm2")
        else:
            norm(d + ".ln1", s + ".norm1")
            norm(d + ".ln2", s + ".norm2")
            lin(d + ".qkv", s + ".attn.qkv")
            lin(d + ".fc1", s + ".mlp.fc1")
        lin(d + ".proj", s + ".attn.proj")
        lin(d + ".fc2", s + ".mlp.fc2")

    # ---- PPuE FFN (K = 899 padded to a multiple of 8 with zero columns) ----
    w1 = f["neck.ffn_layer.lin1.weight"]
    kpad = (cfg.ppue_dim + 7) // 8 * 8
    w1p = torch.zeros(w1.shape[0], kpad, device=device)
    w1p[:, :cfg.ppue_dim] = w1
    W("ffn.w1", w1p)
    F32("ffn.b1", f["neck.ffn_layer.lin1.bias"])
    W("ffn.w2", f["neck.ffn_layer.lin2.weight"])
    F32("ffn.b2", f["neck.ffn_layer.lin2.bias"])

    # ---- DMA ----
    kpe = key_pe_table(C, g, device)                              # [N, C]
    for j in range(cfg.dma_depth):
        s, d = "neck.att.layers.%d" % j, "dma%d" % j
        sa = s + ".self_attn"
        if j == 0:      # layer 0: q, k and v all read the PPuE tokens (no positional term): one [3C, C] projection
            W(d + ".sa.qkv.w", torch.cat([f[sa + ".q_proj.weight"], f[sa + ".k_proj.weight"], f[sa + ".v_proj.weight"]], 0))
            F32(d + ".sa.qkv.b", torch.cat([f[sa + ".q_proj.bias"], f[sa + ".k_proj.bias"], f[sa + ".v_proj.bias"]], 0))
        lin(d + ".sa.o", sa + ".out_proj")
        norm(d + ".n1", s + ".norm1")
        t2i, i2t = s + ".cross_attn_token_to_image", s + ".cross_attn_image_to_token"
        lin(d + ".t2i.q", t2i + ".q_proj")
        W(d + ".img.w", torch.cat([f[t2i + ".k_proj.weight"], f[t2i + ".v_proj.weight"], f[i2t + ".q_proj.weight"]], 0))
        Ci = C // 2
        tabk = kpe @ f[t2i + ".k_proj.weight"].t() + f[t2i + ".k_proj.bias"]
        tabv = f[t2i + ".v_proj.bias"].view(1, Ci).expand(N, Ci)
        tabq = kpe @ f[i2t + ".q_proj.weight"].t() + f[i2t + ".q_proj.bias"]
        F32(d + ".img.tab", _pad_table(torch.cat([tabk, tabv, tabq], 1)))
        lin(d + ".t2i.o", t2i + ".out_proj")
        norm(d + ".n2", s + ".norm2")
        W(d + ".mlp.w1", f[s + ".mlp.lin1.weight"])
        F32(d + ".mlp.b1", f[s + ".mlp.lin1.bias"])
        W(d + ".mlp.w2", f[s + ".mlp.lin2.weight"])
        F32(d + ".mlp.b2", f[s + ".mlp.lin2.bias"])
        norm(d + ".n3", s + ".norm3")
        # every projection that reads the token state after norm3 as ONE GEMM over A = [tokens + PE | tokens] (K = 2C):
        # image->token k (tokens + PE) and v (tokens) of this layer, then the next layer's self-attention q | k (tokens + PE) and
        # v (tokens), or after the last layer the final attention's q (tokens + PE).  Block rows [W 0] / [0 W].
        blocks = [(f[i2t + ".k_proj.weight"], 0), (f[i2t + ".v_proj.weight"], 1)]       # (weight, 0 = reads tokens + PE / 1 = reads tokens)
        bias = [f[i2t + ".k_proj.bias"], f[i2t + ".v_proj.bias"]]
        if j + 1 < cfg.dma_depth:
            nsa = "neck.att.layers.%d.self_attn" % (j + 1)
            blocks += [(f[nsa + ".q_proj.weight"], 0), (f[nsa + ".k_proj.weight"], 0), (f[nsa + ".v_proj.weight"], 1)]
            bias += [f[nsa + ".q_proj.bias"], f[nsa + ".k_proj.bias"], f[nsa + ".v_proj.bias"]]
        else:
            fq = "neck.att.final_attn_token_to_image.q_proj"
            blocks += [(f[fq + ".weight"], 0)]
            bias += [f[fq + ".bias"]]
        rows = []
        for wblk, side in blocks:
            z = torch.zeros_like(wblk)
            rows.append(torch.cat([wblk, z], 1) if side == 0 else torch.cat([z, wblk], 1))
        W(d + ".tok.w", torch.cat(rows, 0))
        F32(d + ".tok.b", torch.cat(bias, 0))
        lin(d + ".i2t.o", i2t + ".out_proj")
        norm(d + ".n4", s + ".norm4")
    fa = "neck.att.final_attn_token_to_image"
    W("dmaf.img.w", torch.cat([f[fa + ".k_proj.weight"], f[fa + ".v_proj.weight"]], 0))
    F32("dmaf.img.tab", _pad_table(torch.cat([kpe @ f[fa + ".k_proj.weight"].t() + f[fa + ".k_proj.bias"],
                                               f[fa + ".v_proj.bias"].view(1, C // 2).expand(N, C // 2)], 1)))
    lin("dmaf.o", fa + ".out_proj")
    norm("dmaf.n", "neck.att.norm_final_attn")

    # ---- neck pyramid ----
    def convT(dst, src):      # [cin, cout, 2, 2] -> [(kh,kw,cout), cin]; bias repeated per (kh,kw)
        w = f[src + ".weight"]
        W(dst + ".w", w.permute(2, 3, 1, 0).reshape(4 * w.shape[1], w.shape[0]))
        F32(dst + ".b", f[src + ".bias"].repeat(4))

    def conv1(dst, src):      # [cout, cin, 1, 1] -> [cout, cin]
        W(dst + ".w", f[src + ".weight"].reshape(f[src + ".weight"].shape[0], -1))
        F32(dst + ".b", f[src + ".bias"])

    def conv1_gn(dst, src, gn):   # 1x1 conv applied to GroupNorm(1, C)(x) with x stored un-normalised (gemm.cuh Epi::gn_in):
        w = f[src + ".weight"].reshape(f[src + ".weight"].shape[0], -1)      # W (gamma (x - mean) rstd + beta) + b
        wq = (w * f[gn + ".weight"].view(1, -1)).to(bf)                       #   = rstd (W' x) - mean rstd rowsum(W') + (W beta + b)
        out[dst + ".w"] = wq.contiguous()
        F32(dst + ".wg", wq.float().sum(dim=1))                               # row sums of the ROUNDED W': consistent with the MMA
        F32(dst + ".b", w @ f[gn + ".bias"] + f[src + ".bias"])

    convT("d4.a", "neck.down_4.0")
    norm("d4.gn1", "neck.down_4.1")
    convT("d4.b", "neck.down_4.3")
    conv1_gn("d4.c", "neck.down_4.5", "neck.down_4.4")
    norm("d4.gn3", "neck.down_4.6")
    convT("d8.a", "neck.down_8.0")
    conv1_gn("d8.b", "neck.down_8.2", "neck.down_8.1")
    norm("d8.gn2", "neck.down_8.3")
    conv1("d16.a", "neck.down_16.0")
    norm("d16.gn1", "neck.down_16.1")
    w = f["neck.down_32.0.weight"]                                # [cout, C, 2, 2] -> [cout, (kh,kw,C)]
    W("d32.a.w", w.permute(0, 2, 3, 1).reshape(w.shape[0], 4 * C))
    F32("d32.a.b", f["neck.down_32.0.bias"])
    conv1_gn("d32.b", "neck.down_32.2", "neck.down_32.1")
    norm("d32.gn2", "neck.down_32.3")

    # ---- head ----
    hc = cfg.head_channels
    fw = f["head.fusion_conv.conv.weight"].reshape(hc, 4 * hc)
    for i in range(4):
        conv1("hd.c%d" % i, "head.convs.%d.conv" % i)
        W("hd.f%d.w" % i, fw[:, i * hc:(i + 1) * hc])
    F32("hd.f.b", f["head.fusion_conv.conv.bias"])
    W("hd.q.w1", f["head.ffn_layer.lin1.weight"])
    F32("hd.q.b1", f["head.ffn_layer.lin1.bias"])
    W("hd.q.w2", f["head.ffn_layer.lin2.weight"])
    F32("hd.q.b2", f["head.ffn_layer.lin2.bias"])
    F32("hd.seg.w", f["head.conv_seg.weight"].reshape(hc))
    scalars["hd.seg.b"] = float(f["head.conv_seg.bias"].item())
    return out, scalars
