"""state_dict layout of VitMultiGaussianVector_ed_Model and deterministic synthetic weights.

`param_spec` reproduces the reference state_dict keys/shapes (probed from
reference isegm/model/is_vpu_model.py:140-186 and its sub-modules) so that a reference
checkpoint / state_dict loads verbatim with strict=True.

`synthetic_state_dict` is the "random-init weights" generator used by tests and bench.py: there
is no network for checkpoints, and the reference's own init cannot be replayed on the GPU box
(the reference does not travel), so BOTH sides load this state_dict.  Values are a pure
function of (key, shape, seed) so the golden fixtures generated in the build container match
what the GPU box regenerates.
"""
import math
import zlib
from collections import OrderedDict

import torch

from .config import VPUConfig


def param_spec(cfg: VPUConfig):
    """OrderedDict key -> (shape, kind) with kind in {'param','buffer'} in reference order."""
    C, p, N = cfg.embed_dim, cfg.patch, cfg.num_tokens
    s = OrderedDict()

    def lin(prefix, out_f, in_f):
        s[prefix + ".weight"] = ((out_f, in_f), "param")
        s[prefix + ".bias"] = ((out_f,), "param")

    def norm(prefix, c):
        s[prefix + ".weight"] = ((c,), "param")
        s[prefix + ".bias"] = ((c,), "param")

    def conv(prefix, out_c, in_c, k):
        s[prefix + ".weight"] = ((out_c, in_c, k, k), "param")
        s[prefix + ".bias"] = ((out_c,), "param")

    def convT(prefix, in_c, out_c, k):
        s[prefix + ".weight"] = ((in_c, out_c, k, k), "param")
        s[prefix + ".bias"] = ((out_c,), "param")

    conv("patch_embed_coords.proj", C, 3, p)
    s["backbone.cls_token"] = ((1, 1, C), "param")
    s["backbone.pos_embed"] = ((1, N + 1, C), "param")
    conv("backbone.patch_embed.proj", C, 3, p)
    for i in range(cfg.depth):
        b = "backbone.blocks.%d" % i
        norm(b + ".norm1", C)
        norm(b + ".norm2", C)
        lin(b + ".attn.qkv", 3 * C, C)
        lin(b + ".attn.proj", C, C)
        lin(b + ".mlp.fc1", 4 * C, C)
        lin(b + ".mlp.fc2", C, 4 * C)
    norm("backbone.fc_norm", C)
    lin("backbone.head", 1000, C)

    lin("neck.ffn_layer.lin1", cfg.ppue_ffn_dim, cfg.ppue_dim)
    lin("neck.ffn_layer.lin2", C, cfg.ppue_ffn_dim)

    def attn(prefix, internal):
        lin(prefix + ".q_proj", internal, C)
        lin(prefix + ".k_proj", internal, C)
        lin(prefix + ".v_proj", internal, C)
        lin(prefix + ".out_proj", C, internal)

    for j in range(cfg.dma_depth):
        l = "neck.att.layers.%d" % j
        attn(l + ".self_attn", C)
        norm(l + ".norm1", C)
        attn(l + ".cross_attn_token_to_image", C // 2)
        norm(l + ".norm2", C)
        lin(l + ".mlp.lin1", cfg.dma_mlp_dim, C)
        lin(l + ".mlp.lin2", C, cfg.dma_mlp_dim)
        norm(l + ".norm3", C)
        norm(l + ".norm4", C)
        attn(l + ".cross_attn_image_to_token", C // 2)
    attn("neck.att.final_attn_token_to_image", C // 2)
    norm("neck.att.norm_final_attn", C)

    d4, d8, d32 = cfg.down_4_chan, cfg.down_8_chan, cfg.down_32_chan
    o = cfg.out_dims
    convT("neck.down_4.0", C, d4, 2)
    norm("neck.down_4.1", d4)
    convT("neck.down_4.3", d4, d4 // 2, 2)
    norm("neck.down_4.4", d4 // 2)
    conv("neck.down_4.5", o[0], d4 // 2, 1)
    norm("neck.down_4.6", o[0])
    convT("neck.down_8.0", C, d8, 2)
    norm("neck.down_8.1", d8)
    conv("neck.down_8.2", o[1], d8, 1)
    norm("neck.down_8.3", o[1])
    conv("neck.down_16.0", o[2], C, 1)
    norm("neck.down_16.1", o[2])
    conv("neck.down_32.0", d32, C, 2)
    norm("neck.down_32.1", d32)
    conv("neck.down_32.2", o[3], d32, 1)
    norm("neck.down_32.3", o[3])

    ch = cfg.head_channels
    s["head.logit_scale"] = ((), "param")
    conv("head.conv_seg", 1, ch, 1)
    for i in range(4):
        conv("head.convs.%d.conv" % i, ch, o[i], 1)
    conv("head.fusion_conv.conv", ch, 4 * ch, 1)
    convT("head.up_conv1.0", ch, ch // 2, 2)
    norm("head.up_conv1.1", ch // 2)
    conv("head.up_conv1.2", ch // 2, ch // 2, 1)
    norm("head.up_conv1.3", ch // 2)
    convT("head.up_conv2.0", ch // 2, ch // 4, 2)
    norm("head.up_conv2.1", ch // 4)
    conv("head.up_conv2.2", ch // 4, ch // 4, 1)
    norm("head.up_conv2.3", ch // 4)
    # reference hard-codes d_model=768 (swin_transformer.py:668); generalised to embed_dim for L/H
    lin("head.ffn_layer.lin1", 2 * C, C)
    lin("head.ffn_layer.lin2", ch, 2 * C)
    s["pe_layer.positional_encoding_gaussian_matrix"] = ((2, C // 2), "buffer")
    for i in range(4):
        s["point_embeddings.%d.weight" % i] = ((1, C), "param")
    s["not_a_point_embed.weight"] = ((1, C), "param")
    conv("head_aux", 1, 128, 1)
    return s


def _gen(key, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def synthetic_tensor(key, shape, seed=0, scheme="reference_init"):
    """Deterministic fp32 tensor for one state_dict entry (see module docstring).

    scheme="reference_init" follows the distributions the reference's constructors draw from:
    xavier-uniform for every ViT Linear and the image patch embed (models_vit.py:169-188), torch's
    default U(+-1/sqrt(fan_in)) for the DMA / neck / head Linear and Conv layers (plain nn.Linear,
    nn.Conv2d, nn.ConvTranspose2d; the ConvModule stand-in of SURVEY.md 8c), N(0, .02) pos_embed.
    Unlike the reference, biases and norm affines are small random values instead of 0 / 1 so that
    every parameter of the path is exercised by the parity tests.
    scheme="xavier" uses xavier-uniform everywhere (about 8x larger logits: the stress set)."""
    g = _gen(key, seed)
    if key == "head.logit_scale":
        return torch.tensor(math.log(1 / 0.07), dtype=torch.float32)
    dead_table = key.split(".")[0] in ("point_embeddings", "not_a_point_embed")     # nn.Embedding: N(0, 1)
    if len(shape) >= 2 and key.endswith(".weight") and not dead_table:
        rf = 1
        for d in shape[2:]:
            rf *= d
        fan_in, fan_out = shape[1] * rf, shape[0] * rf
        if key == "backbone.patch_embed.proj.weight" and scheme != "xavier":
            fan_out = shape[0]                              # xavier on w.view(C, -1) (models_vit.py:169-171)
        if scheme == "xavier" or key.startswith("backbone."):
            a = math.sqrt(6.0 / (fan_in + fan_out))
        else:
            a = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * a
    if key.endswith("pos_embed") or key.endswith("cls_token"):
        return torch.randn(shape, generator=g, dtype=torch.float32) * 0.02
    if len(shape) == 1 and key.endswith(".weight"):          # LayerNorm / GroupNorm gamma
        return 1.0 + 0.05 * torch.randn(shape, generator=g, dtype=torch.float32)
    if key.endswith(".bias"):
        return 0.02 * torch.randn(shape, generator=g, dtype=torch.float32)
    return torch.randn(shape, generator=g, dtype=torch.float32)


def synthetic_state_dict(cfg: VPUConfig, seed=0, scheme="reference_init"):
    return OrderedDict((k, synthetic_tensor(k, shape, seed, scheme)) for k, (shape, _) in param_spec(cfg).items())
