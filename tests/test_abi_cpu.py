"""CPU: the C-ABI library loads and exports every symbol include/vpuformer_b200.h declares; host-side
logic (packing shapes, state_dict surface, argument validation) -- no compute calls without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from pvpuformer_b200 import build as vbuild, lib as L
from pvpuformer_b200.config import make_config
from pvpuformer_b200.weights import param_spec, synthetic_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    vbuild.build()
    return L.load()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "vpuformer_b200.h")).read()
    declared = set(re.findall(r"\b(vpu_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.vpu_version() == 1


def test_handle_lifecycle_and_error_reporting_without_gpu(lib):
    c = make_config("vit_base")
    d = L.VpuDims(c.img_size, c.patch, c.embed_dim, c.depth, c.num_heads, c.num_max_points, c.dma_depth, c.dma_heads,
                  c.dma_mlp_dim, c.ppue_ffn_dim, c.head_channels, (ctypes.c_int32 * 4)(*c.out_dims), 5.0)
    h = ctypes.c_void_p()
    assert lib.vpu_create(ctypes.byref(h), ctypes.byref(d)) == 0
    assert lib.vpu_workspace_bytes(h, 2) > 0
    off, nb = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.vpu_workspace_lookup(h, 2, b"X", ctypes.byref(off), ctypes.byref(nb)) == 0
    assert nb.value == 2 * 784 * 768 * 4 and off.value % 1024 == 0
    assert lib.vpu_workspace_lookup(h, 2, b"nope", ctypes.byref(off), ctypes.byref(nb)) != 0
    assert b"nope" in lib.vpu_last_error()
    assert lib.vpu_finalize(h) != 0                       # nothing bound yet
    assert b"not bound" in lib.vpu_last_error()
    lib.vpu_destroy(h)
    bad = L.VpuDims(448, 16, 700, 12, 12, 24, 3, 8, 1024, 2048, 256, (ctypes.c_int32 * 4)(128, 256, 512, 1024), 5.0)
    assert lib.vpu_create(ctypes.byref(h), ctypes.byref(bad)) != 0
    assert b"embed_dim" in lib.vpu_last_error()


@pytest.mark.parametrize("arch", ["vit_base", "vit_huge"])
def test_module_surface_and_packing_shapes(arch):
    from pvpuformer_b200.model import build_model
    from pvpuformer_b200.packing import pack_weights
    cfg = make_config(arch)
    spec = param_spec(cfg)
    m = build_model(arch)
    assert list(m.state_dict().keys()) == list(spec.keys())
    assert m.with_prev_mask and m.with_aux_output
    assert m.backbone.patch_embed.grid_size == (cfg.grid, cfg.grid)
    assert m.backbone.pos_embed.shape == (1, cfg.num_tokens + 1, cfg.embed_dim)
    assert "params" in m._config
    if arch == "vit_base":
        sd = synthetic_state_dict(cfg, 0)
        m.load_state_dict(sd, strict=True)
        packed, scalars = pack_weights(m.state_dict(), cfg, torch.device("cpu"))
        C, N = cfg.embed_dim, cfg.num_tokens
        assert packed["pe.w"].shape == (C, 10 * cfg.patch ** 2) and packed["pe.w"].dtype == torch.bfloat16
        assert packed["pe.w_lo"].shape == (C, 6 * cfg.patch ** 2)
        assert packed["pe.tab"].shape == (N, C)
        assert packed["ffn.w1"].shape == (2048, 904) and not packed["ffn.w1"][:, 899:].any()
        assert packed["dma0.img.w"].shape == (3 * C // 2, C) and packed["dma0.img.tab"].shape == (N + 128, 3 * C // 2)
        # merged token projections (packing.py): one GEMM over [tokens + PE | tokens] must reproduce the reference's separate
        # q / k / v projections of transformer.py:436-461 (block rows [W 0] / [0 W])
        assert packed["dma0.sa.qkv.w"].shape == (3 * C, C)
        assert packed["dma0.tok.w"].shape == (4 * C, 2 * C) and packed["dma2.tok.w"].shape == (3 * C // 2, 2 * C)
        assert torch.equal(packed["dma0.img.tab"][N:], packed["dma0.img.tab"][:128])          # periodic padding of the table
        g = torch.Generator().manual_seed(5)
        tok, pe = torch.randn(7, C, generator=g), torch.randn(7, C, generator=g)
        both = torch.cat([tok + pe, tok], 1)
        lin = lambda x, key: x @ sd[key + ".weight"].t() + sd[key + ".bias"]
        l0, l1 = "neck.att.layers.0.", "neck.att.layers.1."
        ref = torch.cat([lin(tok + pe, l0 + "cross_attn_image_to_token.k_proj"), lin(tok, l0 + "cross_attn_image_to_token.v_proj"),
                         lin(tok + pe, l1 + "self_attn.q_proj"), lin(tok + pe, l1 + "self_attn.k_proj"), lin(tok, l1 + "self_attn.v_proj")], 1)
        got = both @ packed["dma0.tok.w"].float().t() + packed["dma0.tok.b"]
        assert (got - ref).abs().max().item() < 2e-2 * ref.abs().max().item()                # bf16 weights
        l2 = "neck.att.layers.2."
        ref = torch.cat([lin(tok + pe, l2 + "cross_attn_image_to_token.k_proj"), lin(tok, l2 + "cross_attn_image_to_token.v_proj"),
                         lin(tok + pe, "neck.att.final_attn_token_to_image.q_proj")], 1)
        got = both @ packed["dma2.tok.w"].float().t() + packed["dma2.tok.b"]
        assert (got - ref).abs().max().item() < 2e-2 * ref.abs().max().item()
        assert packed["d4.a.w"].shape == (4 * cfg.down_4_chan, C)
        assert packed["d32.a.w"].shape == (cfg.down_32_chan, 4 * C)
        assert "hd.seg.b" in scalars
        # normalisation fold: W/std applied to a raw image minus the fold equals W applied to (img-mean)/std
        img = torch.rand(3, 16, 16)
        w = sd["backbone.patch_embed.proj.weight"]
        mean = torch.tensor(cfg.norm_mean).view(3, 1, 1)
        std = torch.tensor(cfg.norm_std).view(3, 1, 1)
        direct = (w * ((img - mean) / std)).sum(dim=(1, 2, 3))
        folded = ((w / std) * img).sum(dim=(1, 2, 3)) - (w * (mean / std)).sum(dim=(1, 2, 3))
        assert torch.allclose(direct, folded, atol=1e-4)


def test_unsupported_configs_fail_loudly():
    from pvpuformer_b200.model import VitMultiGaussianVector_ed_Model, build_model
    with pytest.raises(NotImplementedError):
        VitMultiGaussianVector_ed_Model(random_split=True, use_disks=True, with_prev_mask=True)
    with pytest.raises(NotImplementedError):
        VitMultiGaussianVector_ed_Model(use_disks=False, with_prev_mask=True)
    m = build_model("vit_base")
    with pytest.raises(L.VpuError):
        m(torch.zeros(1, 4, 448, 448), torch.zeros(1, 2, 3))           # CPU tensor => loud failure, no fallback


def test_host_scribble_selection_matches_oracle():
    """Product host code vs the oracle's restatement of reference ops.py:245-295 under the same seed."""
    import random
    from oracle import vpu_oracle as vo
    from pvpuformer_b200 import host_prompts
    from tests import golden_util as gu
    g = gu.load("vit_base_scribble")
    for b in range(3):
        random.seed(11 + b)
        a = host_prompts.scribble_select(g["scribbles"][b][0], g["rects"][b][0])
        random.seed(11 + b)
        sx, sy = vo.scribble_select(g["scribbles"][b][0], g["rects"][b][0])
        assert np.array_equal(a[0], sx) and np.array_equal(a[1], sy)
    z = host_prompts.scribble_select(np.zeros((1000, 2)), np.zeros(4))
    assert (z == host_prompts.INT_MIN).all()
