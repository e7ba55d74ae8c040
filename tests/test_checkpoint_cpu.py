"""Checkpoint ingestion (pvpuformer_b200/checkpoint.py <- isegm/inference/utils.py:21-46, utils/serialization.py:7-69,
model/modeling/pos_embed.py:75-128)."""
import os

import pytest
import torch

from oracle import ref_harness as rh
from pvpuformer_b200 import checkpoint as ck
from pvpuformer_b200.config import make_config
from pvpuformer_b200.model import REFERENCE_CLASS, build_model
from pvpuformer_b200.weights import synthetic_state_dict


def test_config_has_the_reference_serialize_format():
    m = build_model("vit_base")
    c = m._config
    assert c["class"] == REFERENCE_CLASS
    p = c["params"]
    assert p["num_max_points"] == {"type": "builtin", "value": 24, "specified": False}
    assert p["use_disks"] == {"type": "builtin", "value": True, "specified": True}
    assert p["backbone_params"]["value"]["embed_dim"] == 768 and p["backbone_params"]["specified"]


def test_checkpoint_round_trip_through_a_file(tmp_path):
    cfg = make_config("vit_base")
    sd = synthetic_state_dict(cfg, 3)
    src = build_model("vit_base", state_dict=sd)
    path = os.path.join(tmp_path, "ckpt.pth")
    torch.save({"state_dict": src.state_dict(), "config": src._config}, path)        # utils/misc.py:31-33
    m = ck.load_is_model(path, "cpu")
    assert not m.training and all(not p.requires_grad for p in m.parameters())
    got = m.state_dict()
    assert list(got) == list(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    assert m._config["params"]["residual"]["value"] is True and m.cfg.embed_dim == 768
    first, models = ck.load_is_model([torch.load(path, weights_only=False)] * 2, "cpu")
    assert first is models[0] and len(models) == 2
    with pytest.raises(NotImplementedError):
        ck.load_is_model({"config": {"class": "isegm.model.is_plainvit_model.PlainVitModel", "params": {}}, "state_dict": {}}, "cpu")


def test_interpolate_pos_embed_matches_the_formula():
    m = build_model("vit_base")
    C = 768
    ckpt = {"pos_embed": torch.randn(1, 1 + 14 * 14, C)}                               # MAE ViT-B/16 pre-trained at 224 px
    want_tokens = torch.nn.functional.interpolate(ckpt["pos_embed"][:, 1:].reshape(1, 14, 14, C).permute(0, 3, 1, 2), size=(28, 28),
                                                  mode="bicubic", align_corners=False).permute(0, 2, 3, 1).flatten(1, 2)
    cls_tok = ckpt["pos_embed"][:, :1].clone()
    ck.interpolate_pos_embed(m.backbone, ckpt)
    assert ckpt["pos_embed"].shape == (1, 785, C)
    assert torch.equal(ckpt["pos_embed"][:, :1], cls_tok) and torch.equal(ckpt["pos_embed"][:, 1:], want_tokens)
    ck.interpolate_pos_embed_inference(m.backbone, (448, 448), "cpu")                  # equal grid: the reference's no-op
    with pytest.raises(NotImplementedError):
        ck.interpolate_pos_embed_inference(m.backbone, (672, 672), "cpu")


@pytest.mark.reference
def test_checkpoint_written_by_the_reference_loads_and_resamples_like_the_reference():
    """Build container only: a checkpoint the UNMODIFIED reference writes (state_dict + @serialize config) loads here with
    identical parameters, and interpolate_pos_embed equals the reference's."""
    ref = rh.build_reference_model("vit_base")
    ckpt = {"state_dict": ref.state_dict(), "config": ref._config}
    m = ck.load_is_model(ckpt, "cpu")
    got, want = m.state_dict(), ref.state_dict()
    assert list(got) == list(want) and all(torch.equal(got[k], want[k]) for k in want)
    assert m.with_prev_mask and m.cfg.depth == 12 and m.cfg.norm_radius == 5
    from isegm.model.modeling.pos_embed import interpolate_pos_embed as ref_interp
    a = {"pos_embed": torch.randn(1, 197, 768)}
    b = {"pos_embed": a["pos_embed"].clone()}
    ref_interp(ref.backbone, a)
    ck.interpolate_pos_embed(m.backbone, b)
    assert torch.equal(a["pos_embed"], b["pos_embed"])


@pytest.mark.reference
def test_checkpoint_written_here_loads_in_the_reference():
    rh.import_reference()
    from isegm.utils.serialization import load_model as ref_load_model
    m = build_model("vit_base", state_dict=synthetic_state_dict(make_config("vit_base"), 1))
    r = ref_load_model(m._config, False)
    r.load_state_dict(m.state_dict(), strict=True)
    assert type(r).__name__ == "VitMultiGaussianVector_ed_Model" and r.with_prev_mask
