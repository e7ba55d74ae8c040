"""CPU: the oracle restatement (oracle/vpu_oracle.py) against the golden vectors the UNMODIFIED
reference produced in the build container (tests/golden, oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import vpu_oracle as vo
from pvpuformer_b200.config import make_config
from pvpuformer_b200.weights import synthetic_state_dict
from tests import golden_util as gu

_SD = {}


def _sd(arch):
    if arch not in _SD:
        _SD.clear()
        _SD[arch] = synthetic_state_dict(make_config(arch), 0)
    return _SD[arch]


@pytest.mark.parametrize("name", ["vit_base_clicks", "vit_base_box", "vit_base_scribble", "vit_base_manyclicks"])
def test_oracle_matches_reference_golden_vit_base(name):
    cfg = make_config("vit_base")
    image4, points, prompts, t = gu.case_inputs(name)
    g = gu.load(name)
    taps = {}
    gu.seed_scribble()
    with torch.no_grad():
        out = vo.forward(_sd("vit_base"), cfg, image4, points, prompts, t, taps=taps)
    B = image4.shape[0]
    # bit-exact: rasterisation / indexing
    assert np.array_equal(taps["coord_features"][:, 1:].numpy().astype(np.uint8), gu.unpack_disks(g, B))
    assert np.array_equal(taps["coord_features"][:, 0].numpy(), image4[:, 3].numpy())
    ref_ppue = g["ppue"]
    assert np.array_equal(taps["ppue"].float().numpy() != 0, ref_ppue != 0)
    assert np.abs(taps["ppue"].float().numpy() - ref_ppue).max() <= 1e-5
    # fp32 restatement of an fp32 reference: same op order => expect ~1e-6, gate at 1e-4
    assert np.abs(taps["seg_lowres"].numpy() - g["seg_lowres"]).max() < 1e-4
    assert np.abs(taps["aux_lowres"][:, [0, 1, 23, 24, 25, 47]].numpy() - g["aux_lowres_sel"]).max() < 1e-4
    assert np.abs(out["instances"][:, :, ::4, ::4].numpy() - g["instances_s4"]).max() < 1e-4
    assert np.abs(out["instances"][:, 0, 100, :].numpy() - g["instances_row100"]).max() < 1e-4
    assert np.abs(out["instances_aux"][:, [0, 24], ::8, ::8].numpy() - g["aux_s8_sel"]).max() < 1e-4
    assert np.abs(taps["backbone_features"][:, ::49, ::16].numpy() - g["backbone_slice"]).max() < 1e-3
    assert np.abs(taps["q_out"][:, :, ::16].numpy() - g["q_out_slice"]).max() < 1e-3


@pytest.mark.parametrize("name", ["vit_large_box", "vit_large_scribble"])
def test_oracle_matches_reference_golden_vit_large_mixed_prompts(name):
    """ViT-Large with box / scribble prompts through PPuE."""
    cfg = make_config("vit_large")
    image4, points, prompts, t = gu.case_inputs(name)
    g = gu.load(name)
    taps = {}
    gu.seed_scribble()
    with torch.no_grad():
        out = vo.forward(_sd("vit_large"), cfg, image4, points, prompts, t, taps=taps)
    B = image4.shape[0]
    assert np.array_equal(taps["coord_features"][:, 1:].numpy().astype(np.uint8), gu.unpack_disks(g, B))
    assert np.array_equal(taps["ppue"].float().numpy() != 0, g["ppue"] != 0)
    assert np.abs(taps["ppue"].float().numpy() - g["ppue"]).max() <= 1e-5
    assert np.abs(taps["seg_lowres"].numpy() - g["seg_lowres"]).max() < 1e-4
    assert np.abs(out["instances"][:, :, ::4, ::4].numpy() - g["instances_s4"]).max() < 1e-4
    assert np.abs(out["instances_aux"][:, [0, 24], ::8, ::8].numpy() - g["aux_s8_sel"]).max() < 1e-4


@pytest.mark.parametrize("arch", ["vit_large", "vit_huge"])
def test_oracle_matches_reference_golden_large_huge(arch):
    cfg = make_config(arch)
    image4, points, prompts, t = gu.case_inputs(arch + "_clicks")
    g = gu.load(arch + "_clicks")
    taps = {}
    with torch.no_grad():
        out = vo.forward(_sd(arch), cfg, image4, points, prompts, t, taps=taps)
    assert np.abs(taps["seg_lowres"].numpy() - g["seg_lowres"]).max() < 1e-4
    assert np.abs(out["instances"][:, 0, 100, :].numpy() - g["instances_row100"]).max() < 1e-4
    assert np.abs(taps["aux_lowres"][:, [0, 1, 23, 24, 25, 47]].numpy() - g["aux_lowres_sel"]).max() < 1e-4


def test_oracle_matches_reference_golden_config3_sample_of_each_prompt_type():
    """BASELINE.json configs[2] fixture (ViT-L, batch 32 = 11 clicks + 11 boxes + 10 scribbles, oracle/make_golden_config3.py): the
    oracle on the first two samples of every sub-batch against the unmodified reference's stride-8 logits."""
    from oracle.make_golden_config3 import SPLIT, inputs
    cfg = make_config("vit_large")
    g = gu.load("vit_large_config3")
    image4, pts, _ = inputs()
    for t, a, b in SPLIT:
        prompts = None
        if t != 0:
            prompts = (torch.from_numpy(g["t%d_prompt_points" % t])[:2], torch.from_numpy(g["t%d_boxes" % t])[:2],
                       [g["t%d_scribbles" % t][:2], g["t%d_rects" % t][:2]])
        gu.seed_scribble()
        with torch.no_grad():
            out = vo.forward(_sd("vit_large"), cfg, image4[a:a + 2], pts[a:a + 2], prompts, t)
        assert np.abs(out["instances"][:, :, ::8, ::8].numpy() - g["instances_s8"][a:a + 2]).max() < 1e-4, t
        assert np.abs(out["instances_aux"][:, [0, 24], ::16, ::16].numpy() - g["aux_s16_sel"][a:a + 2]).max() < 1e-4, t


def test_ppue_quirks():
    """Reference quirks: corner-drop, trunc toward zero, '>' bounds, peak 2.0."""
    vx, vy = vo.ppue_click_row([224.9, 10.2])
    assert vx[224] == 2.0 and vx[215] > 0 and vx[214] == 0 and vx[233] > 0 and vx[234] == 0
    assert vy[10] == 2.0 and vy[0] == 0 and vy[1] > 0 and vy[19] > 0 and vy[20] == 0
    vx, vy = vo.ppue_click_row([5, 440])          # ul out in x, br out in y -> dropped
    assert not vx.any() and not vy.any()
    vx, vy = vo.ppue_click_row([438, 438])         # br == 448 is still "in" ('>' test)
    assert vx[438] == 2.0 and vx[447] > 0
    vx, vy = vo.ppue_click_row([-0.5, 100])        # trunc toward zero -> 0
    assert vx[0] == 2.0
    rows = vo.ppue(torch.tensor([[[10., 20, 0], [-1, -1, -1]]]))
    assert rows.shape == (1, 48, 899)
    assert rows[0, 0, 896] == 1 and rows[0, 1:24, 898].eq(1).all() and rows[0, 24:, 898].eq(1).all()
    assert rows[0, 1:].sum() == 47


def test_disk_maps_edges():
    pts = torch.tensor([[[0., 0, 0], [-1, -1, -1], [447, 447, 1], [3.5, 2.5, 2]]])
    d = vo.disk_maps(pts, 448, 448)
    assert d.shape == (1, 2, 448, 448)
    assert d[0, 0, 0, 5] == 1 and d[0, 0, 0, 6] == 0 and d[0, 0, 3, 4] == 1 and d[0, 0, 4, 4] == 0
    assert d[0, 1, 447, 442] == 1 and d[0, 1, 447, 441] == 0
    assert d[0, 0].sum() == 26            # quarter disk incl. axes at a corner
    empty = vo.disk_maps(torch.full((2, 2, 3), -1.0), 448, 448)
    assert empty.sum() == 0


def test_oracle_training_shape_case_losses_match_reference_golden():
    """batch 12, points [12,48,3]; the oracle forward + the restated losses reproduce the loss values
    the reference's own loss classes gave on the reference's outputs."""
    from oracle import cases, losses as ol
    cfg = make_config("vit_base")
    image4, pts, gt = cases.train12_inputs()
    g = gu.load("vit_base_train12")
    taps = {}
    with torch.no_grad():
        out = vo.forward(_sd("vit_base"), cfg, image4, pts, taps=taps)
        ls = ol.training_losses(out, gt)
    assert np.array_equal(np.packbits(taps["ppue"].numpy() != 0), g["ppue_support_packed"])
    assert np.abs(out["instances"][:, 0, 100, :].numpy() - g["instances_row100"]).max() < 1e-4
    assert np.abs(ls["nfl"].numpy() - g["loss_nfl"]).max() < 1e-5
    assert abs(float(ls["dice"]) - float(g["loss_dice"])) < 1e-5
    assert np.abs(ls["bce_aux"].numpy() - g["loss_bce_aux"]).max() < 1e-5
