"""Reference import harness (TEST INFRASTRUCTURE ONLY -- never imported by the product).

Imports the UNMODIFIED reference from /root/reference (read-only) with the stand-in modules
under oracle/shims, and builds VitMultiGaussianVector_ed_Model for
ViT-B/L/H.  Only usable in the build container: /root/reference does not exist on the GPU
box, so nothing under tests -m gpu / smoke() / bench.py may import this module.  It is used
(a) to pin oracle/vpu_oracle.py (the travelling CPU restatement) and (b) by
oracle/make_golden.py to generate tests/golden/*.npz.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("VPU_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "isegm"))


def import_reference():
    """Put shims + reference on sys.path and return the `isegm` package."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for p in (_SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    # numpy aliases removed in numpy>=1.24 and used by the reference
    # (inference/utils.py:100, is_model.py:159, pos_embed.py:56).  np.bool is NOT aliased.
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float
    import isegm  # noqa: F401
    # isegm/data was never committed upstream (.gitignore:110-111): stub the names imported by
    # inference/utils.py:6-7 so the predictor plumbing imports.
    if "isegm.data" not in sys.modules:
        data = types.ModuleType("isegm.data")
        data.__path__ = []
        ds = types.ModuleType("isegm.data.datasets")
        for name in ("GrabCutDataset", "BerkeleyDataset", "DavisDataset", "SBDEvaluationDataset",
                     "PascalVocDataset", "BraTSDataset", "ssTEMDataset", "OAIZIBDataset",
                     "HARDDataset", "ADE20kDataset", "COCOMValDataset", "DavisDataset585",
                     "LoveDADataset", "BSDataset", "SADataset", "LvisDataset", "CocoLvisDataset"):
            setattr(ds, name, type(name, (), {}))
        data.datasets = ds
        sys.modules["isegm.data"] = data
        sys.modules["isegm.data.datasets"] = ds
    return isegm


ARCHS = {
    # name: (patch, embed_dim, depth, heads)   reference models_vit.py:306-319
    "vit_base": (16, 768, 12, 12),
    "vit_large": (16, 1024, 24, 16),
    "vit_huge": (14, 1280, 32, 16),
}


def build_reference_model(arch="vit_base", img_size=448, eval_mode=True):
    """Build the reference model exactly as models/iSegNet/vpu_base448_cocolvis.py:17-56 does
    (upsample='x1' => channels=256), for B/L/H.  For L/H the head's hard-coded d_model=768 FFN
    (swin_transformer.py:668,717-721) is replaced by FFNBlock(C, 2C, 256) -- the only sensible
    reading, stated in DESIGN.md."""
    import_reference()
    import torch  # noqa: F401
    from isegm.model.is_vpu_model import VitMultiGaussianVector_ed_Model
    from isegm.model.modeling.common import FFNBlock
    from isegm.model.modeling.transformer_helper.cross_entropy_loss import CrossEntropyLoss

    patch, C, depth, heads = ARCHS[arch]
    backbone_params = dict(img_size=(img_size, img_size), patch_size=(patch, patch), in_chans=3,
                           embed_dim=C, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True)
    neck_params = dict(in_dim=C, out_dims=[128, 256, 512, 1024], img_size=(img_size, img_size))
    head_params = dict(in_channels=[128, 256, 512, 1024], in_index=[0, 1, 2, 3], dropout_ratio=0.1,
                       num_classes=1, loss_decode=CrossEntropyLoss(), align_corners=False,
                       upsample='x1', ed_loss=True, channels=256)
    model = VitMultiGaussianVector_ed_Model(
        use_disks=True, norm_radius=5, with_prev_mask=True, backbone_params=backbone_params,
        neck_params=neck_params, head_params=head_params, random_split=False, residual=True,
        with_aux_output=True)
    if C != 768:
        model.head.ffn_layer = FFNBlock(C, 2 * C, 256)
        model.head.d_model = C
    if eval_mode:
        model.eval()
    for p in model.parameters():
        p.requires_grad_(False)
    return model
