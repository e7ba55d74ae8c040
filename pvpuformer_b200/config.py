"""Model dimensions of the VPUFormer per-click forward.

Values follow the reference: ViT factories (reference isegm/model/modeling/models_vit.py:306-319),
the single shipped model config (reference models/iSegNet/vpu_base448_cocolvis.py:11-56) and the
neck/head constructors (reference isegm/model/is_vpu_model.py:19-86,
isegm/model/modeling/swin_transformer.py:666-721).
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class VPUConfig:
    arch: str = "vit_base"
    img_size: int = 448
    patch: int = 16
    embed_dim: int = 768
    depth: int = 12
    num_heads: int = 12
    num_max_points: int = 24          # is_vpu_model.py:144
    dma_depth: int = 3                # is_vpu_model.py:45-53
    dma_heads: int = 8
    dma_mlp_dim: int = 1024           # SimpleFPN.hide_dim
    ppue_ffn_dim: int = 2048          # SimpleFPN.hide_dim * 2
    head_channels: int = 256          # upsample='x1'
    out_dims: tuple = (128, 256, 512, 1024)
    norm_radius: int = 5
    norm_mean: tuple = (.485, .456, .406)   # is_model.py:13
    norm_std: tuple = (.229, .224, .225)

    @property
    def grid(self):
        return self.img_size // self.patch

    @property
    def num_tokens(self):
        return self.grid * self.grid

    @property
    def head_dim(self):
        return self.embed_dim // self.num_heads

    @property
    def window_grid(self):
        """tokens per side of one 224-px window (models_vit.py:230-237)."""
        return 224 // self.patch

    @property
    def blocks_per_group(self):
        """models_vit.py:274: every `group`-th block (1-based) is global, others windowed."""
        return 6 if self.depth == 12 else self.depth // 4

    @property
    def ppue_dim(self):
        return 2 * self.img_size + 3      # 899 at 448

    @property
    def num_queries(self):
        return 2 * self.num_max_points    # 48

    @property
    def down_4_chan(self):
        return max(self.out_dims[0] * 2, self.embed_dim // 2)

    @property
    def down_8_chan(self):
        return max(self.out_dims[1], self.embed_dim // 2)

    @property
    def down_32_chan(self):
        return max(self.out_dims[3], self.embed_dim * 2)

    @property
    def g4(self):
        return 4 * self.grid


_ARCHS = {
    "vit_base": dict(patch=16, embed_dim=768, depth=12, num_heads=12),
    "vit_large": dict(patch=16, embed_dim=1024, depth=24, num_heads=16),
    "vit_huge": dict(patch=14, embed_dim=1280, depth=32, num_heads=16),
}


def make_config(arch="vit_base", **overrides):
    kw = dict(_ARCHS[arch])
    kw.update(overrides)
    return VPUConfig(arch=arch, **kw)
