"""Stand-in for `timm` (absent). Only timm.models.layers.{DropPath,to_2tuple,trunc_normal_}
are needed (reference isegm/model/modeling/swin_transformer.py:14)."""
