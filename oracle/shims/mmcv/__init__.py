"""Minimal stand-in for the `mmcv` subset the reference imports on the VPU hot path
(SURVEY.md section 8c). TEST INFRASTRUCTURE ONLY. ConvModule semantics relied on:
norm_cfg=None => Conv2d(bias=True) stored as `.conv`, followed by ReLU(inplace=True)
(reference swin_transformer.py:680-695, decode_head.py:56)."""


def load(*a, **k):
    raise NotImplementedError


def jit(*a, **k):
    def deco(f):
        return f
    return deco
