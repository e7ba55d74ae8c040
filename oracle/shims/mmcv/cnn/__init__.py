import torch.nn as nn
from ..utils import Registry

MODELS = Registry('model')


class ConvModule(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias='auto', conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'),
                 inplace=True, **kwargs):
        super().__init__()
        assert norm_cfg is None, "shim supports norm_cfg=None only"
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                              dilation=dilation, groups=groups, bias=True if bias == 'auto' else bias)
        self.with_activation = act_cfg is not None
        if self.with_activation:
            assert act_cfg['type'] == 'ReLU'
            self.activate = nn.ReLU(inplace=inplace)

    def forward(self, x):
        x = self.conv(x)
        if self.with_activation:
            x = self.activate(x)
        return x


def build_conv_layer(cfg, *args, **kwargs):
    return nn.Conv2d(*args, **kwargs)


def build_norm_layer(cfg, num_features, postfix=''):
    return 'ln', nn.LayerNorm(num_features)
