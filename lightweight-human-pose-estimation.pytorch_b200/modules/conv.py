"""Layer-block builders with the reference's names and state_dict layout (reference modules/conv.py:4-32).

These only create the parameter containers (so `state_dict()` keys/shapes and seeded initialisation are
identical to the reference and `load_state` works unchanged); inference never calls their torch
forward -- lwpose_b200.engine walks them and records CUDA launches instead."""
from torch import nn


def _stack(layers):
    return nn.Sequential(*[l for l in layers if l is not None])


def conv(in_channels, out_channels, kernel_size=3, padding=1, bn=True, dilation=1, stride=1, relu=True, bias=True):
    """Conv2d [+ BatchNorm2d] [+ ReLU] -> indices 0, 1, 2 of the Sequential."""
    return _stack([
        nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, bias=bias),
        nn.BatchNorm2d(out_channels) if bn else None,
        nn.ReLU(inplace=True) if relu else None,
    ])


def _depthwise(channels, kernel_size, stride, padding, dilation):
    return nn.Conv2d(channels, channels, kernel_size, stride, padding, dilation=dilation, groups=channels, bias=False)


def conv_dw(in_channels, out_channels, kernel_size=3, padding=1, stride=1, dilation=1):
    """depthwise 3x3 + BN + ReLU, pointwise 1x1 + BN + ReLU -> indices 0..5."""
    return _stack([
        _depthwise(in_channels, kernel_size, stride, padding, dilation), nn.BatchNorm2d(in_channels),
        nn.ReLU(inplace=True),
        nn.Conv2d(in_channels, out_channels, 1, 1, 0, bias=False), nn.BatchNorm2d(out_channels),
        nn.ReLU(inplace=True),
    ])


def conv_dw_no_bn(in_channels, out_channels, kernel_size=3, padding=1, stride=1, dilation=1):
    """depthwise 3x3 + ELU, pointwise 1x1 + ELU -> indices 0..3."""
    return _stack([
        _depthwise(in_channels, kernel_size, stride, padding, dilation), nn.ELU(inplace=True),
        nn.Conv2d(in_channels, out_channels, 1, 1, 0, bias=False), nn.ELU(inplace=True),
    ])
