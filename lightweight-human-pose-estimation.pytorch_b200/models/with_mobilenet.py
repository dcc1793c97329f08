"""Drop-in mirror of the reference's models/with_mobilenet.py:7-123.

Same class names, constructor arguments, attribute names (`model`, `cpm`, `initial_stage`,
`refinement_stages`), state_dict keys/shapes and seeded initialisation as the reference, so
`modules.load_state.load_state(net, checkpoint)` and `torch.manual_seed` behave identically.  The
forward pass is NOT torch: `PoseEstimationWithMobileNet.forward` hands the input to
lwpose_b200.engine, which runs the hand-written sm_100a kernels (tcgen05 implicit-GEMM convs,
depthwise / stem CUDA-core kernels).  There is no CPU path: a CPU tensor raises.
"""
import torch
from torch import nn

from ..modules.conv import conv, conv_dw, conv_dw_no_bn

# MobileNetV1 trunk up to conv5_5: (out_channels, stride, dilation) per depthwise-separable block
_BACKBONE = [(64, 1, 1), (128, 2, 1), (128, 1, 1), (256, 2, 1), (256, 1, 1), (512, 1, 1), (512, 1, 2),
             (512, 1, 1), (512, 1, 1), (512, 1, 1), (512, 1, 1)]


class _EngineOnly(nn.Module):
    """Sub-blocks only hold parameters; the whole-network engine executes them."""

    def forward(self, *args, **kwargs):
        raise RuntimeError("%s holds parameters only; call PoseEstimationWithMobileNet.forward (CUDA engine)"
                           % type(self).__name__)


class Cpm(_EngineOnly):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.align = conv(in_channels, out_channels, kernel_size=1, padding=0, bn=False)
        self.trunk = nn.Sequential(*[conv_dw_no_bn(out_channels, out_channels) for _ in range(3)])
        self.conv = conv(out_channels, out_channels, bn=False)


def _head(in_channels, mid_channels, out_channels):
    return nn.Sequential(conv(in_channels, mid_channels, kernel_size=1, padding=0, bn=False),
                         conv(mid_channels, out_channels, kernel_size=1, padding=0, bn=False, relu=False))


class InitialStage(_EngineOnly):
    def __init__(self, num_channels, num_heatmaps, num_pafs):
        super().__init__()
        self.trunk = nn.Sequential(*[conv(num_channels, num_channels, bn=False) for _ in range(3)])
        self.heatmaps = _head(num_channels, 512, num_heatmaps)
        self.pafs = _head(num_channels, 512, num_pafs)


class RefinementStageBlock(_EngineOnly):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.initial = conv(in_channels, out_channels, kernel_size=1, padding=0, bn=False)
        self.trunk = nn.Sequential(conv(out_channels, out_channels),
                                   conv(out_channels, out_channels, dilation=2, padding=2))


class RefinementStage(_EngineOnly):
    def __init__(self, in_channels, out_channels, num_heatmaps, num_pafs):
        super().__init__()
        self.trunk = nn.Sequential(*[RefinementStageBlock(in_channels if i == 0 else out_channels, out_channels)
                                     for i in range(5)])
        self.heatmaps = _head(out_channels, out_channels, num_heatmaps)
        self.pafs = _head(out_channels, out_channels, num_pafs)


class PoseEstimationWithMobileNet(nn.Module):
    """forward(x: float32 cuda [N,3,H,W], H and W multiples of 8) -> [hm_0, paf_0, ..., hm_R, paf_R],
    each NCHW float32 [N, 19 or 38, H/8, W/8] (reference models/with_mobilenet.py:114-123).

    `precision`: 'tf32' (default: fp32 storage, tensor-core TF32 products, fp32 accumulate) or 'bf16'."""

    def __init__(self, num_refinement_stages=1, num_channels=128, num_heatmaps=19, num_pafs=38):
        super().__init__()
        blocks = [conv(3, 32, stride=2, bias=False)]
        cin = 32
        for cout, stride, dilation in _BACKBONE:
            blocks.append(conv_dw(cin, cout, stride=stride, dilation=dilation, padding=dilation))
            cin = cout
        self.model = nn.Sequential(*blocks)
        self.cpm = Cpm(512, num_channels)
        self.initial_stage = InitialStage(num_channels, num_heatmaps, num_pafs)
        self.refinement_stages = nn.ModuleList()
        for _ in range(num_refinement_stages):
            self.refinement_stages.append(
                RefinementStage(num_channels + num_heatmaps + num_pafs, num_channels, num_heatmaps, num_pafs))
        self.num_channels, self.num_heatmaps, self.num_pafs = num_channels, num_heatmaps, num_pafs
        self.precision = "tf32"
        self._engine = None

    # --- engine management ---------------------------------------------------------------------
    def engine(self):
        from .. import engine as _engine
        if self._engine is None:
            self._engine = _engine.NetEngine(self)
        return self._engine

    def refresh(self):
        """Drop packed weights / plans (call after changing parameters in place)."""
        self._engine = None
        return self

    def _apply(self, fn, *args, **kwargs):
        self._engine = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._engine = None
        return super().load_state_dict(*args, **kwargs)

    def train(self, mode=True):
        if mode:
            raise RuntimeError("lwpose_b200 implements the inference hot path only (eval mode)")
        return super().train(mode)

    def forward(self, x):
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise RuntimeError("lwpose_b200 has no CPU path: PoseEstimationWithMobileNet.forward needs a CUDA tensor "
                               "(call net.cuda() and x.cuda() first)")
        return self.engine().forward(x, precision=self.precision)
