"""model.yolo2 -- Darknet-19 backbone plugin, B200-native.

Drop-in for the plugin surface of ruiminshen/yolo2-pytorch `model/yolo2.py` (reference file:line
cited per item): same constructor contract `Darknet(config_channels, anchors, num_cls, stride=2,
ratio=1)` (:69), same `forward(x[B,3,H,W] fp32) -> [B, A*(5+C), H/32, W/32] fp32` (:125-130), same
state_dict key names (`layers1.N.conv.weight`, `layers1.N.bn.running_mean`, `passthrough.*`,
`layers3.1.conv.bias`, ...), `init()` (:117-123), `scope()` (:132-133), `get_mapper()` (:135-137) and
the module-level `reorg()` (:33-46).

The modules here only HOLD parameters (so `.cuda()`, `.state_dict()`, `load_state_dict()`,
`.parameters()` and torch.optim work unchanged).  No torch.nn forward is ever executed: the forward
pass is the kernel chain in b200.engine (tcgen05 implicit-GEMM convs with fused BN + leaky-ReLU,
fp16 NHWC activations, in-place concat).  There is no CPU fallback.
"""
import torch
import torch.nn as nn

import model
from b200 import engine as _engine
from b200 import ops as _ops
from b200 import train_engine as _train

settings = {
    'size': (416, 416),
}


def reorg(x, stride_h=2, stride_w=2):
    """Space-to-depth with offset-major channel order: out[b,(sh*stride_w+sw)*C+c,h',w'] =
    x[b,c,h'*stride_h+sh,w'*stride_w+sw]  (reference model/yolo2.py:33-46), one CUDA kernel."""
    return _ops.reorg_f32_nchw(x.contiguous().float(), stride_h, stride_w)


class MaxPool2d(nn.Module):
    """Placeholder that keeps the reference's nn.Sequential indices (model/yolo2.py:79,86,97); the
    2x2 pooling itself runs inside the engine (fused into the first conv, or yb_maxpool2x2_f16)."""
    is_pool = True

    def __init__(self, kernel_size=2):
        nn.Module.__init__(self)
        if kernel_size != 2:
            raise ValueError('only MaxPool2d(2) is on the Darknet path')
        self.kernel_size = kernel_size

    def forward(self, x):
        y = _ops.maxpool2x2(x.permute(0, 2, 3, 1).contiguous().half())
        return y.permute(0, 3, 1, 2).float()


class Conv2d(nn.Module):
    """Parameter holder for conv(k, stride 1, pad (k-1)//2, bias = not bn) -> BatchNorm2d(momentum
    0.01) -> LeakyReLU(0.1)  (reference model/yolo2.py:49-65)."""
    is_pool = False

    def __init__(self, in_channels, out_channels, kernel_size, padding=0, stride=1, bn=True, act=True):
        nn.Module.__init__(self)
        if isinstance(padding, bool):
            padding = (kernel_size - 1) // 2 if padding else 0
        if stride != 1 or padding != (kernel_size - 1) // 2 or kernel_size not in (1, 3):
            raise ValueError('B200 Conv2d supports k in {1,3}, stride 1, "same" padding (got k=%d stride=%d pad=%d)'
                             % (kernel_size, stride, padding))
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding=padding, bias=not bn)
        self.has_bn, self.has_act = bool(bn), bool(act)
        if bn:
            self.bn = nn.BatchNorm2d(out_channels, momentum=0.01)
        self._unit = None

    def forward(self, x):
        """Stand-alone use of one unit on an fp32 NCHW tensor (tests, pruning tools).  Layout/dtype
        conversion at this boundary is plain data movement; the conv runs on the tcgen05 kernel."""
        if self.training and self.has_bn:
            raise NotImplementedError('training-mode BatchNorm is not part of this build (SURVEY 8f/next)')
        if self._unit is None:
            self._unit = _engine.ConvUnit(self.conv, self.bn if self.has_bn else None, self.has_act)
        u = self._unit
        if u.cin % 32 != 0:
            raise ValueError('stand-alone Conv2d needs Cin %% 32 == 0 (the 3-channel first layer is fused in Darknet.forward)')
        u.refresh()
        y = _ops.conv_bn_act(x.permute(0, 2, 3, 1).contiguous().half(), u.w16, u.scale, u.shift, u.slope, out_mode=_ops.OUT_F32_NCHW)
        return y


class _DarknetTrainFunction(torch.autograd.Function):
    """Training-mode forward/backward of the whole backbone as one autograd node: forward runs the
    train-mode kernel chain (batch-statistics BatchNorm, running-stat update), backward the explicit
    backward chain (b200.train_engine) and hands every parameter its fp32 gradient."""

    @staticmethod
    def forward(ctx, dnn, x, *params):
        feature, saved = dnn.trainer.forward(x)
        ctx.dnn, ctx.saved = dnn, saved
        return feature

    @staticmethod
    def backward(ctx, dfeature):
        grads = ctx.dnn.trainer.backward(ctx.saved, dfeature)
        ctx.saved = None
        out = []
        for name, p in ctx.dnn.named_parameters():
            g = grads.get(name)
            out.append(g.view_as(p) if g is not None else None)
        return (None, None) + tuple(out)


class Darknet(nn.Module):
    def __init__(self, config_channels, anchors, num_cls, stride=2, ratio=1):
        nn.Module.__init__(self)
        if stride != 2:
            raise ValueError('Darknet (B200): passthrough stride must be 2')
        self.stride = stride
        bn = config_channels.config.getboolean('batch_norm', 'enable')
        cc = config_channels

        def unit(group, seq, width, k):
            seq.append(Conv2d(cc.channels, cc(width, '%s.%d.conv.weight' % (group, len(seq))), k, bn=bn, padding=True))

        # layers1: C32 P C64 P | C128 c64 C128 P | C256 c128 C256 P | C512 c256 C512 c256 C512
        width = int(32 * ratio)
        seq = []
        for _ in range(2):
            unit('layers1', seq, width, 3)
            seq.append(MaxPool2d(2))
            width *= 2
        for _ in range(2):
            unit('layers1', seq, width, 3)
            unit('layers1', seq, width // 2, 1)
            unit('layers1', seq, width, 3)
            seq.append(MaxPool2d(2))
            width *= 2
        for _ in range(2):
            unit('layers1', seq, width, 3)
            unit('layers1', seq, width // 2, 1)
        unit('layers1', seq, width, 3)
        self.layers1 = nn.Sequential(*seq)
        c_trunk16 = cc.channels

        # layers2: P C1024 c512 C1024 c512 C1024 C1024 C1024
        width *= 2
        seq = [MaxPool2d(2)]
        for _ in range(2):
            unit('layers2', seq, width, 3)
            unit('layers2', seq, width // 2, 1)
        for _ in range(3):
            unit('layers2', seq, width, 3)
        self.layers2 = nn.Sequential(*seq)
        c_trunk32 = cc.channels

        self.passthrough = Conv2d(c_trunk16, cc(int(64 * ratio), 'passthrough.conv.weight'), 1, bn=bn)
        c_cat = cc.channels * stride * stride + c_trunk32

        seq = [Conv2d(c_cat, cc(int(1024 * ratio), 'layers3.0.conv.weight'), 3, bn=bn, padding=True)]
        seq.append(Conv2d(cc.channels, model.output_channels(len(anchors), num_cls), 1, bn=False, act=False))
        self.layers3 = nn.Sequential(*seq)

        self.init()
        self._engine = None
        self._trainer = None

    def init(self):
        """kaiming-normal conv weights, BN gamma = 1, beta = 0 (reference model/yolo2.py:117-123)."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    @property
    def engine(self):
        if self._engine is None:
            self._engine = _engine.DarknetEngine(self)
        return self._engine

    @property
    def trainer(self):
        if self._trainer is None:
            self._trainer = _train.DarknetTrainer(self.engine)
        return self._trainer

    def forward(self, x):
        if self.training:
            # batch-statistics BatchNorm + autograd through the explicit backward chain
            return _DarknetTrainFunction.apply(self, x, *[p for _, p in self.named_parameters()])
        return self.engine.forward(x).clone()

    def scope(self, name):
        """'layers1.4.conv.weight' -> 'layers1.4' (reference model/yolo2.py:132-133)."""
        return name.rsplit('.', 2)[0]

    def get_mapper(self, index):
        """Channel map of the reorg node for the pruning tools (reference model/yolo2.py:135-137)."""
        if index == 94:
            n = self.stride * self.stride
            return lambda indices, channels: torch.cat([indices + k * channels for k in range(n)])
