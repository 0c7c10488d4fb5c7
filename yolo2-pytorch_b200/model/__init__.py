"""model -- detection head helpers of the plugin surface, B200-native.

Mirrors the public names of ruiminshen/yolo2-pytorch `model/__init__.py` (reference file:line per
item) so `detect.py`/`train.py`-style callers keep working: `ConfigChannels` (:29-43),
`output_channels` (:46-50), `meshgrid` (:53-56), `Inference` (:110-135), `_inference` (:170-179),
`loss` (:138-167).  The decode (sigmoid xy/objectness, exp wh * anchors, class softmax) is ONE CUDA
kernel (yb_decode_fwd); nothing here computes on the CPU.
"""
import logging

import torch
import torch.nn as nn

from b200 import ops as _ops


class ConfigChannels(object):
    """Plugin constructor contract (reference model/__init__.py:29-43): tracks the current channel
    count; `cc(default, key)` returns the layer's output channels -- `state_dict[key].size(0)` when a
    checkpoint is supplied (channel-pruned models), else `default`."""

    def __init__(self, config, state_dict=None, channels=3):
        self.config = config
        self.state_dict = state_dict
        self.channels = channels

    def __call__(self, default, name, fn=lambda var: var.size(0)):
        if self.state_dict is None:
            self.channels = default
            return default
        found = fn(self.state_dict[name])
        if found != default:
            logging.warning('%s: change number of output channels from %d to %d' % (name, default, found))
        self.channels = found
        return found


def output_channels(num_anchors, num_cls):
    """Head width A*(5+C), or A*5 for single-class models (reference model/__init__.py:46-50)."""
    return num_anchors * (5 + num_cls) if num_cls > 1 else num_anchors * 5


def meshgrid(rows, cols, swap=False):
    """Cell coordinates exactly as the reference builds them (model/__init__.py:53-56): row k is
    (k // rows, k % rows) -- a true (row, col) grid only when rows == cols.  Host helper; the decode
    kernel derives the same pair from the cell index."""
    k = torch.arange(0, rows * cols)
    a, b = k // rows, k % rows
    return torch.stack([b, a] if swap else [a, b], 1)


class Inference(nn.Module):
    """backbone + anchor-box decode (reference model/__init__.py:110-135).  forward(x) returns the
    same 7-tuple: (feature, iou, center_offset, size_norm, yx_min, yx_max, logits)."""

    def __init__(self, config, dnn, anchors):
        nn.Module.__init__(self)
        self.config = config
        self.dnn = dnn
        self.anchors = anchors
        self._anchors_dev = None
        self.last_prob = None

    def _anchors_on(self, device):
        if self._anchors_dev is None or self._anchors_dev.device != device:
            self._anchors_dev = self.anchors.detach().to(device=device, dtype=torch.float32).contiguous()
        return self._anchors_dev

    def forward(self, x):
        feature = self.dnn(x)
        if not feature.is_cuda:
            raise RuntimeError('model.Inference (B200): the backbone must return a CUDA tensor; there is no CPU fallback')
        anchors = self._anchors_on(feature.device)
        a = anchors.size(0)
        per = feature.size(1) // a
        num_cls = per - 5 if per > 5 else 1
        out = _ops.decode(feature.contiguous().float(), anchors, num_cls, with_prob=True)
        self.last_prob = out['prob']
        return feature, out['iou'], out['center_offset'], out['size_norm'], out['yx_min'], out['yx_max'], out.get('logits')


def _inference(inference, tensor):
    """Tuple -> dict (reference model/__init__.py:170-179).  One extra key, `prob`, carries the
    class softmax the decode kernel already produced (the reference recomputes it at detect.py:152)."""
    feature, iou, center_offset, size_norm, yx_min, yx_max, logits = inference(tensor)
    pred = dict(feature=feature, iou=iou, center_offset=center_offset, size_norm=size_norm, yx_min=yx_min, yx_max=yx_max)
    if logits is not None:
        pred['logits'] = logits
    prob = getattr(inference, 'last_prob', None)
    if prob is None and hasattr(inference, 'module'):
        prob = getattr(inference.module, 'last_prob', None)
    if prob is not None:
        pred['prob'] = prob
    return pred


def loss(anchors, data, pred, threshold):
    """Region loss (reference model/__init__.py:138-167).  The fused target-assignment + loss
    kernels are the next row of the hot-path table; this build fails loudly rather
    than fall back to torch ops."""
    raise NotImplementedError('model.loss (B200): region-loss kernels are not part of this build yet')
