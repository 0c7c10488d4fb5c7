#!/usr/bin/env python
"""Golden fixture for the region loss, produced by EXECUTING the reference's own `model.loss` (model/__init__.py:138-167,
with iou_match :59-73, fit_positive :76-95, fill_norm :98-103) on CPU.

The module was written for torch 0.3.1 and does not run on torch >= 0.4 as is.  The reference SOURCE is executed
unmodified; two pieces of torch-0.3.1 behaviour that modern torch dropped are supplied from outside:

  1. masks stay masks: in 0.3.1 `torch.prod(yx_min < yx_max, -1)` is a ByteTensor that indexes as a mask
     (model/__init__.py:80,91).  Modern torch promotes the product to int64, which would index by POSITION (the silent
     mis-masking SURVEY 8c records).  The `torch` name seen by the reference functions is a proxy whose `prod` keeps
     uint8/bool inputs as uint8.
  2. broadcasting masks: in 0.3.1 `x[mask]` with a mask of shape [B, cells, A, 1] on x [B, cells, A, 2 or C] was
     `masked_select` with broadcasting (model/__init__.py:154,155,160,162).  `torch.unsqueeze` of a mask returns a Tensor
     subclass whose `__torch_function__` turns exactly that indexing into `masked_select(x, mask.expand_as(x))`.

Everything else (IoU matrices, argmax ties, scatter of positives, log / floor targets, mse_loss(size_average=False),
cross_entropy mean, the division by B*cells*A) is the reference's code on torch 2.x CPU kernels.

    python tests/golden/make_golden_loss.py          # build container only (needs /root/reference)
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.filterwarnings('ignore')
import make_golden as G  # noqa: E402
from oracle import yolo2_oracle as O  # noqa: E402


class BroadcastMask(torch.Tensor):
    """uint8/bool mask that broadcasts when used as an index (torch 0.3.1 `x[mask]` == masked_select with broadcasting)."""

    @classmethod
    def __torch_function__(cls, func, types_, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is torch.Tensor.__getitem__ and isinstance(args[1], BroadcastMask):
            x, mask = args[0], args[1].as_subclass(torch.Tensor)
            return torch.masked_select(x.as_subclass(torch.Tensor) if isinstance(x, BroadcastMask) else x, mask.bool().expand_as(x))
        with torch._C.DisableTorchFunctionSubclass():
            return func(*args, **kwargs)


class Torch031(types.ModuleType):
    """`torch` as the reference's functions see it: everything forwards to the real module except the two behaviours above."""

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def prod(t, *a, **kw):
        out = torch.prod(t, *a, **kw)
        return out.to(torch.uint8) if t.dtype in (torch.uint8, torch.bool) else out

    @staticmethod
    def unsqueeze(t, dim):
        out = torch.unsqueeze(t, dim)
        return out.as_subclass(BroadcastMask) if t.dtype in (torch.uint8, torch.bool) else out


def main():
    model, utils, detect = G.import_reference()
    shim = Torch031('torch')
    fns = {}
    for name in ('iou_match', 'fit_positive', 'fill_norm', 'square', 'loss'):
        f = getattr(model, name)
        g = dict(f.__globals__)
        g['torch'] = shim
        fns[name] = types.FunctionType(f.__code__, g, name, f.__defaults__, f.__closure__)
    for f in fns.values():                       # the functions call each other through their globals
        f.__globals__.update(fns)
    anchors = O.anchors_yolo_voc()
    out = {}
    for tag, (b, s, slots, seed, one_hot) in dict(a=(4, 13, 16, 0, False), b=(7, 19, 6, 1, False), c=(2, 10, 1, 2, False), d=(3, 13, 5, 3, True)).items():
        size = s * 32
        g = torch.Generator().manual_seed(100 + seed)
        feature = torch.randn(b, 125, s, s, generator=g) * 1.5
        inference = model.Inference(G.make_config(1), lambda t: t, anchors)
        pred = model._inference(inference, feature.clone().requires_grad_(True))
        tgt = O.synth_targets(b, size, size, slots=slots, seed=20 + seed)
        data = O.norm_data(tgt, size, size, s, s)
        if one_hot:                                   # train/cross_entropy = 0: one-hot float class targets [B,G,C]
            data = dict(data, cls=torch.nn.functional.one_hot(data['cls'], 20).float())
        losses, debug = fns['loss'](anchors, data, pred, 0.6)
        total = sum(losses[k] * O.HPARAM_DEFAULT[k] for k in losses)
        grad, = torch.autograd.grad(total, pred['feature'])
        out[tag + '_feature'] = feature.numpy()
        out[tag + '_dims'] = np.array([b, s, slots, seed, int(one_hot)])
        for k, v in losses.items():
            out[tag + '_loss_' + k] = np.float64(v.item())
        out[tag + '_positive'] = debug['positive'].numpy().astype(np.uint8)
        out[tag + '_negative'] = debug['negative'].numpy().astype(np.uint8)
        out[tag + '_iou'] = debug['iou'].numpy()
        out[tag + '_grad'] = grad.numpy()
    path = os.path.join(HERE, 'loss.npz')
    np.savez_compressed(path, **out)
    print('loss.npz %.1f KB' % (os.path.getsize(path) / 1024), {k: float(v) for k, v in out.items() if '_loss_' in k and k.startswith('a_')})


if __name__ == '__main__':
    main()
