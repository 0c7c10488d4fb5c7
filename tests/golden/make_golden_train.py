#!/usr/bin/env python
"""Golden fixture for ONE TRAINING STEP, produced by EXECUTING the reference: its own `model.yolo2.Darknet` in train()
mode (batch-statistics BatchNorm, momentum 0.01; model/yolo2.py:49-65,125-130), `model.Inference` decode, `model.loss`
(under the two torch-0.3.1 shims of make_golden_loss.py), the hparam-weighted sum (train.py:348-349) and torch autograd.

Stored for a 4 x 3 x 128 x 128 batch: head feature, the five loss terms, for every parameter the gradient's L2 norm and
its first 16 elements (202 MB of gradients are not shipped), the full gradient of the small tensors (BN gamma / beta,
head bias), and the BatchNorm running statistics after the step.

    python tests/golden/make_golden_train.py          # build container only (needs /root/reference)
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
import make_golden_loss as L  # noqa: E402
from oracle import yolo2_oracle as O  # noqa: E402


def main():
    model, utils, detect = G.import_reference()
    shim = L.Torch031('torch')
    fns = {}
    for name in ('iou_match', 'fit_positive', 'fill_norm', 'square', 'loss'):
        f = getattr(model, name)
        g = dict(f.__globals__)
        g['torch'] = shim
        fns[name] = types.FunctionType(f.__code__, g, name, f.__defaults__, f.__closure__)
    for f in fns.values():
        f.__globals__.update(fns)
    sd = O.make_state_dict(seed=0)
    dnn, anchors, config = G.build_ref_darknet(model, sd)
    dnn.train()
    b, size = 4, 128
    s = size // 32
    x = O.synth_images(b, size, size, seed=12)
    data = O.norm_data(O.synth_targets(b, size, size, slots=6, seed=13), size, size, s, s)
    inference = model.Inference(config, dnn, anchors)
    inference.train()
    pred = model._inference(inference, x)
    losses, _ = fns['loss'](anchors, data, pred, 0.6)
    total = sum(losses[k] * O.HPARAM_DEFAULT[k] for k in losses)
    dnn.zero_grad()
    total.backward()
    out = dict(feature=pred['feature'].detach().numpy())
    for k, v in losses.items():
        out['loss_' + k] = np.float64(v.item())
    for name, p in dnn.named_parameters():
        gr = p.grad.detach()
        out['gnorm_' + name] = np.float64(gr.double().norm().item())
        out['ghead_' + name] = gr.flatten()[:16].numpy()
        if gr.numel() <= 2048:
            out['gfull_' + name] = gr.numpy()
    for name, buf in dnn.named_buffers():
        if 'running' in name:
            out['buf_' + name] = buf.detach().numpy()
    path = os.path.join(HERE, 'train_step.npz')
    np.savez_compressed(path, **out)
    print('train_step.npz %.1f KB' % (os.path.getsize(path) / 1024), {k: float(v) for k, v in out.items() if k.startswith('loss_')})


if __name__ == '__main__':
    main()
