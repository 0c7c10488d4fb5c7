#!/usr/bin/env python
"""Golden fixture for BASELINE configs[2] ("C3": Darknet-19 416x416 batch-64 training step), produced by EXECUTING the reference on
CPU exactly like make_golden_train.py (its own Darknet in train() mode, Inference decode, model.loss under the two torch-0.3.1 shims,
hparam-weighted sum, autograd) -- at the configuration's real size: 64 x 3 x 416 x 416, 16 ground-truth slots per image.

Stored: the head feature of images 0 and 63, the five loss terms, for every parameter the gradient's L2 norm and its first 16
elements, the full gradient of the small tensors (BN gamma / beta, head bias), BatchNorm running statistics after the step.

    python tests/golden/make_golden_c3.py          # build container only (needs /root/reference; ~10 GB RAM, a few minutes)
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

BATCH, SIZE, SLOTS, SEED_X, SEED_T = 64, 416, 16, 64, 65


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
    s = SIZE // 32
    x = O.synth_images(BATCH, SIZE, SIZE, seed=SEED_X)
    data = O.norm_data(O.synth_targets(BATCH, SIZE, SIZE, slots=SLOTS, seed=SEED_T), SIZE, SIZE, s, s)
    inference = model.Inference(config, dnn, anchors)
    inference.train()
    pred = model._inference(inference, x)
    losses, debug = fns['loss'](anchors, data, pred, 0.6)
    total = sum(losses[k] * O.HPARAM_DEFAULT[k] for k in losses)
    dnn.zero_grad()
    total.backward()
    feature = pred['feature'].detach().numpy()
    out = dict(feature=feature[[0, BATCH - 1]], feature_absmax=np.abs(feature).reshape(BATCH, -1).max(1),
               positives=np.int64(debug['positive'].sum().item()), negatives=np.int64(debug['negative'].sum().item()))
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
    path = os.path.join(HERE, 'c3_train64.npz')
    np.savez_compressed(path, **out)
    print('c3_train64.npz %.1f KB' % (os.path.getsize(path) / 1024), {k: float(v) for k, v in out.items() if k.startswith('loss_')})


if __name__ == '__main__':
    main()
