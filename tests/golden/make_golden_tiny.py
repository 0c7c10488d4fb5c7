#!/usr/bin/env python
"""Golden fixture for the Tiny YOLOv2 plugin, produced by EXECUTING the reference's `model.yolo2.Tiny`
(model/yolo2.py:140-173) on CPU with the oracle's deterministic synthetic weights:

    python tests/golden/make_golden_tiny.py          # build container only (needs /root/reference)

Stores the feature map at 64x64 and 416x416 plus every conv unit's output at 64x64 (the reference is imported with the
same in-memory `async` shim as make_golden.py; nothing is copied)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402
from oracle import yolo2_oracle as O  # noqa: E402


def main():
    model, utils, detect = G.import_reference()
    config = G.make_config(1)
    anchors = O.anchors_yolo_voc()
    sd = O.make_tiny_state_dict(seed=0)
    net = model.yolo2.Tiny(model.ConfigChannels(config), anchors, 20)
    res = net.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all(k.endswith('num_batches_tracked') for k in res.missing_keys), res
    net.eval()
    outs = {}
    hooks = [m.register_forward_hook(lambda mod, inp, out, name=name: outs.__setitem__(name, out.detach().clone()))
             for name, m in net.layers.named_children() if isinstance(m, model.yolo2.Conv2d)]
    with torch.no_grad():
        f64 = net(O.synth_images(1, 64, 64, seed=10))
        acts = {'act_layers.' + k: v.numpy() for k, v in outs.items()}
        f416 = net(O.synth_images(1, 416, 416, seed=0))
    for h in hooks:
        h.remove()
    path = os.path.join(HERE, 'tiny.npz')
    np.savez_compressed(path, feature64=f64.numpy(), feature416=f416.numpy(), **acts)
    print('tiny.npz %.1f KB' % (os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
