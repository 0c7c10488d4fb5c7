#!/usr/bin/env python
"""Golden fixture for the ResNet plugin, produced by EXECUTING the reference's `model.resnet` (model/resnet.py:28-178:
resnet18 -> BasicBlock, resnet50 -> Bottleneck) on CPU with the oracle's deterministic synthetic weights:

    python tests/golden/make_golden_resnet.py        # build container only (needs /root/reference)

Stores the head feature at 64x64 (resnet18, resnet50) and 416x416 (resnet18) plus every block's output at 64x64.  The reference
is imported with make_golden.py's in-memory shims plus one alias for the `nn.init.kaiming_normal` name this torch removed
(resnet.py:119); nothing is copied."""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402
from oracle import yolo2_oracle as O  # noqa: E402


def run(model, config, anchors, name, sizes, acts_at):
    import model.resnet
    sd = O.make_resnet_state_dict(name, seed=0)
    net = getattr(model.resnet, name)(model.ConfigChannels(config), anchors, 20)
    res = net.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all(k.endswith('num_batches_tracked') for k in res.missing_keys), res
    net.eval()
    outs = {}
    hooks = []
    for lname in ('layer1', 'layer2', 'layer3', 'layer4'):
        for bname, m in getattr(net, lname).named_children():
            hooks.append(m.register_forward_hook(lambda mod, inp, out, key='%s.%s' % (lname, bname): outs.__setitem__(key, out.detach().clone())))
    hooks.append(net.maxpool.register_forward_hook(lambda mod, inp, out: outs.__setitem__('maxpool', out.detach().clone())))
    rec = {}
    with torch.no_grad():
        for size, seed in sizes:
            f = net(O.synth_images(1, size, size, seed=seed))
            rec['%s_feature%d' % (name, size)] = f.numpy()
            if size == acts_at:
                rec.update({'%s_act_%s' % (name, k): v.numpy() for k, v in outs.items()})
            # the restatement must agree with the executed reference to fp32 rounding
            o = O.resnet_forward(sd, O.synth_images(1, size, size, seed=seed), name)
            err = ((o - f).norm() / f.norm()).item()
            assert err < 1e-5, (name, size, err)
    for h in hooks:
        h.remove()
    return rec


def main():
    model, utils, detect = G.import_reference()
    if not hasattr(nn.init, 'kaiming_normal'):
        nn.init.kaiming_normal = nn.init.kaiming_normal_
    config = G.make_config(1)
    config.read_dict({'model': {'pretrained': '0'}})
    anchors = O.anchors_yolo_voc()
    rec = {}
    rec.update(run(model, config, anchors, 'resnet18', [(64, 10), (416, 0)], 64))
    rec.update(run(model, config, anchors, 'resnet50', [(64, 10)], None))
    path = os.path.join(HERE, 'resnet.npz')
    np.savez_compressed(path, **rec)
    print('resnet.npz %.1f KB' % (os.path.getsize(path) / 1024), sorted(rec)[:6])


if __name__ == '__main__':
    main()
