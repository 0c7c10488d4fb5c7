#!/usr/bin/env python
"""Per-layer diagnosis of the training path vs the oracle (train-mode BN): forward activations, batch statistics,
then gradients (dz per layer is not exposed; parameter-gradient cosines are printed per tensor)."""
import configparser
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'yolo2-pytorch_b200'))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from oracle import yolo2_oracle as O  # noqa: E402
import model  # noqa: E402
import model.yolo2  # noqa: E402


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    cfg = configparser.ConfigParser()
    cfg.read_dict({'batch_norm': {'enable': '1'}})
    anchors = O.anchors_yolo_voc()
    sd0 = O.make_state_dict(0)
    s = size // 32
    x = O.synth_images(b, size, size, seed=12)
    data = O.norm_data(O.synth_targets(b, size, size, slots=6, seed=13), size, size, s, s)
    sd = {k: (v.clone().requires_grad_(True) if 'running' not in k else v.clone()) for k, v in sd0.items()}
    collect, stats = {}, {}
    f_ref = O.darknet_forward(sd, x, collect=collect, train=True, stats=stats)
    losses, _ = O.loss(anchors, data, O.decode(f_ref, anchors), 0.6)
    O.loss_total(losses).backward()

    dnn = model.yolo2.Darknet(model.ConfigChannels(cfg), anchors, 20)
    dnn.load_state_dict(sd0, strict=False)
    dnn = dnn.cuda().train()
    feature, saved = dnn.trainer.forward(x.cuda())
    print('feature rel %.3e' % rel(feature, f_ref))
    for key, su in saved.units.items():
        u = su.u
        z = su.z.float()
        y = (z - su.mean) * su.invstd * u.bn.weight.detach() + u.bn.bias.detach()
        a = torch.where(y > 0, y, y * 0.1).permute(0, 3, 1, 2)
        m_ref, v_ref = stats[key]
        print('%-12s act rel %.3e | mean rel %.3e | invstd rel %.3e | min var %.3e' % (
            key, rel(a, collect[key]), rel(su.mean, m_ref), rel(su.invstd, 1.0 / torch.sqrt(v_ref + 1e-5)), v_ref.min().item()))
    pred = dict(feature=feature.detach().requires_grad_(True))
    l_gpu, _ = model.loss(anchors, {k: v.cuda() for k, v in data.items()}, pred, 0.6)
    total = sum(l_gpu[k] * O.HPARAM_DEFAULT[k] for k in l_gpu)
    total.backward()
    # use the ORACLE's dfeature so that backward is compared on identical upstream gradients
    f2 = f_ref.detach().clone().requires_grad_(True)
    l2, _ = O.loss(anchors, data, O.decode(f2, anchors), 0.6)
    O.loss_total(l2).backward()
    print('dfeature rel (own feature) %.3e' % rel(pred['feature'].grad, f2.grad))
    grads = dnn.trainer.backward(saved, f2.grad.cuda())
    for name in sorted(grads, key=lambda n: list(sd0).index(n) if n in sd0 else 1e9):
        g, r = grads[name].float().cpu().flatten(), sd[name].grad.flatten()
        cos = (torch.dot(g, r) / (g.norm() * r.norm() + 1e-30)).item()
        print('%-28s cos %.5f  rel %.3e  |ref| %.3e' % (name, cos, ((g - r).norm() / (r.norm() + 1e-30)).item(), r.norm().item()))


if __name__ == '__main__':
    main()
