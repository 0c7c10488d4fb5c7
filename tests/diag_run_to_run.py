"""Diagnostic (not a test): run-to-run variation of one training step's gradients (same process, same data, fresh model each time)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'yolo2-pytorch_b200')); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402
import test_ddp_nccl as T  # noqa: E402
import model  # noqa: E402
import train as yb_train  # noqa: E402
from b200 import ddp  # noqa: E402

batch = T._batch()
runs = []
for rep in range(3):
    dnn, anchors = T._build(0, True)
    inference = model.Inference(T._config(), dnn, anchors).cuda().train()
    opt = torch.optim.SGD(dnn.parameters(), 0.0)
    with ddp.local_only():
        out = yb_train.iterate(inference, opt, anchors, T._config(0.4), T._shard(batch, 0), reducer=False)
    torch.cuda.synchronize()
    g = {n: p.grad.detach().float().cpu().clone() for n, p in dnn.named_parameters()}
    g['__feature'] = out['pred']['feature'].float().cpu().clone()
    g['__loss'] = torch.stack([out['loss'][k].float().cpu() for k in sorted(out['loss'])])
    runs.append(g)
worst = (0.0, None)
for n in runs[0]:
    for other in runs[1:]:
        e = ((other[n] - runs[0][n]).norm() / runs[0][n].norm().clamp_min(1e-30)).item()
        if e > worst[0]:
            worst = (e, n)
print('feature bitwise equal:', torch.equal(runs[0]['__feature'], runs[1]['__feature']), 'max diff', (runs[0]['__feature'] - runs[1]['__feature']).abs().max().item(),
      'loss', runs[0]['__loss'].tolist(), runs[1]['__loss'].tolist())
for n in ['layers3.1.conv.bias', 'layers3.1.conv.weight', 'layers3.0.bn.weight', 'layers3.0.conv.weight', 'layers2.7.conv.weight', 'layers2.1.conv.weight', 'passthrough.conv.weight',
          'layers1.16.conv.weight', 'layers1.12.conv.weight', 'layers1.8.conv.weight', 'layers1.4.conv.weight', 'layers1.2.conv.weight', 'layers1.0.bn.weight', 'layers1.0.conv.weight']:
    print('  %-26s %.3e' % (n, ((runs[1][n] - runs[0][n]).norm() / runs[0][n].norm().clamp_min(1e-30)).item()))
print('run-to-run worst gradient rel L2 %.3e (%s); layers1.0.conv.weight %.3e' % (worst[0], worst[1],
      ((runs[1]['layers1.0.conv.weight'] - runs[0]['layers1.0.conv.weight']).norm() / runs[0]['layers1.0.conv.weight'].norm()).item()))
