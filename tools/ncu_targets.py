#!/usr/bin/env python
"""Workloads for the ncu captures (tools/ncu_capture.sh): a few eager warm-up passes, then ONE pass between cudaProfilerStart/Stop
(run ncu with `--profile-from-start off`), so the capture holds exactly one step's launches in network order.

    python tools/ncu_targets.py infer       # C2: B=32 @ 416 backbone + decode + filter/NMS (every kernel of the step)
    python tools/ncu_targets.py strict      # the same with precision='strict'
    python tools/ncu_targets.py train       # C3: B=64 @ 416 forward + region loss + backward (no optimizer)
    python tools/ncu_targets.py mobilenet   # C5: MobileNet B=32 @ 416
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'yolo2-pytorch_b200'))
import torch  # noqa: E402

import bench  # noqa: E402


def main(what):
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    g = torch.Generator().manual_seed(0)
    if what in ('infer', 'strict'):
        from b200.pipeline import DetectPipeline
        config, dnn, inference = bench.build_model(dev)
        if what == 'strict':
            dnn.engine.set_precision('strict')
        pipe = DetectPipeline(inference, config, 32, 416, 416, slots=1, lanes=1, use_graph=False).prepare()
        pipe.x[0].copy_(torch.rand(32, 3, 416, 416, generator=g))
        step = lambda: pipe._forward(pipe.x[0])  # noqa: E731
    elif what == 'train':
        import model
        config, dnn, inference, anchors, optimizer = bench._train_setup(dev, capturable=False)
        batch = bench._train_batches(64, 416, 416, dev, g, count=1)[0]
        import train as yb_train

        def step():
            pred = model._inference(inference, batch['tensor'])
            rows, cols = pred['feature'].shape[-2:]
            loss, _ = model.loss(anchors, yb_train.norm_data(batch, 416, 416, rows, cols), pred, 0.6, True)
            total = sum(loss[k] * config.getfloat('hparam', k) for k in loss)
            optimizer.zero_grad()
            total.backward()
    elif what == 'mobilenet':
        import model
        import model.mobilenet
        config = bench.make_config()
        anchors = torch.tensor(bench.ANCHORS_HW)
        dnn = model.mobilenet.MobileNet(model.ConfigChannels(config), anchors, 20).to(dev).eval()
        x = torch.rand(32, 3, 416, 416, generator=g).to(dev)
        step = lambda: dnn(x)  # noqa: E731
    else:
        raise SystemExit(__doc__)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'infer')
