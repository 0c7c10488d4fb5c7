#!/usr/bin/env python
"""Golden fixture for BASELINE configs[1] ("C2": Darknet-19 416x416 batch-32 inference + decode + NMS), produced by EXECUTING the
reference's own modules on CPU: model.yolo2.Darknet (eval) -> model.Inference -> F.softmax -> detect.postprocess (fix = 1) for a
32-image synthetic batch (oracle generator, seed 32; weights: the oracle's deterministic state_dict).

Stored: the head feature of images 0, 15 and 31 in full, max|feature| of every image, and for every image the detections the reference
returns (boxes, classes, scores, objectness of the kept boxes, and for every detection the index of the box among the image's 845 predictions it is a copy of) as ragged arrays -- the GPU test computes the whole batch and compares.

    python tests/golden/make_golden_c2.py          # build container only (needs /root/reference)
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.filterwarnings('ignore')
import make_golden as G  # noqa: E402
from oracle import yolo2_oracle as O  # noqa: E402

BATCH, SEED, KEEP = 32, 32, (0, 15, 31)


def main():
    model, utils, detect = G.import_reference()
    sd = O.make_state_dict(seed=0)
    dnn, anchors, config = G.build_ref_darknet(model, sd)
    x = O.synth_images(BATCH, 416, 416, seed=SEED)
    inference = model.Inference(config, dnn, anchors)
    inference.eval()
    out = dict(images=np.array(KEEP), batch=np.array(BATCH), seed=np.array(SEED))
    det = dict(iou=[], yx_min=[], yx_max=[], cls=[], score=[], box=[])
    n_keep, n_det = [], []
    with torch.no_grad():
        pred = model._inference(inference, x)
        prob = torch.nn.functional.softmax(detect.get_logits(pred), -1)
        for bi in range(BATCH):
            iou, yx_min, yx_max, p = (t[bi].reshape(-1, *t.shape[3:]) if t.dim() > 3 else t[bi].reshape(-1)
                                      for t in (pred['iou'], pred['yx_min'], pred['yx_max'], prob))
            res = detect.postprocess(config, iou, yx_min, yx_max, p)
            if res is None:
                n_keep.append(0); n_det.append(0)
                continue
            n_keep.append(len(res[0])); n_det.append(len(res[3]))
            for name, t in zip(('iou', 'yx_min', 'yx_max', 'cls', 'score'), res):
                det[name].append(t.numpy())
            det['box'].append(G.box_indices(yx_min, yx_max, res[1], res[2]))
    feature = pred['feature'].numpy()
    out['feature'] = feature[list(KEEP)]
    out['feature_absmax'] = np.abs(feature).reshape(BATCH, -1).max(1)
    out['n_keep'], out['n_det'] = np.array(n_keep), np.array(n_det)
    for name, parts in det.items():
        out['det_' + name] = np.concatenate(parts, 0)
    path = os.path.join(HERE, 'c2_batch32.npz')
    np.savez_compressed(path, **out)
    print('c2_batch32.npz %.1f KB; kept boxes %d, detections %d' % (os.path.getsize(path) / 1024, sum(n_keep), sum(n_det)))


if __name__ == '__main__':
    main()
