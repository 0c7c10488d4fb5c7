#!/usr/bin/env python
"""Golden fixture for the input pipeline, produced by EXECUTING the reference's `transform.resize.label.rescale`
(transform/resize/label.py:25-31 -> cv2.resize, INTER_LINEAR) and `BGR2RGB` on the repo's own image.jpg and on synthetic
frames.  The module imports `inflection` (absent), so the pure function is extracted with `ast`; cv2 is the real thing.

    python tests/golden/make_golden_resize.py          # build container only (needs /root/reference + cv2)

Large outputs are stored as SHA-256 digests (the comparison is bit-exact anyway); one small case is stored in full."""
import ast
import hashlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import yolo2_oracle as O  # noqa: E402

REF = '/root/reference'
CASES = [(0, 375, 500, 416, 416), (1, 480, 640, 608, 608), (2, 333, 500, 320, 320), (3, 100, 80, 416, 416), (4, 1080, 1920, 416, 416),
         (5, 416, 416, 416, 416), (6, 13, 17, 320, 608), (7, 500, 375, 608, 320)]


def main():
    tree = ast.parse(open(os.path.join(REF, 'transform', 'resize', 'label.py')).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'rescale']
    ns = dict(np=np, cv2=cv2)
    exec(compile(ast.Module(body=fn, type_ignores=[]), os.path.join(REF, 'transform/resize/label.py'), 'exec'), ns)
    rescale = ns['rescale']
    out = {}
    img = cv2.imread(os.path.join(REF, 'image.jpg'))                      # BGR uint8 [576, 768, 3]
    g = np.random.RandomState(7)
    yx_min = (g.rand(5, 2) * np.array([400, 500])).astype(np.float32)
    yx_max = yx_min + (g.rand(5, 2) * 150 + 10).astype(np.float32)
    r, a, b = rescale(img, yx_min.copy(), yx_max.copy(), 416, 416)
    rgb = cv2.cvtColor(r, cv2.COLOR_BGR2RGB)
    out['jpg_sha_bgr'] = np.frombuffer(hashlib.sha256(r.tobytes()).digest(), np.uint8)
    out['jpg_sha_rgb'] = np.frombuffer(hashlib.sha256(rgb.tobytes()).digest(), np.uint8)
    out['jpg_src_sha'] = np.frombuffer(hashlib.sha256(img.tobytes()).digest(), np.uint8)
    out['jpg_yx_min_in'], out['jpg_yx_max_in'], out['jpg_yx_min'], out['jpg_yx_max'] = yx_min, yx_max, a, b
    out['jpg_crop'] = r[100:132, 200:232].copy()
    for seed, h0, w0, h, w in CASES:
        src = O.synth_frame(seed, h0, w0)
        r, _, _ = rescale(src, np.zeros((1, 2), np.float32), np.ones((1, 2), np.float32), h, w)
        out['case%d_sha' % seed] = np.frombuffer(hashlib.sha256(r.tobytes()).digest(), np.uint8)
        out['case%d_dims' % seed] = np.array([h0, w0, h, w])
    small = O.synth_frame(9, 37, 53)
    out['small_src'] = small
    out['small_out'], _, _ = rescale(small, np.zeros((1, 2), np.float32), np.ones((1, 2), np.float32), 64, 96)
    path = os.path.join(HERE, 'resize.npz')
    np.savez_compressed(path, **out)
    print('resize.npz %.1f KB' % (os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
