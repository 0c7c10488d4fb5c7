#!/usr/bin/env python
"""Golden fixture for the training-side input pipeline, produced by EXECUTING the reference's own functions with cv2:
`transform.augmentation.flip_horizontally` (transform/augmentation.py:87-95) and `transform.resize.label.random_crop` -> `resize` ->
`rescale` (transform/resize/label.py:25-31,44-46,58-75), i.e. the default `augmentation` flip followed by `resize_train = RandomCrop`
(config.ini:47-48).  The modules import `inflection` (absent), so the pure functions are extracted with `ast`; numpy / cv2 are real.
`np.random.rand(4)` is seeded per case and the four draws are stored, so the GPU test replays the same window.

    python tests/golden/make_golden_augment.py          # build container only (needs /root/reference + cv2)
"""
import ast
import configparser
import hashlib
import inspect
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import yolo2_oracle as O  # noqa: E402

REF = '/root/reference'
CASES = [(0, 375, 500, 416, 416, True), (1, 480, 640, 608, 608, False), (2, 333, 500, 320, 320, True), (3, 120, 90, 416, 416, False),
         (4, 720, 1280, 416, 416, True), (5, 97, 131, 320, 608, True)]


def extract(path, names):
    tree = ast.parse(open(path).read())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    ns = dict(np=np, cv2=cv2, inspect=inspect)
    exec(compile(ast.Module(body=fns, type_ignores=[]), path, 'exec'), ns)
    return ns


def extract_with_classes(path, names, extra):
    tree = ast.parse(open(path).read())
    nodes = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    ns = dict(np=np, cv2=cv2, inspect=inspect)
    ns.update(extra)
    exec(compile(ast.Module(body=nodes, type_ignores=[]), path, 'exec'), ns)
    return ns


def main():
    lab = extract(os.path.join(REF, 'transform', 'resize', 'label.py'), ('rescale', 'resize', 'random_crop'))
    aug = extract(os.path.join(REF, 'transform', 'augmentation.py'), ('flip_horizontally',))
    config = configparser.ConfigParser()
    config.read_dict({'data': {'resize': 'rescale'}, 'augmentation': {'random_crop': '1'}})
    out = dict(cases=np.array([c[:5] + (int(c[5]),) for c in CASES]))
    for seed, h0, w0, h, w, flip in CASES:
        src = O.synth_frame(seed, h0, w0)
        g = np.random.RandomState(100 + seed)
        n = 1 + seed % 4
        yx_min = (g.rand(n, 2) * np.array([h0 * 0.5, w0 * 0.5]) + np.array([h0 * 0.1, w0 * 0.1])).astype(np.float32)
        yx_max = (yx_min + g.rand(n, 2) * np.array([h0 * 0.3, w0 * 0.3]) + 4).astype(np.float32)
        out['c%d_yx_min_in' % seed], out['c%d_yx_max_in' % seed] = yx_min.copy(), yx_max.copy()
        image, a, b = src, yx_min.copy(), yx_max.copy()
        if flip:
            image, a, b = aug['flip_horizontally'](image, a, b)
        np.random.seed(200 + seed)
        out['c%d_draws' % seed] = np.random.rand(4)
        np.random.seed(200 + seed)                         # random_crop draws the same four numbers
        image, a, b = lab['random_crop'](config, image, a, b, h, w)
        out['c%d_sha' % seed] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(image).tobytes()).digest(), np.uint8)
        out['c%d_yx_min' % seed], out['c%d_yx_max' % seed] = a, b
    # flip alone (same-size output), stored in full for one small frame
    small = O.synth_frame(9, 37, 53)
    bmin, bmax = np.array([[3.0, 5.0], [10.0, 20.5]], np.float32), np.array([[30.0, 40.0], [33.25, 50.0]], np.float32)
    f, a, b = aug['flip_horizontally'](small, bmin.copy(), bmax.copy())
    out['flip_src'], out['flip_out'], out['flip_min_in'], out['flip_max_in'], out['flip_min'], out['flip_max'] = small, f, bmin, bmax, a, b
    # rotation (transform/augmentation.py:28-76): the reference's Rotator + random_rotate executed with cv2, angles drawn by random.uniform
    import random
    rot = extract_with_classes(os.path.join(REF, 'transform', 'augmentation.py'), ('Rotator', 'random_rotate'), dict(random=random))
    cfg_rot = configparser.ConfigParser()
    cfg_rot.read_dict({'augmentation': {'random_rotate': '-7 7'}})
    for seed, h0, w0 in ((0, 120, 160), (1, 333, 500), (2, 97, 61), (3, 416, 416)):
        src = O.synth_frame(20 + seed, h0, w0)
        g = np.random.RandomState(300 + seed)
        yx_min = (g.rand(3, 2) * np.array([h0 * 0.5, w0 * 0.5])).astype(np.float32)
        yx_max = (yx_min + g.rand(3, 2) * np.array([h0 * 0.4, w0 * 0.4]) + 2).astype(np.float32)
        random.seed(400 + seed)
        image, a, b = rot['random_rotate'](cfg_rot, src, yx_min.copy(), yx_max.copy())
        out['r%d_dims' % seed] = np.array([h0, w0, image.shape[0], image.shape[1]])
        out['r%d_sha' % seed] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(image).tobytes()).digest(), np.uint8)
        out['r%d_yx_min_in' % seed], out['r%d_yx_max_in' % seed], out['r%d_yx_min' % seed], out['r%d_yx_max' % seed] = yx_min, yx_max, a, b
    out['rot_cases'] = np.array([0, 1, 2, 3])
    # `fixed` (transform/resize/image.py:36-46), shrinking cases (warpAffine runs INTER_AREA as INTER_LINEAR)
    fx = extract(os.path.join(REF, 'transform', 'resize', 'image.py'), ('fixed',))
    for seed, h0, w0, h, w in ((0, 375, 500, 320, 320), (1, 480, 640, 416, 416), (2, 900, 500, 416, 608)):
        src = O.synth_frame(30 + seed, h0, w0)
        r = fx['fixed'](src, h, w)
        out['f%d_dims' % seed] = np.array([h0, w0, h, w])
        out['f%d_sha' % seed] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(r).tobytes()).digest(), np.uint8)
    path = os.path.join(HERE, 'augment.npz')
    np.savez_compressed(path, **out)
    print('augment.npz %.1f KB' % (os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
