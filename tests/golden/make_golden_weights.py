#!/usr/bin/env python
"""Golden fixture for the Darknet `.weights` head permutation, produced by EXECUTING the reference's own
`transpose_weight` / `transpose_bias` (convert_darknet_torch.py:37-57).  The module itself cannot be imported
(humanize / utils.train imports), so the two pure functions are extracted from its source with `ast` and exec'd.

    python tests/golden/make_golden_weights.py          # build container only (needs /root/reference)
"""
import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'


def main():
    tree = ast.parse(open(os.path.join(REF, 'convert_darknet_torch.py')).read())
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ('transpose_weight', 'transpose_bias')]
    ns = dict(torch=torch)
    exec(compile(ast.Module(body=wanted, type_ignores=[]), os.path.join(REF, 'convert_darknet_torch.py'), 'exec'), ns)
    out = {}
    g = torch.Generator().manual_seed(42)
    for tag, anchors, classes, cin in (('voc', 5, 20, 16), ('coco', 5, 80, 8), ('one', 3, 1, 4)):
        per = 5 + classes
        w = torch.randn(anchors * per, cin, 1, 1, generator=g)
        b = torch.randn(anchors * per, generator=g)
        out['w_in_' + tag], out['b_in_' + tag] = w.numpy(), b.numpy()
        out['w_out_' + tag] = ns['transpose_weight'](w, anchors).contiguous().numpy()
        out['b_out_' + tag] = ns['transpose_bias'](b, anchors).contiguous().numpy()
        out['anchors_' + tag] = np.int64(anchors)
    np.savez_compressed(os.path.join(HERE, 'darknet_weights.npz'), **out)
    print('darknet_weights.npz %.1f KB' % (os.path.getsize(os.path.join(HERE, 'darknet_weights.npz')) / 1024))


if __name__ == '__main__':
    main()
