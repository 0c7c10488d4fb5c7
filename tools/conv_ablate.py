#!/usr/bin/env python
"""Which pipe bounds the tcgen05 conv kernel?  Times a few layer shapes with parts of the kernel
switched off (flags bits 24..27: 1 = no A (im2col) loads, 2 = no B (weight) loads, 4 = no MMA,
8 = no epilogue stores).  Results of ablated runs are garbage; only the timing matters."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'yolo2-pytorch_b200'))
import torch  # noqa: E402
from b200 import ops  # noqa: E402

CASES = [  # h, cin, cout, k, bn, mt, pair  (bn 0 = no forcing: the small-K kernel for cin 32)
    (208, 32, 64, 3, 0, 0, 0),
    (13, 1024, 1024, 3, 256, 2, 1), (13, 1024, 1024, 3, 256, 1, 1), (13, 1024, 1024, 3, 128, 1, 1),
    (52, 128, 256, 3, 256, 1, 1), (52, 128, 256, 3, 256, 1, 2), (52, 128, 256, 3, 128, 2, 1),
    (26, 256, 512, 3, 256, 1, 1), (104, 64, 128, 3, 128, 2, 1), (208, 32, 64, 3, 64, 2, 1),
    (26, 512, 256, 1, 256, 1, 1),
]
# `python tools/conv_ablate.py net`: every distinct Darknet-19 layer shape at batch 32 with the library's own tile choice
NET_CASES = [(208, 32, 64, 3, 0, 0, 0), (104, 64, 128, 3, 0, 0, 0), (104, 128, 64, 1, 0, 0, 0), (52, 128, 256, 3, 0, 0, 0), (52, 256, 128, 1, 0, 0, 0),
             (26, 256, 512, 3, 0, 0, 0), (26, 512, 256, 1, 0, 0, 0), (26, 512, 64, 1, 0, 0, 0), (13, 512, 1024, 3, 0, 0, 0), (13, 1024, 512, 1, 0, 0, 0),
             (13, 1024, 1024, 3, 0, 0, 0), (13, 1280, 1024, 3, 0, 0, 0)]
ABL = [(0, 'full'), (8, 'no-store'), (1, 'no-A'), (2, 'no-B'), (3, 'no-A,B'), (4, 'no-MMA'), (12, 'no-MMA,store'), (7, 'none(A,B,MMA)'), (15, 'empty')]


def main():
    b = 32
    cases = NET_CASES if len(sys.argv) > 1 and sys.argv[1] == 'net' else CASES
    for h, cin, cout, k, bn, mt, pr in cases:
        xs = [torch.randn(b, h, h, cin, device='cuda').half() for _ in range(3)]
        w = (torch.randn(cout, k, k, cin, device='cuda') * 0.05).half()
        sc, sh = torch.ones(cout, device='cuda'), torch.zeros(cout, device='cuda')
        out = torch.empty(b, h, h, cout, device='cuda', dtype=torch.float16)
        line = '%3dx%-3d cin%-4d cout%-4d k%d bn%d mt%d p%d: ' % (h, h, cin, cout, k, bn, mt, pr)
        for code, name in ABL:
            flags = ((ops.conv_force_bn(bn) | ops.conv_force_mt(mt) | ops.conv_force_pair(pr)) if bn else 0) | (code << 24)
            for i in range(3):
                ops.conv_bn_act(xs[i % 3], w, sc, sh, 0.1, out=out, flags=flags)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(int(4e6))
            s.record()
            for i in range(10):
                ops.conv_bn_act(xs[i % 3], w, sc, sh, 0.1, out=out, flags=flags)
            e.record()
            torch.cuda.synchronize()
            us = s.elapsed_time(e) * 100
            line += '%s=%.0f  ' % (name, us)
            if code == 0:
                line += '(%.0f TF/s)  ' % (2.0 * b * h * h * cin * cout * k * k / us / 1e6)
        print(line, flush=True)


if __name__ == '__main__':
    main()
