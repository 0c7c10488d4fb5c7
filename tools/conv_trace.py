#!/usr/bin/env python
"""Per-event timeline of block 0 of a tcgen05 conv launch (yb_conv_set_trace): prints clock64 deltas
for the TMA producer, the MMA issuer and the epilogue so stalls in the pipeline hand-shake show up."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'yolo2-pytorch_b200'))
import torch  # noqa: E402
from b200 import lib, ops  # noqa: E402


def run(h, cin, cout, k, flags, b=32, n=48):
    x = torch.randn(b, h, h, cin, device='cuda').half()
    w = (torch.randn(cout, k, k, cin, device='cuda') * 0.05).half()
    sc, sh = torch.ones(cout, device='cuda'), torch.zeros(cout, device='cuda')
    oh = h // 2 if flags & ops.CONV_POOL2X2 else h
    out = torch.empty(b, oh, oh, cout, device='cuda', dtype=torch.float16)
    for _ in range(3):
        ops.conv_bn_act(x, w, sc, sh, 0.1, out=out, flags=flags)
    tr = torch.zeros(768, dtype=torch.int64, device='cuda')
    lib.load().yb_conv_set_trace(tr.data_ptr())
    ops.conv_bn_act(x, w, sc, sh, 0.1, out=out, flags=flags)
    torch.cuda.synchronize()
    lib.load().yb_conv_set_trace(None)
    t = tr.cpu().view(3, 256)
    t0 = int(t[t > 0].min())
    for role, name in enumerate(('producer', 'mma', 'epilogue')):
        v = [int(c) - t0 for c in t[role].tolist() if c > 0][:n]
        d = [v[i] - v[i - 1] for i in range(1, len(v))]
        print('  %-9s first@%d deltas(cycles): %s' % (name, v[0] if v else -1, ' '.join(str(c) for c in d)), flush=True)


def main():
    for fl, knm in ((0, 'halo'), (ops.CONV_POOL2X2, 'halo+pool')):
        for code, nm in ((0, 'full'), (13, 'empty'), (4, 'no-MMA')):
            print('208x208 cin32 cout64 k3 %s [%s]' % (knm, nm))
            run(208, 32, 64, 3, (code << 24) | fl)
    if '--all' in sys.argv:
        for code, nm in ((0, 'full'), (15, 'empty')):
            print('208x208 cin32 cout64 k3 small-K im2col [%s]' % nm)
            run(208, 32, 64, 3, (code << 24) | ops.CONV_C32_IM2COL)
            print('13x13 cin1024 cout1024 k3 bn256 mt2 [%s]' % nm)
            run(13, 1024, 1024, 3, ops.conv_force_bn(256) | ops.conv_force_mt(2) | ops.conv_force_pair(1) | (code << 24))


if __name__ == '__main__':
    main()
