#!/usr/bin/env python
"""Bring-up check for the halo-tile Cin=32 kernel: error against the CUDA-core reference kernel for both
LBO/SBO orientations of the un-swizzled A descriptor, then timings (plain / fused pool / im2col kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'yolo2-pytorch_b200'))
import torch  # noqa: E402
from b200 import ops  # noqa: E402


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(4e6))
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1000.0 / n


def main():
    torch.manual_seed(0)
    for (b, h, w, cout) in ((1, 16, 8, 64), (2, 37, 21, 64), (2, 32, 32, 48)):
        x = torch.randn(b, h, w, 32, device='cuda').half()
        wt = (torch.randn(cout, 3, 3, 32, device='cuda') * 0.08).half()
        sc, sh = torch.rand(cout, device='cuda') + 0.5, torch.randn(cout, device='cuda') * 0.1
        ref = ops.conv_bn_act(x, wt, sc, sh, 0.1, ref=True).float()
        for name, fl in (('default', 0), ('swapped', ops.CONV_C32_SWAP), ('im2col', ops.CONV_C32_IM2COL)):
            try:
                y = ops.conv_bn_act(x, wt, sc, sh, 0.1, flags=fl).float()
                torch.cuda.synchronize()
                err = ((y - ref).abs().max() / ref.abs().max()).item()
            except RuntimeError as e:
                err = str(e)[:80]
            print('shape %s cout %d  %-8s rel err %s' % ((b, h, w), cout, name, err), flush=True)
    b, h = 32, 208
    xs = [torch.randn(b, h, h, 32, device='cuda').half() for _ in range(3)]
    wt = (torch.randn(64, 3, 3, 32, device='cuda') * 0.08).half()
    sc, sh = torch.ones(64, device='cuda'), torch.zeros(64, device='cuda')
    out = torch.empty(b, h, h, 64, device='cuda', dtype=torch.float16)
    outp = torch.empty(b, h // 2, h // 2, 64, device='cuda', dtype=torch.float16)
    it = [0]

    def run(fl, o):
        it[0] += 1
        ops.conv_bn_act(xs[it[0] % 3], wt, sc, sh, 0.1, out=o, flags=fl)
    for code, nm in ((0, 'full'), (8, 'no-store'), (1, 'no-A'), (4, 'no-MMA'), (5, 'no-A,MMA'), (13, 'empty')):
        print('halo %-9s plain %.1f us   fused-pool %.1f us' % (nm, timeit(lambda: run(code << 24, out)),
                                                             timeit(lambda: run((code << 24) | ops.CONV_POOL2X2, outp))), flush=True)
    print('im2col small-K kernel %.1f us; generic kernel %.1f us; maxpool alone %.1f us' % (
        timeit(lambda: run(ops.CONV_C32_IM2COL, out)), timeit(lambda: run(ops.CONV_NO_SMALLK, out)), timeit(lambda: ops.maxpool2x2(out, out=outp))))


if __name__ == '__main__':
    main()
