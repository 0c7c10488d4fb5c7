#!/usr/bin/env python
"""Tile-shape sweep of the tcgen05 conv kernel over the Darknet-19 layer shapes at a given batch.

For every distinct (H, Cin, Cout, k) of the backbone it times each legal (BLOCK_N, M-subtiles)
configuration with CUDA events (inputs rotate over enough copies to exceed L2 for the small
layers) and writes gpurun_out/conv_sweep.json + a table.  Used to derive the host-side heuristic
in conv_igemm_forward; also checks every configuration against the default one (max abs diff).

    python tools/conv_sweep.py [--batch 32] [--size 416] [--iters 20]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'yolo2-pytorch_b200'))

import torch  # noqa: E402
from b200 import ops  # noqa: E402

# (H=W divisor of the input size, cin, cout, k)
SHAPES = [(2, 32, 64, 3), (4, 64, 128, 3), (4, 128, 64, 1), (8, 128, 256, 3), (8, 256, 128, 1), (16, 256, 512, 3), (16, 512, 256, 1),
          (16, 512, 64, 1), (32, 512, 1024, 3), (32, 1024, 512, 1), (32, 1024, 1024, 3), (32, 1280, 1024, 3), (32, 1024, 125, 1)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--size', type=int, default=416)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--no-pair', action='store_true')
    a = ap.parse_args()
    dev = 'cuda'
    results = []
    for div, cin, cout, k in SHAPES:
        h = a.size // div
        copies = max(2, min(8, int(200e6 / (a.batch * h * h * cin * 2)) + 1))
        xs = [torch.randn(a.batch, h, h, cin, device=dev).half() for _ in range(copies)]
        w = (torch.randn(cout, k, k, cin, device=dev) * (2.0 / (cin * k * k)) ** 0.5).half()
        scale = torch.rand(cout, device=dev) + 0.5
        shift = torch.randn(cout, device=dev) * 0.1
        head = cout == 125
        mode = ops.OUT_F32_NCHW if head else ops.OUT_F16_NHWC
        flops = 2.0 * a.batch * h * h * cin * cout * k * k
        base = None
        row = dict(shape='%dx%d cin%d cout%d k%d' % (h, h, cin, cout, k), gflop=flops / 1e9, configs={})
        for bn in (64, 128, 256):
            if bn > max(64, cout) and not (bn == 128 and cout == 125):
                continue
            if cout <= 64 and bn != 64:
                continue
            for mt, pr in ((1, 1), (2, 1), (1, 2), (2, 2)):
                flags = ops.conv_force_bn(bn) | (mt << 20) | ops.conv_force_pair(pr)
                if pr == 2 and a.no_pair:
                    continue
                try:
                    y = ops.conv_bn_act(xs[0], w, scale, shift, 0.1, out_mode=mode, flags=flags)
                    torch.cuda.synchronize()
                except RuntimeError as e:
                    row['configs']['bn%d_mt%d%s' % (bn, mt, 'p' if pr == 2 else '')] = dict(error=str(e)[:120])
                    continue
                if base is None:
                    base = y.clone()
                diff = (y.float() - base.float()).abs().max().item()
                out = torch.empty_like(y)
                for i in range(3):
                    ops.conv_bn_act(xs[i % copies], w, scale, shift, 0.1, out=out, out_mode=mode, flags=flags)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda._sleep(int(6e6))     # let the host run ahead so launches are back to back
                s.record()
                for i in range(a.iters):
                    ops.conv_bn_act(xs[i % copies], w, scale, shift, 0.1, out=out, out_mode=mode, flags=flags)
                e.record()
                torch.cuda.synchronize()
                us = s.elapsed_time(e) / a.iters * 1e3
                row['configs']['bn%d_mt%d%s' % (bn, mt, 'p' if pr == 2 else '')] = dict(us=us, tflops=flops / us / 1e6, maxdiff_vs_first=diff)
        best = min((v['us'], kk) for kk, v in row['configs'].items() if 'us' in v)
        row['best'] = best[1]
        results.append(row)
        print('%-30s %7.2f GF  ' % (row['shape'], row['gflop']) + '  '.join(
            '%s:%.0f%s' % (kk, v['us'], '' if v['maxdiff_vs_first'] < 1e-2 else ' DIFF!') if 'us' in v else '%s:ERR' % kk
            for kk, v in row['configs'].items()) + '   best=' + row['best'], flush=True)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'conv_sweep.json'), 'w') as f:
        json.dump(results, f, indent=1)
    tot_best = sum(min(v['us'] for v in r['configs'].values() if 'us' in v) * n for r, n in zip(results, [1, 2, 1, 2, 1, 3, 2, 1, 3, 2, 2, 1, 1]))
    print('sum of best per-layer times over the 22 tcgen05 layers: %.1f us' % tot_best)


if __name__ == '__main__':
    main()
