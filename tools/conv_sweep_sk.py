#!/usr/bin/env python
"""Tile shape x stream-K sweep over the Darknet-19 layers whose tile count does not fill 148 SMs.
Prints microseconds per launch (batch 32) for every (BLOCK_N, M-subtiles) with plain tiles and with
stream-K, plus the library's own choice with and without a workspace."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'yolo2-pytorch_b200'))
import torch  # noqa: E402
from b200 import ops  # noqa: E402

LAYERS = [(13, 1024, 1024, 3), (13, 512, 1024, 3), (13, 1280, 1024, 3), (13, 1024, 512, 1), (13, 1024, 125, 1), (26, 256, 512, 3),
          (26, 512, 256, 1), (52, 128, 256, 3), (52, 256, 128, 1), (104, 64, 128, 3), (104, 128, 64, 1)]


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
    b = 32
    ws = ops.conv_workspace('cuda')
    rows = []
    for h, cin, cout, k in LAYERS:
        xs = [torch.randn(b, h, h, cin, device='cuda').half() for _ in range(3)]
        w = (torch.randn(cout, k, k, cin, device='cuda') * 0.05).half()
        sc, sh = torch.ones(cout, device='cuda'), torch.zeros(cout, device='cuda')
        f32 = cout % 8 != 0
        out = torch.empty(b, cout, h, h, device='cuda') if f32 else torch.empty(b, h, h, cout, device='cuda', dtype=torch.float16)
        om = ops.OUT_F32_NCHW if f32 else ops.OUT_F16_NHWC
        it = [0]

        def run(flags, wsp):
            it[0] += 1
            ops.conv_bn_act(xs[it[0] % 3], w, sc, sh, 0.1, out=out, out_mode=om, flags=flags, workspace=wsp)
        rec = {'shape': '%dx%d cin%d cout%d k%d' % (h, h, cin, cout, k)}
        rec['auto_plain'] = timeit(lambda: run(ops.CONV_NO_STREAMK, None))
        rec['auto_ws'] = timeit(lambda: run(0, ws))
        for bn in (64, 128, 256):
            if bn > 64 and bn // 2 >= cout:
                continue
            for mt in (1, 2):
                base = ops.conv_force_bn(bn) | ops.conv_force_mt(mt) | ops.conv_force_pair(1)
                rec['bn%d_mt%d' % (bn, mt)] = timeit(lambda: run(base | ops.CONV_NO_STREAMK, None))
                try:
                    rec['bn%d_mt%d_sk' % (bn, mt)] = timeit(lambda: run(base | ops.CONV_FORCE_STREAMK, ws))
                except RuntimeError as e:
                    rec['bn%d_mt%d_sk' % (bn, mt)] = None
        rows.append(rec)
        print(' '.join('%s=%s' % (k2, v if isinstance(v, str) else ('%.1f' % v if v is not None else 'n/a')) for k2, v in rec.items()), flush=True)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, 'gpurun_out', 'conv_sweep_sk.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
