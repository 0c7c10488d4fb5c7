#!/usr/bin/env python
"""Which pipe bounds the tcgen05 weight-gradient kernel?  Times the Darknet-19 wgrad shapes (batch 64) with parts of the
kernel switched off through YB_WGRAD_SKIP (1 = no x loads, 2 = no dz loads, 4 = no MMA, 8 = no stores) and with the
split-K factor forced through YB_WGRAD_SPLITS.  Results of ablated runs are garbage; only the timing matters."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'yolo2-pytorch_b200'))
import torch  # noqa: E402
from b200 import ops  # noqa: E402

CASES = [(208, 32, 64, 3), (104, 64, 128, 3), (104, 128, 64, 1), (52, 128, 256, 3), (26, 256, 512, 3), (26, 512, 256, 1), (13, 512, 1024, 3),
         (13, 1024, 1024, 3), (13, 1280, 1024, 3), (13, 1024, 512, 1)]
ABL = [(0, 'full'), (1, 'no-x'), (3, 'no-loads')]


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(2e6))
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1000.0 / n


def check():
    """every variant against torch on a small case (fp32 reference of the fp16 operands)"""
    torch.manual_seed(0)
    for h, cin, cout, k, b in ((13, 128, 192, 3, 3), (10, 64, 64, 1, 5), (26, 256, 128, 3, 2)):
        x = torch.randn(b, h, h, cin, device='cuda').half()
        dz = (torch.randn(b, h, h, cout, device='cuda') * 0.1).half()
        ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, k, k), dz.float().permute(0, 3, 1, 2), padding=(k - 1) // 2)
        ref = ref.permute(0, 2, 3, 1).contiguous()
        for cfg in range(6):
            os.environ['YB_WGRAD_CFG'] = str(cfg)
            dw = torch.empty(cout, k, k, cin, device='cuda', dtype=torch.float32)
            ops.call('yb_conv_wgrad', x, dz, dw, b, h, h, cin, cout, k, cin, cout)
            err = ((dw - ref).norm() / ref.norm()).item()
            print('check %dx%d cin%d cout%d k%d cfg%d: rel L2 err %.2e %s' % (h, h, cin, cout, k, cfg, err, 'OK' if err < 2e-3 else 'FAIL'), flush=True)
    del os.environ['YB_WGRAD_CFG']


def main():
    check()
    b = int(os.environ.get('B', '64'))
    only = os.environ.get('ONLY')
    for h, cin, cout, k in CASES:
        if only and only != '%dx%d' % (h, cin):
            continue
        x = torch.randn(b, h, h, cin, device='cuda').half()
        dz = torch.randn(b, h, h, cout, device='cuda').half()
        dw = torch.empty(cout, k, k, cin, device='cuda', dtype=torch.float32)
        gflop = 2.0 * b * h * h * cin * cout * k * k / 1e9

        def run():
            ops.call('yb_conv_wgrad', x, dz, dw, b, h, h, cin, cout, k, cin, cout)
        line = '%3dx%-3d cin%-4d cout%-4d k%d (%.0f GF): ' % (h, h, cin, cout, k, gflop)
        for code, name in ABL:
            os.environ['YB_WGRAD_SKIP'] = str(code)
            line += '%s=%.0f  ' % (name, timeit(run))
        os.environ['YB_WGRAD_SKIP'] = '0'
        line += '| splits: '
        for sp in (1, 2, 4, 16):
            os.environ['YB_WGRAD_SPLITS'] = str(sp)
            line += '%d=%.0f  ' % (sp, timeit(run))
        del os.environ['YB_WGRAD_SPLITS']
        line += '| cfg(kp,stages,nmax): '
        for cfg, name in ((0, '32,4,512'), (1, '64,2,512'), (2, '64,4,256'), (3, '128,2,256'), (4, '64,3,384'), (5, '128,3,128')):
            os.environ['YB_WGRAD_CFG'] = str(cfg)
            line += '%s=%.0f  ' % (name, timeit(run))
        del os.environ['YB_WGRAD_CFG']
        us = timeit(run)
        print(line + '| auto %.0f us = %.0f TF/s' % (us, gflop / us / 1e3 * 1e3), flush=True)


if __name__ == '__main__':
    main()
