#!/usr/bin/env python
"""GPU bring-up diagnostics for the tcgen05 conv kernel: runs a ladder of cases from a plain GEMM
(1x1, tiled TMA) up to 3x3 im2col layers, never stops at a failure, and writes one JSON record per
case to gpurun_out/diag.json.  Each case compares against (a) the CUDA-core reference kernel on the
same fp16 operands and (b) torch fp32 conv2d on the fp16-rounded operands (oracle arithmetic).
Every case runs in its own subprocess so a trapped kernel (poisoned context) cannot hide the rest.

    python tools/gpu_diag.py            # run all cases
    python tools/gpu_diag.py --case 3   # run one case in-process (used by the parent)
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'yolo2-pytorch_b200'))
sys.path.insert(0, ROOT)

# name, B, H, W, Cin, Cout, k, flags(dict)
CASES = [
    dict(name='gemm_1x1_tiled_64x128', b=1, h=8, w=16, cin=64, cout=128, k=1, tiled=1),
    dict(name='gemm_1x1_tiled_K256', b=2, h=16, w=16, cin=256, cout=128, k=1, tiled=1),
    dict(name='gemm_1x1_im2col', b=2, h=16, w=16, cin=256, cout=128, k=1),
    dict(name='conv3x3_identityish', b=1, h=16, w=16, cin=64, cout=128, k=3, structured=1),
    dict(name='conv3x3_64_128', b=2, h=16, w=16, cin=64, cout=128, k=3),
    dict(name='conv3x3_tail_13x13', b=3, h=13, w=13, cin=128, cout=256, k=3),
    dict(name='conv3x3_bn64', b=2, h=26, w=26, cin=128, cout=64, k=3),
    dict(name='conv3x3_bn256', b=2, h=13, w=13, cin=256, cout=512, k=3, wide=1),
    dict(name='conv3x3_cin32_bk32', b=1, h=32, w=32, cin=32, cout=64, k=3),
    dict(name='conv1x1_head_nchw', b=2, h=13, w=13, cin=1024, cout=125, k=1, nchw=1, noact=1),
    dict(name='conv3x3_multi_tile_persist', b=8, h=52, w=52, cin=128, cout=256, k=3),
    dict(name='conv3x3_big_13x13x1024', b=32, h=13, w=13, cin=1024, cout=1024, k=3),
    dict(name='conv3x3_big_wide', b=32, h=13, w=13, cin=1024, cout=1024, k=3, wide=1),
    dict(name='conv_chan_slice', b=2, h=13, w=13, cin=1024, cout=1024, k=3, ch_off=256, y_ld=1280),
    dict(name='conv_cin1280', b=2, h=13, w=13, cin=1280, cout=1024, k=3),
    # M-subtiles = 2 (256 x BN CTA tiles)
    dict(name='mt2_1x1_tiled', b=2, h=16, w=16, cin=256, cout=128, k=1, tiled=1, mt=2),
    dict(name='mt2_3x3_bn128', b=2, h=16, w=16, cin=64, cout=128, k=3, mt=2),
    dict(name='mt2_3x3_tail_13x13_bn256', b=3, h=13, w=13, cin=128, cout=256, k=3, mt=2, bn=256),
    dict(name='mt2_3x3_odd_subtile_tail', b=5, h=13, w=13, cin=128, cout=256, k=3, mt=2, bn=128),
    dict(name='mt2_cin32_bn64', b=1, h=32, w=32, cin=32, cout=64, k=3, mt=2),
    dict(name='mt2_head_nchw', b=2, h=13, w=13, cin=1024, cout=125, k=1, nchw=1, noact=1, mt=2),
    dict(name='mt2_big_bn256', b=32, h=13, w=13, cin=1024, cout=1024, k=3, mt=2, bn=256),
    dict(name='mt2_multi_tile_persist', b=8, h=52, w=52, cin=128, cout=256, k=3, mt=2, bn=128),
    # CTA pairs (cta_group::2), index 23..
    dict(name='pair_1x1_tiled_single_tile', b=1, h=16, w=16, cin=64, cout=128, k=1, tiled=1, pair=2, mt=1, bn=128),
    dict(name='pair_1x1_im2col', b=2, h=16, w=16, cin=256, cout=128, k=1, pair=2, mt=1, bn=128),
    dict(name='pair_3x3_bn128', b=2, h=16, w=16, cin=64, cout=128, k=3, pair=2, mt=1, bn=128),
    dict(name='pair_3x3_bn256_tail', b=3, h=13, w=13, cin=128, cout=512, k=3, pair=2, mt=1, bn=256),
    dict(name='pair_3x3_bn64', b=2, h=26, w=26, cin=128, cout=64, k=3, pair=2, mt=1, bn=64),
    dict(name='pair_mt2_bn256', b=5, h=13, w=13, cin=256, cout=512, k=3, pair=2, mt=2, bn=256),
    dict(name='pair_mt2_bn128_persist', b=8, h=52, w=52, cin=128, cout=256, k=3, pair=2, mt=2, bn=128),
    dict(name='pair_cin32', b=1, h=32, w=32, cin=32, cout=64, k=3, pair=2, mt=1, bn=64),
    dict(name='pair_head_nchw', b=2, h=13, w=13, cin=1024, cout=125, k=1, nchw=1, noact=1, pair=2, mt=1, bn=128),
    dict(name='pair_big', b=32, h=13, w=13, cin=1024, cout=1024, k=3, pair=2, mt=2, bn=256),
    dict(name='pair_chan_slice', b=2, h=13, w=13, cin=1024, cout=1024, k=3, ch_off=256, y_ld=1280, pair=2, mt=1, bn=256),
]


def run_case(idx):
    import torch
    from b200 import ops, lib
    c = CASES[idx]
    torch.manual_seed(idx)
    dev = 'cuda'
    b, h, w, cin, cout, k = c['b'], c['h'], c['w'], c['cin'], c['cout'], c['k']
    if c.get('structured'):
        # x = small integers, w = delta on one tap per output channel -> output is a shifted copy
        x = torch.randint(-4, 5, (b, h, w, cin), device=dev).half()
        wt = torch.zeros(cout, k, k, cin, device=dev)
        for co in range(cout):
            wt[co, (co // cin) % k, (co // (cin * k)) % k, co % cin] = 1.0
        wt = wt.half()
    else:
        x = (torch.randn(b, h, w, cin, device=dev)).half()
        wt = (torch.randn(cout, k, k, cin, device=dev) * (2.0 / (cin * k * k)) ** 0.5).half()
    scale = torch.rand(cout, device=dev) + 0.5
    shift = torch.randn(cout, device=dev) * 0.1
    slope = 1.0 if c.get('noact') else 0.1
    out_mode = ops.OUT_F32_NCHW if c.get('nchw') else ops.OUT_F16_NHWC
    flags = (ops.CONV_A_TILED if c.get('tiled') else 0) | (ops.CONV_WIDE_N if c.get('wide') else 0)
    flags |= ops.conv_force_mt(c.get('mt', 0)) | ops.conv_force_bn(c.get('bn', 0)) | ops.conv_force_pair(c.get('pair', 0))
    y_ld = c.get('y_ld', cout)
    ch_off = c.get('ch_off', 0)

    def alloc():
        if out_mode == ops.OUT_F16_NHWC:
            return torch.full((b, h, w, y_ld), -7.0, dtype=torch.float16, device=dev)
        return torch.full((b, cout, h, w), -7.0, dtype=torch.float32, device=dev)

    y_ref = ops.conv_bn_act(x, wt, scale, shift, slope, out=alloc(), out_mode=out_mode, y_ch_off=ch_off, ref=True)
    torch.cuda.synchronize()
    rec = dict(name=c['name'], case=c)
    try:
        y = ops.conv_bn_act(x, wt, scale, shift, slope, out=alloc(), out_mode=out_mode, y_ch_off=ch_off, flags=flags)
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        rec['error'] = repr(e)
        rec['debug_word'] = ['%#x' % (v & 0xffffffff) for v in lib.debug_read()]
        return rec
    # oracle arithmetic: fp32 conv on the same fp16-rounded operands
    xo = x.float().permute(0, 3, 1, 2)
    wo = wt.float().permute(0, 3, 1, 2)
    yo = torch.nn.functional.conv2d(xo, wo, padding=(k - 1) // 2) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    yo = torch.where(yo > 0, yo, yo * slope)
    if out_mode == ops.OUT_F16_NHWC:
        got = y[..., ch_off:ch_off + cout].float().permute(0, 3, 1, 2)
        refk = y_ref[..., ch_off:ch_off + cout].float().permute(0, 3, 1, 2)
        untouched_ok = bool((y[..., :ch_off] == -7).all() and (y[..., ch_off + cout:] == -7).all())
    else:
        got, refk = y, y_ref
        untouched_ok = True
    denom = yo.abs().max().item()
    err = (got - yo).abs()
    rec.update(max_ref=denom, rel_vs_oracle=err.max().item() / denom, rel_refkernel_vs_oracle=(refk - yo).abs().max().item() / denom,
               rel_vs_refkernel=(got - refk).abs().max().item() / denom, untouched_ok=untouched_ok,
               nan=int(torch.isnan(got).sum().item()), unwritten=int((got == -7).sum().item()))
    if rec['rel_vs_oracle'] > 2e-3:
        # where is it wrong?  per-pixel-row / per-channel error maps help tell a layout bug from a pipeline bug
        bad = (err > 2e-3 * denom)
        rec['bad_frac'] = bad.float().mean().item()
        rec['bad_by_channel_first16'] = bad.float().mean(dim=(0, 2, 3))[:16].tolist()
        rec['bad_by_row_img0'] = bad[0].float().mean(dim=(0, 2)).tolist()[:32]
        rec['bad_by_col_img0'] = bad[0].float().mean(dim=(0, 1)).tolist()[:32]
        rec['sample_got'] = got[0, :4, :2, :6].tolist()
        rec['sample_ref'] = yo[0, :4, :2, :6].tolist()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--case', type=int, default=-1)
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'diag.json'))
    ap.add_argument('--start', type=int, default=0, help='first case index to run')
    ap.add_argument('--inproc', action='store_true', help='run cases [start, end) in THIS process (fast; a trap ends the run)')
    ap.add_argument('--end', type=int, default=len(CASES))
    a = ap.parse_args()
    if a.case >= 0:
        print('DIAG_JSON ' + json.dumps(run_case(a.case)))
        return
    if a.inproc:
        keys = ('name', 'error', 'rel_vs_oracle', 'rel_vs_refkernel', 'untouched_ok', 'nan', 'unwritten', 'bad_frac', 'debug_word',
                'bad_by_channel_first16', 'bad_by_row_img0', 'sample_got', 'sample_ref')
        for i in range(a.start, min(a.end, len(CASES))):
            print('running %d %s' % (i, CASES[i]['name']), flush=True)
            rec = run_case(i)
            print(json.dumps({k: v for k, v in rec.items() if k in keys}), flush=True)
            if 'error' in rec:
                break
        return
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    results = []
    for i, c in enumerate(CASES):
        if i < a.start:
            continue
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--case', str(i)], capture_output=True, text=True, timeout=300)
            rec = None
            for line in r.stdout.splitlines():
                if line.startswith('DIAG_JSON '):
                    rec = json.loads(line[len('DIAG_JSON '):])
            if rec is None:
                rec = dict(name=c['name'], error='no result', rc=r.returncode, stderr=r.stderr[-1500:])
        except subprocess.TimeoutExpired:
            rec = dict(name=c['name'], error='timeout (300 s)')
        results.append(rec)
        brief = {k: v for k, v in rec.items() if k in ('name', 'error', 'rel_vs_oracle', 'rel_vs_refkernel', 'rel_refkernel_vs_oracle',
                                                        'untouched_ok', 'nan', 'unwritten', 'bad_frac', 'debug_word')}
        print(json.dumps(brief), flush=True)
        with open(a.out, 'w') as f:
            json.dump(results, f, indent=1)


if __name__ == '__main__':
    main()
