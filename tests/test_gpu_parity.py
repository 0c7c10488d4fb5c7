"""GPU parity tests: the CUDA path (through the C ABI via b200.ops / the plugin modules) against
the CPU oracle (oracle/yolo2_oracle.py) on identical seeded inputs, and against the committed golden
fixtures produced by the reference itself (tests/golden/*.npz).

Tolerances (BASELINE.json north_star): conv/BN activations within 1e-3 relative (max|d|/max|ref|);
NMS survivor indices, reorg, pooling, filtering: bit-exact; decode within 1e-5 relative.
"""
import configparser
import os

import numpy as np
import pytest
import torch

from oracle import yolo2_oracle as O

pytestmark = pytest.mark.gpu

DEV = 'cuda'


def rel_err(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def rel_l2(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()


MEASURED = {}


def record(name, value):
    """Measured parity figures of this run -> gpurun_out/parity_measured.json (copied to profiles/ when committed)."""
    import json
    MEASURED[name] = value
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(root, 'gpurun_out', 'parity_measured.json'), 'w') as f:
        json.dump(MEASURED, f, indent=1, sort_keys=True)


# Contract (BASELINE.json north_star): conv/BN activations within 1e-3 relative of the reference's fp32 -- asserted for
# precision='strict'.  The default precision='fast' (fp16 operands, one tensor-core pass) has a documented end-to-end drift of up to
# 2.5e-3 (23 layers x 2 operand roundings of 2e-4 each, tools/error_budget.py); its per-layer error stays inside 1e-3.
TOL_CONTRACT = 1e-3
TOL_FAST_E2E = 2.5e-3


def make_config(fix):
    config = configparser.ConfigParser()
    config.read_dict({'batch_norm': {'enable': '1'},
                      'detect': {'threshold': '0.3', 'threshold_cls': '0.005', 'fix': str(int(fix)), 'overlap': '0.45'}})
    return config


@pytest.fixture(scope='module')
def ops():
    from b200 import ops
    return ops


# ------------------------------------------------------------------------------------------------
# pure data movement: bit-exact
# ------------------------------------------------------------------------------------------------
def test_reorg_f32_golden(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, 'reorg.npz'))
    y = ops.reorg_f32_nchw(torch.from_numpy(g['x']).to(DEV))
    assert np.array_equal(y.cpu().numpy(), g['y'])


def test_reorg_public_function_large():
    import model.yolo2
    x = torch.randn(4, 64, 26, 26, device=DEV)
    assert torch.equal(model.yolo2.reorg(x).cpu(), O.reorg(x.cpu()))


def test_reorg_f16_into_concat_slice(ops):
    x = torch.randn(3, 26, 26, 64, device=DEV).half()
    cat = torch.full((3, 13, 13, 1280), 5.0, dtype=torch.float16, device=DEV)
    ops.reorg_f16(x, cat, 0)
    ref = O.reorg(x.float().permute(0, 3, 1, 2).cpu()).permute(0, 2, 3, 1).half()
    assert torch.equal(cat[..., :256].cpu(), ref)
    assert bool((cat[..., 256:] == 5).all())


def test_maxpool_exact(ops):
    x = torch.randn(2, 52, 52, 128, device=DEV).half()
    y = ops.maxpool2x2(x)
    ref = torch.nn.functional.max_pool2d(x.float().permute(0, 3, 1, 2).cpu(), 2).permute(0, 2, 3, 1).half()
    assert torch.equal(y.cpu(), ref)


# ------------------------------------------------------------------------------------------------
# IoU / NMS / filter: bit-exact indices
# ------------------------------------------------------------------------------------------------
def test_iou_known_answers(golden_dir):
    import utils.iou.torch as iou
    g = np.load(os.path.join(golden_dir, 'iou.npz'))
    t = lambda k: torch.from_numpy(g[k]).to(DEV)
    m0 = iou.iou_matrix(t('c_min'), t('c_max'), t('d_min'), t('d_max')).cpu().numpy()
    np.testing.assert_almost_equal(m0, np.zeros((1, 8), np.float32))                                      # utils/iou/torch.py:79-95
    m1 = iou.iou_matrix(t('a_min'), t('a_max'), t('b_min'), t('b_max')).cpu().numpy()
    np.testing.assert_almost_equal(m1, np.array([[1 / 7] * 4, [4 / 16] * 4], np.float32))                 # :97-113
    assert np.array_equal(m1, g['m1'])
    mb = iou.batch_iou_matrix(t('r_min'), t('r_max'), t('s_min'), t('s_max')).cpu().numpy()
    assert np.array_equal(mb, g['mb'])                                                                    # bit-exact vs the reference
    # batch_iou_pair known answers (utils/iou/torch.py:236-289): box1 tiled over cells, box2 tiled over the batch
    for bbox1, bbox2, ans in (([(1, 1, 2, 2)], [(0, 0, 1, 1), (0, 1, 1, 2), (0, 2, 1, 3), (1, 0, 2, 1), (2, 0, 3, 1), (1, 2, 2, 3), (2, 1, 3, 2), (2, 2, 3, 3)], [[0] * 8]),
                              ([(1, 1, 3, 3), (0, 0, 4, 4)], [(0, 0, 2, 2), (2, 0, 4, 2), (0, 2, 2, 4), (2, 2, 4, 4)], [[1 / 7] * 4, [4 / 16] * 4])):
        b1 = np.tile(np.reshape(np.array(bbox1, np.float32), [-1, 1, 4]), [1, len(bbox2), 1])
        b2 = np.tile(np.reshape(np.array(bbox2, np.float32), [1, -1, 4]), [len(bbox1), 1, 1])
        pair = iou.batch_iou_pair(*(torch.from_numpy(np.ascontiguousarray(a)).to(DEV) for a in (b1[..., :2], b1[..., 2:], b2[..., :2], b2[..., 2:])))
        np.testing.assert_almost_equal(pair.cpu().numpy(), np.array(ans, np.float32))


@pytest.mark.parametrize('tag', list('abcdef'))
def test_nms_golden_exact(golden_dir, tag):
    import utils.postprocess
    g = np.load(os.path.join(golden_dir, 'nms.npz'))
    t = lambda k: torch.from_numpy(g[k + '_' + tag]).to(DEV)
    keep = utils.postprocess.nms(t('score'), t('yx_min'), t('yx_max'), float(g['overlap_' + tag]))
    assert keep == g['keep_' + tag].tolist()


def test_nms_empty_and_stress():
    import utils.postprocess
    assert utils.postprocess.nms(torch.zeros(0, device=DEV), torch.zeros(0, 2, device=DEV), torch.zeros(0, 2, device=DEV)) == []
    for n, seed, limit in ((4096, 21, 200), (845, 22, 200), (5000, 23, 500), (33, 24, 5)):
        score, a, b = O.synth_boxes(n, seed)
        ref = O.nms(score.numpy(), a.numpy(), b.numpy(), 0.45, limit)
        got = utils.postprocess.nms(score.to(DEV), a.to(DEV), b.to(DEV), 0.45, limit)
        assert got == ref, (n, seed)


@pytest.mark.parametrize('fix', [1, 0])
def test_filter_and_postprocess_golden(golden_dir, fix):
    import detect
    g = np.load(os.path.join(golden_dir, 'postprocess.npz'))
    d = np.load(os.path.join(golden_dir, 'decode.npz'))
    cfg = make_config(fix)
    for img in (0, 1):
        iou = torch.from_numpy(d['iou'][img]).reshape(-1).to(DEV)
        yx_min = torch.from_numpy(d['yx_min'][img]).reshape(-1, 2).to(DEV)
        yx_max = torch.from_numpy(d['yx_max'][img]).reshape(-1, 2).to(DEV)
        prob = torch.from_numpy(d['prob'][img]).reshape(-1, 20).to(DEV)
        tag = 'fix%d_img%d_' % (fix, img)
        fv = detect.filter_visible(cfg, iou, yx_min, yx_max, prob)
        for name, t in zip(('iou', 'yx_min', 'yx_max', 'prob', 'prob_cls', 'cls'), fv):
            assert np.array_equal(t.cpu().numpy(), g[tag + 'fv_' + name]), (name, img)
        res = detect.postprocess(cfg, iou, yx_min, yx_max, prob)
        assert (res is None) == bool(g[tag + 'none'])
        if res is not None:
            for name, t in zip(('iou', 'yx_min', 'yx_max', 'cls', 'score'), res):
                ref = g[tag + name]
                assert t.shape == ref.shape, name
                if name == 'cls':
                    assert np.array_equal(t.cpu().numpy(), ref)
                else:
                    np.testing.assert_allclose(t.cpu().numpy(), ref, rtol=1e-6, atol=0, err_msg=name)


def test_postprocess_none():
    import detect
    cfg = make_config(0)
    res = detect.postprocess(cfg, torch.full((845,), 0.1, device=DEV), torch.zeros(845, 2, device=DEV), torch.ones(845, 2, device=DEV),
                             torch.full((845, 20), 0.05, device=DEV))
    assert res is None


# ------------------------------------------------------------------------------------------------
# decode + softmax
# ------------------------------------------------------------------------------------------------
def test_decode_golden(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, 'decode.npz'))
    out = ops.decode(torch.from_numpy(g['feature']).to(DEV), torch.from_numpy(g['anchors']).to(DEV), 20)
    for k in ('iou', 'center_offset', 'size_norm', 'yx_min', 'yx_max', 'logits', 'prob'):
        assert rel_err(out[k], torch.from_numpy(g[k])) <= 1e-5, k
    assert torch.equal(out['logits'].cpu(), torch.from_numpy(g['logits']))          # pure move
    assert torch.equal(out['size_norm'].cpu(), torch.from_numpy(g['size_norm']))    # pure move


@pytest.mark.parametrize('shape', [(32, 13), (3, 19), (5, 10)])
def test_decode_vs_oracle(ops, shape):
    b, s = shape
    feat = torch.randn(b, 125, s, s, generator=torch.Generator().manual_seed(b * 100 + s)) * 2
    anchors = O.anchors_yolo_voc()
    ref = O.decode(feat, anchors)
    ref['prob'] = O.class_prob(ref)
    out = ops.decode(feat.to(DEV), anchors.to(DEV), 20)
    for k in ('iou', 'center_offset', 'size_norm', 'yx_min', 'yx_max', 'logits', 'prob'):
        assert rel_err(out[k], ref[k]) <= 1e-5, k


# ------------------------------------------------------------------------------------------------
# convolutions (tcgen05 implicit GEMM) vs the oracle arithmetic
# ------------------------------------------------------------------------------------------------
CONV_CASES = [
    # b, h, w, cin, cout, k, flags-name
    (2, 16, 16, 256, 128, 1, 'tiled'),
    (2, 16, 16, 256, 128, 1, ''),
    (2, 16, 16, 64, 128, 3, ''),
    (3, 13, 13, 128, 256, 3, ''),
    (2, 26, 26, 128, 64, 3, ''),
    (2, 13, 13, 256, 512, 3, 'wide'),
    (1, 32, 32, 32, 64, 3, ''),
    (3, 21, 19, 32, 64, 3, ''),            # small-K kernel, ragged last tile, tiles crossing image boundaries
    (2, 20, 20, 32, 48, 3, ''),            # small-K kernel, Cout < 64 (TMA store clips the channel box)
    (3, 21, 19, 32, 64, 3, 'plainstore'),  # small-K kernel with per-thread stores
    (3, 21, 19, 32, 64, 3, 'generic'),     # same shape through the generic kernel
    (3, 21, 19, 32, 64, 3, 'im2col'),      # small-K im2col kernel (the default for this shape is the halo-tile kernel)
    (2, 20, 20, 32, 48, 3, 'im2col'),
    (2, 48, 40, 32, 64, 3, ''),            # halo-tile kernel, several full tiles per image
    (8, 52, 52, 128, 256, 3, ''),
    (32, 13, 13, 512, 1024, 3, ''),
    (2, 13, 13, 1280, 1024, 3, ''),
]


@pytest.mark.parametrize('case', CONV_CASES)
def test_conv_unit_vs_oracle(ops, case):
    b, h, w, cin, cout, k, fl = case
    gen = torch.Generator().manual_seed(cin * 7 + cout)
    x = torch.randn(b, cin, h, w, generator=gen)
    wt = torch.randn(cout, cin, k, k, generator=gen) * (2.0 / (cin * k * k)) ** 0.5
    sd = {'u.conv.weight': wt, 'u.bn.weight': torch.rand(cout, generator=gen) + 0.5, 'u.bn.bias': torch.randn(cout, generator=gen) * 0.1,
          'u.bn.running_mean': torch.randn(cout, generator=gen) * 0.1, 'u.bn.running_var': torch.rand(cout, generator=gen) + 0.5}
    ref = O.conv_unit(x, sd, 'u', k, True, True)                      # fp32 oracle on fp32 operands
    scale, shift = ops.bn_fold(*(sd['u.bn.' + n].to(DEV) for n in ('weight', 'bias', 'running_mean', 'running_var')))
    w16 = ops.pack_weight_f16(wt.to(DEV))
    flags = {'tiled': ops.CONV_A_TILED, 'wide': ops.CONV_WIDE_N, '': 0, 'plainstore': ops.CONV_C32_IM2COL | ops.CONV_PLAIN_STORE,
             'generic': ops.CONV_NO_SMALLK, 'im2col': ops.CONV_C32_IM2COL}[fl]
    y = ops.conv_bn_act(x.to(DEV).permute(0, 2, 3, 1).contiguous().half(), w16, scale, shift, 0.1, flags=flags)
    err = rel_err(y.permute(0, 3, 1, 2), ref)
    assert err <= 1e-3, 'rel err %.3e' % err


SK_CASES = [
    # b, h, w, cin, cout, k, bn, mt   (bn 0 = library's choice); every case is forced onto the stream-K path
    (32, 13, 13, 512, 1024, 3, 0, 0),      # layers2.x: 88 tiles over 148 CTAs, every tile cut once or twice
    (2, 13, 13, 1024, 1024, 3, 256, 2),    # 8 tiles x 144 K-blocks: each tile is summed from ~19 CTAs
    (8, 26, 26, 256, 512, 3, 256, 1),      # two accumulator stages: dump / collect overlap the next segment
    (4, 52, 52, 128, 256, 3, 128, 2),
    (3, 13, 13, 1024, 512, 1, 0, 0),       # 1x1
    (5, 19, 17, 64, 72, 3, 64, 1),         # ragged rows, Cout not a multiple of the tile, BK = 64 taps
    (5, 19, 17, 96, 136, 3, 128, 1),       # BK = 32 path, two column tiles with a ragged second one
]


@pytest.mark.parametrize('case', SK_CASES)
def test_conv_streamk_vs_oracle(ops, case):
    b, h, w, cin, cout, k, bn, mt = case
    gen = torch.Generator().manual_seed(cin * 3 + cout + h)
    x = torch.randn(b, cin, h, w, generator=gen)
    wt = torch.randn(cout, cin, k, k, generator=gen) * (2.0 / (cin * k * k)) ** 0.5
    sd = {'u.conv.weight': wt, 'u.bn.weight': torch.rand(cout, generator=gen) + 0.5, 'u.bn.bias': torch.randn(cout, generator=gen) * 0.1,
          'u.bn.running_mean': torch.randn(cout, generator=gen) * 0.1, 'u.bn.running_var': torch.rand(cout, generator=gen) + 0.5}
    ref = O.conv_unit(x, sd, 'u', k, True, True)
    scale, shift = ops.bn_fold(*(sd['u.bn.' + n].to(DEV) for n in ('weight', 'bias', 'running_mean', 'running_var')))
    w16 = ops.pack_weight_f16(wt.to(DEV))
    ws = ops.conv_workspace(DEV)
    flags = ops.CONV_FORCE_STREAMK | (ops.conv_force_bn(bn) | ops.conv_force_mt(mt) | ops.conv_force_pair(1) if bn else 0)
    x16 = x.to(DEV).permute(0, 2, 3, 1).contiguous().half()
    for rep in range(3):      # repeated launches reuse the workspace: the flags must come back to zero every time
        y = ops.conv_bn_act(x16, w16, scale, shift, 0.1, flags=flags, workspace=ws)
        err = rel_err(y.permute(0, 3, 1, 2), ref)
        assert err <= 1e-3, 'launch %d: rel err %.3e' % (rep, err)
    assert int(ws[:4096].view(torch.int32).abs().sum().item()) == 0, 'stream-K flags not reset'
    # and the split changes nothing beyond fp32 summation order
    y0 = ops.conv_bn_act(x16, w16, scale, shift, 0.1, flags=ops.CONV_NO_STREAMK)
    assert rel_err(y, y0) <= 2e-3


def test_conv_streamk_head_fp32_nchw(ops):
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(32, 1024, 13, 13, generator=gen)
    wt = torch.randn(125, 1024, 1, 1, generator=gen) * (1.0 / 1024) ** 0.5
    bias = torch.randn(125, generator=gen) * 0.1
    ref = torch.nn.functional.conv2d(x, wt, bias)
    w16 = ops.pack_weight_f16(wt.to(DEV))
    ws = ops.conv_workspace(DEV)
    y = ops.conv_bn_act(x.to(DEV).permute(0, 2, 3, 1).contiguous().half(), w16, torch.ones(125, device=DEV), bias.to(DEV), 1.0,
                        out_mode=ops.OUT_F32_NCHW, flags=ops.CONV_FORCE_STREAMK, workspace=ws)
    assert rel_err(y, ref) <= 1e-3


@pytest.mark.parametrize('shape', [(2, 32, 24), (3, 22, 18), (1, 208, 208)])
def test_conv_c32_fused_maxpool(ops, shape):
    """layers1.2 + the MaxPool2d after it in one launch == the two separate kernels, bit for bit (max of fp16 values)."""
    b, h, w = shape
    gen = torch.Generator().manual_seed(h)
    x16 = torch.randn(b, h, w, 32, generator=gen).half().to(DEV)
    wt = torch.randn(64, 32, 3, 3, generator=gen) * (2.0 / 288) ** 0.5
    w16 = ops.pack_weight_f16(wt.to(DEV))
    scale, shift = (torch.rand(64, generator=gen) + 0.5).to(DEV), (torch.randn(64, generator=gen) * 0.1).to(DEV)
    full = ops.conv_bn_act(x16, w16, scale, shift, 0.1)
    two_step = ops.maxpool2x2(full)
    fused = ops.conv_bn_act(x16, w16, scale, shift, 0.1, flags=ops.CONV_POOL2X2)
    assert fused.shape == two_step.shape
    assert torch.equal(fused, two_step)
    ref = torch.nn.functional.max_pool2d(torch.nn.functional.leaky_relu(
        torch.nn.functional.conv2d(x16.float().permute(0, 3, 1, 2).cpu(), wt.half().float(), padding=1) * scale.cpu()[None, :, None, None]
        + shift.cpu()[None, :, None, None], 0.1), 2)
    assert rel_err(fused.permute(0, 3, 1, 2), ref) <= 1e-3


def test_conv_head_fp32_nchw_and_slice(ops):
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 1024, 13, 13, generator=gen)
    wt = torch.randn(125, 1024, 1, 1, generator=gen) * (1.0 / 1024) ** 0.5
    bias = torch.randn(125, generator=gen) * 0.1
    ref = torch.nn.functional.conv2d(x, wt, bias)
    y = ops.conv_bn_act(x.to(DEV).permute(0, 2, 3, 1).contiguous().half(), ops.pack_weight_f16(wt.to(DEV)),
                        torch.ones(125, device=DEV), bias.to(DEV), 1.0, out_mode=ops.OUT_F32_NCHW)
    assert rel_err(y, ref) <= 1e-3
    # channel-slice output into a wider buffer (in-place concat)
    wt2 = torch.randn(64, 1024, 1, 1, generator=gen) * (1.0 / 1024) ** 0.5
    buf = torch.full((2, 13, 13, 320), 3.0, dtype=torch.float16, device=DEV)
    ops.conv_bn_act(x.to(DEV).permute(0, 2, 3, 1).contiguous().half(), ops.pack_weight_f16(wt2.to(DEV)), torch.ones(64, device=DEV),
                    torch.zeros(64, device=DEV), 0.1, out=buf, y_ch_off=128)
    ref2 = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x, wt2), 0.1)
    assert rel_err(buf[..., 128:192].permute(0, 3, 1, 2), ref2) <= 1e-3
    assert bool((buf[..., :128] == 3).all()) and bool((buf[..., 192:] == 3).all())


def test_conv0_vs_oracle(ops):
    sd = O.make_state_dict(0)
    x = O.synth_images(2, 64, 96, seed=3)
    ref = torch.nn.functional.max_pool2d(O.conv_unit(x, sd, 'layers1.0', 3, True, True), 2)
    scale, shift = ops.bn_fold(*(sd['layers1.0.bn.' + n].to(DEV) for n in ('weight', 'bias', 'running_mean', 'running_var')))
    y = ops.conv0_bn_leaky_pool(x.to(DEV), sd['layers1.0.conv.weight'].to(DEV), scale, shift, 0.1)
    assert rel_err(y.permute(0, 3, 1, 2), ref) <= 1e-3


def test_conv0_u8_frames_vs_oracle(ops):
    """Raw uint8 NHWC frames: the kernel's 1/255 scaling == torchvision ToTensor (reference detect.py:144-145)."""
    sd = O.make_state_dict(0)
    frames = torch.randint(0, 256, (2, 64, 96, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(9))
    x = frames.permute(0, 3, 1, 2).float() / 255
    ref = torch.nn.functional.max_pool2d(O.conv_unit(x, sd, 'layers1.0', 3, True, True), 2)
    scale, shift = ops.bn_fold(*(sd['layers1.0.bn.' + n].to(DEV) for n in ('weight', 'bias', 'running_mean', 'running_var')))
    y = ops.conv0_u8_bn_leaky_pool(frames.to(DEV), sd['layers1.0.conv.weight'].to(DEV), scale, shift, 0.1)
    assert rel_err(y.permute(0, 3, 1, 2), ref) <= 1e-3


@pytest.mark.parametrize('cfg', [(128, 2), (256, 2), (256, 1), (64, 2)])
def test_conv_tile_shapes_agree(ops, cfg):
    """Every CTA tile shape (BLOCK_N x M-subtiles) computes the same result as the oracle arithmetic."""
    bn, mt = cfg
    gen = torch.Generator().manual_seed(bn + mt)
    cout = 64 if bn == 64 else 512
    x = torch.randn(5, 256, 13, 13, generator=gen)
    wt = torch.randn(cout, 256, 3, 3, generator=gen) * (2.0 / (256 * 9)) ** 0.5
    ref = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x, wt, padding=1), 0.1)
    y = ops.conv_bn_act(x.to(DEV).permute(0, 2, 3, 1).contiguous().half(), ops.pack_weight_f16(wt.to(DEV)), torch.ones(cout, device=DEV),
                        torch.zeros(cout, device=DEV), 0.1, flags=ops.conv_force_bn(bn) | ops.conv_force_mt(mt))
    assert rel_err(y.permute(0, 3, 1, 2), ref) <= 1e-3


SPLIT_CASES = [
    # b, h, w, cin, cout, k, split_a, split_w, src_lo
    (2, 16, 16, 64, 128, 3, True, True, True),
    (3, 13, 13, 128, 256, 3, True, False, True),
    (3, 13, 13, 128, 256, 3, False, True, False),
    (2, 26, 26, 256, 128, 1, True, True, True),
    (2, 26, 26, 256, 128, 1, False, True, True),     # input buffer holds [hi | lo] but the unit reads hi only
    (32, 13, 13, 512, 1024, 3, True, True, True),
    (3, 21, 19, 32, 64, 3, False, True, False),      # Cin = 32 leaves the halo-tile kernel when an operand is split
    (5, 19, 17, 96, 136, 3, True, True, True),       # BK = 32 path, ragged column tile
]


@pytest.mark.parametrize('case', SPLIT_CASES)
def test_conv_split_precision_vs_oracle(ops, case):
    """yb_conv_bn_act_split_fwd: with an operand split into fp16 hi + lo the unit reproduces the fp32 oracle on that operand to
    ~1e-6; an operand that stays fp16 is compared with the oracle run on the fp16-rounded operand.  Also checks the [hi | lo]
    output: hi is the fp16 rounding of the result, hi + lo carries it to ~2^-21."""
    b, h, w, cin, cout, k, split_a, split_w, src_lo = case
    gen = torch.Generator().manual_seed(cin * 5 + cout + k)
    x = torch.randn(b, cin, h, w, generator=gen)
    wt = torch.randn(cout, cin, k, k, generator=gen) * (2.0 / (cin * k * k)) ** 0.5
    sd = {'u.bn.weight': torch.rand(cout, generator=gen) + 0.5, 'u.bn.bias': torch.randn(cout, generator=gen) * 0.1,
          'u.bn.running_mean': torch.randn(cout, generator=gen) * 0.1, 'u.bn.running_var': torch.rand(cout, generator=gen) + 0.5}
    sd['u.conv.weight'] = wt if split_w else wt.half().float()
    ref = O.conv_unit(x if split_a else x.half().float(), sd, 'u', k, True, True)
    scale, shift = ops.bn_fold(*(sd['u.bn.' + n].to(DEV) for n in ('weight', 'bias', 'running_mean', 'running_var')))
    xn = x.permute(0, 2, 3, 1).contiguous()
    hi = xn.half()
    src = torch.cat([hi, (xn - hi.float()).half()], -1).contiguous().to(DEV) if src_lo else hi.to(DEV)
    w16 = ops.pack_weight_split_f16(wt.to(DEV), split_a, split_w)
    assert w16.shape == (cout, k, k, cin * (1 + int(split_a) + int(split_w)))
    out = torch.full((b, h, w, 2 * cout + 8), 7.0, dtype=torch.float16, device=DEV)
    ops.conv_bn_act_split(src, w16, scale, shift, 0.1, out, a_channels=cin * (2 if split_a else 1), y_ch_off=0, lo_ch_off=cout)
    got_hi, got_lo = out[..., :cout].float().cpu(), out[..., cout:2 * cout].float().cpu()
    assert bool((out[..., 2 * cout:] == 7).all()), 'wrote outside its channel slices'
    full = (got_hi + got_lo).permute(0, 3, 1, 2)
    err = rel_err(full, ref)
    assert err <= 2e-5, 'hi + lo rel err %.3e' % err
    # hi is the fp16 rounding of the result: the residual is at most half an ulp of hi (2^-11 relative; 2^-25 absolute below the normal range)
    assert bool((got_lo.abs() <= got_hi.abs() * 2.0 ** -11 * 1.001 + 2.0 ** -25).all()), 'lo exceeds half an ulp of hi'
    # fp32 NCHW output of the same operands (the head's mode)
    y32 = torch.empty(b, cout, h, w, dtype=torch.float32, device=DEV)
    ops.conv_bn_act_split(src, w16, scale, shift, 0.1, y32, a_channels=cin * (2 if split_a else 1), out_mode=ops.OUT_F32_NCHW)
    assert rel_err(y32, ref) <= 2e-5


def test_maxpool_split_exact(ops):
    gen = torch.Generator().manual_seed(77)
    v = torch.randn(3, 12, 10, 64, generator=gen)
    v[0, :2, :2, :8] = 1.0                               # ties: the first element of the window wins
    hi = v.half()
    lo = (v - hi.float()).half()
    x = torch.cat([hi, lo], -1).contiguous().to(DEV)
    out = torch.empty(3, 6, 5, 128, dtype=torch.float16, device=DEV)
    ops.maxpool2x2_split(x, 64, out)
    val = (hi.float() + lo.float()).permute(0, 3, 1, 2)
    ref = torch.nn.functional.max_pool2d(val, 2)
    got = (out[..., :64].float() + out[..., 64:].float()).permute(0, 3, 1, 2).cpu()
    assert torch.equal(got, ref)


# ------------------------------------------------------------------------------------------------
# whole backbone through the plugin surface
# ------------------------------------------------------------------------------------------------
def _build_darknet(precision):
    import model
    import model.yolo2
    cfg = make_config(1)
    cfg.read_dict({'b200': {'precision': precision}})
    dnn = model.yolo2.Darknet(model.ConfigChannels(cfg), O.anchors_yolo_voc(), 20)
    dnn.load_state_dict(O.make_state_dict(0), strict=False)
    dnn = dnn.to(DEV).eval()
    assert dnn.engine.precision == precision
    return dnn


@pytest.fixture(scope='module')
def darknet_strict():
    """precision='strict' ([b200] precision in the INI): split-precision operands, the mode the 1e-3 contract is asserted in."""
    return _build_darknet('strict')


@pytest.fixture(scope='module')
def darknet():
    import model
    import model.yolo2
    cfg = make_config(1)
    dnn = model.yolo2.Darknet(model.ConfigChannels(cfg), O.anchors_yolo_voc(), 20)
    dnn.load_state_dict(O.make_state_dict(0), strict=False)
    return dnn.to(DEV).eval()


@pytest.mark.parametrize('precision', ['strict', 'fast'])
def test_darknet_64_every_layer_golden(darknet, darknet_strict, golden_dir, precision):
    """Each unit on the GPU vs the REFERENCE's own activation for the same image (golden), fed by
    the GPU's own previous layer (so this is the end-to-end drift, layer by layer)."""
    dnn, tol = (darknet_strict, TOL_CONTRACT) if precision == 'strict' else (darknet, TOL_FAST_E2E)
    g = np.load(os.path.join(golden_dir, 'darknet_64.npz'))
    x = O.synth_images(1, 64, 64, seed=10).to(DEV)
    collect = {}
    feature = dnn.engine.forward(x, collect=collect).clone()
    worst = 0.0
    for key, act in collect.items():
        if 'act_' + key not in g:
            continue
        e = rel_err(act.permute(0, 3, 1, 2), torch.from_numpy(g['act_' + key]))
        worst = max(worst, e)
        assert e <= tol, '%s rel err %.3e' % (key, e)
    e = rel_err(feature, torch.from_numpy(g['feature']))
    record('darknet64_%s' % precision, dict(worst_layer=worst, feature=e))
    assert e <= tol


@pytest.mark.parametrize('precision', ['strict', 'fast'])
def test_darknet_416_feature_golden(darknet, darknet_strict, golden_dir, precision):
    dnn, tol = (darknet_strict, TOL_CONTRACT) if precision == 'strict' else (darknet, TOL_FAST_E2E)
    g = np.load(os.path.join(golden_dir, 'darknet_416.npz'))
    x = O.synth_images(1, 416, 416, seed=0).to(DEV)
    f = dnn(x)
    e = rel_err(f, torch.from_numpy(g['feature']))
    record('darknet416_%s' % precision, e)
    assert f.shape == (1, 125, 13, 13)
    assert e <= tol


def test_strict_pipeline_with_and_without_collect_agree(darknet_strict):
    """The strict forward takes a different route when a test inspects every layer (no fused pool on layers1.2): same feature."""
    x = O.synth_images(2, 96, 96, seed=21).to(DEV)
    f0 = darknet_strict.engine.forward(x).clone()
    f1 = darknet_strict.engine.forward(x, collect={}).clone()
    assert torch.equal(f0, f1)


def test_darknet_per_layer_on_oracle_inputs(darknet, ops):
    """every unit on the ORACLE's own input to that layer, rel <= 1e-3."""
    sd = O.make_state_dict(0)
    x = O.synth_images(2, 128, 128, seed=4)
    collect = {}
    O.darknet_forward(sd, x, collect=collect)
    layers = O.darknet19_layers()
    eng = darknet.engine
    eng.refresh()
    units = dict(zip([l['key'] for l in layers], eng.units1 + eng.units2 + [eng.unit_pt] + eng.units3))
    prev = x
    worst = 0.0
    for l in layers:
        key = l['key']
        if key == 'layers1.0':
            prev = torch.nn.functional.max_pool2d(collect[key], 2)
            continue
        if key == 'passthrough':
            inp = collect['layers1.16']
        elif key == 'layers2.1':
            inp = torch.nn.functional.max_pool2d(collect['layers1.16'], 2)
        elif key == 'layers3.0':
            inp = torch.cat([O.reorg(collect['passthrough']), collect['layers2.7']], 1)
        else:
            inp = prev
        u = units[key]
        mode = ops.OUT_F32_NCHW if key == 'layers3.1' else ops.OUT_F16_NHWC
        y = ops.conv_bn_act(inp.to(DEV).permute(0, 2, 3, 1).contiguous().half(), u.w16, u.scale, u.shift, u.slope, out_mode=mode)
        got = y if mode == ops.OUT_F32_NCHW else y.permute(0, 3, 1, 2)
        e = rel_err(got, collect[key])
        worst = max(worst, e)
        assert e <= 1e-3, '%s rel err %.3e' % (key, e)
        prev = torch.nn.functional.max_pool2d(collect[key], 2) if l['pool_after'] else collect[key]
    print('per-layer worst rel %.3e' % worst)


def test_inference_and_postprocess_batch(darknet):
    """C2 pipeline through the plugin surface: Inference -> postprocess_batch.  Decode + NMS are
    checked exactly against the oracle run on the GPU's own feature map (identical inputs)."""
    import detect
    import model
    cfg = make_config(1)
    anchors = O.anchors_yolo_voc()
    inference = model.Inference(cfg, darknet, anchors).eval()
    x = O.synth_images(4, 416, 416, seed=7).to(DEV)
    pred = model._inference(inference, x)
    assert pred['feature'].shape == (4, 125, 13, 13) and pred['iou'].shape == (4, 169, 5) and pred['logits'].shape == (4, 169, 5, 20)
    ref = O.decode(pred['feature'].cpu(), anchors)
    for k in ('iou', 'center_offset', 'yx_min', 'yx_max'):
        assert rel_err(pred[k], ref[k]) <= 1e-5, k
    results = detect.postprocess_batch(cfg, pred)
    for bi, res in enumerate(results):
        iou, a, b, p = (pred[k][bi].reshape(-1, *pred[k].shape[3:]).cpu() if pred[k].dim() > 3 else pred[k][bi].reshape(-1).cpu()
                        for k in ('iou', 'yx_min', 'yx_max', 'prob'))
        exp = O.postprocess(iou, a, b, p, True, 0.3, 0.005, 0.45)
        assert (res is None) == (exp is None)
        if res is not None:
            for name, t, r in zip(('iou', 'yx_min', 'yx_max', 'cls', 'score'), res, exp):
                assert t.shape == r.shape, (bi, name)
                if name == 'cls':
                    assert torch.equal(t.cpu(), r)
                else:
                    np.testing.assert_allclose(t.cpu().numpy(), r.numpy(), rtol=1e-6, atol=0)


@pytest.mark.parametrize('case', [(4, 13, 16, 0), (7, 19, 6, 1), (2, 10, 1, 2)])
def test_region_loss_values_masks_and_gradient(case):
    """K8/K9 vs the oracle's restated model.loss (itself pinned against the executed reference, see oracle header):
    the five scalars, positive/negative masks, matched IoU and d(total)/dfeature via autograd."""
    import model
    b, s, g, seed = case
    anchors = O.anchors_yolo_voc()
    gen = torch.Generator().manual_seed(seed)
    feat = (torch.randn(b, 125, s, s, generator=gen) * 0.7)
    data = O.norm_data(O.synth_targets(b, s * 32, s * 32, slots=g, seed=seed + 5), s * 32, s * 32, s, s)
    # oracle
    f_ref = feat.clone().requires_grad_(True)
    l_ref, dbg_ref = O.loss(anchors, data, O.decode(f_ref, anchors), 0.6)
    O.loss_total(l_ref).backward()
    # CUDA
    f_gpu = feat.to(DEV).requires_grad_(True)
    l_gpu, dbg = model.loss(anchors, {k: v.to(DEV) for k, v in data.items()}, dict(feature=f_gpu), 0.6)
    total = sum(l_gpu[k] * O.HPARAM_DEFAULT[k] for k in l_gpu)
    total.backward()
    assert torch.equal(dbg['positive'].cpu().bool(), dbg_ref['positive'])
    assert torch.equal(dbg['negative'].cpu().bool(), dbg_ref['negative'])
    np.testing.assert_allclose(dbg['iou'].cpu().numpy(), dbg_ref['iou'].numpy(), rtol=1e-5, atol=1e-6)
    for k in ('foreground', 'background', 'center', 'size', 'cls'):
        assert abs(l_gpu[k].item() - l_ref[k].item()) <= 1e-4 * abs(l_ref[k].item()) + 1e-9, (k, l_gpu[k].item(), l_ref[k].item())
    assert rel_err(f_gpu.grad, f_ref.grad) <= 1e-4


@pytest.mark.parametrize('tag', list('abcd'))
def test_region_loss_vs_executed_reference_golden(golden_dir, tag):
    """The loss kernels directly against what the reference's own model.loss returned (tests/golden/make_golden_loss.py): 13x13
    G=16 (a), 19x19 (b), one ground-truth slot (c) and the one-hot class branch train/cross_entropy = 0 (d, model/__init__.py:156-160)."""
    import model
    g = np.load(os.path.join(golden_dir, 'loss.npz'))
    b, s, slots, seed, one_hot = (int(v) for v in g[tag + '_dims'])
    anchors = O.anchors_yolo_voc()
    data = O.norm_data(O.synth_targets(b, s * 32, s * 32, slots=slots, seed=20 + seed), s * 32, s * 32, s, s)
    f = torch.from_numpy(g[tag + '_feature']).to(DEV).requires_grad_(True)
    losses, dbg = model.loss(anchors, {k: v.to(DEV) for k, v in data.items()}, dict(feature=f), 0.6, cross_entropy=not one_hot)
    sum(losses[k] * O.HPARAM_DEFAULT[k] for k in losses).backward()
    for k, v in losses.items():
        ref = float(g[tag + '_loss_' + k])
        assert abs(v.item() - ref) <= 1e-4 * abs(ref) + 1e-9, (k, v.item(), ref)
    assert np.array_equal(dbg['positive'].cpu().numpy().astype(np.uint8), g[tag + '_positive'])
    assert np.array_equal(dbg['negative'].cpu().numpy().astype(np.uint8), g[tag + '_negative'])
    np.testing.assert_allclose(dbg['iou'].cpu().numpy(), g[tag + '_iou'], rtol=1e-5, atol=1e-6)
    assert rel_err(f.grad, torch.from_numpy(g[tag + '_grad'])) <= 1e-4


def test_region_loss_single_class_head():
    """A*5-channel head (model.output_channels with one category, model/__init__.py:46-50): no class targets, no class term."""
    import model
    anchors = O.anchors_yolo_voc()
    b, s = 3, 13
    feat = torch.randn(b, 25, s, s, generator=torch.Generator().manual_seed(4)) * 0.7
    data = O.norm_data(O.synth_targets(b, s * 32, s * 32, slots=8, seed=31), s * 32, s * 32, s, s)
    f_ref = feat.clone().requires_grad_(True)
    l_ref, dbg_ref = O.loss(anchors, data, O.decode(f_ref, anchors), 0.6)
    assert 'cls' not in l_ref
    O.loss_total(l_ref).backward()
    f_gpu = feat.to(DEV).requires_grad_(True)
    l_gpu, dbg = model.loss(anchors, {k: v.to(DEV) for k, v in data.items()}, dict(feature=f_gpu), 0.6)
    assert set(l_gpu) == set(l_ref)
    sum(l_gpu[k] * O.HPARAM_DEFAULT[k] for k in l_gpu).backward()
    assert torch.equal(dbg['positive'].cpu().bool(), dbg_ref['positive']) and torch.equal(dbg['negative'].cpu().bool(), dbg_ref['negative'])
    for k in l_ref:
        assert abs(l_gpu[k].item() - l_ref[k].item()) <= 1e-4 * abs(l_ref[k].item()) + 1e-9, k
    assert rel_err(f_gpu.grad, f_ref.grad) <= 1e-4


def test_cpu_input_fails_loudly(darknet):
    with pytest.raises(RuntimeError):
        darknet(torch.zeros(1, 3, 64, 64))


# ------------------------------------------------------------------------------------------------
# training path: weight gradient kernel, train-mode forward, full backward vs the oracle's autograd
# ------------------------------------------------------------------------------------------------
WGRAD_CASES = [  # b, h, cin, cout, k, dz_ld
    (2, 16, 64, 128, 3, 128), (3, 13, 256, 512, 3, 512), (1, 32, 32, 64, 3, 64), (2, 26, 512, 64, 1, 64),
    (2, 13, 1024, 125, 1, 128), (8, 52, 128, 256, 3, 256), (4, 13, 1280, 1024, 3, 1024),
]


@pytest.mark.parametrize('case', WGRAD_CASES)
def test_conv_wgrad_vs_torch(ops, case):
    b, h, cin, cout, k, dz_ld = case
    gen = torch.Generator().manual_seed(cin + cout + k)
    x = torch.randn(b, h, h, cin, generator=gen).half()
    dz = torch.zeros(b, h, h, dz_ld).half()
    dz[..., :cout] = (torch.randn(b, h, h, cout, generator=gen) * 0.1).half()
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, k, k), dz[..., :cout].float().permute(0, 3, 1, 2),
                                      padding=(k - 1) // 2)
    dw_krsc = torch.empty(cout, k, k, cin, dtype=torch.float32, device=DEV)
    ops.call('yb_conv_wgrad', x.to(DEV), dz.to(DEV), dw_krsc, b, h, h, cin, cout, k, cin, dz_ld)
    dw = torch.empty(cout, cin, k, k, dtype=torch.float32, device=DEV)
    ops.call('yb_unpack_wgrad', dw_krsc, dw, cout, cin, k, 1.0)
    err = rel_err(dw, ref)
    assert err <= 2e-3, 'rel err %.3e' % err


def _oracle_train_step(sd_in, x, data, anchors):
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.clone()) for k, v in sd_in.items()}
    stats = {}
    feature = O.darknet_forward(sd, x, train=True, stats=stats)
    losses, _ = O.loss(anchors, data, O.decode(feature, anchors), 0.6)
    total = O.loss_total(losses)
    total.backward()
    return feature.detach(), {k: v.item() for k, v in losses.items()}, {k: v.grad for k, v in sd.items() if v.requires_grad}, stats


TRAIN_UNIT_CASES = [  # b, h, cin, cout, k, pooled, direct+pooled grads (branch point)
    (4, 16, 64, 128, 3, False, False), (4, 16, 128, 64, 1, False, False), (2, 26, 256, 512, 3, True, False),
    (2, 26, 256, 512, 3, True, True), (8, 13, 512, 1024, 3, False, False), (2, 32, 32, 64, 3, True, False),
]


@pytest.mark.parametrize('case', TRAIN_UNIT_CASES)
def test_training_unit_forward_backward(ops, case):
    """One model.yolo2.Conv2d unit in TRAIN mode (conv -> batch-stat BN -> leaky [-> MaxPool2d(2)]) against torch
    autograd on the same fp16-representable inputs: activation, batch statistics, dgamma, dbeta, dW and dx.
    (End-to-end train-mode comparisons are dominated by the chaotic sensitivity of batch-stat BN to ANY fp16
    rounding -- see DESIGN.md -- so parity is asserted per unit, as SURVEY appendix A does for inference.)"""
    b, h, cin, cout, k, pooled, branch = case
    gen = torch.Generator().manual_seed(cin * 3 + cout + k + int(pooled))
    x = (torch.randn(b, cin, h, h, generator=gen) + 0.3).half().float()
    wt = (torch.randn(cout, cin, k, k, generator=gen) * (2.0 / (cin * k * k)) ** 0.5).half().float()
    gamma = torch.rand(cout, generator=gen) + 0.5
    beta = torch.randn(cout, generator=gen) * 0.1
    oh = h // 2 if pooled else h
    g_out = (torch.randn(b, cout, oh, oh, generator=gen) * 0.05).half().float()       # gradient w.r.t. the (pooled) output
    g_dir = (torch.randn(b, cout, h, h, generator=gen) * 0.05).half().float() if branch else None
    # ---- torch autograd reference (fp32) ----
    xr, wr, gr, br = (t.clone().requires_grad_(True) for t in (x, wt, gamma, beta))
    z = torch.nn.functional.conv2d(xr, wr, padding=(k - 1) // 2)
    # the CUDA path stores the raw conv output in fp16; mimic that storage rounding (straight-through in backward) so
    # that both sides take the leaky-ReLU slope decision on the same values -- otherwise ~1e-3 of the activations that
    # lie within fp16 rounding of 0 flip slope (1 vs 0.1) and dominate the gradient comparison
    z = z + (z.detach().half().float() - z.detach())
    y = torch.nn.functional.leaky_relu(torch.nn.functional.batch_norm(z, None, None, gr, br, True, 0.0, 1e-5), 0.1)
    out = torch.nn.functional.max_pool2d(y, 2) if pooled else y
    obj = (out * g_out).sum() + ((y * g_dir).sum() if branch else 0.0)
    obj.backward()
    # ---- CUDA path (same call sequence as b200.train_engine) ----
    xd = x.to(DEV).permute(0, 2, 3, 1).contiguous().half()
    w16 = ops.pack_weight_f16(wt.to(DEV))
    one, zero = torch.ones(cout, device=DEV), torch.zeros(cout, device=DEV)
    zd = ops.conv_bn_act(xd, w16, one, zero, 1.0)
    rows = b * h * h
    sums = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    mean, invstd = torch.empty(cout, device=DEV), torch.empty(cout, device=DEV)
    rm, rv = torch.zeros(cout, device=DEV), torch.ones(cout, device=DEV)
    ops.call('yb_bn_stats', zd, cout, rows, cout, sums)
    ops.call('yb_bn_finalize', sums, rows, cout, 1e-5, 0.01, rm, rv, mean, invstd)
    gd, bd = gamma.to(DEV), beta.to(DEV)
    a = torch.empty(b, oh, oh, cout, dtype=torch.float16, device=DEV)
    ops.call('yb_bn_act_apply', zd, cout, mean, invstd, gd, bd, 0.1, a, cout, 0, b, h, h, cout, int(pooled))
    assert rel_err(a.permute(0, 3, 1, 2), out) <= 3e-3
    zf = z.detach()
    assert rel_err(mean, zf.mean(dim=(0, 2, 3))) <= 1e-3 and rel_err(invstd, 1.0 / torch.sqrt(zf.var(dim=(0, 2, 3), unbiased=False) + 1e-5)) <= 1e-3
    np.testing.assert_allclose(rm.cpu().numpy(), 0.01 * zf.mean(dim=(0, 2, 3)).numpy(), rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(rv.cpu().numpy(), (0.99 + 0.01 * zf.var(dim=(0, 2, 3), unbiased=True)).numpy(), rtol=2e-3, atol=1e-6)
    go = g_out.to(DEV).permute(0, 2, 3, 1).contiguous().half()
    gdir = g_dir.to(DEV).permute(0, 2, 3, 1).contiguous().half() if branch else None
    da, dap = (gdir, go) if pooled else (go, None)
    window = 1 if pooled else 0
    args = (zd, cout, mean, invstd, gd, bd, 0.1, da, 0 if da is None else cout, 0, dap, 0 if dap is None else cout, 0, b, h, h, cout, window, sums)
    ops.call('yb_bn_act_bwd', 0, *args, None, 0, 1)
    dgamma, dbeta = torch.empty(cout, device=DEV), torch.empty(cout, device=DEV)
    ops.call('yb_bn_param_grad', sums, cout, dgamma, dbeta, 0, 1.0)
    dz = torch.empty(b, h, h, cout, dtype=torch.float16, device=DEV)
    ops.call('yb_bn_act_bwd', 1, *args, dz, cout, 1)
    assert rel_err(dgamma, gr.grad) <= 5e-3 and rel_err(dbeta, br.grad) <= 5e-3
    dw_krsc = torch.empty(cout, k, k, cin, dtype=torch.float32, device=DEV)
    ops.call('yb_conv_wgrad', xd, dz, dw_krsc, b, h, h, cin, cout, k, cin, cout)
    dw = torch.empty(cout, cin, k, k, dtype=torch.float32, device=DEV)
    ops.call('yb_unpack_wgrad', dw_krsc, dw, cout, cin, k, 1.0)
    # gradients downstream of the leaky kink: an activation within fp16 rounding of 0 flips its slope (1 vs 0.1) for that
    # single element, so dW / dx are compared in relative L2 (robust to isolated flips) with a loose max-norm bound
    assert rel_l2(dw, wr.grad) <= 5e-3 and rel_err(dw, wr.grad) <= 5e-2
    wd = torch.empty(cin, k, k, cout, dtype=torch.float16, device=DEV)
    ops.call('yb_pack_weight_dgrad_f16', wt.to(DEV), wd, cout, cin, k, cout)
    dx = ops.conv_bn_act(dz, wd, torch.ones(cin, device=DEV), torch.zeros(cin, device=DEV), 1.0)
    assert rel_l2(dx.permute(0, 3, 1, 2), xr.grad) <= 5e-3 and rel_err(dx.permute(0, 3, 1, 2), xr.grad) <= 1e-1


def test_conv0_training_pieces(ops):
    """layers1.0 in train mode: raw conv output from the fp32 image and its weight gradient."""
    gen = torch.Generator().manual_seed(77)
    x = torch.rand(2, 3, 64, 96, generator=gen)
    w = torch.randn(32, 3, 3, 3, generator=gen) * 0.3
    z = torch.empty(2, 64, 96, 32, dtype=torch.float16, device=DEV)
    ops.call('yb_conv0_raw_fwd', x.to(DEV), w.to(DEV), z, 2, 64, 96, 32)
    ref = torch.nn.functional.conv2d(x, w, padding=1)
    assert rel_err(z.permute(0, 3, 1, 2), ref) <= 1e-3
    dz = (torch.randn(2, 64, 96, 32, generator=gen) * 0.1).half()
    dw = torch.empty(32, 3, 3, 3, dtype=torch.float32, device=DEV)
    ops.call('yb_conv0_wgrad', x.to(DEV), dz.to(DEV), dw, 2, 64, 96)
    ref_dw = torch.nn.grad.conv2d_weight(x, (32, 3, 3, 3), dz.float().permute(0, 3, 1, 2), padding=1)
    assert rel_err(dw, ref_dw) <= 1e-3
    # raw output with the batch statistics fused in (yb_conv0_raw_stats_fwd): the same z bit for bit, sums == yb_bn_stats of it; the
    # second size gives every CTA several tiles (the statistics accumulate in registers across them) and accumulates on top of the first call
    for bsz, hh, ww in ((2, 64, 96), (24, 416, 416)):
        xs = torch.rand(bsz, 3, hh, ww, generator=gen).to(DEV)
        z0 = torch.empty(bsz, hh, ww, 32, dtype=torch.float16, device=DEV)
        z1 = torch.full((bsz, hh, ww, 32), float('nan'), dtype=torch.float16, device=DEV)
        ops.call('yb_conv0_raw_fwd', xs, w.to(DEV), z0, bsz, hh, ww, 32)
        sums = torch.zeros(64, dtype=torch.float64, device=DEV)
        ref_sums = torch.zeros(64, dtype=torch.float64, device=DEV)
        for _ in range(2):
            ops.call('yb_conv0_raw_stats_fwd', xs, w.to(DEV), z1, sums, bsz, hh, ww, 32)
            ops.call('yb_bn_stats', z0, 32, bsz * hh * ww, 32, ref_sums)
        assert torch.equal(z0, z1)
        zd = z0.double().reshape(-1, 32)
        exact = torch.cat([zd.sum(0), (zd * zd).sum(0)]) * 2
        assert ((sums - exact).abs() / exact.abs().clamp_min(1.0)).max().item() <= 1e-4          # fp32 partial sums per thread / CTA
        assert ((ref_sums - exact).abs() / exact.abs().clamp_min(1.0)).max().item() <= 1e-4


def test_conv0_wgrad_with_fused_bn_backward(ops):
    """yb_conv0_wgrad_bn (dz formed in shared memory from z and the pooled gradient) against the two-kernel path it replaces:
    yb_bn_act_bwd mode 1 -> dz in memory -> yb_conv0_wgrad.  Both round dz to fp16 before the tensor-core pass; only the order of the
    final atomics differs."""
    gen = torch.Generator().manual_seed(78)
    for b, h, w in ((2, 64, 96), (5, 160, 160)):
        x = torch.rand(b, 3, h, w, generator=gen).to(DEV)
        z = (torch.randn(b, h, w, 32, generator=gen) * 0.7).half().to(DEV)
        z[0, :2, :2, :8] = 0.25                                           # tied window: the first maximum takes the gradient
        dap_buf = (torch.randn(b, h // 2, w // 2, 48, generator=gen) * 0.05).half().to(DEV)      # gradient at channel offset 8 of a wider buffer
        mean = (torch.randn(32, generator=gen) * 0.1).to(DEV)
        invstd = (torch.rand(32, generator=gen) + 0.8).to(DEV)
        gamma = (torch.rand(32, generator=gen) + 0.5).to(DEV)
        gamma[3] = -0.7                                                   # a negative scale reverses the order inside the window
        beta = (torch.randn(32, generator=gen) * 0.2).to(DEV)
        sums = torch.zeros(64, dtype=torch.float64, device=DEV)
        args = (z, 32, mean, invstd, gamma, beta, 0.1, None, 0, 0, dap_buf, 48, 8, b, h, w, 32, 1, sums)
        ops.call('yb_bn_act_bwd', 0, *args, None, 0, 1)
        dz = torch.empty(b, h, w, 32, dtype=torch.float16, device=DEV)
        ops.call('yb_bn_act_bwd', 1, *args, dz, 32, 1)
        dw_ref = torch.empty(32, 3, 3, 3, dtype=torch.float32, device=DEV)
        ops.call('yb_conv0_wgrad', x, dz, dw_ref, b, h, w)
        dw = torch.full((32, 3, 3, 3), float('nan'), dtype=torch.float32, device=DEV)
        ops.call('yb_conv0_wgrad_bn', x, z, dap_buf, 48, 8, mean, invstd, gamma, beta, 0.1, sums, dw, b, h, w)
        assert rel_err(dw, dw_ref) <= 1e-5, (b, h, w, rel_err(dw, dw_ref))
        # and against fp32 autograd of the whole unit tail: a = maxpool(leaky(bn(z))), loss = sum(a * g)
        zr = z.float().permute(0, 3, 1, 2).cpu().requires_grad_(True)
        m, v = zr.mean(dim=(0, 2, 3)), zr.var(dim=(0, 2, 3), unbiased=False)
        # the kernels take (mean, invstd) as given: use the batch's own so that autograd's BN backward is the same function
        mean_b, invstd_b = m.detach().to(DEV), (1.0 / torch.sqrt(v.detach() + 1e-5)).to(DEV)
        a = torch.nn.functional.max_pool2d(torch.nn.functional.leaky_relu(
            torch.nn.functional.batch_norm(zr, None, None, gamma.cpu(), beta.cpu(), True, 0.0, 1e-5), 0.1), 2)
        gref = dap_buf[..., 8:40].float().permute(0, 3, 1, 2).cpu()
        (a * gref).sum().backward()
        dw_auto = torch.nn.grad.conv2d_weight(x.cpu(), (32, 3, 3, 3), zr.grad, padding=1)
        sums.zero_()
        args = (z, 32, mean_b, invstd_b, gamma, beta, 0.1, None, 0, 0, dap_buf, 48, 8, b, h, w, 32, 1, sums)
        ops.call('yb_bn_act_bwd', 0, *args, None, 0, 1)
        ops.call('yb_conv0_wgrad_bn', x, z, dap_buf, 48, 8, mean_b, invstd_b, gamma, beta, 0.1, sums, dw, b, h, w)
        assert rel_l2(dw.cpu(), dw_auto) <= 5e-3, rel_l2(dw.cpu(), dw_auto)


def test_training_step_vs_oracle():
    """C3-style step at a small size through the plugin surface: train-mode forward (batch-stat BN), region
    loss, full backward, gradients on every parameter.  End to end, batch-stat BN over 22 random layers amplifies
    ANY fp16 rounding chaotically (an fp32 oracle whose activations/weights are merely rounded to fp16 deviates by
    2.4e-2 from itself, tests/diag_train.py + DESIGN.md), so this is a wiring / sanity check with loose bounds; the
    numerical parity of every kernel is asserted per unit in test_training_unit_forward_backward."""
    import model
    import model.yolo2
    cfg = make_config(1)
    anchors = O.anchors_yolo_voc()
    sd0 = O.make_state_dict(0)
    b, size = 4, 128
    s = size // 32
    x = O.synth_images(b, size, size, seed=12)
    data = O.norm_data(O.synth_targets(b, size, size, slots=6, seed=13), size, size, s, s)
    f_ref, l_ref, g_ref, stats = _oracle_train_step(sd0, x, data, anchors)

    dnn = model.yolo2.Darknet(model.ConfigChannels(cfg), anchors, 20)
    dnn.load_state_dict(sd0, strict=False)
    dnn = dnn.to(DEV).train()
    inference = model.Inference(cfg, dnn, anchors).train()
    pred = model._inference(inference, x.to(DEV))
    losses, _ = model.loss(anchors, {k: v.to(DEV) for k, v in data.items()}, pred, 0.6)
    total = sum(losses[k] * O.HPARAM_DEFAULT[k] for k in losses)
    total.backward()
    e_f = rel_err(pred['feature'], f_ref)
    print('train forward feature rel %.3e' % e_f)
    assert e_f <= 1e-1
    for k in l_ref:
        assert abs(losses[k].item() - l_ref[k]) <= 5e-2 * abs(l_ref[k]) + 1e-7, (k, losses[k].item(), l_ref[k])
    worst = (1.0, None)
    for name, p in dnn.named_parameters():
        assert p.grad is not None, name
        g, r = p.grad.detach().float().cpu().flatten(), g_ref[name].flatten()
        cos = torch.dot(g, r) / (g.norm() * r.norm() + 1e-30)
        rel = ((g - r).norm() / (r.norm() + 1e-30)).item()
        if cos.item() < worst[0]:
            worst = (cos.item(), name, rel)
        assert cos.item() >= 0.85, '%s: cosine %.4f rel %.3e' % (name, cos.item(), rel)
    print('worst gradient cosine %.5f (%s, rel %.3e)' % worst)
    # running statistics after one step (momentum 0.01, unbiased variance)
    for key, (mean, var) in stats.items():
        n = b * f_ref.shape[-1] ** 2 if False else None
        rm = dict(dnn.named_buffers())[key + '.bn.running_mean'].cpu()
        exp = 0.99 * sd0[key + '.bn.running_mean'] + 0.01 * mean.detach()
        assert rel_err(rm, exp) <= 2e-2, key


def test_c3_batch64_training_step_vs_executed_reference(golden_dir):
    """BASELINE configs[2] at its real size: one 64 x 3 x 416 x 416 training step (train-mode forward with batch-statistics BatchNorm,
    region loss, full backward) against the SAME step executed with the reference's own modules on CPU
    (tests/golden/make_golden_c3.py).

    What can be asserted end to end: train-mode BatchNorm removes each channel's batch mean, so a perturbation made on the un-centred
    conv output grows by sqrt(1 + mu^2/sigma^2) ~ 1.25 PER LAYER relative to the centred signal -- with this untrained network (random BN
    parameters) fp16 storage alone moves an fp32 forward by 3.7e-2 at the head (tools/train_error_budget.py, pure CPU fp32 arithmetic
    with only the roundings added; profiles/r02_train_error_budget.txt); the GPU measures 4.7e-2.  The target assignment, the five loss
    terms (<= 1e-2; measured <= 5.1e-3), the running statistics (1e-3) and the positive / negative counts survive that; gradient norms
    agree to ~10 %.  Kernel-level parity of the same step is asserted where it is well posed: every unit on the reference's own input
    (test_training_per_layer_on_oracle_inputs, 2e-3) and per unit against autograd (test_training_unit_forward_backward)."""
    import model
    import model.yolo2
    g = np.load(os.path.join(golden_dir, 'c3_train64.npz'))
    cfg = make_config(1)
    anchors = O.anchors_yolo_voc()
    sd0 = O.make_state_dict(0)
    b, size = 64, 416
    s = size // 32
    x = O.synth_images(b, size, size, seed=64)
    data = O.norm_data(O.synth_targets(b, size, size, slots=16, seed=65), size, size, s, s)
    dnn = model.yolo2.Darknet(model.ConfigChannels(cfg), anchors, 20)
    dnn.load_state_dict(sd0, strict=False)
    dnn = dnn.to(DEV).train()
    inference = model.Inference(cfg, dnn, anchors).train()
    pred = model._inference(inference, x.to(DEV))
    losses, dbg = model.loss(anchors, {k: v.to(DEV) for k, v in data.items()}, pred, 0.6)
    sum(losses[k] * O.HPARAM_DEFAULT[k] for k in losses).backward()
    f = pred['feature'].detach().cpu()
    e_f = max(((f[bi] - torch.from_numpy(g['feature'][slot])).abs().max() / float(g['feature_absmax'][bi])).item() for slot, bi in enumerate((0, b - 1)))
    e_loss = {k: abs(losses[k].item() - float(g['loss_' + k])) / abs(float(g['loss_' + k])) for k in losses}
    npos, nneg = int(dbg['positive'].sum().item()), int(dbg['negative'].sum().item())
    worst_norm, worst_head = (0.0, None), (1.0, None)
    small = 0.0
    for name, p in dnn.named_parameters():
        assert p.grad is not None, name
        gr = p.grad.detach().float().cpu()
        rn = abs(gr.double().norm().item() / float(g['gnorm_' + name]) - 1.0)
        if rn > worst_norm[0]:
            worst_norm = (rn, name)
        h, hr = gr.flatten()[:16], torch.from_numpy(g['ghead_' + name])
        cos = (torch.dot(h, hr) / (h.norm() * hr.norm() + 1e-30)).item()
        if cos < worst_head[0]:
            worst_head = (cos, name)
        if 'gfull_' + name in g:
            small = max(small, rel_l2(gr, torch.from_numpy(g['gfull_' + name])))
    e_run = max(rel_err(buf, torch.from_numpy(g['buf_' + name])) for name, buf in dnn.named_buffers() if 'running' in name)
    record('c3_batch64_train', dict(feature=e_f, losses=e_loss, positives=(npos, int(g['positives'])), negatives=(nneg, int(g['negatives'])),
                                    worst_grad_norm=worst_norm, worst_grad_head_cosine=worst_head, small_grads_rel_l2=small, running_stats=e_run))
    assert npos == int(g['positives'])
    assert abs(nneg - int(g['negatives'])) <= 1e-3 * int(g['negatives'])
    assert e_f <= 1e-1, e_f
    for k, v in e_loss.items():
        assert v <= 1e-2, (k, v)
    assert worst_norm[0] <= 0.2, worst_norm
    assert worst_head[0] >= 0.75, worst_head
    assert e_run <= 1e-3, e_run


def test_training_per_layer_on_oracle_inputs(ops):
    """Train-mode analogue of test_darknet_per_layer_on_oracle_inputs at 416 x 416: every unit (conv -> batch statistics -> normalise
    + leaky [+ pool]) on the REFERENCE arithmetic's own input to that unit; batch mean 2e-3 of the channel spread (the statistics are those of the fp16-stored conv output), variance 1e-3, activation 2e-3.  This is the
    well-posed form of train-mode parity: end to end the same roundings compound by ~1.25x per layer (see the C3 test)."""
    import model
    import model.yolo2
    cfg = make_config(1)
    anchors = O.anchors_yolo_voc()
    sd = O.make_state_dict(0)
    x = O.synth_images(4, 416, 416, seed=9)
    collect, stats = {}, {}
    with torch.no_grad():
        O.darknet_forward(sd, x, collect=collect, train=True, stats=stats)
    dnn = model.yolo2.Darknet(model.ConfigChannels(cfg), anchors, 20)
    dnn.load_state_dict(sd, strict=False)
    dnn = dnn.to(DEV).train()
    eng, tr = dnn.engine, dnn.trainer
    eng.refresh(force=True)
    layers = O.darknet19_layers()
    units = dict(zip([l['key'] for l in layers], eng.units1 + eng.units2 + [eng.unit_pt] + eng.units3))
    prev, worst_a, worst_s = x, (0.0, None), (0.0, None)
    for l in layers:
        key = l['key']
        if key == 'layers3.1':
            break
        if key == 'layers1.0':
            prev = torch.nn.functional.max_pool2d(collect[key], 2)
            continue
        if key == 'passthrough':
            inp = collect['layers1.16']
        elif key == 'layers2.1':
            inp = torch.nn.functional.max_pool2d(collect['layers1.16'], 2)
        elif key == 'layers3.0':
            inp = torch.cat([O.reorg(collect['passthrough']), collect['layers2.7']], 1)
        else:
            inp = prev
        u = units[key]
        b, _, h, w = inp.shape
        z = tr._raw_conv(u, inp.to(DEV).permute(0, 2, 3, 1).contiguous().half(), key=key)
        mean, invstd = tr._bn_forward(key, u, z, b * h * w)
        a = tr._apply(u, z, mean, invstd, b, h, w, False)
        m_ref, v_ref = stats[key]
        e_m = ((mean.cpu() - m_ref).abs().max() / v_ref.sqrt().max()).item()          # mean error relative to the channel spread
        e_v = rel_err(1.0 / (invstd.cpu() ** 2) - 1e-5, v_ref)
        e_a = rel_err(a.permute(0, 3, 1, 2), collect[key])
        if max(e_m, e_v) > worst_s[0]:
            worst_s = (max(e_m, e_v), key)
        if e_a > worst_a[0]:
            worst_a = (e_a, key)
        assert e_m <= 2e-3 and e_v <= 1e-3, '%s batch statistics: mean %.3e var %.3e' % (key, e_m, e_v)
        assert e_a <= 2e-3, '%s activation rel err %.3e' % (key, e_a)
        prev = torch.nn.functional.max_pool2d(collect[key], 2) if l['pool_after'] else collect[key]
    record('train_per_layer_on_oracle_inputs_416', dict(worst_activation=worst_a, worst_statistic=worst_s))


def test_grad_guard_flags_and_clears_non_finite_gradients(ops):
    """yb_grad_guard: a clean buffer is left alone (flag 0); one inf / NaN anywhere (body or the non-multiple-of-4 tail) raises the flag
    and zeroes the buffer; and a training step whose loss weights overflow the fp16 gradient range leaves finite (zero) gradients."""
    import model
    import train as yb_train
    for n, bad_at, bad in ((4099, None, 0.0), (4099, 17, float('inf')), (4099, 4098, float('nan')), (64, 5, float('-inf'))):
        buf = torch.randn(n, device=DEV)
        keep = buf.clone()
        if bad_at is not None:
            buf[bad_at] = bad
        found = torch.full((), 5.0, device=DEV)
        ops.call('yb_grad_guard', buf, n, found, 1)
        if bad_at is None:
            assert found.item() == 0.0 and torch.equal(buf, keep)
        else:
            assert found.item() == 1.0 and bool((buf == 0).all())
    cfg = make_config(1)
    cfg.read_dict({'model': {'threshold': '0.6'}, 'hparam': {'foreground': '1e30', 'background': '1e30', 'center': '1', 'size': '1', 'cls': '1'},
                   'train': {'cross_entropy': '1'}})
    anchors = O.anchors_yolo_voc()
    dnn = model.yolo2.Darknet(model.ConfigChannels(cfg), anchors, 20)
    dnn.load_state_dict(O.make_state_dict(0), strict=False)
    dnn = dnn.to(DEV).train()
    inference = model.Inference(cfg, dnn, anchors).train()
    before = {n: p.detach().clone() for n, p in dnn.named_parameters()}
    opt = torch.optim.Adam(dnn.parameters(), 1e-3, fused=True)
    x = O.synth_images(2, 64, 64, seed=3)
    t = O.synth_targets(2, 64, 64, slots=3, seed=4)
    yb_train.iterate(inference, opt, anchors, cfg, dict(tensor=x, yx_min=t['yx_min'], yx_max=t['yx_max'], cls=t['cls']))
    assert dnn.trainer.found_inf.item() == 1.0
    for n, p in dnn.named_parameters():
        assert bool(torch.isfinite(p.grad).all()) and torch.equal(p.detach(), before[n]), n      # fused Adam skipped the step


def test_graphed_training_step_matches_eager():
    """train.GraphedStep (whole iteration replayed as one CUDA graph) against eager train.iterate from the same
    initial state: same first-step loss, parameters and running statistics after three SGD steps agree (the only
    run-to-run difference is the order of the fp32 atomics in the weight-gradient kernel)."""
    import model
    import model.yolo2
    import train as yb_train
    cfg = make_config(1)
    cfg.read_dict({'model': {'threshold': '0.6'}, 'hparam': {k: str(v) for k, v in O.HPARAM_DEFAULT.items()}, 'train': {'cross_entropy': '1'}})
    anchors = O.anchors_yolo_voc()
    sd0 = O.make_state_dict(0)
    b, size = 4, 128
    batches = []
    for i in range(2):
        t = O.synth_targets(b, size, size, slots=6, seed=21 + i)
        batches.append(dict(tensor=O.synth_images(b, size, size, seed=31 + i).to(DEV), yx_min=t['yx_min'].to(DEV), yx_max=t['yx_max'].to(DEV),
                            cls=t['cls'].to(DEV)))

    def run(graphed):
        dnn = model.yolo2.Darknet(model.ConfigChannels(cfg), anchors, 20)
        dnn.load_state_dict(sd0, strict=False)
        dnn = dnn.to(DEV).train()
        inference = model.Inference(cfg, dnn, anchors).train()
        opt = torch.optim.SGD(dnn.parameters(), 1e-3, momentum=0.9)
        step = yb_train.GraphedStep(inference, opt, anchors, cfg) if graphed else (lambda d: yb_train.iterate(inference, opt, anchors, cfg, d))
        losses = []
        for i in range(3):
            out = step(batches[i % 2])
            losses.append(float(out['loss_total'].item()))
        if graphed:
            assert step.launches > 0 and len(step.graphs) == 1
            step.finish()
        return losses, {k: v.detach().float().cpu().clone() for k, v in dnn.state_dict().items()}

    l_e, sd_e = run(False)
    l_g, sd_g = run(True)
    print('eager losses %s graphed losses %s' % (l_e, l_g))
    # Not bit-identical, and not even close to it: the float shared-memory atomics of the BN statistics make two eager
    # runs differ in the last bit, and the step is discontinuous in such perturbations (max-pool ties, fp16 rounding,
    # the region loss's best-IoU matching and its IoU < 0.6 "negative" threshold), so a 1-ulp change moves the loss by
    # ~1e-2 (measured).  The check is therefore structural: same loss scale, the three updates point the same way, the
    # running statistics agree, and the step counter advanced exactly three times (no warm-up step leaked).
    for a, g in zip(l_e, l_g):
        assert abs(a - g) <= 0.1 * abs(a)
    for k in sd_e:
        if k.endswith('num_batches_tracked'):
            assert int(sd_e[k]) == int(sd_g[k]) == 3, k
            continue
        de, dg = (sd_e[k] - sd0[k].float()).flatten(), (sd_g[k] - sd0[k].float()).flatten()
        assert de.norm().item() > 0 and dg.norm().item() > 0, k
        if 'running' in k:
            assert (sd_e[k] - sd_g[k]).norm().item() <= 0.05 * sd_e[k].norm().item() + 1e-6, k
        elif k.endswith('conv.weight'):
            cos = (torch.dot(de, dg) / (de.norm() * dg.norm())).item()
            assert cos >= 0.8, '%s: update cosine %.3f' % (k, cos)
            assert 0.5 <= (dg.norm() / de.norm()).item() <= 2.0, k


@pytest.mark.parametrize('case', [(3, 26, 26, 64, 192, 3), (2, 13, 13, 256, 512, 1), (5, 10, 14, 128, 1024, 3), (64, 13, 13, 512, 1024, 3),
                                  (3, 32, 24, 32, 64, 3), (6, 208, 208, 32, 64, 3), (2, 48, 40, 32, 32, 3)])      # the last three: the Cin = 32 halo-tile kernel
def test_conv_fused_batch_statistics(ops, case):
    """yb_conv_bn_act_stats_fwd: same z as the plain kernel, bit for bit, and per-channel sum / sum of squares of the
    stored fp16 values equal to a separate yb_bn_stats pass (float partial sums in a different order: 1e-5)."""
    b, h, w, cin, cout, k = case
    gen = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(b, h, w, cin, generator=gen).half().to(DEV)
    wt = (torch.randn(cout, cin, k, k, generator=gen) * (2.0 / (cin * k * k)) ** 0.5)
    w16 = ops.pack_weight_f16(wt.to(DEV))
    one, zero = torch.ones(cout, device=DEV), torch.zeros(cout, device=DEV)
    z_plain = ops.conv_bn_act(x, w16, one, zero, 1.0)
    sums = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    z = ops.conv_bn_act_stats(x, w16, one, zero, 1.0, sums)
    assert torch.equal(z, z_plain)
    zf = z.double().reshape(-1, cout)
    exact = torch.cat([zf.sum(0), (zf * zf).sum(0)])
    scale = exact[cout:].sqrt().repeat(2) * (b * h * w) ** 0.5 + 1e-12      # ~ rows * rms: the natural size of both sums
    assert ((sums - exact).abs() / scale).max().item() <= 1e-5
    if 256 % (cout // 8) == 0:                                              # shapes the stand-alone statistics kernel takes
        ref = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
        ops.call('yb_bn_stats', z, cout, b * h * w, cout, ref)
        assert ((ref - exact).abs() / scale).max().item() <= 1e-5


@pytest.mark.parametrize('size', [320, 608])
def test_training_step_multiscale_shapes(size):
    """BASELINE configs[3] trains at {320, 416, 608}: one graphed step per size through the same GraphedStep object (one
    CUDA graph per input shape, shared optimizer state); grids of 10x10 / 19x19 cells exercise the ragged-tile paths of
    every training kernel.  Checks: finite loss of the usual size, every parameter moved, running statistics updated."""
    import model
    import model.yolo2
    import train as yb_train
    cfg = make_config(1)
    cfg.read_dict({'model': {'threshold': '0.6'}, 'hparam': {k: str(v) for k, v in O.HPARAM_DEFAULT.items()}, 'train': {'cross_entropy': '1'}})
    anchors = O.anchors_yolo_voc()
    sd0 = O.make_state_dict(0)
    dnn = model.yolo2.Darknet(model.ConfigChannels(cfg), anchors, 20)
    dnn.load_state_dict(sd0, strict=False)
    dnn = dnn.to(DEV).train()
    inference = model.Inference(cfg, dnn, anchors).train()
    opt = torch.optim.SGD(dnn.parameters(), 1e-3, momentum=0.9)
    step = yb_train.GraphedStep(inference, opt, anchors, cfg)
    losses = []
    for i, s in enumerate((size, 416, size)):
        t = O.synth_targets(2, s, s, slots=5, seed=40 + i)
        out = step(dict(tensor=O.synth_images(2, s, s, seed=50 + i).to(DEV), yx_min=t['yx_min'].to(DEV), yx_max=t['yx_max'].to(DEV), cls=t['cls'].to(DEV)))
        assert (out['rows'], out['cols']) == (s // 32, s // 32)
        losses.append(float(out['loss_total'].item()))
    assert len(step.graphs) == 2 and all(l == l and 0.0 < l < 10.0 for l in losses), losses
    sd = dnn.state_dict()
    for k, v in sd0.items():
        if k.endswith('conv.weight') or k.endswith('running_mean'):
            assert torch.isfinite(sd[k]).all() and not torch.equal(sd[k].float().cpu(), v.float()), k
    assert int(sd['layers1.0.bn.num_batches_tracked']) == 3


def test_training_repacks_operands_under_fused_optimizer(ops):
    """torch.optim.Adam(fused=True) updates parameters without advancing their version counters; the training path must
    still see the new weights (it re-packs every step), and switching to eval() must re-fold / re-pack too."""
    import model
    import model.yolo2
    import train as yb_train
    cfg = make_config(1)
    cfg.read_dict({'model': {'threshold': '0.6'}, 'hparam': {k: str(v) for k, v in O.HPARAM_DEFAULT.items()}, 'train': {'cross_entropy': '1'}})
    anchors = O.anchors_yolo_voc()
    dnn = model.yolo2.Darknet(model.ConfigChannels(cfg), anchors, 20)
    dnn.load_state_dict(O.make_state_dict(0), strict=False)
    dnn = dnn.to(DEV).train()
    inference = model.Inference(cfg, dnn, anchors).train()
    opt = torch.optim.Adam(dnn.parameters(), 1e-3, fused=True)
    t = O.synth_targets(2, 64, 64, slots=4, seed=3)
    batch = dict(tensor=O.synth_images(2, 64, 64, seed=4).to(DEV), yx_min=t['yx_min'].to(DEV), yx_max=t['yx_max'].to(DEV), cls=t['cls'].to(DEV))
    w0 = dnn.layers2[1].conv.weight.detach().clone()
    yb_train.iterate(inference, opt, anchors, cfg, batch)
    w1 = dnn.layers2[1].conv.weight.detach().clone()
    assert not torch.equal(w0, w1)
    yb_train.iterate(inference, opt, anchors, cfg, batch)          # this forward must have packed w1
    unit = dnn.engine.units2[0]
    assert torch.equal(unit.w16, ops.pack_weight_f16(w1.contiguous()))
    w2 = dnn.layers2[1].conv.weight.detach().clone()
    dnn.eval(); inference.eval()
    dnn(batch['tensor'])
    assert torch.equal(unit.w16, ops.pack_weight_f16(w2.contiguous()))


# ------------------------------------------------------------------------------------------------
# MobileNet plugin (BASELINE configs[4])
# ------------------------------------------------------------------------------------------------
def test_mobilenet_plugin_vs_reference_golden(golden_dir):
    import model
    import model.mobilenet
    import utils
    g = np.load(os.path.join(golden_dir, 'mobilenet.npz'))
    cfg = make_config(1)
    cls = utils.parse_attr('model.mobilenet.MobileNet')
    net = cls(model.ConfigChannels(cfg), O.anchors_yolo_voc(), 20)
    res = net.load_state_dict(O.make_mobilenet_state_dict(0), strict=False)
    assert not res.unexpected_keys and not res.missing_keys
    net = net.to(DEV).eval()
    f64 = net(O.synth_images(1, 64, 64, seed=10).to(DEV))
    f416 = net(O.synth_images(1, 416, 416, seed=0).to(DEV))
    e64, e416 = rel_err(f64, torch.from_numpy(g['feature64'])), rel_err(f416, torch.from_numpy(g['feature416']))
    assert f416.shape == (1, 125, 13, 13) and e64 <= 3e-3 and e416 <= 3e-3
    # strict precision ([hi | lo] activations, split-precision pointwise convs): the 1e-3 contract
    net.set_precision('strict')
    s64 = rel_err(net(O.synth_images(1, 64, 64, seed=10).to(DEV)), torch.from_numpy(g['feature64']))
    s416 = rel_err(net(O.synth_images(1, 416, 416, seed=0).to(DEV)), torch.from_numpy(g['feature416']))
    record('mobilenet_golden', dict(feature64=e64, feature416=e416, strict_feature64=s64, strict_feature416=s416))
    assert s64 <= TOL_CONTRACT and s416 <= TOL_CONTRACT, (s64, s416)
    net.set_precision('fast')
    # through the detection head: Inference + postprocess_batch run on any plugin backbone
    import detect
    inference = model.Inference(cfg, net, O.anchors_yolo_voc()).eval()
    pred = model._inference(inference, O.synth_images(3, 416, 416, seed=2).to(DEV))
    assert len(detect.postprocess_batch(cfg, pred)) == 3


def test_pack_weights_batch_matches_per_unit(ops):
    """yb_pack_weights_batch (one launch for all units of a training step) == yb_pack_weight_f16 + yb_pack_weight_dgrad_f16 per unit, bit for bit,
    including the zero-padded filters of the head and channel counts that do not fill a tile."""
    from b200.train_engine import PackPlan
    g = torch.Generator().manual_seed(5)
    shapes = [(64, 32, 3, 0), (128, 64, 3, 0), (64, 128, 1, 0), (1024, 1280, 3, 0), (125, 1024, 1, 128), (40, 24, 3, 48), (72, 520, 1, 0), (256, 16, 3, 0)]
    ws = [torch.randn(co, ci, k, k, generator=g).to(DEV) for co, ci, k, _ in shapes]
    plan = PackPlan([('u%d' % i, w, True, True, cp) for i, (w, (_, _, _, cp)) in enumerate(zip(ws, shapes))], torch.device(DEV))
    for t in list(plan.fwd.values()) + list(plan.dgrad.values()):
        t.fill_(float('nan'))
    plan.run()
    for i, (w, (co, ci, k, cp)) in enumerate(zip(ws, shapes)):
        ref_f = ops.pack_weight_f16(w, 0)
        cpad = max(co, cp)
        ref_d = torch.empty(ci, k, k, cpad, dtype=torch.float16, device=DEV)
        ops.call('yb_pack_weight_dgrad_f16', w, ref_d, co, ci, k, cpad)
        assert torch.equal(plan.fwd['u%d' % i], ref_f), shapes[i]
        assert torch.equal(plan.dgrad['u%d' % i], ref_d), shapes[i]
    # forward-only / dgrad-only entries
    plan2 = PackPlan([('a', ws[0], True, False, 0), ('b', ws[1], False, True, 0)], torch.device(DEV))
    plan2.run()
    assert torch.equal(plan2.fwd['a'], ops.pack_weight_f16(ws[0], 0)) and 'a' not in plan2.dgrad and 'b' not in plan2.fwd


def test_resnet_kernels_vs_torch(ops):
    """Stem 7x7 s2 + BN + ReLU, max-pool 3x3 s2 p1 (bit-exact), x[::2, ::2] (bit-exact) and relu(a + b) against torch on the same fp16 inputs;
    and the identity the stride-2 blocks rely on: conv3x3(stride 1)[::2, ::2] == conv3x3(stride 2)."""
    g = torch.Generator().manual_seed(3)
    for (b, h, w) in ((2, 64, 96), (1, 416, 416)):
        x = torch.rand(b, 3, h, w, generator=g)
        wt = torch.randn(64, 3, 7, 7, generator=g) * 0.1
        scale, shift = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
        y = torch.empty(b, h // 2, w // 2, 64, dtype=torch.float16, device=DEV)
        ops.call('yb_stem7x7_bn_relu_fwd', x.to(DEV), wt.to(DEV), scale.to(DEV), shift.to(DEV), y, b, h, w)
        ref = torch.relu(torch.nn.functional.conv2d(x.double(), wt.double(), None, 2, 3) * scale.double()[None, :, None, None] + shift.double()[None, :, None, None])
        assert rel_err(y.permute(0, 3, 1, 2), ref.float()) <= 6e-4               # one fp16 rounding of the output
    for (b, h, w, c) in ((2, 32, 48, 64), (1, 13, 27, 8), (3, 208, 208, 64)):
        x = torch.randn(b, h, w, c, generator=g).half().to(DEV)
        oh, ow = (h + 1) // 2, (w + 1) // 2
        y = torch.empty(b, oh, ow, c, dtype=torch.float16, device=DEV)
        ops.call('yb_maxpool3x3_s2_f16', x, y, b, h, w, c)
        ref = torch.nn.functional.max_pool2d(x.permute(0, 3, 1, 2).float(), 3, 2, 1).permute(0, 2, 3, 1)
        assert torch.equal(y.float(), ref)
        ops.call('yb_subsample2_f16', x, y, b, h, w, c)
        assert torch.equal(y, x[:, ::2, ::2, :])
    a = torch.randn(5, 13, 13, 512, generator=g).half().to(DEV)
    r = torch.randn(5, 13, 13, 512, generator=g).half().to(DEV)
    out = torch.empty_like(a)
    ops.call('yb_add_relu_f16', a, r, out, a.numel())
    assert torch.equal(out, torch.relu(a.float() + r.float()).half())
    ops.call('yb_add_relu_f16', a, r, a, a.numel())                               # in place, as the blocks use it
    assert torch.equal(a, out)
    x = torch.randn(1, 64, 26, 26, generator=g)
    wt = torch.randn(128, 64, 3, 3, generator=g)
    s1 = torch.nn.functional.conv2d(x, wt, None, 1, 1)[:, :, ::2, ::2]
    s2 = torch.nn.functional.conv2d(x, wt, None, 2, 1)
    assert s1.shape == s2.shape and (s1 - s2).abs().max() <= 1e-4 * s2.abs().max()


def test_resnet_plugin_vs_reference_golden(golden_dir):
    """model.resnet.resnet18 / resnet50 on the B200 kernels vs the EXECUTED reference (tests/golden/make_golden_resnet.py): stem pool and
    every block's output at 64x64, head feature at 64 and 416.  fp16 operands through 17 (resnet18) / 49 (resnet50) convs: <= 3e-3."""
    import detect
    import model
    import model.resnet
    import utils
    g = np.load(os.path.join(golden_dir, 'resnet.npz'))
    cfg = make_config(1)
    anchors = O.anchors_yolo_voc()
    rec = {}
    for name in ('resnet18', 'resnet50'):
        net = utils.parse_attr('model.resnet.' + name)(model.ConfigChannels(cfg), anchors, 20)
        res = net.load_state_dict(O.make_resnet_state_dict(name, 0), strict=False)
        assert not res.unexpected_keys and not res.missing_keys
        net = net.to(DEV).eval()
        f64 = net(O.synth_images(1, 64, 64, seed=10).to(DEV))
        rec[name + '_feature64'] = rel_err(f64, torch.from_numpy(g[name + '_feature64']))
        if name == 'resnet18':
            f416 = net(O.synth_images(1, 416, 416, seed=0).to(DEV))
            rec['resnet18_feature416'] = rel_err(f416, torch.from_numpy(g['resnet18_feature416']))
            assert f416.shape == (1, 125, 13, 13)
            # block by block at 64x64 through the model's own block runner
            x = O.synth_images(1, 64, 64, seed=10).to(DEV)
            scale, shift = net._fold('bn1', net.bn1)
            stem = torch.empty(1, 32, 32, 64, dtype=torch.float16, device=DEV)
            from b200 import ops as _ops
            _ops.call('yb_stem7x7_bn_relu_fwd', x, net.conv1.weight.detach().contiguous(), scale, shift, stem, 1, 64, 64)
            cur = torch.empty(1, 16, 16, 64, dtype=torch.float16, device=DEV)
            _ops.call('yb_maxpool3x3_s2_f16', stem, cur, 1, 32, 32, 64)
            worst = (rel_err(cur.permute(0, 3, 1, 2), torch.from_numpy(g['resnet18_act_maxpool'])), 'maxpool')
            for lname in ('layer1', 'layer2', 'layer3', 'layer4'):
                for bname, blk in getattr(net, lname).named_children():
                    key = '%s.%s' % (lname, bname)
                    cur = net._block(key, blk, cur)
                    e = rel_err(cur.permute(0, 3, 1, 2), torch.from_numpy(g['resnet18_act_' + key]))
                    worst = max(worst, (e, key))
            rec['resnet18_worst_block'] = list(worst)
            assert worst[0] <= 3e-3, worst
            inference = model.Inference(cfg, net, anchors).eval()
            pred = model._inference(inference, O.synth_images(3, 416, 416, seed=2).to(DEV))
            assert len(detect.postprocess_batch(cfg, pred)) == 3
    record('resnet_golden', rec)
    assert all(v <= 3e-3 for k, v in rec.items() if 'feature' in k), rec


def test_c5_mobilenet_batch32_vs_oracle():
    """BASELINE configs[4] at its real size: MobileNet backbone on 32 x 3 x 416 x 416, head feature vs the oracle (itself pinned to the
    reference's MobileNet by mobilenet.npz) and the detection chain on top.  27
    conv layers drift 1.5e-3 .. 2.5e-3 end to end in the default `fast` mode (asserted <= 3e-3, measured value recorded); `strict` must meet 1e-3."""
    import detect
    import model
    import model.mobilenet
    cfg = make_config(1)
    anchors = O.anchors_yolo_voc()
    sd = O.make_mobilenet_state_dict(0)
    net = model.mobilenet.MobileNet(model.ConfigChannels(cfg), anchors, 20)
    net.load_state_dict(sd, strict=False)
    net = net.to(DEV).eval()
    x = O.synth_images(32, 416, 416, seed=50)
    with torch.no_grad():
        ref = O.mobilenet_forward(sd, x)
    inference = model.Inference(cfg, net, anchors).eval()
    pred = model._inference(inference, x.to(DEV))
    f = pred['feature'].cpu()
    e = rel_err(f, ref)
    per_image = max(((f[i] - ref[i]).abs().max() / ref[i].abs().max()).item() for i in range(32))
    results = detect.postprocess_batch(cfg, pred)
    assert f.shape == (32, 125, 13, 13) and e <= 3e-3 and per_image <= 4e-3
    assert len(results) == 32
    net.set_precision('strict')
    fs = model._inference(inference, x.to(DEV))['feature'].cpu()
    es = rel_err(fs, ref)
    per_image_s = max(((fs[i] - ref[i]).abs().max() / ref[i].abs().max()).item() for i in range(32))
    record('c5_mobilenet_batch32', dict(feature=e, worst_image=per_image, strict_feature=es, strict_worst_image=per_image_s,
                                        detections=sum(0 if r is None else len(r[3]) for r in results)))
    assert es <= TOL_CONTRACT and per_image_s <= TOL_CONTRACT, (es, per_image_s)


def test_mobilenet_training_kernels_vs_torch(ops):
    """The three training kernels of the MobileNet plugin against torch autograd on fp16-representable data: depthwise data gradient and
    weight gradient (stride 1 and 2), first-layer (stride-2) weight gradient."""
    for (b, h, c, stride) in ((2, 12, 64, 1), (2, 12, 64, 2), (1, 26, 256, 2), (3, 13, 1024, 1)):
        gen = torch.Generator().manual_seed(h + c + stride)
        a = torch.randn(b, c, h, h, generator=gen).half().float()
        w = (torch.randn(c, 1, 3, 3, generator=gen) * 0.3)
        dz = torch.randn(b, c, h // stride, h // stride, generator=gen).half().float()
        ar, wr = a.clone().requires_grad_(True), w.clone().requires_grad_(True)
        torch.nn.functional.conv2d(ar, wr, None, stride, 1, groups=c).backward(dz)
        a16 = a.permute(0, 2, 3, 1).contiguous().half().to(DEV)
        dz16 = dz.permute(0, 2, 3, 1).contiguous().half().to(DEV)
        w9 = w.view(c, 9).contiguous().to(DEV)
        da = torch.empty(b, h, h, c, dtype=torch.float16, device=DEV)
        ops.call('yb_dwconv3x3_dgrad', dz16, w9, da, b, h, h, c, stride)
        assert rel_err(da.permute(0, 3, 1, 2), ar.grad) <= 2e-3, (b, h, c, stride)
        dw = torch.full((c, 9), 7.0, dtype=torch.float32, device=DEV)
        ops.call('yb_dwconv3x3_wgrad', a16, dz16, dw, b, h, h, c, stride)
        assert rel_err(dw.view(c, 1, 3, 3), wr.grad) <= 1e-4, (b, h, c, stride)
        # raw forward = the conv itself
        z = torch.empty(b, h // stride, h // stride, c, dtype=torch.float16, device=DEV)
        ops.call('yb_dwconv3x3_raw_fwd', a16, w9, z, b, h, h, c, stride)
        ref = torch.nn.functional.conv2d(a, w, None, stride, 1, groups=c)
        assert rel_err(z.permute(0, 3, 1, 2), ref) <= 1e-3
    gen = torch.Generator().manual_seed(3)
    x = torch.rand(2, 3, 32, 48, generator=gen)
    w0 = torch.randn(32, 3, 3, 3, generator=gen) * 0.2
    dz = torch.randn(2, 32, 16, 24, generator=gen).half().float()
    wr = w0.clone().requires_grad_(True)
    torch.nn.functional.conv2d(x, wr, None, 2, 1).backward(dz)
    dw = torch.full((32, 3, 3, 3), 5.0, dtype=torch.float32, device=DEV)
    ops.call('yb_mb_conv0_wgrad', x.to(DEV), dz.permute(0, 2, 3, 1).contiguous().half().to(DEV), dw, 2, 32, 48)
    assert rel_err(dw, wr.grad) <= 1e-4
    z = torch.empty(2, 16, 24, 32, dtype=torch.float16, device=DEV)
    ops.call('yb_mb_conv0_raw_fwd', x.to(DEV), w0.to(DEV), z, 2, 32, 48)
    assert rel_err(z.permute(0, 3, 1, 2), torch.nn.functional.conv2d(x, w0, None, 2, 1)) <= 1e-3


def test_mobilenet_training_step_vs_oracle_and_descent():
    """model.mobilenet.MobileNet in train() mode: one step (train-mode forward with batch statistics at momentum 0.1, region loss, full
    backward through 27 BatchNorm layers) against the oracle's arithmetic with torch autograd on CPU, and 15 SGD steps on one batch reduce
    the loss.  27 train-mode BatchNorm layers amplify the fp16 roundings like Darknet-19's 22 do (see the C3 test): the head feature moves
    by ~7e-2 and the first layers' gradients decorrelate (cosine 0.7), so the tight bounds sit where the chain is short -- the head and the
    last unit -- and on the running statistics; the kernels themselves are pinned against autograd in
    test_mobilenet_training_kernels_vs_torch."""
    import model
    import model.mobilenet
    import train as yb_train
    cfg = make_config(1)
    cfg.read_dict({'model': {'threshold': '0.6'}, 'hparam': {'foreground': '5', 'background': '1', 'center': '1', 'size': '1', 'cls': '1'},
                   'train': {'cross_entropy': '1'}})
    anchors = O.anchors_yolo_voc()
    sd0 = O.make_mobilenet_state_dict(0)
    b, size = 8, 160
    s = size // 32
    x = O.synth_images(b, size, size, seed=80)
    tgt = O.synth_targets(b, size, size, slots=5, seed=81)
    data = O.norm_data(tgt, size, size, s, s)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.clone()) for k, v in sd0.items()}
    stats = {}
    f_ref = O.mobilenet_forward(sd, x, train=True, stats=stats)
    l_ref, _ = O.loss(anchors, data, O.decode(f_ref, anchors), 0.6)
    O.loss_total(l_ref).backward()
    net = model.mobilenet.MobileNet(model.ConfigChannels(cfg), anchors, 20)
    net.load_state_dict(sd0, strict=False)
    net = net.to(DEV).train()
    inference = model.Inference(cfg, net, anchors).train()
    pred = model._inference(inference, x.to(DEV))
    losses, _ = model.loss(anchors, {k: v.to(DEV) for k, v in data.items()}, pred, 0.6)
    sum(losses[k] * O.HPARAM_DEFAULT[k] for k in losses).backward()
    e_f = rel_err(pred['feature'], f_ref)
    e_loss = {k: abs(losses[k].item() - l_ref[k].item()) / abs(l_ref[k].item()) for k in losses}
    worst_cos, worst_rel, late_cos = (1.0, None), (0.0, None), (1.0, None)
    for name, p in net.named_parameters():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), name
        g, r = p.grad.detach().float().cpu().flatten(), sd[name].grad.flatten()
        cos = (torch.dot(g, r) / (g.norm() * r.norm() + 1e-30)).item()
        rel = ((g - r).norm() / (r.norm() + 1e-30)).item()
        if cos < worst_cos[0]:
            worst_cos = (cos, name)
        if rel > worst_rel[0]:
            worst_rel = (rel, name)
        if (name.startswith('layers.14') or name.startswith('layers.13.pw')) and cos < late_cos[0]:
            late_cos = (cos, name)
    e_run = 0.0
    bufs = dict(net.named_buffers())
    for prefix, (mean, var) in stats.items():
        exp = 0.9 * sd0[prefix + '.running_mean'] + 0.1 * mean.detach()          # nn.BatchNorm2d default momentum 0.1 (model/mobilenet.py:28)
        e_run = max(e_run, rel_err(bufs[prefix + '.running_mean'].cpu(), exp))
    record('mobilenet_train_step', dict(feature=e_f, losses=e_loss, worst_grad_cosine=worst_cos, late_grad_cosine=late_cos, worst_grad_rel_l2=worst_rel,
                                        running_mean=e_run))
    assert e_f <= 0.15, e_f
    for k, v in e_loss.items():
        assert v <= 0.15, (k, v)
    assert late_cos[0] >= 0.9, late_cos
    assert worst_cos[0] >= 0.5, worst_cos
    assert e_run <= 1e-2, e_run
    opt = torch.optim.SGD(net.parameters(), 1e-3, momentum=0.9)
    batch = dict(tensor=x, yx_min=tgt['yx_min'], yx_max=tgt['yx_max'], cls=tgt['cls'])
    hist = [float(yb_train.iterate(inference, opt, anchors, cfg, batch)['loss_total'].item()) for _ in range(15)]
    assert hist[-1] < 0.9 * hist[0], hist
    net.eval()
    f = net(x[:2].to(DEV))
    assert f.shape == (2, 125, s, s) and bool(torch.isfinite(f).all())


def test_mobilenet_depthwise_vs_torch(ops):
    for (b, h, c, stride) in ((2, 13, 1024, 1), (3, 52, 128, 2), (2, 7, 64, 1), (1, 60, 32, 1)):          # strip lengths 4 and 8, ragged last strips
        gen = torch.Generator().manual_seed(h * c)
        x16 = torch.randn(b, h, h, c, generator=gen).half().to(DEV)
        w = torch.randn(c, 1, 3, 3, generator=gen) * 0.3
        scale, shift = (torch.rand(c, generator=gen) + 0.5).to(DEV), (torch.randn(c, generator=gen) * 0.1).to(DEV)
        if h % stride:
            continue
        y = torch.empty(b, h // stride, h // stride, c, dtype=torch.float16, device=DEV)
        ops.call('yb_dwconv3x3_bn_relu_fwd', x16, w.view(c, 9).contiguous().to(DEV), scale, shift, y, b, h, h, c, stride)
        ref = torch.relu(torch.nn.functional.conv2d(x16.float().cpu().permute(0, 3, 1, 2), w, stride=stride, padding=1, groups=c)
                         * scale.cpu()[None, :, None, None] + shift.cpu()[None, :, None, None])
        assert rel_err(y.permute(0, 3, 1, 2), ref) <= 1e-3, (b, h, c, stride)
    for (b, h, c, stride) in ((2, 26, 256, 1), (2, 26, 256, 2), (1, 104, 64, 2)):
        gen = torch.Generator().manual_seed(c + stride)
        x = torch.randn(b, c, h, h, generator=gen).half().float()
        w = torch.randn(c, 1, 3, 3, generator=gen) * 0.4
        scale, shift = torch.rand(c, generator=gen) + 0.5, torch.randn(c, generator=gen) * 0.1
        ref = torch.relu(torch.nn.functional.conv2d(x, w, None, stride, 1, groups=c) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
        y = torch.empty(b, h // stride, h // stride, c, dtype=torch.float16, device=DEV)
        ops.call('yb_dwconv3x3_bn_relu_fwd', x.to(DEV).permute(0, 2, 3, 1).contiguous().half(), w.to(DEV).view(c, 9).contiguous(), scale.to(DEV),
                 shift.to(DEV), y, b, h, h, c, stride)
        assert rel_err(y.permute(0, 3, 1, 2), ref) <= 1e-3


# ------------------------------------------------------------------------------------------------
# Tiny YOLOv2 plugin (reference model/yolo2.py:140-173)
# ------------------------------------------------------------------------------------------------
def test_maxpool_stride1_padded_exact(ops):
    """ConstantPad2d((0,1,0,1), float32 min) + MaxPool2d(2, stride=1): bit-exact against torch on fp16 values."""
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(3, 13, 13, 64, generator=gen).half()
    ref = torch.nn.functional.max_pool2d(torch.nn.functional.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1), value=O.FLOAT32_MIN), 2, stride=1)
    y = ops.maxpool2x2_s1(x.to(DEV))
    assert torch.equal(y.float().cpu().permute(0, 3, 1, 2), ref)


def test_tiny_plugin_vs_reference_golden(golden_dir):
    import model
    import model.yolo2
    g = np.load(os.path.join(golden_dir, 'tiny.npz'))
    cfg = make_config(1)
    anchors = O.anchors_yolo_voc()
    net = model.yolo2.Tiny(model.ConfigChannels(cfg), anchors, 20)
    res = net.load_state_dict(O.make_tiny_state_dict(0), strict=False)
    assert not res.unexpected_keys and all(k.endswith('num_batches_tracked') for k in res.missing_keys)
    assert net.scope('layers.4.conv.weight') == 'layers.4'
    net = net.to(DEV).eval()
    for size, key in ((64, 'feature64'), (416, 'feature416')):
        x = O.synth_images(1, size, size, seed=10 if size == 64 else 0)
        f = net(x.to(DEV))
        e = rel_err(f, torch.from_numpy(g[key]))
        print('tiny %d feature rel err %.3e' % (size, e))
        assert f.shape == g[key].shape and e <= 3e-3
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 64, 64))


# ------------------------------------------------------------------------------------------------
# Evaluation matching on the device (reference eval.py:57-75)
# ------------------------------------------------------------------------------------------------
def test_maxpool_stride1_backward_vs_torch(ops):
    """yb_maxpool2x2_s1_bwd_f16 (training of Tiny): gradient routed to the first maximum of every window, vs torch autograd through
    F.pad(-inf-like) + max_pool2d(2, stride=1) on fp16-representable data with deliberate ties."""
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(3, 13, 13, 64, generator=gen).half()
    x[0, :4, :4, :16] = 0.5                                  # ties: the first element of the window wins
    dy = torch.randn(3, 13, 13, 64, generator=gen).half()
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    y = torch.nn.functional.max_pool2d(torch.nn.functional.pad(xr, (0, 1, 0, 1), value=float(np.finfo(np.float32).min)), 2, stride=1)
    y.backward(dy.float().permute(0, 3, 1, 2))
    dx = torch.empty_like(x, device=DEV)
    ops.call('yb_maxpool2x2_s1_bwd_f16', x.to(DEV), dy.to(DEV), dx, 3, 13, 13, 64)
    ref = xr.grad.permute(0, 2, 3, 1)
    assert rel_err(dx, ref) <= 2e-3                          # sums of up to four fp16 gradients, rounded once to fp16


def test_tiny_training_step_vs_oracle_and_descent():
    """model.yolo2.Tiny in train() mode (the reference's default `model/dnn`, config.ini:25; model/yolo2.py:140-173): one step --
    train-mode forward (batch-statistics BN, the 16-filter first layer on the zero-padded 32-filter kernels, MaxPool2d(2) x5, pad + stride-1
    pool), region loss, full backward -- against the oracle's arithmetic with torch autograd on CPU: feature, losses, every parameter
    gradient (cosine / relative L2; 8 BatchNorm layers amplify fp16 roundings far less than Darknet-19's 22), running statistics; the
    padding channels stay exactly zero; and 15 SGD steps on one batch reduce the loss."""
    import model
    import model.yolo2
    import train as yb_train
    cfg = make_config(1)
    cfg.read_dict({'model': {'threshold': '0.6'}, 'hparam': {'foreground': '5', 'background': '1', 'center': '1', 'size': '1', 'cls': '1'},
                   'train': {'cross_entropy': '1'}})
    anchors = O.anchors_yolo_voc()
    sd0 = O.make_tiny_state_dict(0)
    b, size = 8, 160
    s = size // 32
    x = O.synth_images(b, size, size, seed=70)
    data = O.norm_data(O.synth_targets(b, size, size, slots=5, seed=71), size, size, s, s)
    # oracle step
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.clone()) for k, v in sd0.items()}
    stats = {}
    f_ref = O.tiny_forward(sd, x, train=True, stats=stats)
    l_ref, _ = O.loss(anchors, data, O.decode(f_ref, anchors), 0.6)
    O.loss_total(l_ref).backward()
    # CUDA step
    dnn = model.yolo2.Tiny(model.ConfigChannels(cfg), anchors, 20)
    dnn.load_state_dict(sd0, strict=False)
    dnn = dnn.to(DEV).train()
    inference = model.Inference(cfg, dnn, anchors).train()
    pred = model._inference(inference, x.to(DEV))
    losses, _ = model.loss(anchors, {k: v.to(DEV) for k, v in data.items()}, pred, 0.6)
    sum(losses[k] * O.HPARAM_DEFAULT[k] for k in losses).backward()
    e_f = rel_err(pred['feature'], f_ref)
    e_loss = {k: abs(losses[k].item() - l_ref[k].item()) / abs(l_ref[k].item()) for k in losses}
    worst_cos, worst_rel = (1.0, None), (0.0, None)
    for name, p in dnn.named_parameters():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), name
        g, r = p.grad.detach().float().cpu().flatten(), sd[name].grad.flatten()
        cos = (torch.dot(g, r) / (g.norm() * r.norm() + 1e-30)).item()
        rel = ((g - r).norm() / (r.norm() + 1e-30)).item()
        if cos < worst_cos[0]:
            worst_cos = (cos, name)
        if rel > worst_rel[0]:
            worst_rel = (rel, name)
    e_run = 0.0
    for key, (mean, var) in stats.items():
        rm = dict(dnn.named_buffers())[key + '.bn.running_mean'].cpu()
        exp = 0.99 * sd0[key + '.bn.running_mean'] + 0.01 * mean.detach()
        e_run = max(e_run, rel_err(rm, exp))
    record('tiny_train_step', dict(feature=e_f, losses=e_loss, worst_grad_cosine=worst_cos, worst_grad_rel_l2=worst_rel, running_mean=e_run))
    tr = dnn.trainer
    assert bool((tr._zero_bufs[('z0', b, size, size)][..., 16:] == 0).all()) and bool((tr._zero_bufs[('a0', b, size, size)][..., 16:] == 0).all())
    assert e_f <= 3e-2, e_f
    for k, v in e_loss.items():
        assert v <= 2e-2, (k, v)
    assert worst_cos[0] >= 0.95, worst_cos
    assert worst_rel[0] <= 0.3, worst_rel
    assert e_run <= 1e-3, e_run
    # descent on one batch
    opt = torch.optim.SGD(dnn.parameters(), 1e-3, momentum=0.9)
    batch = dict(tensor=x, yx_min=O.synth_targets(b, size, size, slots=5, seed=71)['yx_min'], yx_max=O.synth_targets(b, size, size, slots=5, seed=71)['yx_max'],
                 cls=O.synth_targets(b, size, size, slots=5, seed=71)['cls'])
    hist = [float(yb_train.iterate(inference, opt, anchors, cfg, batch)['loss_total'].item()) for _ in range(15)]
    assert hist[-1] < 0.9 * hist[0], hist
    # and back to inference on the trained weights (operand caches are dropped by train(False))
    dnn.eval()
    f = dnn(x[:2].to(DEV))
    assert f.shape == (2, 125, s, s) and bool(torch.isfinite(f).all())


def test_eval_matching_vs_reference_golden(golden_dir):
    """yb_eval_match (one launch for the whole ragged batch) and the per-class drop-in eval.matching: true-positive
    flags bit-identical to the reference's, incl. images without ground truth / without detections."""
    import eval as yb_eval
    g = np.load(os.path.join(golden_dir, 'eval.npz'))
    tags = sorted({k.split('_')[0] for k in g.files if k.startswith('case')})
    cases = [{k[len(t) + 1:]: g[k] for k in g.files if k.startswith(t + '_')} for t in tags]
    det_off = np.cumsum([0] + [c['det_cls'].shape[0] for c in cases])
    gt_off = np.cumsum([0] + [c['gt_cls'].shape[0] for c in cases])
    cat = lambda key, dt: torch.from_numpy(np.concatenate([c[key] for c in cases]).astype(dt)).to(DEV)
    tp = yb_eval.matching_batch(cat('det_min', np.float32), cat('det_max', np.float32), cat('det_cls', np.int32), torch.from_numpy(det_off),
                                cat('gt_min', np.float32), cat('gt_max', np.float32), cat('gt_cls', np.int32), torch.from_numpy(gt_off), 20, 0.5)
    assert np.array_equal(tp.cpu().numpy().astype(bool), np.concatenate([c['tp'] for c in cases]))
    # per-class drop-in, as Eval.filter_cls calls it
    case = cases[1]
    for c in range(int(case['num_cls'])):
        dm, gm = case['det_cls'] == c, case['gt_cls'] == c
        got = yb_eval.matching(torch.from_numpy(case['gt_min'][gm]).to(DEV), torch.from_numpy(case['gt_max'][gm]).to(DEV),
                               torch.from_numpy(case['det_min'][dm]).to(DEV), torch.from_numpy(case['det_max'][dm]).to(DEV), 0.5)
        assert got.dtype == bool and np.array_equal(got, case['tp'][dm])
    # a larger random ragged batch against the oracle restatement
    rows, expect = [], []
    for i in range(24):
        case = O.synth_eval_case(100 + i, n_det=40 + 7 * i, n_gt=1 + i, num_cls=20)
        rows.append(case)
        tp_i = np.zeros(case['det_cls'].numel(), dtype=bool)
        for c in range(20):
            dm, gm = case['det_cls'] == c, case['gt_cls'] == c
            tp_i[dm.numpy()] = O.eval_matching(case['gt_min'][gm], case['gt_max'][gm], case['det_min'][dm], case['det_max'][dm], 0.45)
        expect.append(tp_i)
    det_off = np.cumsum([0] + [r['det_cls'].numel() for r in rows])
    gt_off = np.cumsum([0] + [r['gt_cls'].numel() for r in rows])
    tcat = lambda key: torch.cat([r[key] for r in rows]).to(DEV)
    tp = yb_eval.matching_batch(tcat('det_min'), tcat('det_max'), tcat('det_cls'), torch.from_numpy(det_off), tcat('gt_min'), tcat('gt_max'),
                                tcat('gt_cls'), torch.from_numpy(gt_off), 20, 0.45)
    assert np.array_equal(tp.cpu().numpy().astype(bool), np.concatenate(expect))


# ------------------------------------------------------------------------------------------------
# GPU input pipeline (reference transform/resize/label.py:25-31 + transform/image.py:27-29)
# ------------------------------------------------------------------------------------------------
def test_resize_batch_bit_exact_vs_cv2_golden(golden_dir):
    """yb_resize_batch_u8: a ragged batch of frames -> network size in one launch, bit-identical to cv2.resize (goldens made by
    the reference's rescale with cv2), BGR->RGB swap and box scaling included; then straight into the uint8 model input."""
    import hashlib
    import transform
    import transform.resize.image
    import transform.resize.label
    g = np.load(os.path.join(golden_dir, 'resize.npz'))
    out = transform.resize.image.rescale(g['small_src'], 64, 96)
    assert out.is_cuda and np.array_equal(out.cpu().numpy(), g['small_out'])
    for h, w in ((416, 416), (608, 608), (320, 320), (320, 608), (608, 320)):
        seeds = [s for s in range(8) if tuple(int(v) for v in g['case%d_dims' % s][2:]) == (h, w)]
        if not seeds:
            continue
        frames = [torch.from_numpy(O.synth_frame(s, int(g['case%d_dims' % s][0]), int(g['case%d_dims' % s][1]))) for s in seeds]
        batch = transform.resize_batch(frames, h, w, bgr2rgb=False).cpu().numpy()
        for i, s in enumerate(seeds):
            assert hashlib.sha256(batch[i].tobytes()).digest() == g['case%d_sha' % s].tobytes(), s
        swapped = transform.resize_batch(frames, h, w, bgr2rgb=True).cpu().numpy()
        assert np.array_equal(swapped, batch[..., ::-1])
    # labels: same float32 arithmetic as numpy
    src = O.synth_frame(3, 100, 80)
    yx_min = np.array([[3.5, 7.25], [40.0, 11.0], [0.0, 0.0]], np.float32)
    yx_max = np.array([[30.0, 50.5], [99.0, 79.0], [0.0, 0.0]], np.float32)
    r_ref, a_ref, b_ref = O.rescale_label(src, yx_min.copy(), yx_max.copy(), 416, 416)
    r, a, b = transform.resize.label.rescale(src, yx_min, yx_max, 416, 416)
    assert np.array_equal(r.cpu().numpy(), r_ref) and np.array_equal(a.cpu().numpy(), a_ref) and np.array_equal(b.cpu().numpy(), b_ref)
    # the resized RGB uint8 batch is a valid model input (ToTensor's 1/255 is applied by the first conv kernel)
    import model
    import model.yolo2
    cfg = make_config(1)
    anchors = O.anchors_yolo_voc()
    dnn = model.yolo2.Darknet(model.ConfigChannels(cfg), anchors, 20)
    dnn.load_state_dict(O.make_state_dict(0), strict=False)
    dnn = dnn.to(DEV).eval()
    frames = [torch.from_numpy(O.synth_frame(20 + i, 90 + 13 * i, 120 + 7 * i)) for i in range(3)]
    u8 = transform.resize_batch(frames, 64, 64, bgr2rgb=True)
    f_u8 = dnn(u8)
    f_f32 = dnn((u8.float() / 255.0).permute(0, 3, 1, 2).contiguous())
    assert rel_err(f_u8, f_f32) <= 2e-3


def test_flip_and_random_crop_bit_exact_vs_executed_reference(golden_dir):
    """The training-side resize (reference `flip_horizontally` -> `random_crop` -> `rescale`, executed with cv2 by
    tests/golden/make_golden_augment.py): every output frame bit-identical (SHA-256), boxes equal in float32, through the batched launch,
    through the single-image mirrors with the reference's own random draws, and the stand-alone flip."""
    import hashlib
    import configparser
    import transform
    import transform.augmentation
    import transform.resize.label
    g = np.load(os.path.join(golden_dir, 'augment.npz'))
    cfg = configparser.ConfigParser()
    cfg.read_dict({'augmentation': {'random_crop': '1', 'random_flip_horizontally': '0.5'}})
    by_size = {}
    for seed, h0, w0, h, w, flip in g['cases'].tolist():
        src = O.synth_frame(seed, h0, w0)
        yx_min, yx_max = g['c%d_yx_min_in' % seed], g['c%d_yx_max_in' % seed]
        # boxes after the flip decide the window (the reference crops the flipped image)
        fmin, fmax = yx_min.copy(), yx_max.copy()
        if flip:
            t = w0 - fmin[:, 1]
            fmin[:, 1] = w0 - fmax[:, 1]
            fmax[:, 1] = t
        window, margin = transform.resize.label.crop_window(1.0, fmin, fmax, (h0, w0), g['c%d_draws' % seed])
        out, a, b = transform.resize_batch([torch.from_numpy(src)], h, w, bgr2rgb=False, yx_min=torch.from_numpy(yx_min)[None], yx_max=torch.from_numpy(yx_max)[None],
                                           flip=[bool(flip)], crop=[window], margin=[margin.tolist()])
        assert hashlib.sha256(out[0].cpu().numpy().tobytes()).digest() == g['c%d_sha' % seed].tobytes(), 'case %d pixels' % seed
        assert np.array_equal(a[0].cpu().numpy(), g['c%d_yx_min' % seed]) and np.array_equal(b[0].cpu().numpy(), g['c%d_yx_max' % seed]), 'case %d boxes' % seed
        by_size.setdefault((h, w), []).append((seed, src, yx_min, yx_max, bool(flip), window, margin))
        # the single-image mirror with the reference's own np.random draw
        image, bmin, bmax = src, yx_min.copy(), yx_max.copy()
        if flip:
            image, bmin, bmax = transform.augmentation.flip_horizontally(image, bmin, bmax)
        np.random.seed(200 + seed)
        image, bmin, bmax = transform.resize.label.random_crop(cfg, image, bmin, bmax, h, w)
        assert hashlib.sha256(image.cpu().numpy().tobytes()).digest() == g['c%d_sha' % seed].tobytes(), 'case %d (mirror)' % seed
        assert np.array_equal(bmin.cpu().numpy(), g['c%d_yx_min' % seed])
    # several ragged frames of one target size in ONE launch
    (h, w), group = max(by_size.items(), key=lambda kv: len(kv[1]))
    slots = max(len(c[2]) for c in group)
    pmin = torch.zeros(len(group), slots, 2)
    pmax = torch.zeros(len(group), slots, 2)
    for i, c in enumerate(group):
        pmin[i, :len(c[2])], pmax[i, :len(c[3])] = torch.from_numpy(c[2]), torch.from_numpy(c[3])
    out, a, b = transform.resize_batch([torch.from_numpy(c[1]) for c in group], h, w, bgr2rgb=False, yx_min=pmin, yx_max=pmax,
                                       flip=[c[4] for c in group], crop=[c[5] for c in group], margin=[c[6].tolist() for c in group])
    for i, c in enumerate(group):
        assert hashlib.sha256(out[i].cpu().numpy().tobytes()).digest() == g['c%d_sha' % c[0]].tobytes()
        assert np.array_equal(a[i, :len(c[2])].cpu().numpy(), g['c%d_yx_min' % c[0]])
    # stand-alone flip
    f, a, b = transform.augmentation.flip_horizontally(g['flip_src'], g['flip_min_in'].copy(), g['flip_max_in'].copy())
    assert np.array_equal(f.cpu().numpy(), g['flip_out']) and np.array_equal(a.cpu().numpy(), g['flip_min']) and np.array_equal(b.cpu().numpy(), g['flip_max'])


def test_rotate_and_fixed_bit_exact_vs_executed_reference(golden_dir):
    """cv2.warpAffine's 8-bit bilinear path on the device (yb_warp_affine_u8): the reference's `random_rotate` (Rotator: canvas grown to the
    rotated hull, zero fill, box hulls) with its own `random.uniform` draw, and `transform.resize.image.fixed` when it shrinks -- frames
    bit-identical (SHA-256) to the reference functions executed with cv2 (tests/golden/make_golden_augment.py), boxes equal in float32."""
    import hashlib
    import random
    import configparser
    import transform.augmentation
    import transform.resize.image
    g = np.load(os.path.join(golden_dir, 'augment.npz'))
    cfg = configparser.ConfigParser()
    cfg.read_dict({'augmentation': {'random_rotate': '-7 7'}})
    for seed in g['rot_cases'].tolist():
        h0, w0, h1, w1 = g['r%d_dims' % seed].tolist()
        src = O.synth_frame(20 + seed, h0, w0)
        random.seed(400 + seed)
        image, a, b = transform.augmentation.random_rotate(cfg, src, g['r%d_yx_min_in' % seed].copy(), g['r%d_yx_max_in' % seed].copy())
        assert tuple(image.shape) == (h1, w1, 3), (seed, image.shape)
        assert hashlib.sha256(image.cpu().numpy().tobytes()).digest() == g['r%d_sha' % seed].tobytes(), 'rotation case %d pixels' % seed
        assert np.array_equal(a, g['r%d_yx_min' % seed]) and np.array_equal(b, g['r%d_yx_max' % seed]), 'rotation case %d boxes' % seed
    for seed in range(3):
        h0, w0, h, w = g['f%d_dims' % seed].tolist()
        r = transform.resize.image.fixed(O.synth_frame(30 + seed, h0, w0), h, w)
        assert tuple(r.shape) == (h, w, 3)
        assert hashlib.sha256(r.cpu().numpy().tobytes()).digest() == g['f%d_sha' % seed].tobytes(), 'fixed case %d' % seed
    with pytest.raises(NotImplementedError):
        transform.resize.image.fixed(O.synth_frame(1, 100, 100), 416, 416)


def test_collate_gpu_batch_and_training_step_from_uint8_frames():
    """utils.data.Collate: a list of decoded BGR frames of different sizes + ragged labels -> one GPU batch at the scheduled
    size (frames bit-identical to cv2.resize + BGR2RGB, boxes scaled like transform.resize.label.rescale, labels zero-padded
    like padding_labels), ToTensor on the device bit-identical to torchvision's `.float().div(255)`, and the batch drives a
    training step as is."""
    import model
    import model.yolo2
    import train as yb_train
    import transform
    import utils.data as ud
    samples = []
    for i, (h, w, n) in enumerate(((90, 120, 2), (75, 100, 0), (64, 64, 3))):
        g = np.random.RandomState(60 + i)
        lo = (g.rand(n, 2) * [h * 0.5, w * 0.5]).astype(np.float32)
        samples.append(dict(image=O.synth_frame(30 + i, h, w), yx_min=lo, yx_max=lo + (g.rand(n, 2) * [h * 0.4, w * 0.4] + 4).astype(np.float32),
                            cls=g.randint(0, 20, n)))
    collate = ud.Collate([(64, 64)], maintain=3, seed=1)
    batch = collate(samples)
    assert batch['tensor'].shape == (3, 64, 64, 3) and batch['tensor'].dtype == torch.uint8 and batch['yx_min'].shape == (3, 3, 2)
    for i, smp in enumerate(samples):
        ref_img, a, b = O.rescale_label(smp['image'], smp['yx_min'].copy(), smp['yx_max'].copy(), 64, 64)
        assert np.array_equal(batch['tensor'][i].cpu().numpy(), ref_img[..., ::-1])
        n = len(smp['cls'])
        assert np.array_equal(batch['yx_min'][i, :n].cpu().numpy(), a) and np.array_equal(batch['yx_max'][i, :n].cpu().numpy(), b)
        assert float(batch['yx_max'][i, n:].abs().sum()) == 0.0 and batch['cls'][i, :n].tolist() == smp['cls'].tolist()
    x = transform.to_tensor(batch['tensor'])
    # the reference's ToTensor runs on the CPU (a true IEEE division; torch's CUDA `div` by a scalar multiplies by the reciprocal)
    assert torch.equal(x.cpu(), batch['tensor'].cpu().permute(0, 3, 1, 2).float().div(255))
    cfg = make_config(1)
    cfg.read_dict({'model': {'threshold': '0.6'}, 'hparam': {k: str(v) for k, v in O.HPARAM_DEFAULT.items()}, 'train': {'cross_entropy': '1'}})
    anchors = O.anchors_yolo_voc()
    dnn = model.yolo2.Darknet(model.ConfigChannels(cfg), anchors, 20)
    dnn.load_state_dict(O.make_state_dict(0), strict=False)
    dnn = dnn.to(DEV).train()
    inference = model.Inference(cfg, dnn, anchors).train()
    opt = torch.optim.SGD(dnn.parameters(), 1e-4)
    out = yb_train.iterate(inference, opt, anchors, cfg, batch)
    lt = float(out['loss_total'].item())
    assert lt == lt and 0.0 < lt < 10.0 and (out['height'], out['width']) == (64, 64)


def _detection_sets(cfg, pred):
    """Per image: (set of kept box indices, set of (box index, class) detections) from the batched filter + NMS + expansion kernel."""
    import detect
    fix, res = detect._run(cfg, pred['iou'], pred['yx_min'], pred['yx_max'], detect.get_prob(pred), True, True)
    host = {k: res[k].cpu() for k in ('n_keep', 'keep_box', 'n_det', 'det_keep', 'det_cls')}
    out = []
    for bi in range(pred['iou'].size(0)):
        nk, nd = int(host['n_keep'][bi]), int(host['n_det'][bi])
        kbox = host['keep_box'][bi, :nk].long()
        dbox = kbox[host['det_keep'][bi, :nd].long()]
        out.append((set(kbox.tolist()), set(zip(dbox.tolist(), host['det_cls'][bi, :nd].tolist()))))
    return out


def _jaccard(a, b):
    return len(a & b) / float(max(1, len(a | b)))


@pytest.mark.parametrize('precision', ['strict', 'fast'])
def test_c1_single_image_feature_and_detections_vs_executed_reference(golden_dir, precision):
    """BASELINE configs[0] on the GPU: the 416x416 RGB uint8 network input the reference's transform produced from its own
    image.jpg goes into the model as is (ToTensor fused into the first conv kernel); head feature AND the detections (the
    reference returns 586 for these weights) vs the reference's detect.py chain executed on CPU (tests/golden/make_golden_c1.py)."""
    import detect
    import model
    g = np.load(os.path.join(golden_dir, 'c1_image.npz'))
    dnn = _build_darknet(precision)
    tol = TOL_CONTRACT if precision == 'strict' else TOL_FAST_E2E
    cfg = make_config(1)
    anchors = O.anchors_yolo_voc()
    inference = model.Inference(cfg, dnn, anchors).eval()
    pred = model._inference(inference, torch.from_numpy(g['rgb'])[None].contiguous().to(DEV))
    f = pred['feature']
    e = rel_err(f, torch.from_numpy(g['feature']))
    assert f.shape == g['feature'].shape and e <= tol, 'feature rel err %.3e' % e
    res = detect.postprocess_batch(cfg, pred)[0]
    assert res is not None and not bool(g['none'])
    kept, dets = _detection_sets(cfg, pred)[0]
    ref_dets = set(zip(g['det_box'].tolist(), g['det_cls'].tolist()))
    ref_kept = set(g['det_box'].tolist())
    j_keep, j_det = _jaccard(kept, ref_kept), _jaccard(dets, ref_dets)
    record('c1_%s' % precision, dict(feature=e, detections_ref=len(ref_dets), detections_gpu=len(dets), detections_common=len(dets & ref_dets),
                                     kept_ref=len(ref_kept), kept_gpu=len(kept), kept_common=len(kept & ref_kept)))
    assert len(res[3]) == len(dets)
    # Identity of a detection = (which of the 845 predicted boxes, class).  Every decision (0.005 score threshold, IoU 0.45 suppression)
    # is taken on values that differ from the reference's by the feature error above, so borderline boxes may flip; one flipped NMS
    # decision moves all ~3.5 class-detections of that box.  Strict precision must reproduce >= 97 % of the reference's set.
    if precision == 'strict':
        assert j_keep >= 0.97 and j_det >= 0.97, (j_keep, j_det)
    else:
        assert j_keep >= 0.90 and j_det >= 0.90, (j_keep, j_det)
    # scores of the common detections
    iou, yx_min, yx_max, cls, score = (t.cpu() for t in res)
    ref_score = {(int(b), int(c)): float(sc) for b, c, sc in zip(g['det_box'], g['det_cls'], g['det_score'])}
    ref_box = {int(b): (g['det_yx_min'][i], g['det_yx_max'][i]) for i, b in enumerate(g['det_box'])}
    # box = centre +- exp(feature) * anchor / 2: with random weights boxes reach 1e5 cells, so corners are compared relative to the box extent
    host = detect._run(cfg, pred['iou'], pred['yx_min'], pred['yx_max'], detect.get_prob(pred), True, True)[1]
    kbox = host['keep_box'][0, :int(host['n_keep'][0])].long().cpu()
    dbox = kbox[host['det_keep'][0, :int(host['n_det'][0])].long().cpu()]
    worst = 0.0
    for i, (bx, c) in enumerate(zip(dbox.tolist(), cls.tolist())):
        if (bx, c) in ref_score:
            worst = max(worst, abs(float(score[i]) - ref_score[(bx, c)]) / max(ref_score[(bx, c)], 0.005))
            extent = float(np.abs(ref_box[bx][1] - ref_box[bx][0]).max())
            assert float(np.abs(yx_min[i].numpy() - ref_box[bx][0]).max()) <= 2e-2 * extent + 2e-2, (bx, yx_min[i], ref_box[bx])
    assert worst <= 5e-2, worst


def test_c2_batch32_feature_and_detections_vs_executed_reference(golden_dir):
    """BASELINE configs[1] at its real size: 32 x 3 x 416 x 416 through backbone (strict precision) + decode + filter + NMS + per-class
    expansion in one pipeline call; head features of three images vs the reference's (1e-3), and the detection SETS of all 32
    images vs the ones the reference returns (tests/golden/make_golden_c2.py)."""
    import detect
    import model
    g = np.load(os.path.join(golden_dir, 'c2_batch32.npz'))
    b = int(g['batch'])
    dnn = _build_darknet('strict')
    cfg = make_config(1)
    anchors = O.anchors_yolo_voc()
    inference = model.Inference(cfg, dnn, anchors).eval()
    x = O.synth_images(b, 416, 416, seed=int(g['seed'])).to(DEV)
    pred = model._inference(inference, x)
    f = pred['feature'].cpu()
    errs = []
    for slot, bi in enumerate(g['images']):       # per image: max|d| over the image / max|ref| of that image (stricter than the whole-batch norm)
        errs.append(((f[bi] - torch.from_numpy(g['feature'][slot])).abs().max() / float(g['feature_absmax'][bi])).item())
    assert max(errs) <= TOL_CONTRACT, errs
    np.testing.assert_allclose(f.abs().reshape(b, -1).max(1).values.numpy(), g['feature_absmax'], rtol=2e-3)
    sets = _detection_sets(cfg, pred)
    off_d = np.concatenate([[0], np.cumsum(g['n_det'])])
    tot_ref = tot_got = tot_common = kept_ref = kept_got = kept_common = identical = 0
    for bi, (kept, dets) in enumerate(sets):
        sl = slice(off_d[bi], off_d[bi + 1])
        ref_dets = set(zip(g['det_box'][sl].tolist(), g['det_cls'][sl].tolist()))
        ref_kept = set(g['det_box'][sl].tolist())
        tot_ref += len(ref_dets); tot_got += len(dets); tot_common += len(dets & ref_dets)
        kept_ref += len(ref_kept); kept_got += len(kept); kept_common += len(kept & ref_kept)
        identical += int(dets == ref_dets)
    record('c2_batch32_strict', dict(feature=max(errs), detections_ref=tot_ref, detections_gpu=tot_got, detections_common=tot_common,
                                     kept_ref=kept_ref, kept_gpu=kept_got, kept_common=kept_common, images_identical=identical, images=b))
    assert kept_ref == int(g['n_keep'].sum())
    # (box, class) identities over the whole batch: >= 97 % of the union in common (see the C1 test for why not 100 %)
    assert tot_common >= 0.97 * (tot_ref + tot_got - tot_common), (tot_common, tot_ref, tot_got)
    assert kept_common >= 0.97 * (kept_ref + kept_got - kept_common), (kept_common, kept_ref, kept_got)
