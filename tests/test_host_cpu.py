"""CPU tests (no GPU): the C-ABI library builds/loads and exports every symbol include/*.h declares;
the host-side mirror of the reference's plugin/config surface behaves like the reference."""
import collections
import configparser
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'yolo2-pytorch_b200')


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'yolo2_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(yb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from b200 import lib
    handle = lib.load()
    names = declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(handle, name), name
    assert set(lib.SIGNATURES) | {'yb_last_error'} == set(names)
    assert handle.yb_version() >= 100
    assert isinstance(lib.last_error(), str)


def test_no_cpu_fallback():
    from b200 import ops
    import utils.postprocess
    with pytest.raises(RuntimeError):
        ops.decode(torch.zeros(1, 125, 13, 13), torch.ones(5, 2), 20)
    with pytest.raises(RuntimeError):
        utils.postprocess.nms(torch.rand(4), torch.zeros(4, 2), torch.ones(4, 2))
    assert utils.postprocess.nms(torch.zeros(0), torch.zeros(0, 2), torch.zeros(0, 2)) == []   # utils/postprocess.py:35-36


def load_config():
    import utils
    config = configparser.ConfigParser()
    cwd = os.getcwd()
    os.chdir(PKG)
    try:
        utils.load_config(config, ['config.ini', 'config/darknet/yolo-voc.ini'])
    finally:
        os.chdir(cwd)
    return config


def test_config_overlay_and_anchors():
    import utils
    config = load_config()
    assert config.get('model', 'dnn') == 'model.yolo2.Darknet' and config.getboolean('detect', 'fix')
    utils.modify_config(config, 'detect/overlap=0.5')
    assert config.getfloat('detect', 'overlap') == 0.5
    utils.modify_config(config, 'detect/overlap=')
    assert not config.has_option('detect', 'overlap')
    cwd = os.getcwd()
    os.chdir(PKG)
    try:
        anchors = utils.get_anchors(config)
        category = utils.get_category(config)
    finally:
        os.chdir(cwd)
    # TSV columns are width,height; the array is (height, width)  (utils/__init__.py:78-81)
    np.testing.assert_allclose(anchors[0], [1.73145, 1.3221], rtol=1e-6)
    assert anchors.shape == (5, 2) and anchors.dtype == np.float32 and len(category) == 20


def test_plugin_surface_state_dict_and_channels():
    import model
    import model.yolo2
    import utils
    from oracle import yolo2_oracle as O
    config = load_config()
    cls = utils.parse_attr(config.get('model', 'dnn'))
    assert cls is model.yolo2.Darknet
    dnn = cls(model.ConfigChannels(config), O.anchors_yolo_voc(), 20)
    sd = dnn.state_dict()
    ref = O.make_state_dict(0)
    assert set(ref) == {k for k in sd if not k.endswith('num_batches_tracked')}
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    assert sum(p.numel() for p in dnn.parameters()) == 50655389
    res = dnn.load_state_dict(ref, strict=False)             # 0.3.1-era checkpoints lack num_batches_tracked
    assert not res.unexpected_keys and not res.missing_keys
    assert dnn.scope('layers2.3.bn.weight') == 'layers2.3'
    idx = dnn.get_mapper(94)(torch.tensor([1, 5]), 64)
    assert idx.tolist() == [1, 5, 65, 69, 129, 133, 193, 197]
    assert model.output_channels(5, 20) == 125 and model.output_channels(5, 1) == 25
    assert model.meshgrid(2, 3).tolist() == [[0, 0], [0, 1], [1, 0], [1, 1], [2, 0], [2, 1]]   # reference quirk, rows != cols
    # channel-pruned checkpoint: ConfigChannels takes widths from the state_dict (model/__init__.py:35-43)
    pruned = dict(ref)
    pruned['layers1.4.conv.weight'] = ref['layers1.4.conv.weight'][:96]
    for n in ('weight', 'bias', 'running_mean', 'running_var'):
        pruned['layers1.4.bn.' + n] = ref['layers1.4.bn.' + n][:96]
    pruned['layers1.5.conv.weight'] = ref['layers1.5.conv.weight'][:, :96]
    dnn2 = cls(model.ConfigChannels(config, pruned), O.anchors_yolo_voc(), 20)
    assert dnn2.layers1[4].conv.weight.shape[0] == 96 and dnn2.layers1[5].conv.weight.shape[1] == 96
    with pytest.raises(RuntimeError):
        dnn.train()(torch.zeros(1, 3, 32, 32))     # CPU tensor: no fallback in train mode either


def test_mobilenet_plugin_surface():
    import model
    import model.mobilenet
    from oracle import yolo2_oracle as O
    config = load_config()
    net = model.mobilenet.MobileNet(model.ConfigChannels(config), O.anchors_yolo_voc(), 20)
    ref = O.make_mobilenet_state_dict(0)
    sd = net.state_dict()
    assert set(ref) == {k for k in sd if not k.endswith('num_batches_tracked')}
    assert all(tuple(sd[k].shape) == tuple(ref[k].shape) for k in ref)
    assert sum(p.numel() for p in net.parameters()) == 3335101
    assert sd['layers.14.bias'].shape == (125,) and sd['layers.3.dw.conv.weight'].shape == (128, 1, 3, 3)


def test_pack_unit_table_matches_the_header_struct():
    """b200.train_engine.PackPlan builds the device table of yb_pack_weights_batch with numpy: field order, sizes and the 48-byte stride
    must be those of `yb_pack_unit` in include/yolo2_b200.h (3 pointers + 6 ints on LP64)."""
    import re
    from b200.train_engine import PackPlan
    header = open(os.path.join(ROOT, 'include', 'yolo2_b200.h')).read()
    body = re.search(r'typedef struct yb_pack_unit \{(.*?)\} yb_pack_unit;', header, re.S).group(1)
    fields = []
    for decl in body.split(';'):
        decl = decl.strip()
        if not decl:
            continue
        ctype, names = decl.rsplit(' ', 1) if ',' not in decl else (decl.split(' ', 1)[0], decl.split(' ', 1)[1])
        for name in names.split(','):
            name = name.strip()
            fields.append(('ptr' if '*' in ctype or '*' in name else ctype.strip(), name.lstrip('*')))
    assert [k for k, _ in fields] == ['ptr'] * 3 + ['int'] * 6, fields
    dt = PackPlan.DTYPE
    assert dt.itemsize == 48 and len(dt.names) == len(fields)
    assert [dt.fields[n][1] for n in dt.names] == [0, 8, 16, 24, 28, 32, 36, 40, 44]
    assert [dt.fields[n][0].itemsize for n in dt.names] == [8, 8, 8, 4, 4, 4, 4, 4, 4]
    # logical order: weight, forward operand, data-gradient operand, cout, cin, ksize, cout_pad, block0, ci_blocks
    assert [n for _, n in fields] == ['w_oihw', 'out_fwd', 'out_dgrad', 'cout', 'cin', 'ksize', 'cout_pad', 'block0', 'ci_blocks']
    assert dt.names == ('w', 'f', 'd', 'cout', 'cin', 'k', 'cp', 'b0', 'cib')


def test_checkpoint_directory_round_trip_and_torch031_compat(tmp_path):
    """utils.train: the reference's model-directory format (`<step>.pth` = OrderedDict of CPU tensors + `<step>.epoch`, utils/train.py:51-126),
    Saver's keep-N tidy, load_model's latest-step rule, and loading a torch-0.3.1-style state_dict (no num_batches_tracked) into the plugin."""
    import model
    import model.yolo2
    import utils.train
    from oracle import yolo2_oracle as O
    config = load_config()
    dnn = model.yolo2.Darknet(model.ConfigChannels(config), O.anchors_yolo_voc(), 20)
    saver = utils.train.Saver(str(tmp_path), keep=2, logger=None)
    for step, epoch in ((10, 0), (200, 1), (3000, None)):
        saver(utils.train.state_dict_cpu(dnn), step, epoch)
    assert sorted(os.listdir(str(tmp_path))) == ['200.epoch', '200.pth', '3000.pth']
    path, step, epoch = utils.train.load_model(str(tmp_path), logger=None)
    assert (os.path.basename(path), step, epoch) == ('3000.pth', 3000, None)
    assert utils.train.load_model(str(tmp_path), 200, logger=None)[1:] == (200, 1)
    sd, step, epoch = utils.train.load_checkpoint(str(tmp_path), logger=None)
    assert step == 3000 and set(sd) == set(dnn.state_dict())
    # channel-pruned / checkpoint-shaped construction as the reference does it (detect.py:95): ConfigChannels(config, state_dict)
    old = collections.OrderedDict((k, v) for k, v in O.make_state_dict(0).items())           # torch 0.3.1 style: no num_batches_tracked
    assert not any(k.endswith('num_batches_tracked') for k in old)
    dnn2 = model.yolo2.Darknet(model.ConfigChannels(config, old), O.anchors_yolo_voc(), 20)
    utils.train.load_state_dict(dnn2, old)
    assert torch.equal(dnn2.state_dict()['layers2.3.conv.weight'], old['layers2.3.conv.weight'])
    bad = collections.OrderedDict(old)
    bad.pop('layers1.4.bn.running_var')
    with pytest.raises(RuntimeError):
        utils.train.load_state_dict(dnn2, bad)
    assert utils.train.load_sizes(config)[:1] and all(len(hw) == 2 for hw in utils.train.load_sizes(config))
    t = utils.train.Timer(3600, first=True)
    assert t() is True and t() is False and utils.train.Timer(3600, first=False)() is False


def test_resnet_plugin_surface():
    """model.resnet: constructors, torchvision-style state_dict keys, `scope` (reference model/resnet.py:144-158), loud failures."""
    import model
    import model.resnet
    import utils
    from oracle import yolo2_oracle as O
    config = load_config()
    for name, params in (('resnet18', None), ('resnet50', None)):
        net = utils.parse_attr('model.resnet.' + name)(model.ConfigChannels(config), O.anchors_yolo_voc(), 20)
        ref = O.make_resnet_state_dict(name, 0)
        sd = net.state_dict()
        assert set(ref) == {k for k in sd if not k.endswith('num_batches_tracked')}, name
        assert all(tuple(sd[k].shape) == tuple(ref[k].shape) for k in ref)
    net = model.resnet.resnet18(model.ConfigChannels(config), O.anchors_yolo_voc(), 20)
    assert net.scope('layer1.0.conv1.weight') == 'layer1.0.1' and net.scope('layer2.0.bn2.bias') == 'layer2.0.2'
    assert net.scope('layer2.0.downsample.0.weight') == 'layer2.0.downsample' and net.scope('conv.weight') == 'conv'
    assert net.scope('conv1.weight') == '1' and net.scope('bn1.running_mean') == '1'
    assert net.conv.weight.shape == (125, 512, 1, 1) and net.layer2[0].conv1.stride == (2, 2) and net.layer2[0].downsample is not None
    with pytest.raises(RuntimeError):
        net.eval()(torch.zeros(1, 3, 32, 32))       # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        net.train()(torch.zeros(1, 3, 32, 32))


# ------------------------------------------------------------------------------------------------
# Darknet `.weights` importer (reference convert_darknet_torch.py:37-57,93-113)
# ------------------------------------------------------------------------------------------------
def test_darknet_head_permutation_matches_reference_golden():
    """Head rows (x, y, w, h, obj, cls...) -> (obj, y, x, h, w, cls...): fixtures made by executing the reference's
    transpose_weight / transpose_bias (tests/golden/make_golden_weights.py)."""
    from utils import darknet_weights as dw
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'darknet_weights.npz'))
    for tag in ('voc', 'coco', 'one'):
        a = int(g['anchors_' + tag])
        w, b = g['w_in_' + tag], g['b_in_' + tag]
        perm = dw.head_permutation(a, w.shape[0] // a)
        assert np.array_equal(w[perm], g['w_out_' + tag]) and np.array_equal(b[perm], g['b_out_' + tag])
        inv = dw.head_permutation(a, w.shape[0] // a, inverse=True)
        assert np.array_equal(g['w_out_' + tag][inv], w)


def _tiny_template():
    """A 3-unit stand-in with the real key structure: BN unit, BN unit, plain head (A=2, C=3)."""
    import collections
    g = torch.Generator().manual_seed(3)
    sd = collections.OrderedDict()
    for name, cin, cout, k in (('layers1.0', 3, 4, 3), ('passthrough', 4, 6, 1)):
        sd[name + '.conv.weight'] = torch.randn(cout, cin, k, k, generator=g)
        sd[name + '.bn.weight'] = torch.rand(cout, generator=g)
        sd[name + '.bn.bias'] = torch.randn(cout, generator=g)
        sd[name + '.bn.running_mean'] = torch.randn(cout, generator=g)
        sd[name + '.bn.running_var'] = torch.rand(cout, generator=g)
        sd[name + '.bn.num_batches_tracked'] = torch.tensor(7)
    sd['layers3.1.conv.weight'] = torch.randn(16, 6, 1, 1, generator=g)
    sd['layers3.1.conv.bias'] = torch.randn(16, generator=g)
    return sd


def test_darknet_weights_file_layout_and_round_trip(tmp_path):
    """Byte layout written by hand exactly as Darknet stores it (header; per unit beta, gamma, mean, var, weight --
    or bias, weight -- float32 LE; head in Darknet's channel order), then read back; save -> load is the identity."""
    import struct
    from utils import darknet_weights as dw
    sd = _tiny_template()
    a, per = 2, 8
    inv = dw.head_permutation(a, per, inverse=True)
    blob = struct.pack('<4i', 0, 2, 0, 12345)
    for name in ('layers1.0', 'passthrough'):
        for suffix in ('bn.bias', 'bn.weight', 'bn.running_mean', 'bn.running_var', 'conv.weight'):
            blob += sd[name + '.' + suffix].numpy().astype('<f4').tobytes()
    blob += sd['layers3.1.conv.bias'].numpy()[inv].astype('<f4').tobytes()
    blob += sd['layers3.1.conv.weight'].numpy()[inv].astype('<f4').tobytes()
    path = str(tmp_path / 'tiny.weights')
    with open(path, 'wb') as f:
        f.write(blob + b'\0' * 8)                                        # 8 trailing bytes -> reported as remaining
    template = {k: torch.zeros_like(v) for k, v in sd.items()}
    out, info = dw.load_darknet_weights(path, template, a)
    assert (info['major'], info['minor'], info['seen'], info['remaining']) == (0, 2, 12345, 8)
    assert list(out) [:5] == ['layers1.0.bn.bias', 'layers1.0.bn.weight', 'layers1.0.bn.running_mean', 'layers1.0.bn.running_var', 'layers1.0.conv.weight']
    for k, v in sd.items():
        if k.endswith('num_batches_tracked'):
            assert int(out[k]) == 0                                      # not in the file: the template's value is kept
        else:
            assert torch.equal(out[k], v), k
    path2 = str(tmp_path / 'again.weights')
    dw.save_darknet_weights(path2, sd, a, header=dict(major=0, minor=2, revision=0, seen=12345))
    assert open(path2, 'rb').read() == blob
    with open(path2, 'r+b') as f:
        f.truncate(len(blob) - 4)
    with pytest.raises(ValueError):
        dw.load_darknet_weights(path2, template, a)


def test_darknet_weights_full_model_round_trip(tmp_path):
    """The real Darknet-19 key set: 50,655,389 parameters + running statistics survive save -> load_into."""
    import model
    import model.yolo2
    from oracle import yolo2_oracle as O
    from utils import darknet_weights as dw
    config = configparser.ConfigParser()
    config.read_dict({'batch_norm': {'enable': '1'}})
    anchors = O.anchors_yolo_voc()
    dnn = model.yolo2.Darknet(model.ConfigChannels(config), anchors, 20)
    sd = O.make_state_dict(0)
    path = str(tmp_path / 'd19.weights')
    dw.save_darknet_weights(path, sd, len(anchors))
    n_float = sum(v.numel() for k, v in sd.items() if not k.endswith('num_batches_tracked'))
    assert os.path.getsize(path) == 16 + 4 * n_float
    info = dw.load_into(dnn, path, len(anchors))
    assert info['remaining'] == 0 and info['assigned'] == n_float
    got = dnn.state_dict()
    for k, v in sd.items():
        if not k.endswith('num_batches_tracked'):
            assert torch.equal(got[k], v), k


def test_eval_ap_host_functions_match_reference_golden():
    """eval.voc_ap / average_precision / merge_ap (host numpy, reference eval.py:78-121,296-303) against the golden APs."""
    import eval as yb_eval
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'eval.npz'))
    for metric07 in (0, 1):
        config = configparser.ConfigParser()
        config.read_dict({'eval': {'metric07': str(metric07)}})
        for c in range(20):
            ap = yb_eval.average_precision(config, g['sorted_tp_cls%d' % c], int(g['num_cls%d' % c]))
            assert abs(ap - float(g['ap%d_cls%d' % (metric07, c)])) <= 1e-12
    config.read_dict({'eval': {'metric07': '0'}})
    merged = yb_eval.merge_ap(config, [2, 0], [np.array([0.9, 0.1, 0.5]), np.array([0.3])], [np.array([True, False, True]), np.array([False])])
    assert list(merged) == [0] and abs(merged[0] - 1.0) <= 1e-12          # both ground truths found by the two best-scored detections
    with pytest.raises(RuntimeError):
        yb_eval.matching_batch(torch.zeros(1, 2), torch.ones(1, 2), torch.zeros(1), torch.tensor([0, 1]), torch.zeros(1, 2), torch.ones(1, 2),
                               torch.zeros(1), torch.tensor([0, 1]), 1, 0.5)


def test_convert_darknet_torch_cli_round_trip(tmp_path):
    """The converter CLI (reference convert_darknet_torch.py:83-125): .weights -> .pth -> .weights is the identity for
    the network the config describes (model.yolo2.Darknet, yolo-voc anchors, 20 classes)."""
    import subprocess
    import sys
    from oracle import yolo2_oracle as O
    from utils import darknet_weights as dw
    sd = O.make_state_dict(3)
    w0 = str(tmp_path / 'a.weights')
    dw.save_darknet_weights(w0, sd, 5, header=dict(major=0, minor=2, revision=0, seen=777))
    cli = os.path.join(PKG, 'convert_darknet_torch.py')
    env = dict(os.environ, PYTHONPATH=PKG)
    cfg = ['-c', 'config.ini', 'config/darknet/yolo-voc.ini']
    pth = str(tmp_path / 'a.pth')
    subprocess.run([sys.executable, cli, w0, pth] + cfg, check=True, env=env)
    got = torch.load(pth, map_location='cpu')
    for k, v in sd.items():
        assert torch.equal(got[k], v), k
    w1 = str(tmp_path / 'b.weights')
    subprocess.run([sys.executable, cli, pth, w1, '--reverse'] + cfg, check=True, env=env)
    assert open(w1, 'rb').read()[16:] == open(w0, 'rb').read()[16:]


def test_padding_labels_and_size_schedule():
    """utils.data: padding_labels (reference utils/data.py:28-41), load_sizes (utils/train.py:129-131) and the multi-scale
    schedule (Collate.next_size, utils/data.py:135-141): a size is kept for `maintain` further batches; the same seed
    gives every rank the same sequence."""
    import utils.data as ud
    d = ud.padding_labels(dict(yx_min=np.ones((2, 2), np.float32), yx_max=np.ones((2, 2), np.float32), cls=np.array([3, 4]), difficult=np.array([0, 1])), 5)
    assert d['yx_min'].shape == (5, 2) and d['cls'].tolist() == [3, 4, 0, 0, 0] and d['difficult'].tolist() == [0, 1, 0, 0, 0]
    assert float(d['yx_max'][2:].sum()) == 0.0
    config = load_config()
    sizes = ud.load_sizes(config)
    assert (320, 320) in sizes and (416, 416) in sizes and (608, 608) in sizes and all(h % 32 == 0 and w % 32 == 0 for h, w in sizes)
    a, b = ud.SizeSchedule(sizes, maintain=10, seed=5), ud.SizeSchedule(sizes, maintain=10, seed=5)
    seq = [a.next_size() for _ in range(45)]
    assert seq == [b.next_size() for _ in range(45)]                      # ranks stay in step
    runs = [seq[i:i + 11] for i in range(0, 44, 11)]
    assert all(len(set(r)) == 1 for r in runs) and len(set(seq)) > 1     # held for maintain + 1 batches, then redrawn
    # reference semantics, restated: _maintain counts from 0 after a draw
    import random
    rng, ref, m, cur = random.Random(5), [], 10, None
    cnt = m
    for _ in range(45):
        if cnt < m:
            cnt += 1
        else:
            cur, cnt = rng.choice(sizes), 0
        ref.append(cur)
    assert seq == ref


def test_header_is_plain_c(tmp_path):
    """include/yolo2_b200.h is a C header (no torch / C++ types in any signature): it compiles with a C99 compiler in
    pedantic mode and as C++."""
    import shutil
    import subprocess
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc')
    src = tmp_path / 'hdr.c'
    src.write_text('#include "yolo2_b200.h"\nint main(void) { return yb_version() < 0; }\n')
    inc = os.path.join(ROOT, 'include')
    subprocess.run([gcc, '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-fsyntax-only', '-I', inc, str(src)], check=True)
    gxx = shutil.which('g++')
    if gxx is not None:
        subprocess.run([gxx, '-std=c++11', '-fsyntax-only', '-x', 'c++', '-I', inc, str(src)], check=True)


def test_save_darknet_weights_accepts_partial_header(tmp_path):
    """`{'seen': N}` (the natural way to carry the image counter over) must not collide with the defaults."""
    import struct
    from utils import darknet_weights as dw
    sd = {'layers.0.conv.weight': torch.randn(8, 3, 3, 3), 'layers.0.bn.weight': torch.rand(8), 'layers.0.bn.bias': torch.randn(8),
          'layers.0.bn.running_mean': torch.randn(8), 'layers.0.bn.running_var': torch.rand(8) + 0.5,
          'layers.1.conv.weight': torch.randn(10, 8, 1, 1), 'layers.1.conv.bias': torch.randn(10)}
    path = str(tmp_path / 'x.weights')
    dw.save_darknet_weights(path, sd, 2, header={'seen': 12345})
    assert struct.unpack('<4i', open(path, 'rb').read(16)) == (0, 1, 0, 12345)
    dw.save_darknet_weights(path, sd, 2)
    assert struct.unpack('<4i', open(path, 'rb').read(16)) == (0, 1, 0, 0)
    dw.save_darknet_weights(path, sd, 2, header=dict(major=0, minor=2, revision=0, seen=7))
    assert struct.unpack('<4i', open(path, 'rb').read(16)) == (0, 2, 0, 7)
