"""CPU tests (no GPU): the C-ABI library builds/loads and exports every symbol include/*.h declares;
the host-side mirror of the reference's plugin/config surface behaves like the reference."""
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
