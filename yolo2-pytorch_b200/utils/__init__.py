"""utils -- config / plugin-resolution helpers kept call-compatible with the reference
(`utils/__init__.py`: get_anchors :78-81, parse_attr :84-87, load_config :90-94, modify_config
:97-106, ensure_device :109-112, get_category :72-75, get_model_dir :59-63).  Host-only code."""
import configparser
import importlib
import os

import numpy as np
import torch


def _expand(path):
    return os.path.expanduser(os.path.expandvars(path))


def get_model_dir(config):
    return os.path.join(_expand(config.get('config', 'root')), config.get('model', 'name'), config.get('model', 'dnn'))


def get_category(config, cache_dir=None):
    path = _expand(config.get('cache', 'category')) if cache_dir is None else os.path.join(cache_dir, 'category')
    with open(path, 'r') as f:
        return [line.strip() for line in f]


def get_anchors(config, dtype=np.float32):
    """Anchors TSV (header `width<TAB>height`, grid units) -> ndarray [A,2] in (height, width) order."""
    path = _expand(config.get('model', 'anchors'))
    with open(path, 'r') as f:
        header = f.readline().split()
        rows = [line.split() for line in f if line.strip()]
    cols = {name: i for i, name in enumerate(header)}
    return np.array([[r[cols['height']], r[cols['width']]] for r in rows], dtype=dtype)


def parse_attr(s):
    """'model.yolo2.Darknet' -> the class object (dotted-path plugin resolution)."""
    module_name, attr = s.rsplit('.', 1)
    return getattr(importlib.import_module(module_name), attr)


def load_config(config, paths):
    for path in paths:
        path = _expand(path)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        config.read(path)


def modify_config(config, cmd):
    """`section/option=value`; an empty value removes the option."""
    var, value = cmd.split('=', 1)
    section, option = var.split('/')
    if value:
        config.set(section, option, value)
    else:
        try:
            config.remove_option(section, option)
        except (configparser.NoSectionError, configparser.NoOptionError):
            pass


def ensure_device(t, device_id=None, non_blocking=False):
    """Move to the GPU when one is present (the reference's parameter was named `async`, which is
    no longer a legal identifier)."""
    if torch.cuda.is_available():
        t = t.cuda(device_id, non_blocking=non_blocking)
    return t


def abs_mean(data, dtype=np.float32):
    assert isinstance(data, np.ndarray), type(data)
    return np.sum(np.abs(data)) / dtype(data.size)
