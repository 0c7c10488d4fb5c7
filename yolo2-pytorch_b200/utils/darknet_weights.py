"""utils.darknet_weights -- Darknet `.weights` files <-> the plugin's state_dict.

The reference ships this as a CLI (`convert_darknet_torch.py:37-57,83-113`); here it is a pair of functions so that
pretrained Darknet-19 / YOLOv2 weights can be loaded straight into `model.yolo2.Darknet` (and written back).

File format (little endian): int32 major, minor, revision, seen; then, for every conv unit in network order
(`layers1.*`, `layers2.*`, `passthrough`, `layers3.*` -- the order of `state_dict()`):

    with BatchNorm:   bn.bias (beta), bn.weight (gamma), bn.running_mean, bn.running_var, conv.weight [Cout,Cin,k,k]
    without:          conv.bias, conv.weight

all float32.  The detection head (last unit) stores its A*(5+C) output channels per anchor as
(x, y, w, h, objectness, classes...) while this code base decodes (objectness, y, x, h, w, classes...)
(model/__init__.py:123-134 of the reference), so the head's weight and bias rows are permuted on the way in / out.
Pure host code (numpy); no reference source is used at run time.
"""
import collections
import struct

import numpy as np
import torch

_ORDER_BN = ('bn.bias', 'bn.weight', 'bn.running_mean', 'bn.running_var', 'conv.weight')
_ORDER_PLAIN = ('conv.bias', 'conv.weight')


def head_permutation(num_anchors, per_anchor, inverse=False):
    """Row index map of the head: ours[i] = darknet[perm[i]] (inverse: darknet[i] = ours[perm[i]]).
    Darknet per-anchor order (x, y, w, h, obj, cls...) -> (obj, y, x, h, w, cls...)."""
    one = [4, 1, 0, 3, 2] + list(range(5, per_anchor))
    perm = np.concatenate([np.asarray(one) + a * per_anchor for a in range(num_anchors)])
    if inverse:
        inv = np.empty_like(perm)
        inv[perm] = np.arange(perm.size)
        return inv
    return perm


def _units(state_dict):
    """state_dict keys grouped per conv unit, in first-appearance order: {unit: {suffix: key}}."""
    units = collections.OrderedDict()
    for key in state_dict:
        unit, s1, s2 = key.rsplit('.', 2)
        units.setdefault(unit, {})[s1 + '.' + s2] = key
    return units


def read_header(f):
    raw = f.read(16)
    if len(raw) != 16:
        raise ValueError('darknet weights: file shorter than its 16-byte header')
    major, minor, revision, seen = struct.unpack('<4i', raw)
    return dict(major=major, minor=minor, revision=revision, seen=seen)


def load_darknet_weights(path, template_state_dict, num_anchors):
    """Read `path` into a new state_dict shaped like `template_state_dict` (e.g. `dnn.state_dict()`; only shapes and
    key order are used).  Returns (state_dict, info) with info = header fields + `assigned` (floats read) and
    `remaining` (unread bytes, 0 for a matching architecture).  Raises ValueError if the file is too short."""
    units = _units(template_state_dict)
    out = collections.OrderedDict()
    with open(path, 'rb') as f:
        info = read_header(f)
        data = np.fromfile(f, dtype='<f4')
    pos = 0
    last = None
    for unit, group in units.items():
        order = _ORDER_BN if 'bn.weight' in group else _ORDER_PLAIN
        for suffix in order:
            if suffix not in group:
                continue
            key = group[suffix]
            shape = tuple(template_state_dict[key].shape)
            n = int(np.prod(shape))
            if pos + n > data.size:
                raise ValueError('darknet weights: %s needs %d floats but only %d are left (wrong architecture?)' % (key, n, data.size - pos))
            out[key] = torch.from_numpy(data[pos:pos + n].astype(np.float32).reshape(shape).copy())
            pos += n
        last = unit
    # buffers the file does not carry (torch >= 0.4 `num_batches_tracked`) keep the template's value
    for key, v in template_state_dict.items():
        if key not in out and key.endswith('num_batches_tracked'):
            out[key] = v.detach().clone()
    if last is not None:
        wkey, bkey = last + '.conv.weight', last + '.conv.bias'
        rows = out[wkey].shape[0]
        if rows % num_anchors:
            raise ValueError('head has %d channels, not a multiple of %d anchors' % (rows, num_anchors))
        perm = torch.from_numpy(head_permutation(num_anchors, rows // num_anchors))
        out[wkey] = out[wkey][perm].contiguous()
        if bkey in out:
            out[bkey] = out[bkey][perm].contiguous()
    info['assigned'] = pos
    info['remaining'] = int((data.size - pos) * 4)
    return out, info


def save_darknet_weights(path, state_dict, num_anchors, header=None):
    """Inverse of `load_darknet_weights`: write `state_dict` (ours) as a Darknet `.weights` file."""
    hdr = dict(major=0, minor=1, revision=0, seen=0)
    hdr.update(header or {})          # partial headers (e.g. {'seen': N}) keep the defaults for the rest
    header = hdr
    units = _units({k: v for k, v in state_dict.items() if not k.endswith('num_batches_tracked')})
    names = list(units)
    with open(path, 'wb') as f:
        f.write(struct.pack('<4i', int(header['major']), int(header['minor']), int(header['revision']), int(header['seen'])))
        for unit in names:
            group = units[unit]
            order = _ORDER_BN if 'bn.weight' in group else _ORDER_PLAIN
            for suffix in order:
                if suffix not in group:
                    continue
                t = state_dict[group[suffix]].detach().float().cpu()
                if unit == names[-1] and suffix in ('conv.weight', 'conv.bias'):
                    rows = t.shape[0]
                    inv = torch.from_numpy(head_permutation(num_anchors, rows // num_anchors, inverse=True))
                    t = t[inv]
                f.write(np.ascontiguousarray(t.numpy(), dtype='<f4').tobytes())


def load_into(dnn, path, num_anchors):
    """Convenience: read `path` and load it into the module; returns the info dict.  Checkpoints written by
    torch 0.3.1 (no `num_batches_tracked`) and `.weights` files both go through `strict=False` for that one buffer."""
    sd, info = load_darknet_weights(path, dnn.state_dict(), num_anchors)
    missing = dnn.load_state_dict(sd, strict=False)
    bad = [k for k in missing.missing_keys if not k.endswith('num_batches_tracked')] + list(missing.unexpected_keys)
    if bad:
        raise KeyError('darknet weights do not match the module: %s' % bad)
    return info
