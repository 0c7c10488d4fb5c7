"""train -- the training step of the reference (train.py:57-71,338-362) on the B200 path.

`norm_data`, `ensure_model` and `iterate` keep the reference's meaning:
  iterate: forward (train-mode Darknet: batch-statistics BatchNorm) -> region loss -> hparam-weighted sum ->
           zero_grad -> backward (explicit kernel chain + optional NCCL gradient all-reduce) -> optional clip ->
           optimizer.step().
The optimizer / scheduler objects are torch.optim's, built from the same `train/optimizer` lambda strings as the
reference (train.py:270,368).  TensorBoard summaries, checkpoint timers and the data loader are host glue outside the
hot path.  Multi-GPU is one process per GPU (torch.distributed, NCCL) instead of the reference's nn.DataParallel.
"""
import configparser

import torch
import torch.distributed as dist
import torch.nn as nn

import model
from b200 import ddp as _ddp


def norm_data(data, height, width, rows, cols, keys='yx_min, yx_max'):
    """GT pixel coordinates -> grid units (reference train.py:57-62)."""
    out = {key: data[key] for key in data}
    ref = data[keys.split(', ')[0]]
    scale = torch.tensor([rows / height, cols / width], dtype=torch.float32, device=ref.device).view(1, 1, 2)
    for key in keys.split(', '):
        out[key] = out[key] * scale
    return out


def ensure_model(module):
    """reference train.py:65-71: move to the GPU.  Replication is per process here, so nothing is wrapped; the
    gradient exchange is attached by `iterate` when torch.distributed is initialised."""
    if not torch.cuda.is_available():
        raise RuntimeError('train (B200): a CUDA device is required; there is no CPU fallback')
    return module.cuda()


def build_optimizer(config, params, lr):
    """`train/optimizer` is a Python lambda in the INI, exactly as in the reference (config.ini:72, train.py:270)."""
    return eval(config.get('train', 'optimizer'))(params, lr)


def iterate(inference, optimizer, anchors, config, data, reducer=None):
    """One training step (reference Train.iterate, train.py:338-362).  `data`: dict with `tensor` [B,3,H,W] fp32,
    `yx_min`/`yx_max` [B,G,2] in pixels, `cls` [B,G].  Returns the same kind of dict as the reference."""
    dev = torch.device('cuda', torch.cuda.current_device())
    data = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in data.items()}
    tensor = data['tensor']
    height, width = tensor.shape[-2:]
    dnn = inference.dnn
    if reducer is None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        reducer = _ddp.GradientAllReducer()
    dnn.trainer.on_grad = reducer.on_grad if reducer is not None else None
    pred = model._inference(inference, tensor)
    rows, cols = pred['feature'].shape[-2:]
    cross_entropy = config.getboolean('train', 'cross_entropy') if config.has_option('train', 'cross_entropy') else True
    loss, debug = model.loss(anchors, norm_data(data, height, width, rows, cols), pred, config.getfloat('model', 'threshold'), cross_entropy)
    loss_hparam = {key: loss[key] * config.getfloat('hparam', key) for key in loss}
    loss_total = sum(loss_hparam.values())
    optimizer.zero_grad()
    loss_total.backward()
    if reducer is not None:
        reducer.finish()
    try:
        clip = config.getfloat('train', 'clip')
        nn.utils.clip_grad_norm_(inference.parameters(), clip)
    except (configparser.NoOptionError, configparser.NoSectionError):
        pass
    optimizer.step()
    return dict(height=height, width=width, rows=rows, cols=cols, data=data, pred=pred, debug=debug, loss_total=loss_total, loss=loss,
                loss_hparam=loss_hparam)
