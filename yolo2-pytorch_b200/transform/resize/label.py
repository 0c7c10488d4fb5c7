"""transform.resize.label -- `rescale(image, yx_min, yx_max, height, width)` of the reference
(transform/resize/label.py:25-31): resize the image and scale the boxes by (height / _height, width / _width); and `random_crop`
(:58-75), the default training resize (`resize_train = transform.resize.label.RandomCrop`, config.ini:48)."""
import numpy as np
import torch

import transform as _t


def rescale(image, yx_min, yx_max, height, width):
    """image uint8 [h, w, 3]; yx_min / yx_max float32 [G, 2] in pixels of `image` -> (image', yx_min', yx_max') on the GPU."""
    as_t = lambda a: a if torch.is_tensor(a) else torch.from_numpy(a)
    out, a, b = _t.resize_batch([as_t(image)], height, width, bgr2rgb=False, yx_min=as_t(yx_min)[None], yx_max=as_t(yx_max)[None])
    return out[0], a[0], b[0]


class Rescale(object):
    def __call__(self, data, height, width):
        data['image'], data['yx_min'], data['yx_max'] = rescale(data['image'], data['yx_min'], data['yx_max'], height, width)
        return data


def crop_window(scale, yx_min, yx_max, size, draws):
    """The window `random_crop` cuts (reference transform/resize/label.py:62-72) for four uniform draws: the margins are
    scale * draw * (distance of the box hull to the frame border) in float32; the pixels use their truncation.
    Returns ((y0, x0, y1, x1) ints, float32 margin[2] that is subtracted from the boxes)."""
    yx_min, yx_max = np.asarray(yx_min), np.asarray(yx_max)
    dtype = yx_min.dtype
    hull_min, hull_max = np.min(yx_min, 0), np.max(yx_max, 0)
    size = np.array(size, dtype)
    margin = scale * np.asarray(draws).astype(dtype) * np.concatenate([hull_min, size - hull_max], 0)
    lo, hi = margin[:2], size - margin[2:]
    return tuple(int(v) for v in (lo[0], lo[1], hi[0], hi[1])), lo


def random_crop(config, image, yx_min, yx_max, height, width):
    """Same signature and the same `np.random.rand(4)` draw as the reference; the cut and the resize are ONE launch on the GPU."""
    scale = config.getfloat('augmentation', 'random_crop')
    assert 0 < scale <= 1
    as_np = lambda a: a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    as_t = lambda a: a if torch.is_tensor(a) else torch.from_numpy(a)
    window, margin = crop_window(scale, as_np(yx_min), as_np(yx_max), tuple(image.shape[:2]), np.random.rand(4))
    out, a, b = _t.resize_batch([as_t(image)], height, width, bgr2rgb=False, yx_min=as_t(yx_min)[None], yx_max=as_t(yx_max)[None],
                                crop=[window], margin=[margin.tolist()])
    return out[0], a[0], b[0]


class RandomCrop(object):
    def __init__(self, config):
        self.config = config

    def __call__(self, data, height, width):
        data['image'], data['yx_min'], data['yx_max'] = random_crop(self.config, data['image'], data['yx_min'], data['yx_max'], height, width)
        return data
