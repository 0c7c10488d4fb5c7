"""transform.resize.label -- `rescale(image, yx_min, yx_max, height, width)` of the reference
(transform/resize/label.py:25-31): resize the image and scale the boxes by (height / _height, width / _width)."""
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
