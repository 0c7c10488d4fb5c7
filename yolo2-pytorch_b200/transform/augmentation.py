"""transform.augmentation -- the geometric augmentations of the reference that are pure index transforms, on the GPU
(reference transform/augmentation.py:87-116): `flip_horizontally`, `random_flip_horizontally`, `RandomFlipHorizontally`.

The pixels never move on their own: the flip is applied to the SOURCE indexing of the one resize launch that follows
(`transform.resize_batch(..., flip=...)`, yb_resize_aug_batch_u8), which is bit-identical to cv2.flip followed by cv2.resize.  Calling
`flip_horizontally` stand-alone returns the flipped frame (a same-size resize is the identity) and the transformed boxes.
`random_rotate` (cv2.warpAffine with a random fill colour, :61-76) is not part of this build."""
import random

import torch

import transform as _t


def flip_horizontally(image, yx_min, yx_max):
    """image uint8 [h, w, 3]; yx_min / yx_max float32 [G, 2] -> (flipped image, boxes) on the GPU; x' = w - x with min / max swapped."""
    as_t = lambda a: a if torch.is_tensor(a) else torch.from_numpy(a)
    image = as_t(image)
    out, a, b = _t.resize_batch([image], int(image.shape[0]), int(image.shape[1]), bgr2rgb=False, yx_min=as_t(yx_min)[None], yx_max=as_t(yx_max)[None],
                                flip=[True])
    return out[0], a[0], b[0]


def random_flip_horizontally(config, image, yx_min, yx_max):
    """Flips when `random.random() > augmentation/random_flip_horizontally`, exactly the reference's draw (:98-103)."""
    if random.random() > config.getfloat('augmentation', 'random_flip_horizontally'):
        return flip_horizontally(image, yx_min, yx_max)
    return image, yx_min, yx_max


class RandomFlipHorizontally(object):
    def __init__(self, config):
        self.config = config

    def __call__(self, data):
        data['image'], data['yx_min'], data['yx_max'] = random_flip_horizontally(self.config, data['image'], data['yx_min'], data['yx_max'])
        return data
