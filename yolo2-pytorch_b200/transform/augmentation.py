"""transform.augmentation -- the geometric augmentations of the reference that are pure index transforms, on the GPU
(reference transform/augmentation.py:87-116): `flip_horizontally`, `random_flip_horizontally`, `RandomFlipHorizontally`.

The pixels never move on their own: the flip is applied to the SOURCE indexing of the one resize launch that follows
(`transform.resize_batch(..., flip=...)`, yb_resize_aug_batch_u8), which is bit-identical to cv2.flip followed by cv2.resize.  Calling
`flip_horizontally` stand-alone returns the flipped frame (a same-size resize is the identity) and the transformed boxes.
`Rotator` / `random_rotate` (:28-76) run cv2.warpAffine's 8-bit bilinear arithmetic in one kernel (yb_warp_affine_u8, bit-exact); the
2x3 matrix and the rotated box hulls are the reference's own small host-side numpy arithmetic."""
import ctypes
import math
import random

import numpy as np
import torch

import transform as _t
from b200 import lib as _lib
from b200 import ops as _ops


def warp_affine(image, matrix, size, fill=(0, 0, 0)):
    """cv2.warpAffine(image, matrix, size=(width, height), flags=INTER_LINEAR, borderMode=BORDER_CONSTANT, borderValue=fill) for a uint8
    [h, w, 3] frame, on the GPU.  The matrix is inverted on the host exactly the way OpenCV does (imgwarp.cpp), in double."""
    image = image if torch.is_tensor(image) else torch.from_numpy(np.ascontiguousarray(image))
    if image.dtype != torch.uint8 or image.dim() != 3 or image.shape[2] != 3:
        raise ValueError('warp_affine expects a uint8 [h, w, 3] frame')
    dev = torch.device('cuda', torch.cuda.current_device())
    src = image.to(dev).contiguous()
    m = np.array(matrix, np.float64).reshape(2, 3).copy()
    d = m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[1, 1] * d, m[0, 0] * d
    m[0, 0] = a11; m[0, 1] *= -d; m[1, 0] *= -d; m[1, 1] = a22
    b1 = -m[0, 0] * m[0, 2] - m[0, 1] * m[1, 2]
    b2 = -m[1, 0] * m[0, 2] - m[1, 1] * m[1, 2]
    m[0, 2], m[1, 2] = b1, b2
    width, height = int(size[0]), int(size[1])
    out = torch.empty(height, width, 3, dtype=torch.uint8, device=dev)
    minv = (ctypes.c_double * 6)(*[float(v) for v in m.reshape(-1)])
    fill = [int(v) for v in (list(fill) + [0, 0, 0])[:3]] if not np.isscalar(fill) else [int(fill)] * 3
    fill3 = (ctypes.c_int * 3)(*fill)
    _ops.call('yb_warp_affine_u8', src, int(src.shape[0]), int(src.shape[1]), out, height, width, ctypes.byref(minv), ctypes.byref(fill3))
    return out


def rotation_matrix(center_xy, angle, scale=1.0):
    """cv2.getRotationMatrix2D (degrees, counter-clockwise, top-left origin)."""
    a = math.radians(angle)
    alpha, beta = math.cos(a) * scale, math.sin(a) * scale
    cx, cy = center_xy
    return np.array([[alpha, beta, (1 - alpha) * cx - beta * cy], [-beta, alpha, beta * cx + (1 - alpha) * cy]], np.float64)


class Rotator(object):
    """Rotates about (y, x) and grows the canvas to the rotated bounding box (reference transform/augmentation.py:28-59)."""

    def __init__(self, y, x, height, width, angle):
        self._mat = rotation_matrix((x, y), angle, 1.0)
        r = np.abs(self._mat[0, :2])
        new_h, new_w = np.inner(r, [height, width]), np.inner(r, [width, height])
        self._mat[:, 2] += [new_w / 2 - x, new_h / 2 - y]
        self._size = int(new_w), int(new_h)

    def __call__(self, image, fill=None):
        if fill is None:
            fill = np.random.rand(3) * 256
        return warp_affine(image, self._mat, self._size, fill)

    def _rotate_points(self, points):
        pts = np.pad(points, [(0, 0), (0, 1)], 'constant')
        pts[:, 2] = 1
        return np.dot(self._mat, pts.T).T.astype(points.dtype)

    def rotate_points(self, points):
        return self._rotate_points(points[:, ::-1])[:, ::-1]


def random_rotate(config, image, yx_min, yx_max):
    """Same draw (`random.uniform` over `augmentation/random_rotate`), same box hulls as the reference (:61-76); the frame on the GPU."""
    lo, hi = (float(v) for v in config.get('augmentation', 'random_rotate').split())
    angle = random.uniform(lo, hi)
    as_np = lambda a: a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    yx_min, yx_max = as_np(yx_min), as_np(yx_max)
    height, width = image.shape[:2]
    p1, p2 = np.copy(yx_min), np.copy(yx_max)
    p1[:, 0] = yx_max[:, 0]
    p2[:, 0] = yx_min[:, 0]
    rotator = Rotator(height / 2, width / 2, height, width, angle)
    out = rotator(image, fill=0)
    corners = np.reshape(rotator.rotate_points(np.concatenate([yx_min, yx_max, p1, p2], 0)), [4, -1, 2])
    return out, corners.min(0), corners.max(0)


class RandomRotate(object):
    def __init__(self, config):
        self.config = config

    def __call__(self, data):
        data['image'], data['yx_min'], data['yx_max'] = random_rotate(self.config, data['image'], data['yx_min'], data['yx_max'])
        return data


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
