"""transform.resize.image -- `rescale(image, height, width)` of the reference (transform/resize/image.py:23-24:
cv2.resize(image, (width, height))) on the GPU, bit-exact for uint8 HWC images."""
import torch

import transform as _t


def rescale(image, height, width):
    """image: uint8 [h, w, 3] tensor or ndarray -> uint8 CUDA tensor [height, width, 3] (channel order unchanged)."""
    if not torch.is_tensor(image):
        image = torch.from_numpy(image)
    return _t.resize_batch([image], height, width, bgr2rgb=False)[0]


class Rescale(object):
    def __call__(self, image, height, width):
        return rescale(image, height, width)
