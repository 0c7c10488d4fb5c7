"""transform.image -- BGR2RGB of the reference (transform/image.py:27-29) for uint8 HWC CUDA tensors.  Inside
`transform.resize_batch` the swap is fused into the resize kernel; this stand-alone form is a strided view (no copy)."""


class BGR2RGB(object):
    def __call__(self, image):
        return image.flip(-1)
