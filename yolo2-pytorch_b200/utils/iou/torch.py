"""utils.iou.torch -- IoU matrices on the GPU (reference utils/iou/torch.py:47-61 iou_matrix,
:139-153 batch_iou_matrix).  Same operation order as the reference (clamped intersection extents,
union clamped to float32 eps), evaluated by yb_iou_matrix with round-to-nearest, no FMA
contraction, so results are bit-identical to the reference's fp32 arithmetic."""
import numpy as np

from b200 import ops as _ops

_EPS = float(np.finfo(np.float32).eps)


def iou_matrix(yx_min1, yx_max1, yx_min2, yx_max2, min=_EPS):
    """[N1,2] x2 vs [N2,2] x2 -> [N1,N2]."""
    return _ops.iou_matrix(yx_min1.contiguous().float(), yx_max1.contiguous().float(), yx_min2.contiguous().float(),
                           yx_max2.contiguous().float(), min)


def batch_iou_matrix(yx_min1, yx_max1, yx_min2, yx_max2, min=_EPS):
    """[B,N1,2] x2 vs [B,N2,2] x2 -> [B,N1,N2]."""
    return _ops.iou_matrix(yx_min1.contiguous().float(), yx_max1.contiguous().float(), yx_min2.contiguous().float(),
                           yx_max2.contiguous().float(), min)


def batch_iou_pair(yx_min1, yx_max1, yx_min2, yx_max2, min=_EPS):
    """Pairwise IoU of two equally shaped box lists [N,M,2] x2 -> [N,M] (reference utils/iou/torch.py:216-233): the same
    arithmetic as iou_matrix, evaluated by the same kernel on N*M one-by-one problems."""
    shape = yx_min1.shape[:-1]
    flat = lambda t: t.contiguous().float().reshape(-1, 1, 2)
    return _ops.iou_matrix(flat(yx_min1), flat(yx_max1), flat(yx_min2), flat(yx_max2), min).reshape(shape)
