"""Drop-in for the reference's efficientnet/utils/iou.py (get_iou :26-100): yxyx boxes, NaN-free family."""
import torch

from .... import _lib, _tensors as T


def get_iou(boxes1, boxes2, iou_type = 'iou'):
  """
  Args:
    boxes1: [..., [y_min, x_min, y_max, x_max]]
    boxes2: [..., [y_min, x_min, y_max, x_max]]
    iou_type: one of 'iou', 'ciou', 'diou', 'giou'
  Returns:
    IoU: [...,] with the usual broadcasting of the leading dimensions.
  """
  assert iou_type in ('iou', 'giou', 'diou', 'ciou')
  lib = _lib.load()
  b1 = T.to_cuda(boxes1)
  b2 = T.to_cuda(boxes2)
  metric = _lib.METRIC_EFF[iou_type]
  # the anchor-vs-GT pattern of generate_targets: (..., 1, 4) x (n, 4) -> (..., n)
  if b1.dim() >= 2 and b1.shape[-2] == 1 and b2.dim() == 2:
    lead = b1.shape[:-2]
    m = b1.reshape(-1, 4)
    n2 = b2.shape[0]
    out = torch.empty((m.shape[0], n2), dtype=torch.float32, device=b1.device)
    _lib.check(lib.b200_pairwise_iou(T.ptr(m), m.shape[0], T.ptr(b2), n2, metric, T.ptr(out), T.stream_ptr()),
               'get_iou')
    return out.reshape(*lead, n2)
  x1, x2 = torch.broadcast_tensors(b1, b2)
  x1 = x1.contiguous()
  x2 = x2.contiguous()
  out = torch.empty(x1.shape[:-1], dtype=torch.float32, device=x1.device)
  _lib.check(lib.b200_elementwise_iou(T.ptr(x1), T.ptr(x2), out.numel(), metric, T.ptr(out), T.stream_ptr()),
             'get_iou')
  return out
