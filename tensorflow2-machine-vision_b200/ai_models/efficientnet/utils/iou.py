"""Drop-in for the reference's efficientnet/utils/iou.py (get_iou :26-100): yxyx boxes, NaN-free family."""
import torch

from .... import _lib, _tensors as T


def _get_v(b1_height, b1_width, b2_height, b2_width, dv=None):
  """Aspect-ratio term of the CIoU (reference :5-24).  With `dv` (the upstream gradient) also returns the gradient the
  reference's tf.custom_gradient hands back for (b2_height, b2_width): (v, grad_height, grad_width) — what a
  tf.custom_gradient wrapper around the drop-in passes on."""
  lib = _lib.load()
  ts = torch.broadcast_tensors(*[T.to_cuda(t) for t in (b1_height, b1_width, b2_height, b2_width)])
  h1, w1, h2, w2 = [t.contiguous() for t in ts]
  v = torch.empty_like(h1)
  if dv is None:
    _lib.check(lib.b200_ciou_v_grad(T.ptr(h1), T.ptr(w1), T.ptr(h2), T.ptr(w2), None, v.numel(), T.ptr(v), None, None,
                                    T.stream_ptr()), '_get_v')
    return v
  d = torch.broadcast_to(T.to_cuda(dv), h1.shape).contiguous()
  gh, gw = torch.empty_like(h1), torch.empty_like(h1)
  _lib.check(lib.b200_ciou_v_grad(T.ptr(h1), T.ptr(w1), T.ptr(h2), T.ptr(w2), T.ptr(d), v.numel(), T.ptr(v), T.ptr(gh), T.ptr(gw),
                                  T.stream_ptr()), '_get_v')
  return v, gh, gw


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
