"""Drop-in for the reference's utils/tf_iou_utils.py (GetIOU :5-65, GetIOUNMS :67-108, GetIOUNMSByClasses :110-157)."""
import torch

from ... import _lib, _tensors as T


def GetIOU(b1, b2, iou_type='iou'):
  '''
  IOU / DIOU / CIOU, boxes (x1, y1, x2, y2).

  Args:
    b1:(..., b1_num, 1, 4)
    b2:(1, b2_num, 4)
    iou_type: 'iou', 'diou', 'ciou'
  Returns:
    (..., b1_num, b2_num); any other broadcastable pair of shapes is evaluated elementwise.
  '''
  assert iou_type in ['iou','diou','ciou']
  lib = _lib.load()
  b1 = T.to_cuda(b1)
  b2 = T.to_cuda(b2)
  metric = _lib.METRIC_YOLO[iou_type]
  if b1.dim() >= 2 and b1.shape[-2] == 1 and b2.dim() == 3 and b2.shape[0] == 1:
    lead = b1.shape[:-2]
    m = b1.reshape(-1, 4)
    n2 = b2.shape[1]
    out = torch.empty((m.shape[0], n2), dtype=torch.float32, device=b1.device)
    _lib.check(lib.b200_pairwise_iou(T.ptr(m), m.shape[0], T.ptr(b2), n2, metric, T.ptr(out), T.stream_ptr()),
               'GetIOU')
    return out.reshape(*lead, n2)
  x1, x2 = torch.broadcast_tensors(b1, b2)
  x1 = x1.contiguous()
  x2 = x2.contiguous()
  out = torch.empty(x1.shape[:-1], dtype=torch.float32, device=x1.device)
  _lib.check(lib.b200_elementwise_iou(T.ptr(x1), T.ptr(x2), out.numel(), metric, T.ptr(out), T.stream_ptr()),
             'GetIOU')
  return out


def _nms(boxes, scores, classes, max_output_size, iou_threshold, metric, mode, score_threshold=None):
  lib = _lib.load()
  boxes = T.to_cuda(boxes).reshape(-1, 4)
  scores = T.to_cuda(scores).reshape(-1)
  n = boxes.shape[0]
  if scores.shape[0] != n:
    raise ValueError('boxes and scores disagree: %d vs %d' % (n, scores.shape[0]))
  cls = None
  if classes is not None:
    cls = T.to_cuda(classes, torch.int32).reshape(-1)
  max_out = int(min(int(max_output_size), max(n, 1)))
  if max_out < 1:
    return torch.empty((0,), dtype=torch.int32, device=boxes.device)
  seg = torch.tensor([0, n], dtype=torch.int32, device=boxes.device)
  out_idx = torch.empty((max_out,), dtype=torch.int32, device=boxes.device)
  out_cnt = torch.zeros((1,), dtype=torch.int32, device=boxes.device)
  use_thr = 0 if score_threshold is None else 1
  thr = 0.0 if score_threshold is None else float(score_threshold)
  _lib.check(lib.b200_nms(T.ptr(boxes), T.ptr(scores), T.ptr(cls), 0, T.ptr(seg), 1, metric, mode,
                          float(iou_threshold), use_thr, thr, max_out, T.ptr(out_idx), T.ptr(out_cnt),
                          T.stream_ptr()), 'nms')
  k = int(out_cnt.item())
  return out_idx[:k]


def GetIOUNMS(boxes, scores, max_output_size, iou_threshold=0.5, iou_type='iou'):
  '''Class-agnostic greedy NMS; returns int32 indices into `boxes` in descending-score emit order.'''
  assert iou_type in ['iou','diou','ciou']
  return _nms(boxes, scores, None, max_output_size, iou_threshold, _lib.METRIC_YOLO[iou_type], _lib.NMS_AGNOSTIC)


def GetIOUNMSByClasses(boxes, scores, classes, max_output_size, iou_threshold=0.5, iou_type='iou'):
  '''Per-class greedy NMS; returns int32 indices into `boxes` in descending-score emit order.'''
  assert iou_type in ['iou','diou','ciou']
  return _nms(boxes, scores, classes, max_output_size, iou_threshold, _lib.METRIC_YOLO[iou_type], _lib.NMS_BY_CLASS)
