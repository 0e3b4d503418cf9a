"""Drop-in for the reference's efficientnet/utils/nms.py (get_nms :5-61)."""
from .... import _lib
from ...utils.tf_iou_utils import _nms


def get_nms(boxes,
            scores,
            max_output_size,
            iou_threshold=0.5,
            score_threshold=float('-inf'),
            iou_type='diou'):
  '''
  Class-agnostic greedy NMS on yxyx boxes; stops at the first top score below `score_threshold`.

  Returns:
    int32 indices into `boxes`, descending-score emit order.
  '''
  assert iou_type in ('iou', 'giou', 'diou', 'ciou')
  thr = None if score_threshold == float('-inf') else score_threshold
  return _nms(boxes, scores, None, max_output_size, iou_threshold, _lib.METRIC_EFF[iou_type], _lib.NMS_AGNOSTIC,
              score_threshold=thr)
