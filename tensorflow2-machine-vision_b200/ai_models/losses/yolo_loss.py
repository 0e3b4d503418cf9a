"""Drop-in for the reference's losses/yolo_loss.py: Yolov4Loss (keras-yolo3 formulation, :4-159)."""
import numpy as np

from ..utils.tf_yolo_utils import _loss_call


class Yolov4Loss(object):
  '''anchors: flat ascending (9,2) pixel list; layers use anchor_mask [[6,7,8],[3,4,5],[0,1,2]] (:96).'''

  def __init__(self, anchors, classes_num, ignore_thresh=.5, print_loss=False, **args):
    self.anchors = np.asarray(anchors, dtype=np.float32).reshape(-1, 2)
    self.classes_num = classes_num
    self.ignore_thresh = ignore_thresh
    self.print_loss = print_loss

  def call(self, y_true, y_pred):
    '''y_true: 3 x (B,H,W,3,5+C); y_pred: 3 x (B,H,W,3*(5+C)).  input_shape = grid0 * 32 (:98).'''
    anchor_mask = [[6, 7, 8], [3, 4, 5], [0, 1, 2]]
    anchors_wh = np.stack([self.anchors[m] for m in anchor_mask], axis=0)
    h0, w0 = y_pred[0].shape[1], y_pred[0].shape[2]
    image_wh = (w0 * 32, h0 * 32)
    return _loss_call(y_true, y_pred, image_wh, anchors_wh, self.ignore_thresh, 'iou', 1)

  __call__ = call
