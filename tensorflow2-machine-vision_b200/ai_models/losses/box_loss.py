"""Drop-in for the reference's losses/box_loss.py (BoxLoss :3-29): Huber(delta) on non-zero targets / (4 num_pos)."""
import torch

from ... import _tensors as T
from .focal_loss import _partial_sums


class BoxLoss(object):
  """L2 box regression loss."""

  def __init__(self, delta=0.1, **kwargs):
    self.delta = delta

  def call(self, y_true, box_outputs):
    num_positives, box_targets = y_true
    t = T.to_cuda(box_targets)
    mask = torch.zeros(t.shape[:-1] + (1,), dtype=torch.bool, device=t.device)  # positives are not counted here
    sums, _ = _partial_sums([t], None, [mask], [box_outputs], None, 0.25, 1.5, self.delta, 0.0)
    return (sums[1].to(torch.float32) / (float(num_positives) * 4.0))

  __call__ = call
