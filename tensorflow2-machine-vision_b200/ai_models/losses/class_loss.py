"""Drop-in for the reference's losses/class_loss.py (ClassFocalLoss :3-60; used only by the stale demo model)."""
import torch

from ... import _tensors as T
from .focal_loss import _partial_sums


class ClassFocalLoss(object):
  """Per-level focal loss summed (not averaged) and divided by normalizer_l = sum(mask_l) / batch."""

  def __init__(self, alpha, gamma, label_smoothing=0.0, **kwargs):
    self.alpha = alpha
    self.gamma = gamma
    self.label_smoothing = label_smoothing

  def call(self, y_true, y_pred):
    class_targets = y_true
    class_outputs, mask = y_pred
    sums, _ = _partial_sums(None, list(class_targets), None, None, list(class_outputs), self.alpha, self.gamma, 0.1,
                            self.label_smoothing)
    total = torch.zeros((), dtype=torch.float32, device=sums.device)
    for i in range(len(class_targets)):
      m = T.to_cuda(mask[i], torch.float32)
      normalizer = m.sum() / float(m.shape[0])
      term = sums[i].to(torch.float32) / normalizer
      total = total + torch.where(normalizer == 0, torch.zeros_like(term), term)  # divide_no_nan
    return total

  __call__ = call
