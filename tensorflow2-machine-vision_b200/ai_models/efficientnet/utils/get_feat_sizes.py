"""Drop-in for the reference's efficientnet/utils/get_feat_sizes.py (:4-20); host-side integer arithmetic."""
from typing import List, Tuple


def get_feat_sizes(image_size: Tuple[int, int], max_level: int) -> List[Tuple[int, int]]:
  '''Feature-map (height, width) for levels 0..max_level; each level is ceil(previous / 2).'''
  size = (int(image_size[0]), int(image_size[1]))
  sizes = [size]
  for _ in range(max_level):
    size = ((size[0] - 1) // 2 + 1, (size[1] - 1) // 2 + 1)
    sizes.append(size)
  return sizes
