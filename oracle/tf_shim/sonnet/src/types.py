"""type aliases named in signatures of vq_layers.py (annotations only)."""
from typing import Union
FloatLike = Union[float, int]
