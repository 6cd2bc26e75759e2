"""`tensorflow_probability` stand-in: only `tfp.math.clip_by_value_preserve_gradient` (tfp 0.12.1, restated from its
published definition: `t + stop_gradient(clip_by_value(t, lo, hi) - t)`).  TEST INFRASTRUCTURE ONLY."""
import types as _types

import torch as _torch

math = _types.ModuleType('tensorflow_probability.math')


def _clip_by_value_preserve_gradient(t, clip_value_min, clip_value_max, name=None):
    clip_t = _torch.clamp(t, min=clip_value_min, max=clip_value_max)
    return t + (clip_t - t).detach()


math.clip_by_value_preserve_gradient = _clip_by_value_preserve_gradient
