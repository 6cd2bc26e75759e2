"""Mirror of decomp/nerfvq_nfr3/nerfactor/networks/mlp.py:24-50.

Keras builds Dense kernels lazily; here the layers are created either from explicit (kernel, bias)
arrays (`from_arrays`, the checkpoint hand-off of vq_nfr.py:148-155) or with Keras' default
glorot-uniform / zero-bias init on first call.
"""
import math
from typing import List, Optional, Sequence

import numpy as np
import torch

from ... import abi


class Network:
    def __init__(self, widths: Sequence[int], act: Optional[Sequence] = None, skip_at: Optional[Sequence[int]] = None,
                 device='cuda', seed: Optional[int] = None):
        depth = len(widths)
        if act is None:
            act = [None] * depth
        assert len(act) == depth, "If not `None`, `act` must have the save length as `widths`"
        if skip_at is not None and len(skip_at) > 1:
            raise NotImplementedError('one skip connection per network (as every net on the hot path)')
        self.widths = [int(w) for w in widths]
        self.act = list(act)
        self.skip_at = skip_at
        self.device = torch.device(device)
        self.seed = seed
        self.kernels: List[torch.Tensor] = []
        self.biases: List[torch.Tensor] = []
        self._packed: Optional[abi.PackedNet] = None

    # -- construction --------------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, kernels, biases, act, skip_at=None, device='cuda'):
        net = cls([int(np.shape(k)[1]) for k in kernels], act=act, skip_at=skip_at, device=device)
        net.kernels = [torch.as_tensor(np.asarray(k), dtype=torch.float32).to(net.device).contiguous() for k in kernels]
        net.biases = [torch.as_tensor(np.asarray(b), dtype=torch.float32).to(net.device).contiguous() for b in biases]
        net._pack()
        return net

    def build(self, in_dim: int):
        rng = np.random.RandomState(self.seed)
        d, ks, bs = in_dim, [], []
        for i, w in enumerate(self.widths):
            limit = math.sqrt(6.0 / (d + w))      # Keras glorot_uniform
            ks.append(rng.uniform(-limit, limit, size=(d, w)).astype(np.float32))
            bs.append(np.zeros((w,), np.float32))
            d = w + (in_dim if (self.skip_at is not None and i in self.skip_at) else 0)
        self.kernels = [torch.from_numpy(k).to(self.device) for k in ks]
        self.biases = [torch.from_numpy(b).to(self.device) for b in bs]
        self._pack()

    def _pack(self):
        self._packed = abi.PackedNet(self.kernels, self.biases, self.act,
                                     None if self.skip_at is None else int(self.skip_at[0]))

    def weights_updated(self):
        """Call after modifying `kernels` / `biases` in place (optimizer step)."""
        self._packed.repack()

    @property
    def packed(self) -> abi.PackedNet:
        if self._packed is None:
            raise RuntimeError('network not built yet')
        return self._packed

    # -- mlp.py:39-50 ----------------------------------------------------------------------------
    def __call__(self, x, precision='fp32'):
        if self._packed is None:
            self.build(int(x.shape[-1]))
        return self._packed.forward(x, precision=precision)
