"""Mirror of geo/NeuS-ours2/models/fields.py: SDFNetwork (:9-112), RenderingNetwork (:116-172),
SingleVarianceNetwork (:246-253), with the reference's constructor arguments, method names and state_dict keys
(lin{l}.weight_g / .weight_v / .bias), evaluated by the fused tcgen05 MLP kernel (csrc/mlp_tc.cu):

* SDFNetwork.forward / .sdf / .gradient are ONE launch of vqn_sdf_forward.  The gradient is not a second autograd
  pass: the kernel carries (value, d/dx, d/dy, d/dz) jets through the layers (four MMA-tile rows per point).
  `forward_with_gradient` returns sdf, feature vector and gradient from a single launch -- what render_core needs.
* RenderingNetwork.forward (mode 'idr') is vqn_net_forward on rows [feature | points, embed(view_dirs), normals];
  the feature vector is written into those rows by the SDF kernel itself, so the 289-wide concat is never copied.

Parameters are torch CUDA tensors owned by these objects (checkpoints of the reference load with load_state_dict);
`repack()` must be called after they change.  There is no CPU path.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from .. import abi

F32 = torch.float32


class _LinStack:
    """Shared parameter handling: nn.Linear layers with optional weight_norm under the reference's names."""

    def _init_params(self, dims_in, dims_out, weight_norm, device):
        self.weight_norm = bool(weight_norm)
        self.params: Dict[str, torch.Tensor] = {}
        for l, (di, do) in enumerate(zip(dims_in, dims_out)):
            k = 1.0 / math.sqrt(di)
            w = (torch.rand((do, di)) * 2 - 1) * k              # nn.Linear default init
            b = (torch.rand((do,)) * 2 - 1) * k
            self._set_linear(l, w, b, device)

    def _set_linear(self, l, w, b, device):
        w = w.to(device=device, dtype=F32)
        if self.weight_norm:
            self.params['lin%d.weight_g' % l] = w.norm(dim=1, keepdim=True).contiguous()
            self.params['lin%d.weight_v' % l] = w.contiguous()
        else:
            self.params['lin%d.weight' % l] = w.contiguous()
        self.params['lin%d.bias' % l] = b.to(device=device, dtype=F32).contiguous()

    def state_dict(self):
        return dict(self.params)

    def load_state_dict(self, state):
        missing = set(self.params) - set(state)
        extra = set(state) - set(self.params)
        if missing or extra:
            raise KeyError('state_dict mismatch: missing %s, unexpected %s' % (sorted(missing), sorted(extra)))
        for k, v in state.items():
            v = torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v)
            if tuple(v.shape) != tuple(self.params[k].shape):
                raise ValueError('%s: shape %s, expected %s' % (k, tuple(v.shape), tuple(self.params[k].shape)))
            self.params[k] = v.to(device=self.params[k].device, dtype=F32).contiguous()
        self.repack()

    def parameters(self):
        return list(self.params.values())

    def _effective(self, l):
        """[out, in] weight with weight_norm folded (w = g v / |v|, torch.nn.utils.weight_norm dim=0), float64."""
        if self.weight_norm:
            v = self.params['lin%d.weight_v' % l].double()
            g = self.params['lin%d.weight_g' % l].double()
            return g * v / v.norm(dim=1, keepdim=True)
        return self.params['lin%d.weight' % l].double()


class SDFNetwork(_LinStack):
    def __init__(self, d_in, d_out, d_hidden, n_layers, skip_in=(4,), multires=0, bias=0.5, scale=1,
                 geometric_init=True, weight_norm=True, inside_outside=False, device='cuda', precision='tf32x3',
                 grad_mode='reverse'):
        if d_in != 3 or multires <= 0:
            raise NotImplementedError('the fused kernel embeds 3-D points in-kernel: d_in == 3 and multires > 0')
        if float(scale) != 1.0:
            raise NotImplementedError('scale != 1 (every shipped conf uses scale = 1.0)')
        skip_in = tuple(skip_in)
        if len(skip_in) > 1 or any(s < 1 or s >= n_layers for s in skip_in):
            raise NotImplementedError('at most one skip_in layer, inside the hidden stack')
        self.multires, self.skip_in, self.scale, self.precision = int(multires), skip_in, 1.0, precision
        self.grad_mode = grad_mode          # 'reverse' (default) or 'jet': how the fused kernel evaluates grad sdf
        self.device = torch.device(device)
        d0 = 3 + 6 * self.multires
        dims = [d0] + [d_hidden] * n_layers + [d_out]
        self.num_layers = len(dims)
        self.d_feature = d_out - 1
        self.weight_norm = bool(weight_norm)
        self.params = {}
        for l in range(self.num_layers - 1):
            out_dim = dims[l + 1] - dims[0] if (l + 1) in skip_in else dims[l + 1]
            k = 1.0 / math.sqrt(dims[l])
            w = (torch.rand((out_dim, dims[l])) * 2 - 1) * k
            b = (torch.rand((out_dim,)) * 2 - 1) * k
            if geometric_init:                                        # fields.py:45-61
                if l == self.num_layers - 2:
                    sign = -1.0 if inside_outside else 1.0
                    w = torch.randn((out_dim, dims[l])) * 1e-4 + sign * math.sqrt(math.pi) / math.sqrt(dims[l])
                    b = torch.full((out_dim,), -sign * bias)
                elif l == 0:
                    b = torch.zeros(out_dim)
                    w = torch.zeros((out_dim, dims[l]))
                    w[:, :3] = torch.randn((out_dim, 3)) * (math.sqrt(2) / math.sqrt(out_dim))
                elif l in skip_in:
                    b = torch.zeros(out_dim)
                    w = torch.randn((out_dim, dims[l])) * (math.sqrt(2) / math.sqrt(out_dim))
                    w[:, -(dims[0] - 3):] = 0.0
                else:
                    b = torch.zeros(out_dim)
                    w = torch.randn((out_dim, dims[l])) * (math.sqrt(2) / math.sqrt(out_dim))
            self._set_linear(l, w, b, self.device)
        self._trunk = self._feat = None
        self.repack()

    def repack(self):
        """Re-derive the kernel's weight images from the parameters (call after an optimizer step / load)."""
        L = self.num_layers - 1
        ws, bs = [], []
        for l in range(L - 1):
            w = self._effective(l).t()                               # Keras layout [in, out]
            if l in self.skip_in:
                w = w / math.sqrt(2.0)                               # cat([x, inputs]) / sqrt(2)  (fields.py:82)
            ws.append(w.to(F32).contiguous())
            bs.append(self.params['lin%d.bias' % l])
        skip_at = (self.skip_in[0] - 1) if self.skip_in else None
        self._trunk = abi.PackedNet(ws, bs, ['softplus100'] * (L - 1), skip_at=skip_at)
        w_last = self._effective(L - 1).t()                          # [d_hidden, 1 + d_feature]
        b_last = self.params['lin%d.bias' % (L - 1)]
        self._w_sdf = w_last[:, 0].to(F32).contiguous()
        self._b_sdf = b_last[:1].contiguous()
        self._feat = None
        if self.d_feature > 0:
            self._feat = abi.PackedNet([w_last[:, 1:].to(F32).contiguous()], [b_last[1:].contiguous()], ['none'])

    # ---- the reference's methods -----------------------------------------------------------------------------
    def forward_with_gradient(self, x, feat_out: Optional[torch.Tensor] = None, want_grad=True, want_feat=True):
        """(sdf [n,1], feature [n,d_feature] (a view of feat_out when given) or None, gradient [n,3] or None)."""
        x = x.reshape(-1, 3)
        if want_feat and self._feat is not None:
            if feat_out is None:
                feat_out = torch.empty((x.shape[0], self.d_feature), dtype=F32, device=x.device)
        else:
            feat_out = None
        sdf, grad = abi.sdf_forward(self._trunk, self._w_sdf, self._b_sdf, self._feat, self.multires, x,
                                    want_grad=want_grad, feat_out=feat_out, precision=self.precision,
                                    grad_mode=self.grad_mode)
        feat = feat_out[:, :self.d_feature] if feat_out is not None else None
        return sdf, feat, grad

    def forward(self, inputs):
        sdf, feat, _ = self.forward_with_gradient(inputs, want_grad=False)
        return sdf if feat is None else torch.cat([sdf, feat], dim=-1)

    __call__ = forward

    def sdf(self, x):
        return self.forward_with_gradient(x, want_grad=False, want_feat=False)[0]

    def sdf_hidden_appearance(self, x):
        return self.forward(x)

    def gradient(self, x):
        return self.forward_with_gradient(x, want_grad=True, want_feat=False)[2].unsqueeze(1)


class RenderingNetwork(_LinStack):
    ROW = 320           # [feature 256 | points 3, embed(view) 27, normals 3, zero pad]: a multiple of the 64-wide K chunk

    def __init__(self, d_feature, mode, d_in, d_out, d_hidden, n_layers, weight_norm=True, multires_view=0,
                 squeeze_out=True, device='cuda', precision='tf32x3'):
        if mode != 'idr' or d_in != 9:
            raise NotImplementedError("only mode 'idr' with d_in = 9 (every shipped conf)")
        self.mode, self.squeeze_out, self.precision = mode, squeeze_out, precision
        self.multires_view = int(multires_view)
        self.d_feature = int(d_feature)
        self.device = torch.device(device)
        self.side = 9 + 6 * self.multires_view                      # points + embedded view_dirs + normals
        self.row = max(self.ROW, -(-(self.d_feature + self.side) // 64) * 64)
        if self.d_feature % 4 != 0:
            raise NotImplementedError('d_feature must be a multiple of 4 (16-byte aligned row segments)')
        dims = [d_in + d_feature + 6 * self.multires_view] + [d_hidden] * n_layers + [d_out]
        self.num_layers = len(dims)
        self._init_params(dims[:-1], dims[1:], weight_norm, self.device)
        self._net = None
        self.repack()

    def repack(self):
        L = self.num_layers - 1
        ws, bs = [], []
        for l in range(L):
            w = self._effective(l).t()                               # [in, out]
            if l == 0:
                # reference input order [points, view, normals, feature] (fields.py:151) -> [feature | points, view,
                # normals | 0]; rows of the kernel permuted accordingly
                e = 3 + 3 + 6 * self.multires_view
                wp = torch.zeros((self.row, w.shape[1]), dtype=torch.float64, device=w.device)
                wp[:self.d_feature] = w[e + 3:]
                wp[self.d_feature:self.d_feature + self.side] = w[:e + 3]
                w = wp
            ws.append(w.to(F32).contiguous())
            bs.append(self.params['lin%d.bias' % l])
        acts = ['relu'] * (L - 1) + ['sigmoid' if self.squeeze_out else 'none']
        self._net = abi.PackedNet(ws, bs, acts)

    def alloc_rows(self, n, device=None):
        """Row buffer [n, 320]; SDFNetwork.forward_with_gradient(feat_out=rows) fills the feature columns."""
        return torch.empty((n, self.row), dtype=F32, device=device or self.device)

    def forward_rows(self, rows, points, normals, view_dirs):
        abi.neus_color_input(points, view_dirs, normals, self.multires_view, rows, self.d_feature,
                             self.row - self.d_feature)
        return self._net.forward(rows, precision=self.precision)

    def forward(self, points, normals, view_dirs, feature_vectors):
        n = points.shape[0]
        rows = self.alloc_rows(n, points.device)
        fv = feature_vectors.to(F32)
        if fv.stride(1) != 1:
            fv = fv.contiguous()
        abi.copy_cols(fv, fv.stride(0), rows, self.row, n, self.d_feature)
        return self.forward_rows(rows, points, normals, view_dirs)

    __call__ = forward


class SingleVarianceNetwork:
    """fields.py:246-253: forward(x) = ones([len(x), 1]) * exp(10 * variance)."""

    def __init__(self, init_val, device='cuda'):
        self.variance = torch.tensor(float(init_val), dtype=F32, device=device)
        self._host = float(init_val)            # host copy: render_core needs inv_s as a kernel argument (no sync)

    def state_dict(self):
        return {'variance': self.variance}

    def load_state_dict(self, state):
        self.variance = torch.as_tensor(state['variance']).to(device=self.variance.device, dtype=F32).reshape(())
        self._host = float(self.variance)

    def inv_s(self) -> float:
        return float(np.float32(math.exp(self._host * 10.0)))

    def forward(self, x):
        return torch.ones((len(x), 1), dtype=F32, device=x.device) * torch.exp(self.variance * 10.0)

    __call__ = forward
