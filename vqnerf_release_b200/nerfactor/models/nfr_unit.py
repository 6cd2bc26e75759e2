"""Mirror of nerfactor/models/nfr_unit.py::Model -- the warm-up model of the decomposition stage (no VQ layer, no second
branch): fine_enc + bottleneck give the latent z_bias, diff_out / spec_out / rough_out the material, and the point is
shaded under the learnable light (nfr_unit.py:110-129 nets, :146-180 gen_z, :182-271 call, :273-306 _render).

The forward pass is the SAME kernels as the VQ model's main branch: the three `*_out` heads have the shapes of
`diff_main / spec_main / rough_main`, so `call` and `gen_z` run encoder + heads as ONE fused tcgen05 launch
(`abi.mlp_main`) and the shading as one `abi.shade` launch.  Forward only: the warm-up TRAINING loop (trainvali.py) is
outside the hot path."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from ... import abi
from .vq_nfr import Model as VQNfrModel


class Model(VQNfrModel):
    def __init__(self, config=None, debug: bool = False, *, nets: Optional[Dict] = None, light=None,
                 novel_probes: Optional[Dict] = None, device='cuda'):
        """`nets`: name -> (kernels, biases) for fine_enc, bottleneck, diff_out, spec_out, rough_out (nfr_unit.py:110-129);
        missing ones get Keras' default init."""
        given = dict(nets or {})
        base = {k: given[k] for k in ('fine_enc', 'bottleneck') if k in given}
        for src, dst in (('diff_out', 'diff_main'), ('spec_out', 'spec_main'), ('rough_out', 'rough_main')):
            if src in given:
                base[dst] = given[src]
        super().__init__(config, debug, nets=base, light=light, novel_probes=novel_probes, device=device)
        self.net = {'fine_enc': self.net['fine_enc'], 'bottleneck': self.net['bottleneck'],
                    'diff_out': self.net['diff_main'], 'spec_out': self.net['spec_main'],
                    'rough_out': self.net['rough_main']}

    # ------------------------------------------------------------------ fine-grained entry points (nfr_unit.py:329-384)
    def _pred_bias_at(self, pts):
        return abi.pred_enc_at(self.net['fine_enc'].packed, self.net['bottleneck'].packed, self.embedder['xyz'].n_freqs,
                               pts, precision=self.precision)

    def _pred_diff_at(self, z, vq=False):
        return abi.pred_heads(self.net['diff_out'].packed, None, None, z, self.albedo_slope, self.albedo_bias,
                              self.precision)[0]

    def _pred_spec_at(self, z, vq=False):
        return abi.pred_heads(None, self.net['spec_out'].packed, None, z, precision=self.precision)[1]

    def _pred_rough_at(self, z, vq=False):
        return abi.pred_heads(None, None, self.net['rough_out'].packed, z, precision=self.precision)[2]

    def _materials(self, xyz, row_idx, n_act, n_total, want_z=False):
        n_ = self.net
        return abi.mlp_main(n_['fine_enc'].packed, n_['bottleneck'].packed, n_['diff_out'].packed, n_['spec_out'].packed,
                            n_['rough_out'].packed, self.embedder['xyz'].n_freqs, xyz, row_idx=row_idx, n_dev=n_act,
                            n=n_total, slope=self.albedo_slope, bias=self.albedo_bias, want_z=want_z,
                            precision=self.precision)

    # ------------------------------------------------------------------ gen_z (nfr_unit.py:146-180): no host sync
    def gen_z(self, batch, genz=False):
        if self.data_type == 'nerf':
            id_, hw, _, _, _, alpha, pred_alpha, xyz, _, _ = batch
        else:
            id_, hw, _, _, _, alpha, pred_alpha, xyz, _ = batch
        n_total = alpha.shape[0]
        row_idx, n_act = abi.compact_mask(alpha)
        z_bias, basecolor, ks, rough = self._materials(xyz, row_idx, n_act, n_total, want_z=genz)
        albedo, spec, _, _ = abi.material_combine(basecolor, ks, n_dev=n_act)
        sc = lambda v: abi.scatter_rows(v, row_idx, n_total, n_dev=n_act, n=n_total)
        to_vis = {'id': id_, 'hw': hw, 'gt_alpha': alpha, 'pred_alpha': pred_alpha, 'albedo': sc(albedo), 'spec': sc(spec),
                  'rough': sc(rough)}
        if genz:
            to_vis['z_bias'] = sc(z_bias)
        return to_vis

    # ------------------------------------------------------------------ call (nfr_unit.py:182-271), forward only
    def call(self, batch, mode='train', pretrain=False, relight_olat=False, relight_probes=False, save_z=False,
             opt_scale=None, bias_weight=None):
        self._validate_mode(mode)
        id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, lvis = self._unpack(batch, False)
        n_total = alpha.shape[0]
        row_idx, n_act = abi.compact_mask(alpha)
        # loss_kwargs carries the COMPACT rows (boolean_mask'ed gtc / rgb / spec / rough, :232-236): their shape is the
        # foreground count, so this entry point has the reference's one host round trip
        n = int(n_act.item())
        _, basecolor, ks, rough = self._materials(xyz, row_idx, n_act, n)
        albedo, spec, _, _ = abi.material_combine(basecolor, ks)
        gamma = None if self.data_type == 'nerf' else self.gamma
        sh = abi.shade(xyz, rayo, normal, lvis, albedo, spec, rough, self.lxyz, self.lareas, self._lights(False, None),
                       row_idx=row_idx, n=n, n_total=n_total, gamma=gamma, want_split=(mode != 'train'), want_normal=True)
        self._check_numerics(xyz.device)
        idx = row_idx[:n].long()
        rgb_lin = sh['rgb'][:, 0, :]
        loss_kwargs = {'mode': mode, 'pretrain': pretrain, 'gtc': rgb.index_select(0, idx),
                       'rgb': rgb_lin.index_select(0, idx), 'env': self._light, 'spec': spec, 'rough': rough}
        sc = lambda v: abi.scatter_rows(v, row_idx, n_total, n=n)
        fg = (alpha[:, :1] > 0).to(torch.float32)
        to_srgb = self.data_type == 'nerf'
        pred = {'rgb': abi.linear2srgb(rgb_lin) * fg if to_srgb else rgb_lin, 'normal': sh['normal'], 'albedo': sc(albedo),
                'basecolor': sc(basecolor), 'alpha': pred_alpha, 'spec': sc(spec), 'rough': sc(rough), 'ks': sc(ks),
                'xyz': xyz * fg}
        if mode != 'train':
            pred['rgb_spec'], pred['rgb_diff'] = sh['rgb_spec'], sh['rgb_diff']
        gt = {'rgb': rgb, 'normal': normal, 'alpha': alpha, 'xyz': xyz}                      # (:193: un-masked in this model)
        to_vis = {'id': id_, 'hw': hw}
        for k, v in pred.items():
            to_vis['pred_' + k] = v
        for k, v in gt.items():
            to_vis['gt_' + k] = v
        return pred, gt, loss_kwargs, to_vis

    __call__ = call

    # the VQ-stage entry points do not exist on this model
    def fast_embed(self, *a, **k):
        raise AttributeError('nfr_unit.Model has no VQ layer')

    vq_test = vis_mat = init_z = fast_render = fast_embed
