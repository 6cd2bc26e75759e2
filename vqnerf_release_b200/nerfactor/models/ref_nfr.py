"""Mirror of decomp/nerfvq_nfr3/nerfactor/models/ref_nfr.py::Model -- the residual ("reference-RGB") decomposition model
that nerfactor/test.py:181-197 renders for every view before the VQ model.

Same shading path as vq_nfr.Model with one more MLP: `rgb_enc` (3 -> 256 -> 256 -> 256: none, relu, sigmoid;
ref_nfr.py:148) encodes a reference RGB per pixel, and the albedo / roughness heads read the 512-d concat
[z_xyz, z_ref] (`diff_out`, `rough_out`: 512 -> 256 -> 128 (+512) -> out, :149-152) while `spec_out` (ks) keeps the
256-d z_xyz.  `fine_enc`, `bottleneck` and `spec_out` are the frozen VQ-stage networks (:141-146).

Inference surface only (`call` forward, `fast_render`, the `_pred_*_at` entry points): every arithmetic step is a call
into libvqnerf_b200.so -- `vqn_pred_enc_at`, `vqn_net_forward` (tcgen05 kernel; `rgb_enc`'s 3-wide input is padded with
a zero-weight fourth column for the 16-byte row loads), `vqn_material_combine / _edit`, `vqn_shade`.  Training this model
(its own compute_loss, ref_nfr.py:560-700) is outside SURVEY 8's rows.

Batch tuple (ref_nfr.py:87-104): (id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, ref[, lvis]).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from ... import abi
from ..networks import mlp
from .vq_nfr import Model as VQNfrModel


class Model(VQNfrModel):
    def __init__(self, config=None, debug: bool = False, *, nets: Optional[Dict] = None, light=None,
                 novel_probes: Optional[Dict] = None, device='cuda'):
        """`nets`: name -> (kernels, biases) for fine_enc, bottleneck, spec_out (VQ-stage checkpoint, :141-146) and
        rgb_enc, diff_out, rough_out (:148-152); missing ones get Keras' default init."""
        given = dict(nets or {})
        base = {k: given[k] for k in ('fine_enc', 'bottleneck') if k in given}
        if 'spec_out' in given:
            base['spec_main'] = given['spec_out']
        super().__init__(config, debug, nets=base, light=light, novel_probes=novel_probes, device=device)
        z = self.z_dim
        spec = {'rgb_enc': (3, [z, z, z], [None, 'relu', 'sigmoid'], None),
                'diff_out': (2 * z, [z, z // 2, 3], ['relu'] * 2 + ['sigmoid'], [1]),
                'rough_out': (2 * z, [z, z // 2, 1], ['relu'] * 2 + ['sigmoid'], [1])}
        net = {'fine_enc': self.net['fine_enc'], 'bottleneck': self.net['bottleneck'], 'spec_out': self.net['spec_main']}
        for i, (name, (in_dim, widths, act, skip)) in enumerate(spec.items()):
            if name in given:
                ks, bs = [np.asarray(k, np.float32) for k in given[name][0]], list(given[name][1])
            else:
                tmp = mlp.Network(widths, act=act, skip_at=skip, device='cpu', seed=self.seed + 100 + i)
                rng = np.random.RandomState(self.seed + 100 + i)
                ks, bs, d = [], [], in_dim
                for li, w in enumerate(widths):
                    lim = np.sqrt(6.0 / (d + w))
                    ks.append(rng.uniform(-lim, lim, size=(d, w)).astype(np.float32))
                    bs.append(np.zeros((w,), np.float32))
                    d = w + (in_dim if (skip is not None and li in skip) else 0)
                del tmp
            if ks[0].shape[0] != in_dim:
                raise ValueError('%s: first kernel must have %d rows' % (name, in_dim))
            if name == 'rgb_enc':                       # 3 -> 4 input columns, the fourth meets a zero weight row
                ks = [np.concatenate([ks[0], np.zeros((1, ks[0].shape[1]), np.float32)], 0)] + ks[1:]
            net[name] = mlp.Network.from_arrays(ks, bs, act, skip_at=skip, device=self.device)
        self.net = net

    # ------------------------------------------------------------------ fine-grained entry points (ref_nfr.py:472-540)
    def _pred_bias_at(self, pts):
        return super()._pred_enc_at(pts)

    def _pred_ref_at(self, ref):
        ref4 = torch.zeros((ref.shape[0], 4), dtype=torch.float32, device=ref.device)
        ref4[:, :3] = ref
        return self.net['rgb_enc'].packed.forward(ref4, precision=self.precision)

    def _wide_forward(self, name, x):
        """diff_out / rough_out on the 512-d latent.  The tensor-core kernel takes the whole network in one launch; the
        FFMA reference kernel (precision='fp32') is built for inputs up to 256 wide, so that mode chains the per-layer
        Dense kernels of the training path (3xTF32 mma.sync, fp32-level) with the skip concat of mlp.py:44-48."""
        net = self.net[name]
        if self.precision != 'fp32':
            return net.packed.forward(x.contiguous(), precision=self.precision)
        acts = {None: 0, 'relu': 1, 'sigmoid': 2}
        n = x.shape[0]
        x = x.contiguous()
        y = x
        for i, (w, b) in enumerate(zip(net.kernels, net.biases)):
            k, m = w.shape
            ld = (m + 3) // 4 * 4
            out = torch.empty((n, ld), dtype=torch.float32, device=x.device)
            abi.dense_forward(y, y.shape[1], w, b, out, ld, n, k, m, acts[net.act[i]], 1.0, 0.0)
            y = out[:, :m]
            if net.skip_at is not None and i in net.skip_at:
                y = torch.cat([y, x], dim=-1)
            y = y.contiguous()
        return y

    def _pred_diff_at(self, z, vq=False):
        out = self._wide_forward('diff_out', z)
        return out * self.albedo_slope + self.albedo_bias if (self.albedo_slope != 1.0 or self.albedo_bias != 0.0) else out

    def _pred_spec_at(self, z, vq=False):
        return abi.pred_heads(None, self.net['spec_out'].packed, None, z, precision=self.precision)[1]

    def _pred_rough_at(self, z, vq=False):
        return self._wide_forward('rough_out', z)

    # ------------------------------------------------------------------ shared front end of call / fast_render
    def _unpack_ref(self, batch):
        if self.data_type == 'nerf':
            id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, ref, lvis = batch
        else:
            id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, ref = batch
            lvis = None
        return id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, ref, lvis

    def _materials(self, xyz, ref, row_idx, n):
        """z_xyz, ks, z_ref, basecolor, rough, spec, albedo on the n compact foreground rows (:196-208, 338-350)."""
        nf = self.embedder['xyz'].n_freqs
        z_xyz = abi.pred_enc_at(self.net['fine_enc'].packed, self.net['bottleneck'].packed, nf, xyz, row_idx=row_idx, n=n,
                                precision=self.precision)
        ks = self._pred_spec_at(z_xyz)
        z_ref = self._pred_ref_at(torch.index_select(ref, 0, row_idx[:n].long()))
        z_bias = torch.cat([z_xyz, z_ref], dim=-1)
        basecolor = self._pred_diff_at(z_bias)
        rough = self._pred_rough_at(z_bias)
        albedo, spec, _, _ = abi.material_combine(basecolor, ks)
        return ks, basecolor, rough, spec, albedo

    # ------------------------------------------------------------------ call (ref_nfr.py:176-300), forward only
    def call(self, batch, mode='train', relight_olat=False, relight_probes=False, save_z=False, opt_scale=None,
             bias_weight=None):
        self._validate_mode(mode)
        id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, ref, lvis = self._unpack_ref(batch)
        n_total = alpha.shape[0]
        row_idx, n_act = abi.compact_mask(alpha)
        n = int(n_act.item())
        ks, basecolor, rough, spec, albedo = self._materials(xyz, ref, row_idx, n)
        if (opt_scale is not None) and (mode == 'test'):                       # :210-212
            s = torch.as_tensor(np.asarray(opt_scale), dtype=torch.float32).reshape(1, -1).to(xyz.device)
            albedo, spec = albedo * s, spec * s
        gamma = None if self.data_type == 'nerf' else self.gamma
        sh = abi.shade(xyz, rayo, normal, lvis, albedo, spec, rough, self.lxyz, self.lareas,
                       self._lights(relight_probes, None), row_idx=row_idx, n=n, n_total=n_total, gamma=gamma, clip_light0=False,
                       want_split=(mode != 'train'), want_normal=True) if (mode != 'train' or not relight_probes) else None
        if sh is None or (relight_probes and mode != 'train'):
            # the diffuse / specular split is a property of the model light only: probes come from a second, un-split call
            shp = abi.shade(xyz, rayo, normal, lvis, albedo, spec, rough, self.lxyz, self.lareas,
                            self._lights(True, None), row_idx=row_idx, n=n, n_total=n_total, gamma=gamma,
                            clip_light0=False, want_normal=True)
            if sh is None:
                sh = shp
        else:
            shp = sh
        self._check_numerics(xyz.device)
        idx = row_idx[:n].long()
        rgb_lin = sh['rgb'][:, 0, :]
        loss_kwargs = {'mode': mode, 'env': self._light, 'gtc': rgb.index_select(0, idx),
                       'rgb': rgb_lin.index_select(0, idx)}
        sc = lambda v: abi.scatter_rows(v, row_idx, n_total, n=n)
        fg = (alpha[:, :1] > 0).to(torch.float32)
        to_srgb = self.data_type == 'nerf'
        pred = {'rgb': abi.linear2srgb(rgb_lin) * fg if to_srgb else rgb_lin, 'normal': sh['normal'], 'albedo': sc(albedo),
                'alpha': pred_alpha, 'spec': sc(spec), 'rough': sc(rough), 'ks': sc(ks), 'basecolor': sc(basecolor)}
        if mode != 'train':
            pred['rgb_spec'], pred['rgb_diff'] = sh['rgb_spec'], sh['rgb_diff']
        if relight_probes and len(self.novel_probes) > 0:
            rp = shp['rgb'][:, 1:, :]
            pred['rgb_probes'] = abi.linear2srgb(rp) * fg[:, None, :] if to_srgb else rp
        gt = {'rgb': rgb * fg, 'normal': normal * fg, 'alpha': alpha}
        to_vis = {'id': id_, 'hw': hw}
        for k, v in pred.items():
            to_vis['pred_' + k] = v
        for k, v in gt.items():
            to_vis['gt_' + k] = v
        return pred, gt, loss_kwargs, to_vis

    __call__ = call

    # ------------------------------------------------------------------ fast_render (ref_nfr.py:306-417)
    def fast_render(self, batch, mode='train', relight_olat=False, relight_probes=False, opt_scale=None,
                    edit_mask=None, edit_material=None):
        self._validate_mode(mode)
        id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, ref, lvis = self._unpack_ref(batch)
        n_total = alpha.shape[0]
        dev = xyz.device
        row_idx, n_act = abi.compact_mask(alpha)
        n = int(n_act.item())
        ks, basecolor, rough, spec, albedo = self._materials(xyz, ref, row_idx, n)
        if edit_mask is not None:                                              # :352-358
            abi.material_edit(edit_mask, edit_material, row_idx, n_act, albedo, spec, rough)
        gamma = None if self.data_type == 'nerf' else self.gamma
        to_srgb = self.data_type == 'nerf'
        # rgb_pred: the RAW (un-scaled) BRDF under the model light (:360-361, 370-372)
        raw = abi.shade(xyz, rayo, normal, lvis, albedo, spec, rough, self.lxyz, self.lareas, self._lights(False, None),
                        row_idx=row_idx, n=n, n_total=n_total, to_srgb=to_srgb, gamma=gamma, clip_light0=False)
        pred = {'rgb': raw['rgb'][:, 0, :], 'alpha': pred_alpha}
        if relight_probes and len(self.novel_probes) > 0:                      # the opt_scale'd BRDF under the probes (:363-376)
            if opt_scale is not None:
                s = torch.as_tensor(np.asarray(opt_scale), dtype=torch.float32).reshape(1, -1).to(dev)
                albedo, spec = albedo * s, spec * s
            shp = abi.shade(xyz, rayo, normal, lvis, albedo, spec, rough, self.lxyz, self.lareas,
                            self._lights(True, None), row_idx=row_idx, n=n, n_total=n_total, to_srgb=to_srgb, gamma=gamma,
                            clip_light0=False)
            pred['rgb_probes'] = shp['rgb'][:, 1:, :]
        self._check_numerics(dev)
        idx = row_idx[:n].long()
        loss_kwargs = {'mode': mode, 'env': self._light, 'gtc': rgb.index_select(0, idx)}
        fg = (alpha[:, :1] > 0).to(torch.float32)
        gt = {'rgb': rgb * fg, 'normal': normal * fg, 'alpha': alpha}
        to_vis = {'id': id_, 'hw': hw}
        for k, v in pred.items():
            to_vis['pred_' + k] = v
        for k, v in gt.items():
            to_vis['gt_' + k] = v
        return pred, gt, loss_kwargs, to_vis

    # the VQ-stage entry points do not exist on this model
    def fast_embed(self, *a, **k):
        raise AttributeError('ref_nfr.Model has no VQ layer')

    vq_test = vis_mat = init_z = fast_embed
