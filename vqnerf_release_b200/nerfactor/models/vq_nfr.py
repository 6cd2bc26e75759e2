"""Mirror of decomp/nerfvq_nfr3/nerfactor/models/vq_nfr.py::Model -- the drop-in surface of the hot path.

Same method names, argument meaning, batch tuple layout and returned dict keys as the reference
(`Model.__call__/call`, `fast_render`, `fast_embed`, `vq_test`, `vis_mat`, `init_z`, `_pred_*_at`,
`_eval_brdf_at`, `_render`, `_calc_ldir`, `_calc_vdir`, `_normal_correct`, `get_codebook`, `light`, `gamma`),
with torch CUDA tensors instead of eager TF tensors.  Every arithmetic step is a call into
libvqnerf_b200.so (hand-written sm_100a kernels); PyTorch only allocates memory and provides streams.

Batch tuple (datasets/shape_unit.py:93-110, data_type == 'nerf'):
    (id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, lvis)          ref_batch=False
    (id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, ref, lvis)     ref_batch=True
and without the trailing lvis for the other data types.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional

import numpy as np
import torch

from ... import abi
from ... import _lib as L
from ..networks import mlp
from ..networks.embedder import Embedder
from ..networks.vq_layers import VectorQuantizerEMA

_DEFAULTS = {
    'data_type': 'nerf', 'no_brdf_chunk': 'True', 'random_seed': '2', 'pred_brdf': 'True', 'conv_width': '256',
    'mlp_width': '128', 'mlp_chunk': '100000', 'n_freqs_xyz': '10', 'light_h': '16', 'num_embed': '15',
    'commitment_cost': '0.1', 'albedo_slope': '1', 'albedo_bias': '0', 'brdf_chunk_size': '50000',
    'pos_enc': 'True',
    # MLP arithmetic: 'tf32x3' = tcgen05 3xTF32 split with fp32 accumulation (fp32 parity, default),
    # 'fp32' = FFMA on CUDA cores, 'bf16' = tcgen05 bf16 operands (1e-2 budget)
    'precision': 'tf32x3',
}


class _Cfg:
    """Accepts a configparser.ConfigParser (reference style: config.get('DEFAULT', key)) or a plain dict."""

    def __init__(self, config):
        self.c = config if config is not None else {}

    def get(self, key, fallback=None):
        c = self.c
        try:
            if hasattr(c, 'has_option'):
                if c.has_option('DEFAULT', key):
                    return c.get('DEFAULT', key)
            elif key in c:
                return str(c[key])
        except Exception:
            pass
        if fallback is not None:
            return str(fallback)
        return _DEFAULTS.get(key)

    def getint(self, key, fallback=None):
        return int(float(self.get(key, fallback)))

    def getfloat(self, key, fallback=None):
        return float(self.get(key, fallback))

    def getboolean(self, key, fallback=None):
        return str(self.get(key, fallback)).strip().lower() in ('1', 'true', 'yes', 'on')


class Model:
    MODES = ('train', 'vali', 'test', 'render')

    def __init__(self, config=None, debug: bool = False, *, nets: Optional[Dict] = None, light=None, codebook=None,
                 novel_probes: Optional[Dict] = None, device='cuda'):
        """`nets`: dict name -> (kernels, biases) for the 8 networks of vq_nfr.py:135-164 (checkpoint
        hand-off, :148-155); missing nets get Keras' default init.  `light` [16,32,3] (np_light.npy, :747-759),
        `codebook` [K,Z] cluster centres (cluster_center.npy, :761-766), `novel_probes` name -> [16,32,3]."""
        self.config = _Cfg(config)
        cfg = self.config
        self.debug = debug
        self.device = torch.device(device)
        self.data_type = cfg.get('data_type')
        self.no_brdf_chunk = cfg.getboolean('no_brdf_chunk')
        self.seed = cfg.getint('random_seed')
        self.z_dim = cfg.getint('conv_width')
        self.mlp_chunk = cfg.getint('mlp_chunk')
        self.num_embed = cfg.getint('num_embed')
        self.albedo_slope = cfg.getfloat('albedo_slope')
        self.albedo_bias = cfg.getfloat('albedo_bias')
        self.precision = cfg.get('precision')
        light_h = cfg.getint('light_h')
        self.light_res = (light_h, 2 * light_h)
        if light_h * 2 * light_h != 512:
            raise NotImplementedError('the shading kernel is built for the 16x32 probe (light_h=16)')
        lxyz, lareas = abi.gen_light_xyz(*self.light_res)                        # vq_nfr.py:73
        self.lxyz = torch.as_tensor(lxyz, dtype=torch.float32).to(self.device)   # :74
        self.lareas = torch.as_tensor(lareas, dtype=torch.float32).to(self.device)
        self.embedder = {'xyz': Embedder(incl_input=True, in_dims=3, log2_max_freq=cfg.getint('n_freqs_xyz') - 1,
                                         n_freqs=cfg.getint('n_freqs_xyz'))}     # shape.py:82-89
        self.net = self._init_net(nets or {})
        commitment = cfg.getfloat('commitment_cost')
        self.vq_layer = VectorQuantizerEMA(embedding_dim=self.z_dim, num_embeddings=self.num_embed,
                                           commitment_cost=commitment, seed=self.seed, device=self.device)
        # learned tensors (tf.Variable in the reference): caller-visible, checkpointable
        if light is None:
            light = np.full(self.light_res + (3,), 0.5, np.float32)
        self._light = torch.as_tensor(np.asarray(light), dtype=torch.float32).to(self.device).contiguous()
        if codebook is None:
            rng = np.random.RandomState(self.seed)
            codebook = rng.uniform(0, 1, size=(self.num_embed, self.z_dim)).astype(np.float32)
        cb = torch.as_tensor(np.asarray(codebook), dtype=torch.float32)
        if cb.shape != (self.num_embed, self.z_dim):
            raise ValueError('cluster centres must be [num_embed, z_dim]')
        self._codebook = cb.t().contiguous().to(self.device)                     # [Z,K], :764
        self._gamma_index = torch.ones((1,), dtype=torch.float32)
        self._gamma_bias = torch.ones((1,), dtype=torch.float32)
        self.novel_probes = OrderedDict()
        for k, v in (novel_probes or {}).items():
            self.novel_probes[k] = torch.as_tensor(np.asarray(v), dtype=torch.float32).to(self.device).contiguous()
        self.novel_olat = OrderedDict()
        # optional stage timing (bench.py): list of (name, cuda Event) appended by fast_render when not None
        self.stage_events = None

    def _mark(self, name, device):
        if self.stage_events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(device))
            self.stage_events.append((name, ev))

    # ------------------------------------------------------------------ nets (vq_nfr.py:135-164)
    def _init_net(self, given: Dict):
        z = self.z_dim
        w = self.config.getint('mlp_width')
        emb = self.embedder['xyz'].out_dims
        spec = {
            'fine_enc': (emb, [w] * 4, ['relu'] * 4, [2]),                       # nfr_unit.py:121
            'bottleneck': (w, [w, z, z], [None, 'relu', 'sigmoid'], None),       # nfr_unit.py:122
            'diff_main': (z, [z, z // 2, 3], ['relu'] * 2 + ['sigmoid'], [1]),   # nfr_unit.py:115-116
            'spec_main': (z, [z, z // 2, 1], ['relu'] * 2 + ['sigmoid'], [1]),
            'rough_main': (z, [z, z // 2, 1], ['relu'] * 2 + ['sigmoid'], [1]),
            'diff_vq': (z, [z, z // 2, 3], ['relu'] * 2 + ['sigmoid'], [1]),     # vq_nfr.py:141-146
            'spec_vq': (z, [z, z // 2, 3], ['relu'] * 2 + ['sigmoid'], [1]),
            'rough_vq': (z, [z, z // 2, 1], ['relu'] * 2 + ['sigmoid'], [1]),
        }
        net = {}
        for i, (name, (in_dim, widths, act, skip)) in enumerate(spec.items()):
            if name in given:
                ks, bs = given[name]
                net[name] = mlp.Network.from_arrays(ks, bs, act, skip_at=skip, device=self.device)
                if net[name].widths != widths or net[name].packed.in_dim != in_dim:
                    raise ValueError('%s: weights do not match the architecture %s' % (name, widths))
            else:
                net[name] = mlp.Network(widths, act=act, skip_at=skip, device=self.device, seed=self.seed + i)
                net[name].build(in_dim)
        return net

    def _validate_mode(self, mode):
        if mode not in self.MODES:                                               # base.py:106-110
            raise ValueError(mode)

    # ------------------------------------------------------------------ properties
    @property
    def gamma(self):
        """vq_nfr.py:736-745: [gamma_bias, clip(gamma_index, 0, 5)]"""
        idx = float(min(max(float(self._gamma_index[0]), 0.0), 5.0))
        return (float(self._gamma_bias[0]), idx)

    @property
    def light(self):
        """vq_nfr.py:747-759.  The clip(_light, 0, inf) is applied inside the shading kernel (clip_light0)."""
        return self._light

    def get_codebook(self):
        """vq_nfr.py:761-769"""
        return abi.get_codebook(self._codebook)

    # ------------------------------------------------------------------ fine-grained entry points
    def _calc_ldir(self, pts):
        """shape.py:103-110 (materialising; debug sizes).  [N,512,3]"""
        from ..util.math import safe_l2_normalize
        surf2l = self.lxyz.reshape(1, -1, 3) - pts[:, None, :]
        return safe_l2_normalize(surf2l.contiguous(), axis=2)

    @staticmethod
    def _calc_vdir(cam_loc, pts):
        """shape.py:112-119"""
        return abi.l2_normalize_rows((cam_loc - pts).contiguous())

    def _normal_correct(self, normal, surf2c):
        """vq_nfr.py:830-833"""
        cos = (normal * surf2c).sum(-1, keepdim=True)
        return torch.where(cos >= 0, normal, -normal)

    def _pred_enc_at(self, pts):
        """vq_nfr.py:771-784"""
        z = abi.pred_enc_at(self.net['fine_enc'].packed, self.net['bottleneck'].packed, self.embedder['xyz'].n_freqs,
                            pts, precision=self.precision)
        self._check_numerics(pts.device)
        return z

    def _pred_diff_at(self, z, vq=False):
        """vq_nfr.py:786-804"""
        net = self.net['diff_vq' if vq else 'diff_main']
        out = abi.pred_heads(net.packed, None, None, z, self.albedo_slope, self.albedo_bias, self.precision)[0]
        self._check_numerics(z.device)
        return out

    def _pred_spec_at(self, z, vq=False):
        """vq_nfr.py:806-816"""
        net = self.net['spec_vq' if vq else 'spec_main']
        out = abi.pred_heads(None, net.packed, None, z, precision=self.precision)[1]
        self._check_numerics(z.device)
        return out

    def _pred_rough_at(self, z, vq=False):
        """vq_nfr.py:818-828"""
        net = self.net['rough_vq' if vq else 'rough_main']
        out = abi.pred_heads(None, None, net.packed, z, precision=self.precision)[2]
        self._check_numerics(z.device)
        return out

    def _eval_brdf_at(self, pts2l, pts2c, normal, albedo, spec, rough, chunk_size=None):
        """vq_nfr.py:835-874 (chunking is arithmetic-neutral; ignored)"""
        return abi.eval_brdf(pts2l, pts2c, normal, albedo, spec, rough)

    def _render(self, brdf, l, n, light_vis=None, relight_olat=False, relight_probes=False, dst_env=None):
        """vq_nfr.py:694-733; relight_olat is accepted and ignored exactly as in the reference (:733)."""
        light = self.light.clamp(min=0) if dst_env is None else self.novel_probes[dst_env]
        gamma = None if self.data_type == 'nerf' else self.gamma
        rgb = abi.render(brdf, l, n, light_vis, self.lareas, light, gamma)
        rgb_probes = None
        if relight_probes:
            rgb_probes = torch.stack([abi.render(brdf, l, n, light_vis, self.lareas, lt, gamma)
                                      for lt in self.novel_probes.values()], dim=1)
        self._check_numerics(brdf.device)
        return rgb, None, rgb_probes

    def _check_numerics(self, device):
        if self.debug:
            L.Context.get(device).check_numerics(L.stream_ptr(device))

    # ------------------------------------------------------------------ helpers
    def _unpack(self, batch, ref_batch):
        if ref_batch:
            if self.data_type == 'nerf':
                id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, _, lvis = batch
            else:
                id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, _ = batch
                lvis = None
        else:
            if self.data_type == 'nerf':
                id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, lvis = batch
            else:
                id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal = batch
                lvis = None
        return id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, lvis

    def _lights(self, relight_probes, dst_env):
        main = self.light if dst_env is None else self.novel_probes[dst_env]
        if relight_probes and len(self.novel_probes) > 0:
            return torch.cat([main.reshape(1, 512, 3)] + [p.reshape(1, 512, 3) for p in self.novel_probes.values()], 0)
        return main.reshape(1, 512, 3)

    def _thres_mask(self, thres):
        if thres is None:
            return None
        return torch.as_tensor(thres, dtype=torch.float32).reshape(1, self.num_embed)

    # ------------------------------------------------------------------ init_z (vq_nfr.py:183-195)
    def init_z(self, batch):
        if self.data_type == 'nerf':
            id_, hw, _, _, _, alpha, pred_alpha, xyz, _, _ = batch
        else:
            id_, hw, _, _, _, alpha, pred_alpha, xyz, _ = batch
        row_idx, n_act = abi.compact_mask(alpha)
        n = int(n_act.item())             # z_pred [n_active, z] is the return value: its SHAPE is the count (as boolean_mask)
        z_pred = abi.pred_enc_at(self.net['fine_enc'].packed, self.net['bottleneck'].packed,
                                 self.embedder['xyz'].n_freqs, xyz, row_idx=row_idx, n=n, precision=self.precision)
        return {'id': id_, 'hw': hw, 'z_pred': z_pred}

    # ------------------------------------------------------------------ fast_embed (vq_nfr.py:209-256)
    def fast_embed(self, batch, mode='train', thres=None, ref_batch=True, roll=None):
        self._validate_mode(mode)
        id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, lvis = self._unpack(batch, ref_batch)
        n_total = alpha.shape[0]
        row_idx, n_act = abi.compact_mask(alpha)
        # Outside training no output shape depends on the foreground count, so there is no host round trip: the compact
        # rows are computed for row_idx[0 .. n_act) with the count read on the device, rows beyond it are never scattered
        # (the reference syncs in tf.boolean_mask).  The EMA statistics of mode == 'train' need the exact row count.
        n, n_dev = (int(n_act.item()), None) if mode == 'train' else (n_total, n_act)
        z_enc = abi.pred_enc_at(self.net['fine_enc'].packed, self.net['bottleneck'].packed,
                                self.embedder['xyz'].n_freqs, xyz, row_idx=row_idx, n=n, n_dev=n_dev,
                                precision=self.precision)
        z_norm = abi.l2_normalize_rows(z_enc)
        codebook = self.get_codebook()
        vq_outs = self.vq_layer(z_norm, codebook, is_training=(mode == 'train'), thres=self._thres_mask(thres),
                                roll=roll, return_encodings=False, return_distances=False)
        embed_ind = (vq_outs['encoding_indices'] + 1).to(torch.float32)
        pred, gt = {'alpha': pred_alpha}, {'alpha': alpha}
        loss_kwargs = {'mode': mode}
        embed = abi.scatter_rows(embed_ind[:, None], row_idx, n_total, n_dev=n_dev, n=n)
        xyz_s = xyz * (alpha[:, :1] > 0).to(xyz.dtype)           # scatter_nd(ind, boolean_mask(xyz)) without the gather
        to_vis = {'id': id_, 'hw': hw, 'embed': embed, 'xyz': xyz_s}
        for k, v in pred.items():
            to_vis['pred_' + k] = v
        for k, v in gt.items():
            to_vis['gt_' + k] = v
        return pred, gt, loss_kwargs, to_vis

    # ------------------------------------------------------------------ fast_render (vq_nfr.py:262-398)
    def fast_render(self, batch, mode='train', relight_olat=False, relight_probes=False, opt_scale=None,
                    edit_mask=None, edit_material=None, ref_batch=False, dst_env=None, gen_embed=False, thres=None,
                    vis_scale=False, roll=None, peer_image=None, _peer_row_off=0):
        """`peer_image` (vqnerf_release_b200.dist.PeerImage, multi-GPU only): the shaded rows [.., 1+P, 3] are also
        stored into the destination ranks' image buffers from inside the shading kernel (fused gather) and the shard's
        background rows are zero-filled there; frame protocol: `peer_image.begin_frame()` before, `peer_image.barrier()`
        after (see dist.PeerImage)."""
        self._validate_mode(mode)
        id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, lvis = self._unpack(batch, ref_batch)
        n_total = alpha.shape[0]
        dev = xyz.device
        nf = self.embedder['xyz'].n_freqs
        # mask = alpha[:,0] > 0 ; boolean_mask -> device-side compaction, no host sync
        self._mark('begin', dev)
        row_idx, n_act = abi.compact_mask(alpha)
        self._mark('compact', dev)
        n_ = self.net
        z_enc, basecolor, ks, rough = abi.mlp_main(
            n_['fine_enc'].packed, n_['bottleneck'].packed, n_['diff_main'].packed, n_['spec_main'].packed,
            n_['rough_main'].packed, nf, xyz, row_idx=row_idx, n_dev=n_act, n=n_total, slope=self.albedo_slope,
            bias=self.albedo_bias, want_z=gen_embed, precision=self.precision)
        embed_ind = None
        if gen_embed:
            n = int(n_act.item())
            codebook = self.get_codebook()
            vq_outs = self.vq_layer(abi.l2_normalize_rows(z_enc[:n]), codebook, is_training=(mode == 'train'),
                                    thres=self._thres_mask(thres), roll=roll, return_encodings=False,
                                    return_distances=False)
            embed_ind = (vq_outs['encoding_indices'] + 1).to(torch.float32)
        self._mark('mlp_main', dev)
        scale_t = None
        if (opt_scale is not None) and (not vis_scale):
            scale_t = torch.as_tensor(np.asarray(opt_scale), dtype=torch.float32).reshape(-1).to(dev)
        albedo, spec, s_albedo, s_spec = abi.material_combine(basecolor, ks, scale_t, n_dev=n_act)
        if edit_mask is not None:                                              # :293-295, 324-330 (edit.py:219,226)
            abi.material_edit(edit_mask, edit_material, row_idx, n_act, albedo, spec, rough, scale_t, s_albedo, s_spec)
        gamma = None if self.data_type == 'nerf' else self.gamma
        lights = self._lights(relight_probes, dst_env)
        self._mark('combine', dev)
        sh = abi.shade(xyz, rayo, normal, lvis, s_albedo, s_spec, rough, self.lxyz, self.lareas,
                       lights, row_idx=row_idx, n_dev=n_act, n=n_total,
                       n_total=n_total, to_srgb=(self.data_type == 'nerf'), gamma=gamma,
                       clip_light0=(dst_env is None),
                       peer_ptrs=None if peer_image is None else peer_image.peer_ptrs,
                       peer_row0=0 if peer_image is None else peer_image.row0 + _peer_row_off)
        if peer_image is not None:
            peer_image.clear_background(alpha, _peer_row_off)
        self._mark('shade', dev)
        self._check_numerics(dev)
        loss_kwargs = {'mode': mode, 'gtc': rgb}
        sc = lambda v: abi.scatter_rows(v, row_idx, n_total, n_dev=n_act, n=n_total)
        if (opt_scale is not None) and vis_scale:
            s = torch.as_tensor(np.asarray(opt_scale), dtype=torch.float32).reshape(1, -1).to(dev)
            basecolor_v = abi.linear2srgb(basecolor) * s
            spec_v = abi.linear2srgb(spec) * s
        else:
            basecolor_v, spec_v = basecolor, spec
        # the four material maps share the row list: one zero fill + one scatter launch instead of four of each
        m_base, m_alb, m_spec, m_rough = abi.scatter_rows_multi([basecolor_v, albedo, spec_v, rough], row_idx, n_total,
                                                                n_dev=n_act, n=n_total)
        pred = {'alpha': pred_alpha, 'basecolor': m_base, 'albedo': m_alb, 'spec': m_spec, 'rough': m_rough}
        if gen_embed:
            pred['embed'] = abi.scatter_rows(embed_ind[:, None], row_idx, n_total, n=embed_ind.shape[0])
        rgb_all = sh['rgb']                                       # [n_total, 1+P, 3]
        if dst_env is not None:
            pred['rgb'] = rgb_all[:, 0, :]
        if relight_probes and len(self.novel_probes) > 0:
            pred['rgb_probes'] = rgb_all[:, 1:, :]
        # gt rgb is masked like everything else (scatter_nd(ind, boolean_mask(rgb)), :359)
        gt = {'rgb': rgb * (alpha[:, :1] > 0).to(rgb.dtype), 'alpha': alpha}
        self._mark('scatter', dev)
        to_vis = {'id': id_, 'hw': hw}
        for k, v in pred.items():
            to_vis['pred_' + k] = v
        for k, v in gt.items():
            to_vis['gt_' + k] = v
        return pred, gt, loss_kwargs, to_vis

    # ------------------------------------------------------------------ host-buffer entry point
    def fast_render_host(self, batch, n_chunks=8, out=None, **kw):
        """`fast_render` for a batch whose tensors live in (pinned) HOST memory, as the reference's tf.data pipeline
        delivers them (datasets/shape_unit.py:93-110): the view is cut into `n_chunks` contiguous row blocks that are
        copied, shaded and copied back on two alternating CUDA streams, so the H2D transfer of block i+1 overlaps the
        kernels of block i.  Returns the `pred` dict with pinned host tensors (`out` may pass a previous result to
        reuse its buffers).  Same keyword arguments as fast_render."""
        ref_batch = kw.get('ref_batch', False)
        id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, lvis = self._unpack(batch, ref_batch)
        n_total = alpha.shape[0]
        dev = self.device
        if not hasattr(self, '_host_streams'):
            self._host_streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        streams = self._host_streams
        main = torch.cuda.current_stream(dev)
        for s in streams:
            s.wait_stream(main)
        bounds = [(n_total * i) // n_chunks for i in range(n_chunks + 1)]
        res = out if out is not None else {}
        zeros2 = torch.zeros((1, 2), dtype=torch.int32, device=dev)
        for ci in range(n_chunks):
            a, b = bounds[ci], bounds[ci + 1]
            if b <= a:
                continue
            st = streams[ci % 2]
            with torch.cuda.stream(st):
                up = lambda t: None if t is None else t[a:b].to(dev, non_blocking=True)
                d_rayo, d_alpha, d_xyz, d_normal, d_lvis = up(rayo), up(alpha), up(xyz), up(normal), up(lvis)
                # rgb / pred_alpha are pass-through ground truth: they stay on the host
                dummy3 = d_xyz
                sub = [id_, zeros2, d_rayo, dummy3, dummy3, d_alpha, d_alpha, d_xyz, d_normal]
                if ref_batch:
                    sub.append(dummy3)
                if self.data_type == 'nerf':
                    sub.append(d_lvis)
                pred, _, _, _ = self.fast_render(tuple(sub), _peer_row_off=a, **kw) if kw.get('peer_image') is not None \
                    else self.fast_render(tuple(sub), **kw)
                for k, v in pred.items():
                    if k == 'alpha' or not torch.is_tensor(v):
                        continue
                    if k not in res:
                        res[k] = torch.empty((n_total,) + tuple(v.shape[1:]), dtype=v.dtype).pin_memory()
                    res[k][a:b].copy_(v if v.is_contiguous() else v.contiguous(), non_blocking=True)
        for s in streams:
            main.wait_stream(s)
        res['alpha'] = pred_alpha
        return res

    # ------------------------------------------------------------------ vis_mat (vq_nfr.py:400-465)
    def vis_mat(self, batch, mode='train', opt_scale=None, ref_batch=False, thres=None, roll=None):
        self._validate_mode(mode)
        id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, lvis = self._unpack(batch, ref_batch)
        n_total = alpha.shape[0]
        row_idx, n_act = abi.compact_mask(alpha)
        n, n_dev = (int(n_act.item()), None) if mode == 'train' else (n_total, n_act)      # (see fast_embed)
        z_enc = abi.pred_enc_at(self.net['fine_enc'].packed, self.net['bottleneck'].packed,
                                self.embedder['xyz'].n_freqs, xyz, row_idx=row_idx, n=n, n_dev=n_dev,
                                precision=self.precision)
        codebook = self.get_codebook()
        vq_outs = self.vq_layer(abi.l2_normalize_rows(z_enc), codebook, is_training=(mode == 'train'),
                                thres=self._thres_mask(thres), roll=roll, return_encodings=False,
                                return_distances=False)
        embed_ind = (vq_outs['encoding_indices'] + 1).to(torch.float32)
        basecolor, ks, rough = abi.pred_heads(self.net['diff_main'].packed, self.net['spec_main'].packed,
                                              self.net['rough_main'].packed, z_enc, self.albedo_slope,
                                              self.albedo_bias, self.precision, n_dev=n_dev)
        albedo, spec, _, _ = abi.material_combine(basecolor, ks, n_dev=n_dev)
        sc = lambda v: abi.scatter_rows(v, row_idx, n_total, n_dev=n_dev, n=n)
        pred = {'alpha': pred_alpha, 'albedo': sc(albedo), 'spec': sc(spec), 'rough': sc(rough),
                'embed': sc(embed_ind[:, None])}
        gt = {'alpha': alpha}
        to_vis = {'id': id_, 'hw': hw}
        for k, v in pred.items():
            to_vis['pred_' + k] = v
        for k, v in gt.items():
            to_vis['gt_' + k] = v
        return pred, gt, {'mode': mode}, to_vis

    # ------------------------------------------------------------------ vq_test (vq_nfr.py:467-532)
    def vq_test(self, batch, mode='vali', thres=None, roll=None):
        id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, lvis = self._unpack(batch, False)
        n_total = alpha.shape[0]
        row_idx, n_act = abi.compact_mask(alpha)
        n = int(n_act.item())
        z_enc = abi.pred_enc_at(self.net['fine_enc'].packed, self.net['bottleneck'].packed,
                                self.embedder['xyz'].n_freqs, xyz, row_idx=row_idx, n=n, precision=self.precision)
        codebook = self.get_codebook()
        vq_outs = self.vq_layer(abi.l2_normalize_rows(z_enc), codebook, is_training=(mode == 'train'),
                                thres=self._thres_mask(thres), roll=roll, return_encodings=False,
                                return_distances=False)
        z_vq = vq_outs['quantize']
        cnt = torch.bincount(vq_outs['encoding_indices'], minlength=self.num_embed)
        usage = (cnt > 0).to(torch.float32)[None, :]                          # :503
        vq_albedo, vq_spec, vq_rough = abi.pred_heads(self.net['diff_vq'].packed, self.net['spec_vq'].packed,
                                                      self.net['rough_vq'].packed, z_vq, self.albedo_slope,
                                                      self.albedo_bias, self.precision)
        gamma = None if self.data_type == 'nerf' else self.gamma
        sh = abi.shade(xyz, rayo, normal, lvis, vq_albedo, vq_spec, vq_rough, self.lxyz, self.lareas,
                       self._lights(False, None), row_idx=row_idx, n=n, n_total=n_total, gamma=gamma)
        idx = row_idx[:n].long()
        vq_rgb = sh['rgb'][:, 0, :].index_select(0, idx)
        rgb_c = rgb.index_select(0, idx)
        loss_kwargs = {'vqloss': vq_outs['loss'], 'vqrgb': vq_rgb, 'mode': mode, 'gtc': rgb_c, 'rgb': vq_rgb,
                       'usage': usage}
        return {'alpha': pred_alpha}, {'alpha': alpha}, loss_kwargs, {'id': id_, 'hw': hw}

    # ------------------------------------------------------------------ call (vq_nfr.py:534-692)
    def call(self, batch, mode='train', thres=None, full_vis=False, roll=None):
        self._validate_mode(mode)
        id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, lvis = self._unpack(batch, False)
        n_total = alpha.shape[0]
        nf = self.embedder['xyz'].n_freqs
        row_idx, n_act = abi.compact_mask(alpha)
        n = int(n_act.item())     # the VQ statistics need the exact row count (reference: boolean_mask sync)
        z_enc = abi.pred_enc_at(self.net['fine_enc'].packed, self.net['bottleneck'].packed, nf, xyz,
                                row_idx=row_idx, n=n, precision=self.precision)
        codebook = self.get_codebook()
        vq_outs = self.vq_layer(abi.l2_normalize_rows(z_enc), codebook, is_training=(mode == 'train'),
                                thres=self._thres_mask(thres), roll=roll, return_encodings=False,
                                return_distances=False)
        z_vq = vq_outs['quantize']
        embed_ind = (vq_outs['encoding_indices'] + 1).to(torch.float32)
        if mode == 'train':
            self._codebook.copy_(vq_outs['update'])                              # :582-583
        basecolor, ks, rough = abi.pred_heads(self.net['diff_main'].packed, self.net['spec_main'].packed,
                                              self.net['rough_main'].packed, z_enc, self.albedo_slope,
                                              self.albedo_bias, self.precision)
        albedo, spec, _, _ = abi.material_combine(basecolor, ks)
        gamma = None if self.data_type == 'nerf' else self.gamma
        lights = self._lights(False, None)
        sh = abi.shade(xyz, rayo, normal, lvis, albedo, spec, rough, self.lxyz, self.lareas, lights,
                       row_idx=row_idx, n=n, n_total=n_total, gamma=gamma, want_split=(mode != 'train'),
                       want_normal=True)
        vq_albedo, vq_spec, vq_rough = abi.pred_heads(self.net['diff_vq'].packed, self.net['spec_vq'].packed,
                                                      self.net['rough_vq'].packed, z_vq, self.albedo_slope,
                                                      self.albedo_bias, self.precision)
        sh_vq = abi.shade(xyz, rayo, normal, lvis, vq_albedo, vq_spec, vq_rough, self.lxyz, self.lareas, lights,
                          row_idx=row_idx, n=n, n_total=n_total, gamma=gamma)
        self._check_numerics(xyz.device)
        idx = row_idx[:n].long()
        rgb_lin_full = sh['rgb'][:, 0, :]
        rgb_pred_c = rgb_lin_full.index_select(0, idx)
        vq_rgb_full = sh_vq['rgb'][:, 0, :]
        vq_rgb_c = vq_rgb_full.index_select(0, idx)
        loss_kwargs = {'vqloss': vq_outs['loss'], 'vqrgb': vq_rgb_c, 'mode': mode, 'gtc': rgb.index_select(0, idx),
                       'rgb': rgb_pred_c, 'spec': spec, 'rough': rough, 'z': z_vq, 'embed': self._codebook}
        sc = lambda v: abi.scatter_rows(v, row_idx, n_total, n=n)
        to_srgb = self.data_type == 'nerf'
        fg = (alpha[:, :1] > 0).to(torch.float32)
        rgb_out = abi.linear2srgb(rgb_lin_full) * fg if to_srgb else rgb_lin_full   # srgb(0) == 0 anyway
        pred = {'rgb': rgb_out, 'normal': sh['normal'], 'albedo': sc(albedo), 'alpha': pred_alpha, 'spec': sc(spec),
                'rough': sc(rough), 'ks': sc(ks)}
        if mode != 'train':
            pred['rgb_diff'], pred['rgb_spec'] = sh['rgb_diff'], sh['rgb_spec']
        gt = {'rgb': rgb * fg, 'normal': normal * fg, 'alpha': alpha}
        to_vis = {'id': id_, 'hw': hw}
        if full_vis:
            to_vis['enc_z'] = sc(z_enc)
        if mode != 'train':
            pred['embed'] = sc(embed_ind[:, None])
            pred['vq_rgb'] = abi.linear2srgb(vq_rgb_full) * fg if to_srgb else vq_rgb_full
            pred['vq_albedo'], pred['vq_spec'], pred['vq_rough'] = sc(vq_albedo), sc(vq_spec), sc(vq_rough)
        for k, v in pred.items():
            to_vis['pred_' + k] = v
        for k, v in gt.items():
            to_vis['gt_' + k] = v
        return pred, gt, loss_kwargs, to_vis

    __call__ = call


class GraphedFastRender:
    """`Model.fast_render` (+ the device-side barrier of a fused multi-GPU gather) captured ONCE into a CUDA graph and
    replayed.  A pixel shard of a view at 8 GPUs is ~0.6 ms of kernels: launched one by one from Python the ~15 launches
    of a call cost more than they run; as one graph launch the step is GPU-bound again.  Inputs are copied into static
    buffers (or updated in place by the caller through `static`); the returned dict holds the same static output
    tensors on every call.  With `peer_image`, TWO graphs are captured (one per frame buffer of dist.PeerImage) and
    replayed alternately."""

    def __init__(self, model: Model, example_batch, peer_image=None, **kw):
        self.model, self.peer_image = model, peer_image
        dev = model.device
        self.static = tuple(t.clone() if torch.is_tensor(t) else t for t in example_batch)
        self.kw = kw
        self.stream = torch.cuda.Stream(device=dev)
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        n_graphs = 2 if peer_image is not None else 1
        with torch.cuda.stream(self.stream):              # warm-up on the capture stream: the library's per-stream work
            for _ in range(2 * n_graphs):                 # buffers must exist before the capture (no cudaMalloc inside)
                self._run()
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        torch.cuda.synchronize(dev)
        self.graphs, self.outs = [], []
        for _ in range(n_graphs):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.stream, capture_error_mode='thread_local'):
                out = self._run()
            self.graphs.append(g)
            self.outs.append(out)
        self._i = 0

    def _run(self):
        if self.peer_image is not None:
            self.peer_image.begin_frame()
        pred, gt, lk, vis = self.model.fast_render(self.static, peer_image=self.peer_image, **self.kw)
        img = None
        if self.peer_image is not None:
            self.peer_image.barrier()
            img = self.peer_image.tensor
        return pred, img

    def __call__(self, batch=None):
        """Returns (pred, gathered_image_or_None).  `batch`: tensors to copy into the static inputs (None: the caller
        has already updated `self.static` in place)."""
        if batch is not None:
            for dst, src in zip(self.static, batch):
                if torch.is_tensor(dst) and src is not dst:
                    dst.copy_(src, non_blocking=True)
        i = self._i
        self._i = (i + 1) % len(self.graphs)
        self.graphs[i].replay()
        return self.outs[i]
