"""Golden vectors for ref_nfr.Model (the residual model nerfactor/test.py:181-197 renders before the VQ model), produced
by the REFERENCE'S OWN CODE: the unmodified `nerfactor/models/ref_nfr.py::Model.call / fast_render` executed on torch-CPU
float64 through the `oracle/tf_shim` TensorFlow stand-in (see oracle/gen_golden_decomp_ref.py; only the constructor,
which restores checkpoints from disk, is bypassed with `Model.__new__`).

    python oracle/gen_golden_ref_nfr.py       -> tests/golden/ref_nfr_ref.npz       (needs /root/reference)
"""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden_decomp_ref as G  # noqa: E402  (sets up sys.path: tf_shim, the reference tree, the repo root)

tf, O, REF = G.tf, G.O, G.REF
OUT = os.path.join(HERE, '..', 'tests', 'golden', 'ref_nfr_ref.npz')


def ref_batch(n, seed, fg_frac):
    b = O.synth_batch(n, seed, fg_frac=fg_frac)
    b['ref'] = np.random.RandomState(seed + 5).uniform(0, 1, size=(n, 3)).astype(np.float32)
    return b


def build(scene):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        from nerfactor.models.ref_nfr import Model
        from nerfactor.networks import mlp
        from nerfactor.util import io as ioutil
        from brdf.renderer import gen_light_xyz
    cfg = ioutil.read_config(os.path.join(REF, 'nerfactor', 'config', 'vq_nfr.ini'))
    m = Model.__new__(Model)                             # ref_nfr.py:46-131 minus everything read from disk
    m.config, m.debug, m.data_type = cfg, False, 'nerf'
    m.no_brdf_chunk, m.z_dim = True, cfg.getint('DEFAULT', 'conv_width')
    m.white_bg, m.mlp_chunk = cfg.getboolean('DEFAULT', 'white_bg'), cfg.getint('DEFAULT', 'mlp_chunk')
    m.brdf_chunk_size = 50000
    m.embedder = m._init_embedder()
    zd, w = m.z_dim, cfg.getint('DEFAULT', 'mlp_width')
    net = {'fine_enc': mlp.Network([w] * 4, act=['relu'] * 4, skip_at=[2]),
           'bottleneck': mlp.Network([w] + [zd] * 2, act=[None, 'relu', 'sigmoid']),
           'spec_out': mlp.Network([zd, zd // 2, 1], act=['relu'] * 2 + ['sigmoid'], skip_at=[1]),
           'rgb_enc': mlp.Network([zd] + [zd] * 2, act=[None, 'relu', 'sigmoid']),                     # :148
           'diff_out': mlp.Network([zd, zd // 2, 3], act=['relu'] * 2 + ['sigmoid'], skip_at=[1]),     # :149-150
           'rough_out': mlp.Network([zd, zd // 2, 1], act=['relu'] * 2 + ['sigmoid'], skip_at=[1])}    # :151-152
    for name, n_ in scene.nets.items():
        for layer, wt, b in zip(net[name].layers, n_.weights, n_.biases):
            layer.set_weights([wt, b])
    m.net = net
    lxyz, lareas = gen_light_xyz(16, 32)
    m.lxyz = tf.convert_to_tensor(lxyz, dtype=tf.float32)
    m.lareas = tf.convert_to_tensor(lareas, dtype=tf.float32)
    m._light = tf.Variable(scene.light, trainable=True)
    m.light = tf.convert_to_tensor(scene.light, dtype=tf.float32)      # :88: np_light.npy as loaded (no clip in this model)
    m.novel_olat = OrderedDict()
    m.novel_probes = OrderedDict(('p%d' % i, tf.convert_to_tensor(p, dtype=tf.float32)) for i, p in enumerate(scene.probes))
    return m


def main():
    n, seed, n_probes = 61, 11, 2
    rec = {'n': n, 'seed': seed, 'n_probes': n_probes, 'bias_scale': 0.05, 'fg_frac': 0.75}
    tf.set_float(torch.float64)
    try:
        scene = O.synth_scene(seed, n_probes=n_probes, bias_scale=0.05)
        scene.nets = O.make_ref_nfr_nets(seed, 0.05)
        b = ref_batch(n, seed, 0.75)
        m = build(scene)
        t = lambda a: tf.convert_to_tensor(a, dtype=tf.float32)
        batch = ('view0', torch.zeros((n, 2), dtype=torch.int32), t(b['rayo']), t(b['rayd']), t(b['rgb']), t(b['alpha']),
                 t(b['pred_alpha']), t(b['xyz']), t(b['normal']), t(b['ref']), t(b['lvis']))
        np_ = lambda v: v.detach().double().numpy()
        rec['z_ref'] = np_(m._pred_ref_at(batch[9]))
        pred, _, _, _ = m.fast_render(batch, mode='test', relight_probes=True)
        rec['fr_rgb'], rec['fr_rgb_probes'] = np_(pred['rgb']), np_(pred['rgb_probes'])
        pred, _, _, _ = m.fast_render(batch, mode='test', relight_probes=True,
                                      opt_scale=torch.tensor([0.7, 1.1, 1.3], dtype=torch.float64))
        rec['fr_scaled_rgb'], rec['fr_scaled_rgb_probes'] = np_(pred['rgb']), np_(pred['rgb_probes'])
        edit_mask = torch.as_tensor((np.arange(n) % 4 == 1).astype(np.float32)[:, None].repeat(3, 1)).double()
        pred, _, _, _ = m.fast_render(batch, mode='test', relight_probes=True, edit_mask=edit_mask,
                                      edit_material={'diff': [-1.0, 0, 0], 'spec': [0.04, 0.05, 0.06], 'rough': [0.6]})
        rec['fr_edit_rgb'], rec['fr_edit_rgb_probes'] = np_(pred['rgb']), np_(pred['rgb_probes'])
        pred, gt, lk, _ = m.call(batch, mode='vali', relight_probes=True)
        for k in ('rgb', 'normal', 'albedo', 'spec', 'rough', 'ks', 'basecolor', 'rgb_spec', 'rgb_diff', 'rgb_probes'):
            rec['vali_' + k] = np_(pred[k])
        rec['vali_lk_rgb'] = np_(lk['rgb'])
        pred, _, _, _ = m.call(batch, mode='test', opt_scale=torch.tensor([0.7, 1.1, 1.3], dtype=torch.float64))
        rec['test_scaled_rgb'] = np_(pred['rgb'])
    finally:
        tf.set_float(torch.float32)
    np.savez_compressed(OUT, **rec)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
