"""Golden vectors for nfr_unit.Model (the warm-up model of the decomposition stage), produced by the REFERENCE'S OWN CODE:
the unmodified `nerfactor/models/nfr_unit.py::Model.call / gen_z` executed on torch-CPU float64 through the
`oracle/tf_shim` TensorFlow stand-in (see oracle/gen_golden_decomp_ref.py; only the constructor, which reads the light
probes and test lights from disk, is bypassed with `Model.__new__`).

    python oracle/gen_golden_nfr_unit.py       -> tests/golden/nfr_unit_ref.npz       (needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden_decomp_ref as G  # noqa: E402  (sets up sys.path: tf_shim, the reference tree, the repo root)

tf, O, REF = G.tf, G.O, G.REF
OUT = os.path.join(HERE, '..', 'tests', 'golden', 'nfr_unit_ref.npz')


def build(scene):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        from nerfactor.models.nfr_unit import Model
        from nerfactor.networks import mlp
        from nerfactor.util import io as ioutil
        from brdf.renderer import gen_light_xyz
    cfg = ioutil.read_config(os.path.join(REF, 'nerfactor', 'config', 'vq_nfr.ini'))
    m = Model.__new__(Model)                             # nfr_unit.py:44-104 minus everything read from disk
    m.config, m.debug, m.data_type = cfg, False, 'nerf'
    m.no_brdf_chunk, m.z_dim = True, 256
    m.white_bg, m.mlp_chunk = cfg.getboolean('DEFAULT', 'white_bg'), cfg.getint('DEFAULT', 'mlp_chunk')
    m.brdf_chunk_size = 50000
    m.embedder = m._init_embedder()
    w = cfg.getint('DEFAULT', 'mlp_width')
    net = {'diff_out': mlp.Network([256, 128, 3], act=['relu'] * 2 + ['sigmoid'], skip_at=[1]),      # :114-119
           'spec_out': mlp.Network([256, 128, 1], act=['relu'] * 2 + ['sigmoid'], skip_at=[1]),
           'rough_out': mlp.Network([256, 128, 1], act=['relu'] * 2 + ['sigmoid'], skip_at=[1]),
           'fine_enc': mlp.Network([w] * 4, act=['relu'] * 4, skip_at=[2]),                          # :121
           'bottleneck': mlp.Network([w] + [256] * 2, act=[None, 'relu', 'sigmoid'])}                # :122
    src = {'diff_out': 'diff_main', 'spec_out': 'spec_main', 'rough_out': 'rough_main', 'fine_enc': 'fine_enc',
           'bottleneck': 'bottleneck'}
    for name, key in src.items():
        n_ = scene.nets[key]
        for layer, wt, b in zip(net[name].layers, n_.weights, n_.biases):
            layer.set_weights([wt, b])
    m.net = net
    lxyz, lareas = gen_light_xyz(16, 32)
    m.lxyz = tf.convert_to_tensor(lxyz, dtype=tf.float32)
    m.lareas = tf.convert_to_tensor(lareas, dtype=tf.float32)
    m._light = tf.Variable(scene.light, trainable=True)               # read through the clipping `light` property (:320-327)
    m._gamma_index = None
    return m


def main():
    n, seed = 57, 17
    rec = {'n': n, 'seed': seed, 'bias_scale': 0.05, 'fg_frac': 0.7}
    tf.set_float(torch.float64)
    try:
        scene = O.synth_scene(seed, bias_scale=0.05)
        scene.light = scene.light - 0.3                              # some negative texels: the light property clips them
        rec['light_shift'] = -0.3
        b = O.synth_batch(n, seed, fg_frac=0.7)
        m = build(scene)
        t = lambda a: tf.convert_to_tensor(a, dtype=tf.float32)
        batch = ('view0', torch.zeros((n, 2), dtype=torch.int32), t(b['rayo']), t(b['rayd']), t(b['rgb']), t(b['alpha']),
                 t(b['pred_alpha']), t(b['xyz']), t(b['normal']), t(b['lvis']))
        np_ = lambda v: v.detach().double().numpy()
        pred, gt, lk, _ = m.call(batch, mode='vali')
        for k in ('rgb', 'normal', 'albedo', 'spec', 'rough', 'ks', 'basecolor', 'xyz', 'rgb_spec', 'rgb_diff'):
            rec['vali_' + k] = np_(pred[k])
        for k in ('rgb', 'spec', 'rough', 'gtc'):
            rec['vali_lk_' + k] = np_(lk[k])
        pred, _, lk, _ = m.call(batch, mode='train')
        rec['train_rgb'], rec['train_lk_rgb'] = np_(pred['rgb']), np_(lk['rgb'])
        assert 'rgb_spec' not in pred
        tv = m.gen_z(batch, genz=True)
        for k in ('albedo', 'spec', 'rough', 'z_bias'):
            rec['genz_' + k] = np_(tv[k])
    finally:
        tf.set_float(torch.float32)
    np.savez_compressed(OUT, **rec)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
