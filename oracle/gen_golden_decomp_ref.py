"""Golden vectors for the decomp shading path produced by the REFERENCE'S OWN CODE (TEST INFRASTRUCTURE ONLY).

The reference's decomposition stage is TensorFlow 2.4 (not installable here).  With `oracle/tf_shim` (a `tensorflow` /
`tensorflow_probability` / `sonnet` stand-in that maps the ~70 `tf.*` ops these files call onto torch-CPU) first on
`sys.path`, this script imports the UNMODIFIED modules

    nerfactor/models/{vq_nfr,nfr_unit,shape,base}.py   nerfactor/networks/{mlp,seq,base,embedder,vq_layers}.py
    nerfactor/util/{microfacet,math,img}.py             brdf/renderer.py (+ third_party/xiuminglib for sph2cart)

from /root/reference/decomp/nerfvq_nfr3 and EXECUTES their method bodies -- `Model.fast_render`, `Model.call`
(train and vali), `Model.vq_test`, `Model.compute_loss`, `Model._render`, `Model.get_codebook`,
`VectorQuantizerEMA.__call__`, `microfacet.get_brdf`, `Embedder.__call__`, `img.linear2srgb/srgb2linear`,
`gen_light_xyz` -- on the synthetic scene of `oracle/decomp_oracle.py::synth_scene / synth_batch`.  Only the model's
constructor is bypassed (it reads checkpoints, light probes and cluster centres from disk): the instance is created with
`Model.__new__` and given the attributes `__init__` would set (networks built by the reference's own `mlp.Network`,
`_init_embedder`, `VectorQuantizerEMA`, the shipped `config/vq_nfr.ini`).

The code runs twice: with `tf.float32 := torch.float64` (recorded: the high-precision target the float64 oracle must
match to ~1e-10 and the kernels to the north-star tolerances) and with `tf.float32 := torch.float32` (the fp32 op
sequence; recorded for the emulation check).  Gradients come from torch autograd through the same executed code
(`tf.stop_gradient` = detach, `clip_by_value_preserve_gradient` = tfp's published definition).

    python oracle/gen_golden_decomp_ref.py      -> tests/golden/decomp_ref.npz      (needs /root/reference)
"""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('VQN_REFERENCE', '/root/reference') + '/decomp/nerfvq_nfr3'
OUT = os.path.join(HERE, '..', 'tests', 'golden', 'decomp_ref.npz')
sys.dont_write_bytecode = True
sys.path[:0] = [os.path.join(HERE, 'tf_shim'), REF, os.path.join(REF, 'nerfactor'), os.path.join(HERE, '..')]

import tensorflow as tf  # noqa: E402  (the shim)
from oracle import decomp_oracle as O  # noqa: E402


def import_reference():
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        from nerfactor.models.vq_nfr import Model
        from nerfactor.networks import mlp
        from nerfactor.networks.vq_layers import VectorQuantizerEMA
        from nerfactor.networks.embedder import Embedder
        from nerfactor.util import microfacet, math as mathutil, img as imgutil, io as ioutil
        from brdf.renderer import gen_light_xyz
    return dict(Model=Model, mlp=mlp, VQ=VectorQuantizerEMA, Embedder=Embedder, microfacet=microfacet,
                mathutil=mathutil, imgutil=imgutil, ioutil=ioutil, gen_light_xyz=gen_light_xyz)


def build_model(R, scene: O.Scene, data_type='nerf'):
    """What Model.__init__ (vq_nfr.py:41-129) -> ShapeModel.__init__ (shape.py:34-57) -> BaseModel.__init__ would set,
    minus everything read from disk."""
    Model, mlp = R['Model'], R['mlp']
    cfg = R['ioutil'].read_config(os.path.join(REF, 'nerfactor', 'config', 'vq_nfr.ini'))
    m = Model.__new__(Model)
    m.config = cfg
    m.debug = False
    m.data_type = data_type
    m.no_brdf_chunk = cfg.getboolean('DEFAULT', 'no_brdf_chunk', fallback=True)
    m.seed = cfg.getint('DEFAULT', 'random_seed')
    m.z_dim = cfg.getint('DEFAULT', 'conv_width')
    m.white_bg = cfg.getboolean('DEFAULT', 'white_bg')
    m.mlp_chunk = cfg.getint('DEFAULT', 'mlp_chunk')
    m.brdf_chunk_size = cfg.getint('DEFAULT', 'brdf_chunk_size', fallback=50000)
    m.embedder = m._init_embedder()                                            # shape.py:71-101 executed
    m.num_embed = scene.codebook.shape[1]
    mlp_width, zd = cfg.getint('DEFAULT', 'mlp_width'), m.z_dim
    net = {}                                                                   # vq_nfr.py:141-155, nfr_unit.py:110-122
    for name, out in (('diff_vq', 3), ('spec_vq', 3), ('rough_vq', 1), ('diff_main', 3), ('spec_main', 1),
                      ('rough_main', 1)):
        net[name] = mlp.Network([zd, zd // 2, out], act=['relu'] * 2 + ['sigmoid'], skip_at=[1])
    net['fine_enc'] = mlp.Network([mlp_width] * 4, act=['relu'] * 4, skip_at=[2])
    net['bottleneck'] = mlp.Network([mlp_width] + [zd] * 2, act=[None, 'relu', 'sigmoid'])
    for name, n in scene.nets.items():
        assert len(net[name].layers) == len(n.weights)
        for layer, w, b in zip(net[name].layers, n.weights, n.biases):
            layer.set_weights([w, b])
    m.net = net
    m.vq_layer = R['VQ'](embedding_dim=zd, num_embeddings=m.num_embed,
                         commitment_cost=cfg.getfloat('DEFAULT', 'commitment_cost'), seed=m.seed)
    light_h = cfg.getint('DEFAULT', 'light_h')
    lxyz, lareas = R['gen_light_xyz'](light_h, 2 * light_h)                     # float64 numpy, vq_nfr.py:73-75
    m.lxyz = tf.convert_to_tensor(lxyz, dtype=tf.float32)                      # vq_nfr.py:74-75 (float32 values)
    m.lareas = tf.convert_to_tensor(lareas, dtype=tf.float32)
    m._light = tf.Variable(scene.light, trainable=True)
    m._codebook = tf.Variable(scene.codebook, trainable=True)
    if data_type != 'nerf':
        m._gamma_bias = tf.Variable(np.array([scene.gamma[0]], np.float32), trainable=True)
        m._gamma_index = tf.Variable(np.array([scene.gamma[1]], np.float32), trainable=True)
    m.novel_olat = OrderedDict()
    m.novel_probes = OrderedDict()
    if scene.probes is not None:
        for i, p in enumerate(scene.probes):
            m.novel_probes['p%d' % i] = tf.convert_to_tensor(p, dtype=tf.float32)
    return m, lxyz, lareas


def make_batch(b, data_type='nerf'):
    t = lambda a: tf.convert_to_tensor(a, dtype=tf.float32)
    n = b['xyz'].shape[0]
    tup = ['view0', torch.zeros((n, 2), dtype=torch.int32), t(b['rayo']), t(b['rayd']), t(b['rgb']), t(b['alpha']),
           t(b['pred_alpha']), t(b['xyz']), t(b['normal'])]
    if data_type == 'nerf':
        tup.append(t(b['lvis']))
    return tuple(tup)


def project(name, g, rec, prefix):
    """Gradient tensors: small ones whole, big ones as (sum, sum of squares, fixed random projection)."""
    g = g.detach().double().numpy()
    if g.size <= 4096:
        rec[prefix + name] = g
    else:
        rng = np.random.RandomState(g.size % 65521)
        rec[prefix + name + '_stats'] = np.array([g.sum(), (g * g).sum(), (g.ravel() * rng.standard_normal(g.size)).sum()])


def run(R, dtype, n, seed, n_probes, rec, tag):
    tf.set_float(dtype)
    try:
        scene = O.synth_scene(seed, n_probes=n_probes, bias_scale=0.05)
        b = O.synth_batch(n, seed, fg_frac=0.8)
        m, lxyz, lareas = build_model(R, scene)
        batch = make_batch(b)
        np_ = lambda v: v.detach().double().numpy() if v.is_floating_point() else v.detach().numpy()
        if tag == 'f64':
            rec['lxyz'], rec['lareas'] = lxyz, lareas
        # ---- fine-grained functions
        xyz = batch[7]
        rec[tag + '_embed'] = np_(m.embedder['xyz'](xyz))
        rec[tag + '_codebook_norm'] = np_(m.get_codebook())
        rec[tag + '_z_enc'] = np_(m._pred_enc_at(xyz))
        x = tf.convert_to_tensor(np.linspace(-0.25, 1.25, 61), dtype=tf.float32)
        rec[tag + '_srgb_in'] = np_(x)
        rec[tag + '_linear2srgb'] = np_(R['imgutil'].linear2srgb(x))
        rec[tag + '_srgb2linear'] = np_(R['imgutil'].srgb2linear(torch.clamp(x, 0, 1)))
        tiny = tf.convert_to_tensor(np.array([[3e-4, 4e-4, 0.], [0., 0., 0.], [3., 4., 0.]]), dtype=tf.float32)
        rec[tag + '_l2n_tiny'] = np_(R['mathutil'].safe_l2_normalize(tiny, axis=1))
        # get_brdf on the first 8 points, all 512 lights (microfacet.py:9-39 executed)
        surf2l = m._calc_ldir(xyz[:8])
        surf2c = m._calc_vdir(batch[2][:8], xyz[:8])
        nrm = m._normal_correct(batch[8][:8], surf2c)
        g = torch.Generator().manual_seed(5)
        alb, f0 = torch.rand((8, 3), generator=g).to(dtype), torch.rand((8, 3), generator=g).to(dtype)
        rgh = torch.rand((8, 1), generator=g).to(dtype)
        brdf, glossy, diffuse = R['microfacet'].get_brdf(surf2l, surf2c, nrm, albedo=alb, rough=rgh, f0=f0)
        rec[tag + '_brdf_in'] = np.concatenate([np_(alb), np_(f0), np_(rgh)], 1)
        rec[tag + '_brdf'], rec[tag + '_brdf_glossy'] = np_(brdf), np_(glossy)
        rgb8, _, probes8 = m._render(brdf, surf2l, nrm, batch[9][:8], relight_probes=True)
        rec[tag + '_render8'], rec[tag + '_render8_probes'] = np_(rgb8), np_(probes8)
        rgb8n, _, _ = m._render(brdf, surf2l, nrm, None)
        rec[tag + '_render8_nolvis'] = np_(rgb8n)
        # ---- fast_render (vq_nfr.py:262-398 executed)
        pred, gt, _, to_vis = m.fast_render(batch, mode='test', relight_probes=True, gen_embed=True, dst_env='p0')
        for k in ('basecolor', 'albedo', 'spec', 'rough', 'rgb', 'rgb_probes', 'embed'):
            rec['%s_fr_%s' % (tag, k)] = np_(pred[k])
        pred, _, _, _ = m.fast_render(batch, mode='test', relight_probes=False, opt_scale=torch.tensor(
            [0.7, 1.1, 1.3], dtype=dtype), dst_env='p1')
        rec[tag + '_fr_scaled_rgb'] = np_(pred['rgb'])
        edit_mask = torch.as_tensor((np.arange(n) % 3 == 0).astype(np.float32)[:, None].repeat(3, 1)).to(dtype)
        edit_material = {'diff': [0.2, 0.5, 0.1], 'spec': [-1.0, 0.0, 0.0], 'rough': [0.35]}
        pred, _, _, _ = m.fast_render(batch, mode='test', edit_mask=edit_mask, edit_material=edit_material,
                                      dst_env='p0')
        rec[tag + '_fr_edit_rgb'] = np_(pred['rgb'])
        rec[tag + '_fr_edit_albedo'] = np_(pred['albedo'])
        rec[tag + '_fr_edit_rough'] = np_(pred['rough'])
        # ---- fast_embed / vq_test with a dropout threshold
        thres = np.array([0.0] * 3 + [0.5] * 12)
        roll = np.linspace(0.05, 0.95, 15)
        tf.random.queue.append(roll.reshape(1, -1))
        _, _, _, to_vis = m.fast_embed(batch, mode='test', thres=thres, ref_batch=False)
        rec[tag + '_fe_embed'] = np_(to_vis['embed'])
        tf.random.queue.append(roll.reshape(1, -1))
        _, _, lk, _ = m.vq_test(batch, mode='vali', thres=thres)
        rec[tag + '_vqtest_vqrgb'], rec[tag + '_vqtest_usage'] = np_(lk['vqrgb']), np_(lk['usage'])
        loss, ld = m.compute_loss({}, {}, **dict(lk, keep_batch=True))
        rec[tag + '_vqtest_loss'] = np_(loss)
        # ---- call(mode='vali') + compute_loss
        pred, gt, lk, _ = m.call(batch, mode='vali')
        for k in ('rgb', 'albedo', 'spec', 'rough', 'ks', 'rgb_diff', 'rgb_spec', 'embed', 'vq_rgb', 'vq_albedo',
                  'vq_spec', 'vq_rough', 'normal'):
            rec['%s_vali_%s' % (tag, k)] = np_(pred[k])
        loss, ld = m.compute_loss(pred, gt, **dict(lk, keep_batch=True))
        rec[tag + '_vali_loss'] = np_(loss)
        for k, v in ld.items():
            rec['%s_vali_ld_%s' % (tag, k)] = np_(v)
        # ---- two training steps: call(mode='train') + compute_loss + gradients (train_nfr.py:562-576)
        rec['thres'], rec['roll'] = thres, roll
        n_fg = int((b['alpha'][:, 0] > 0).sum())
        if n_fg % 2:                               # the smoothness loss pairs consecutive rows (vq_nfr.py:945-954)
            raise SystemExit('pick n/seed with an even number of foreground rows (got %d)' % n_fg)
        global_bs = n_fg // 2                      # n_rays_per_step = number of PAIRS (shape_unit.py:323-325)
        rec['global_bs'] = global_bs
        for step in range(2):
            tf.random.queue.append(roll.reshape(1, -1))
            pred, gt, lk, _ = m.call(batch, mode='train', thres=thres)
            rec['%s_train%d_rgb' % (tag, step)] = np_(pred['rgb'])
            rec['%s_train%d_vqrgb' % (tag, step)] = np_(lk['vqrgb'])
            rec['%s_train%d_z_vq' % (tag, step)] = np_(lk['z'])
            rec['%s_train%d_vqloss' % (tag, step)] = np_(lk['vqloss'])
            rec['%s_train%d_codebook_after' % (tag, step)] = np_(m._codebook)
            per_example, ld = m.compute_loss(pred, gt, **dict(lk, keep_batch=True))
            rec['%s_train%d_per_example' % (tag, step)] = np_(per_example)
            for k, v in ld.items():
                rec['%s_train%d_ld_%s' % (tag, step, k)] = np_(v)
            weighted = torch.sum(per_example) / global_bs          # tf.nn.compute_average_loss
            variables = [('light', m._light), ('codebook', m._codebook)]
            for name in sorted(m.net):
                for li, layer in enumerate(m.net[name].layers):
                    variables += [('%s_w%d' % (name, li), layer.kernel), ('%s_b%d' % (name, li), layer.bias)]
            grads = torch.autograd.grad(weighted, [v for _, v in variables], allow_unused=True)
            rec['%s_train%d_loss' % (tag, step)] = np_(weighted)
            for (name, _), g_ in zip(variables, grads):
                if name == 'codebook':
                    # tf.sqrt at the K diagonal zeros of the pairwise distances (vq_nfr.py:962): 0 * 0.5/0 = NaN, in
                    # TF as in torch -- recorded as a fact about the reference (oracle/decomp_oracle.py::sim_loss)
                    rec['%s_train%d_dcodebook_nan_frac' % (tag, step)] = float(torch.isnan(g_).double().mean())
                    continue
                project(name, g_, rec, '%s_train%d_d_' % (tag, step))
        rec[tag + '_ema_cluster_hidden'] = np_(m.vq_layer.ema_cluster_size._hidden)
        rec[tag + '_ema_dw_average'] = np_(m.vq_layer.ema_dw.average)
    finally:
        tf.set_float(torch.float32)


def main():
    R = import_reference()
    n, seed, n_probes = 97, 7, 2
    rec = {'n': n, 'seed': seed, 'n_probes': n_probes, 'bias_scale': 0.05, 'fg_frac': 0.8}
    run(R, torch.float64, n, seed, n_probes, rec, 'f64')
    rec32 = {}
    run(R, torch.float32, n, seed, n_probes, rec32, 'f32')
    for k in ('f32_fr_rgb_probes', 'f32_fr_albedo', 'f32_fr_spec', 'f32_fr_rough', 'f32_fr_embed', 'f32_train1_rgb',
              'f32_train1_vqrgb', 'f32_train1_loss', 'f32_train1_codebook_after', 'f32_brdf', 'f32_z_enc'):
        rec[k] = rec32[k].astype(np.float32) if rec32[k].dtype.kind == 'f' else rec32[k]
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **rec)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes,', len(rec), 'arrays')


if __name__ == '__main__':
    main()
