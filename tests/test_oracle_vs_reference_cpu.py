"""The decomp oracle against vectors produced by the REFERENCE'S OWN CODE (tests/golden/decomp_ref.npz).

The vectors come from `oracle/gen_golden_decomp_ref.py`: the unmodified reference modules (models/vq_nfr.py,
networks/{mlp,embedder,vq_layers}.py, util/{microfacet,math,img}.py, brdf/renderer.py) executed on torch-CPU float64
through the `oracle/tf_shim` TensorFlow stand-in.  These tests pin `oracle/decomp_oracle.py` (float64) to them at
~1e-9: same synthetic scene (`synth_scene(7)`, `synth_batch(97, 7)`), same dropout roll.
"""
import os

import numpy as np
import pytest
import torch

from oracle import decomp_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'decomp_ref.npz')
TOL = dict(rtol=1e-9, atol=1e-11)


@pytest.fixture(scope='module')
def ref():
    return np.load(GOLD)


@pytest.fixture(scope='module')
def setup(ref):
    scene = O.synth_scene(int(ref['seed']), n_probes=int(ref['n_probes']), bias_scale=float(ref['bias_scale']))
    batch = O.synth_batch(int(ref['n']), int(ref['seed']), fg_frac=float(ref['fg_frac']))
    return scene, batch


def test_light_geometry_matches_reference(ref):
    xyz, areas = O.gen_light_xyz(16, 32)
    np.testing.assert_allclose(xyz, ref['lxyz'], rtol=1e-13, atol=1e-11)
    np.testing.assert_allclose(areas, ref['lareas'], rtol=1e-13)


def test_primitives_match_reference(ref, setup):
    scene, batch = setup
    d = torch.float64
    xyz = torch.as_tensor(batch['xyz'], dtype=d)
    np.testing.assert_allclose(O.embed(xyz).numpy(), ref['f64_embed'], **TOL)
    np.testing.assert_allclose(O.get_codebook(torch.as_tensor(scene.codebook, dtype=d)).numpy(),
                               ref['f64_codebook_norm'], **TOL)
    np.testing.assert_allclose(O.pred_enc_at(scene.nets, xyz).numpy(), ref['f64_z_enc'], **TOL)
    x = torch.as_tensor(ref['f64_srgb_in'])
    np.testing.assert_allclose(O.linear2srgb(x).numpy(), ref['f64_linear2srgb'], **TOL)
    np.testing.assert_allclose(O.srgb2linear(torch.clamp(x, 0, 1)).numpy(), ref['f64_srgb2linear'], **TOL)
    tiny = torch.as_tensor(np.array([[3e-4, 4e-4, 0.], [0., 0., 0.], [3., 4., 0.]], np.float32)).to(d)
    np.testing.assert_allclose(O.safe_l2_normalize(tiny, 1).numpy(), ref['f64_l2n_tiny'], **TOL)


def test_get_brdf_and_render_match_reference(ref, setup):
    scene, batch = setup
    d = torch.float64
    t = lambda a: torch.as_tensor(a, dtype=d)
    xyz, rayo, normal = t(batch['xyz'][:8]), t(batch['rayo'][:8]), t(batch['normal'][:8])
    lxyz, lareas = t(scene.lxyz.astype(np.float32)), t(scene.lareas.astype(np.float32))
    surf2l, surf2c = O.calc_ldir(lxyz, xyz), O.calc_vdir(rayo, xyz)
    nrm = O.normal_correct(normal, surf2c)
    mat = t(ref['f64_brdf_in'])
    brdf, glossy, _ = O.get_brdf(surf2l, surf2c, nrm, mat[:, 0:3], mat[:, 6:7], mat[:, 3:6])
    np.testing.assert_allclose(brdf.numpy(), ref['f64_brdf'], **TOL)
    np.testing.assert_allclose(glossy.numpy(), ref['f64_brdf_glossy'], **TOL)
    light = torch.clamp(t(scene.light), min=0)
    rgb, probes = O.render(brdf, surf2l, nrm, lareas, light, t(batch['lvis'][:8]), t(scene.probes))
    np.testing.assert_allclose(rgb.numpy(), ref['f64_render8'], **TOL)
    np.testing.assert_allclose(probes.numpy(), ref['f64_render8_probes'], **TOL)
    rgb, _ = O.render(brdf, surf2l, nrm, lareas, light, None, None)
    np.testing.assert_allclose(rgb.numpy(), ref['f64_render8_nolvis'], **TOL)


def test_fast_render_matches_reference(ref, setup):
    scene, batch = setup
    o = O.fast_render(scene, batch, torch.float64, relight_probes=True, gen_embed=True, dst_env=0)
    for k in ('basecolor', 'albedo', 'spec', 'rough', 'rgb', 'rgb_probes'):
        np.testing.assert_allclose(o[k].numpy(), ref['f64_fr_' + k], err_msg=k, **TOL)
    np.testing.assert_array_equal(o['embed'].numpy().astype(np.int64), ref['f64_fr_embed'])
    o = O.fast_render(scene, batch, torch.float64, opt_scale=np.array([0.7, 1.1, 1.3]), dst_env=1)
    np.testing.assert_allclose(o['rgb'].numpy(), ref['f64_fr_scaled_rgb'], **TOL)


def test_fast_render_material_edit_matches_reference(ref, setup):
    scene, batch = setup
    n = int(ref['n'])
    edit_mask = (np.arange(n) % 3 == 0).astype(np.float32)[:, None].repeat(3, 1)
    edit_material = {'diff': [0.2, 0.5, 0.1], 'spec': [-1.0, 0.0, 0.0], 'rough': [0.35]}
    o = O.fast_render(scene, batch, torch.float64, edit_mask=edit_mask, edit_material=edit_material, dst_env=0)
    np.testing.assert_allclose(o['rgb'].numpy(), ref['f64_fr_edit_rgb'], **TOL)
    np.testing.assert_allclose(o['albedo'].numpy(), ref['f64_fr_edit_albedo'], **TOL)
    np.testing.assert_allclose(o['rough'].numpy(), ref['f64_fr_edit_rough'], **TOL)


def test_vq_dropout_paths_match_reference(ref, setup):
    scene, batch = setup
    vq = O.VectorQuantizerEMA(O.Z_DIM, O.NUM_EMBED, O.COMMITMENT_COST, dtype=torch.float64)
    c = O.call_forward(scene, batch, vq, 'vali', thres=ref['thres'], roll=ref['roll'], dtype=torch.float64)
    mask = batch['alpha'][:, 0] > 0
    embed = np.zeros((mask.shape[0], 1), np.int64)
    embed[mask, 0] = c['embed_ind'].numpy()
    np.testing.assert_array_equal(embed, ref['f64_fe_embed'])                      # Model.fast_embed
    np.testing.assert_allclose(c['vq_rgb_linear'].numpy(), ref['f64_vqtest_vqrgb'], **TOL)   # Model.vq_test
    gtc = torch.as_tensor(batch['rgb'][mask], dtype=torch.float64)
    loss, _ = O.compute_loss('vali', gtc, c['vq_rgb_linear'], c['vq_rgb_linear'])
    np.testing.assert_allclose(loss.numpy(), ref['f64_vqtest_loss'], **TOL)
    usage = np.zeros((1, O.NUM_EMBED))
    usage[0, np.unique(c['embed_ind'].numpy() - 1)] = 1
    np.testing.assert_array_equal(usage, ref['f64_vqtest_usage'])


def test_call_vali_matches_reference(ref, setup):
    scene, batch = setup
    vq = O.VectorQuantizerEMA(O.Z_DIM, O.NUM_EMBED, O.COMMITMENT_COST, dtype=torch.float64)
    c = O.call_forward(scene, batch, vq, 'vali', dtype=torch.float64)
    mask = batch['alpha'][:, 0] > 0

    def full(v):
        out = np.zeros((mask.shape[0],) + tuple(v.shape[1:]))
        out[mask] = v.numpy()
        return out
    s = O.linear2srgb
    for k, v in (('rgb', s(c['rgb_linear'])), ('albedo', c['albedo']), ('spec', c['spec']), ('rough', c['rough']),
                 ('ks', c['ks']), ('rgb_diff', c['rgb_diff']), ('rgb_spec', c['rgb_spec']),
                 ('vq_rgb', s(c['vq_rgb_linear'])), ('vq_albedo', c['vq_albedo']), ('vq_spec', c['vq_spec']),
                 ('vq_rough', c['vq_rough']), ('normal', c['normal'])):
        np.testing.assert_allclose(full(v), ref['f64_vali_' + k], err_msg=k, **TOL)
    np.testing.assert_array_equal(full(c['embed_ind'][:, None]).astype(np.int64), ref['f64_vali_embed'])
    gtc = torch.as_tensor(batch['rgb'][mask], dtype=torch.float64)
    loss, ld = O.compute_loss('vali', gtc, c['rgb_linear'], c['vq_rgb_linear'])
    np.testing.assert_allclose(loss.numpy(), ref['f64_vali_loss'], **TOL)
    for k in ('rgb', 'vqrgb', 'chromaticity'):
        np.testing.assert_allclose(ld[k].numpy(), ref['f64_vali_ld_' + k], err_msg=k, **TOL)


def _project(g):
    g = g.detach().double().numpy()
    rng = np.random.RandomState(g.size % 65521)
    return np.array([g.sum(), (g * g).sum(), (g.ravel() * rng.standard_normal(g.size)).sum()])


def test_two_training_steps_match_reference(ref, setup):
    """Model.call(mode='train') + compute_loss + gradients (train_nfr.py:562-576) executed by the reference code, two
    steps (so the EMA debiasing counter and the codebook overwrite matter), against oracle.train_step."""
    scene0, batch = setup
    scene = O.synth_scene(int(ref['seed']), n_probes=0, bias_scale=float(ref['bias_scale']))
    vq = O.VectorQuantizerEMA(O.Z_DIM, O.NUM_EMBED, O.COMMITMENT_COST, dtype=torch.float64)
    gbs = int(ref['global_bs'])
    for step in range(2):
        r = O.train_step(scene, batch, vq, thres=ref['thres'], roll=ref['roll'], global_bs=gbs, dtype=torch.float64)
        p = 'f64_train%d_' % step
        mask = batch['alpha'][:, 0] > 0
        rgb_full = np.zeros((mask.shape[0], 3))
        rgb_full[mask] = O.linear2srgb(r['out']['rgb_linear']).numpy()
        np.testing.assert_allclose(rgb_full, ref[p + 'rgb'], **TOL)
        np.testing.assert_allclose(r['out']['vq_rgb_linear'].numpy(), ref[p + 'vqrgb'], **TOL)
        np.testing.assert_allclose(r['out']['z_vq'].numpy(), ref[p + 'z_vq'], **TOL)
        np.testing.assert_allclose(r['out']['vq_loss'].numpy(), ref[p + 'vqloss'], **TOL)
        np.testing.assert_allclose(r['update'].numpy(), ref[p + 'codebook_after'], **TOL)
        np.testing.assert_allclose(r['per_example'].numpy(), ref[p + 'per_example'], **TOL)
        for k in ('rgb', 'vqrgb', 'vqloss', 'chromaticity', 'chr_smooth', 'sim_smooth', 'lambert', 'loss'):
            np.testing.assert_allclose(r['loss_dict'][k].numpy(), ref[p + 'ld_' + k], err_msg=k, **TOL)
        np.testing.assert_allclose(r['loss'].numpy(), ref[p + 'loss'], **TOL)
        # the reference's own codebook gradient is NaN (tf.sqrt at the diagonal zeros, vq_nfr.py:962) -- the documented
        # deviation of oracle.sim_loss; every other gradient must agree
        assert float(ref[p + 'dcodebook_nan_frac']) > 0
        assert torch.isfinite(r['dcodebook']).all()
        gtol = dict(rtol=1e-7, atol=1e-12)
        np.testing.assert_allclose(r['dlight'].numpy(), ref[p + 'd_light'], **gtol)
        for name, (gw, gb) in r['grads'].items():
            for li, (w, b) in enumerate(zip(gw, gb)):
                for nm, g in (('%s_w%d' % (name, li), w), ('%s_b%d' % (name, li), b)):
                    if g.numel() <= 4096:
                        np.testing.assert_allclose(g.numpy(), ref[p + 'd_' + nm], err_msg=nm, **gtol)
                    else:
                        np.testing.assert_allclose(_project(g), ref[p + 'd_' + nm + '_stats'], err_msg=nm,
                                                   rtol=1e-7, atol=1e-12)
        scene.codebook = r['update'].numpy()            # _codebook.assign(update), kept in float64 like the reference run
    np.testing.assert_allclose(vq.ema_cluster_size.hidden.detach().numpy(), ref['f64_ema_cluster_hidden'], **TOL)
    np.testing.assert_allclose(vq.ema_dw.average.detach().numpy(), ref['f64_ema_dw_average'], **TOL)


def test_float32_emulation_close_to_reference_float32(ref, setup):
    """The same reference code executed in float32 (the TF dtype) vs the oracle's float32 op sequence."""
    scene, batch = setup
    o = O.fast_render(scene, batch, torch.float32, relight_probes=True, gen_embed=True, dst_env=0)
    for k in ('albedo', 'spec', 'rough', 'rgb_probes'):
        np.testing.assert_allclose(o[k].numpy(), ref['f32_fr_' + k], rtol=1e-5, atol=2e-6, err_msg=k)
    np.testing.assert_array_equal(o['embed'].numpy().astype(np.int64), ref['f32_fr_embed'])


def test_training_step_non_nerf_data_matches_reference():
    """data_type != 'nerf' (no light visibility, trainable tone scaling, vq_nfr.py:707, 715-718, 736-745): one training step
    of the reference's own code (oracle/gen_golden_decomp_real.py) against oracle.train_step, including the gradients of
    _gamma_bias / _gamma_index."""
    g = np.load(os.path.join(os.path.dirname(GOLD), 'decomp_real_ref.npz'))
    scene = O.synth_scene(int(g['seed']), bias_scale=float(g['bias_scale']), data_type='real')
    scene.gamma = tuple(float(np.float32(v)) for v in g['gamma'])      # float32 variables in the reference
    batch = O.synth_batch(int(g['n']), int(g['seed']), fg_frac=float(g['fg_frac']), with_lvis=False)
    vq = O.VectorQuantizerEMA(O.Z_DIM, O.NUM_EMBED, O.COMMITMENT_COST, dtype=torch.float64)
    r = O.train_step(scene, batch, vq, thres=g['thres'], roll=g['roll'], global_bs=int(g['global_bs']), dtype=torch.float64)
    mask = batch['alpha'][:, 0] > 0
    rgb_full = np.zeros((mask.shape[0], 3))
    rgb_full[mask] = r['out']['rgb_linear'].numpy()              # no sRGB transform for this data type (:350-356)
    np.testing.assert_allclose(rgb_full, g['train_rgb'], **TOL)
    np.testing.assert_allclose(r['out']['vq_rgb_linear'].numpy(), g['train_vqrgb'], **TOL)
    np.testing.assert_allclose(r['per_example'].numpy(), g['train_per_example'], **TOL)
    np.testing.assert_allclose(r['loss'].numpy(), g['train_loss'], **TOL)
    gtol = dict(rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(r['dgamma'][0].numpy().reshape(-1), g['train_d_gamma_bias'], **gtol)
    np.testing.assert_allclose(r['dgamma'][1].numpy().reshape(-1), g['train_d_gamma_index'], **gtol)
    np.testing.assert_allclose(r['dlight'].numpy(), g['train_d_light'], **gtol)
    for name, (gw, gb) in r['grads'].items():
        for li, b in enumerate(gb):
            np.testing.assert_allclose(b.numpy(), g['train_d_%s_b%d' % (name, li)], err_msg=name, **gtol)
