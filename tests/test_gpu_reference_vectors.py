"""The CUDA path against vectors produced by the REFERENCE'S OWN CODE (tests/golden/decomp_ref.npz).

`oracle/gen_golden_decomp_ref.py` executes the unmodified reference modules (models/vq_nfr.py `fast_render`, `call`,
`vq_test`, `compute_loss`; networks/vq_layers.py; util/microfacet.py; ...) on torch-CPU float64 through the
`oracle/tf_shim` TensorFlow stand-in and records their outputs; the kernels are compared here with those numbers directly
(the oracle is not involved).  Tolerances are the north star's: 1e-4 relative (+ a 5e-6 absolute floor for values near
zero), indices exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import decomp_oracle as O          # synthetic scene / batch generators only
from tests.test_gpu_parity import _batch_tuple, _close, _model_from_scene

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'decomp_ref.npz')


@pytest.fixture(scope='module')
def ref():
    return np.load(GOLD)


@pytest.fixture()
def setup(ref, cuda_dev):
    scene = O.synth_scene(int(ref['seed']), n_probes=int(ref['n_probes']), bias_scale=float(ref['bias_scale']))
    batch = O.synth_batch(int(ref['n']), int(ref['seed']), fg_frac=float(ref['fg_frac']))
    return scene, batch, _model_from_scene(scene, cuda_dev), _batch_tuple(batch, cuda_dev)


def test_fast_render_vs_reference_code(ref, setup, cuda_dev):
    scene, batch, m, bt = setup
    pred, _, _, _ = m.fast_render(bt, mode='test', relight_probes=True, gen_embed=True, dst_env='probe00')
    for k in ('basecolor', 'albedo', 'spec', 'rough', 'rgb', 'rgb_probes'):
        _close(pred[k], ref['f64_fr_' + k], k, rtol=1e-4, atol=5e-6)
    np.testing.assert_array_equal(pred['embed'].cpu().numpy().astype(np.int64), ref['f64_fr_embed'])
    pred, _, _, _ = m.fast_render(bt, mode='test', opt_scale=np.array([0.7, 1.1, 1.3], np.float32), dst_env='probe01')
    _close(pred['rgb'], ref['f64_fr_scaled_rgb'], 'opt_scale rgb', rtol=1e-4, atol=5e-6)


def test_fast_render_material_edit_vs_reference_code(ref, setup, cuda_dev):
    """edit_mask / edit_material (vq_nfr.py:293-295, 324-330; edit.py:219,226)."""
    scene, batch, m, bt = setup
    n = int(ref['n'])
    edit_mask = torch.as_tensor((np.arange(n) % 3 == 0).astype(np.float32)[:, None].repeat(3, 1)).to(cuda_dev)
    edit_material = {'diff': [0.2, 0.5, 0.1], 'spec': [-1.0, 0.0, 0.0], 'rough': [0.35]}
    pred, _, _, _ = m.fast_render(bt, mode='test', edit_mask=edit_mask, edit_material=edit_material, dst_env='probe00')
    _close(pred['rgb'], ref['f64_fr_edit_rgb'], 'edited rgb', rtol=1e-4, atol=5e-6)
    _close(pred['albedo'], ref['f64_fr_edit_albedo'], 'edited albedo', rtol=1e-4, atol=5e-6)
    _close(pred['rough'], ref['f64_fr_edit_rough'], 'edited rough', rtol=1e-4, atol=5e-6)


def test_call_vali_vs_reference_code(ref, setup, cuda_dev):
    scene, batch, m, bt = setup
    pred, gt, lk, _ = m.call(bt, mode='vali')
    for k in ('rgb', 'albedo', 'spec', 'rough', 'ks', 'rgb_diff', 'rgb_spec', 'vq_rgb', 'vq_albedo', 'vq_spec',
              'vq_rough', 'normal'):
        _close(pred[k], ref['f64_vali_' + k], k, rtol=1e-4, atol=5e-6)
    np.testing.assert_array_equal(pred['embed'].cpu().numpy().astype(np.int64), ref['f64_vali_embed'])


def test_dropout_threshold_paths_vs_reference_code(ref, setup, cuda_dev):
    """fast_embed / vq_test with a codeword-dropout threshold and the recorded roll (vq_layers.py:284-290)."""
    scene, batch, m, bt = setup
    _, _, _, to_vis = m.fast_embed(bt, mode='test', thres=ref['thres'], ref_batch=False, roll=ref['roll'])
    np.testing.assert_array_equal(to_vis['embed'].cpu().numpy().astype(np.int64), ref['f64_fe_embed'])
    _, _, lk, _ = m.vq_test(bt, mode='vali', thres=ref['thres'], roll=ref['roll'])
    _close(lk['vqrgb'], ref['f64_vqtest_vqrgb'], 'vq_test vq_rgb', rtol=1e-4, atol=5e-6)
    np.testing.assert_array_equal(lk['usage'].cpu().numpy(), ref['f64_vqtest_usage'])


def test_two_training_steps_vs_reference_code(ref, setup, cuda_dev):
    """train_iter twice (call(mode='train') + compute_loss + gradients + Adam) against the reference code's two steps:
    losses, the EMA codebook overwrite and the light gradient (the Adam update itself is covered by test_gpu_train)."""
    from vqnerf_release_b200.nerfactor import train_nfr as T
    scene, batch, m, bt = setup
    gbs = int(ref['global_bs'])
    opt = T.Adam(learning_rate=5e-4)
    for step in range(1):      # the reference run applies no optimizer step between its two calls: compare step 0
        loss, vis, ld = T.train_iter(m, bt, opt, gbs, thres=ref['thres'], roll=ref['roll'], apply=False)
        p = 'f64_train%d_' % step
        _close(loss, ref[p + 'loss'], 'weighted loss', rtol=1e-4, atol=1e-7)
        _close(vis['pred_vq_rgb_linear'], ref[p + 'vqrgb'], 'vq_rgb', rtol=1e-4, atol=5e-6)
        _close(m._codebook, ref[p + 'codebook_after'], 'EMA codebook', rtol=2e-5, atol=1e-6)
        for k in ('rgb', 'vqrgb', 'chromaticity', 'chr_smooth', 'lambert'):
            _close(ld[k], ref[p + 'ld_' + k].sum(), 'loss_dict[%s]' % k, rtol=1e-4, atol=1e-7)
        st = m._train_state
        d_light = ref[p + 'd_light']
        err = np.abs(st.d_light.cpu().double().numpy().reshape(d_light.shape) - d_light).max()
        assert err <= 2e-4 * np.abs(d_light).max(), 'light gradient: %.3e' % err
        for name in T.NET_ORDER:                       # every bias gradient (the small tensors are recorded whole)
            for i, gb in enumerate(st.dB[name]):
                want = ref['%sd_%s_b%d' % (p, name, i)]
                err = np.abs(gb.cpu().double().numpy() - want).max()
                # + 1e-7: a 1-wide layer's bias gradient is ONE number, a sum over the rows with cancellation (fp32 noise)
                assert err <= 2e-4 * np.abs(want).max() + 1e-7, 'd %s.bias[%d]: %.3e' % (name, i, err)


def test_training_step_non_nerf_data_vs_reference_code(cuda_dev):
    """train_iter on data_type != 'nerf' (no light visibility; rgb = clip((rgb * gamma_bias) ^ clip(gamma_index, 0, 5)) with
    TRAINABLE gamma): loss, rendered colours and the gradients of light / gamma / every bias against one step of the
    reference's own code (tests/golden/decomp_real_ref.npz), eager and graph-replayed."""
    from vqnerf_release_b200.nerfactor import train_nfr as T
    g = np.load(os.path.join(os.path.dirname(GOLD), 'decomp_real_ref.npz'))
    scene = O.synth_scene(int(g['seed']), bias_scale=float(g['bias_scale']), data_type='real')
    scene.gamma = tuple(float(np.float32(v)) for v in g['gamma'])      # float32 variables in the reference
    batch = O.synth_batch(int(g['n']), int(g['seed']), fg_frac=float(g['fg_frac']), with_lvis=False)
    m = _model_from_scene(scene, cuda_dev)
    bt = _batch_tuple(batch, cuda_dev, data_type='real')
    loss, vis, ld = T.train_iter(m, bt, T.Adam(learning_rate=5e-4), int(g['global_bs']), thres=g['thres'], roll=g['roll'],
                                 apply=False)
    st = m._train_state
    assert st.has_gamma
    _close(loss, g['train_loss'], 'weighted loss', rtol=1e-4, atol=1e-7)
    mask = batch['alpha'][:, 0] > 0
    _close(vis['pred_rgb_linear'], g['train_rgb'][mask], 'rgb', rtol=1e-4, atol=5e-6)
    _close(vis['pred_vq_rgb_linear'], g['train_vqrgb'], 'vq_rgb', rtol=1e-4, atol=5e-6)
    dg = st.d_gamma.cpu().double().numpy()
    want = np.array([float(g['train_d_gamma_bias'][0]), float(g['train_d_gamma_index'][0])])
    assert np.abs(dg - want).max() <= 2e-4 * np.abs(want).max(), 'gamma gradient %s vs %s' % (dg, want)
    d_light = g['train_d_light']
    err = np.abs(st.d_light.cpu().double().numpy().reshape(d_light.shape) - d_light).max()
    assert err <= 2e-4 * np.abs(d_light).max(), 'light gradient: %.3e' % err
    for name in T.NET_ORDER:
        for i, gb in enumerate(st.dB[name]):
            want = g['train_d_%s_b%d' % (name, i)]
            err = np.abs(gb.cpu().double().numpy() - want).max()
            assert err <= 2e-4 * np.abs(want).max() + 1e-7, 'd %s.bias[%d]: %.3e' % (name, i, err)
    # the optimizer moves the two tone parameters (they are views of the flat parameter buffer)
    before = st.gamma_par.clone()
    T.train_iter(m, bt, T.Adam(learning_rate=5e-4), int(g['global_bs']), thres=g['thres'], roll=g['roll'])
    assert float((st.gamma_par - before).abs().min()) > 0
    assert abs(m.gamma[0] - float(st.gamma_par[0])) < 1e-7
