"""oracle.ref_fast_render / ref_call against vectors recorded from the reference's own models/ref_nfr.py::Model
(oracle/gen_golden_ref_nfr.py: executed through the tf_shim stand-in, float64)."""
import os

import numpy as np
import pytest
import torch

from oracle import decomp_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_nfr_ref.npz')
TOL = dict(rtol=1e-9, atol=1e-11)


def ref_setup(g):
    seed, n = int(g['seed']), int(g['n'])
    scene = O.synth_scene(seed, n_probes=int(g['n_probes']), bias_scale=float(g['bias_scale']))
    scene.nets = O.make_ref_nfr_nets(seed, float(g['bias_scale']))
    b = O.synth_batch(n, seed, fg_frac=float(g['fg_frac']))
    b['ref'] = np.random.RandomState(seed + 5).uniform(0, 1, size=(n, 3)).astype(np.float32)
    return scene, b


@pytest.fixture(scope='module')
def gold():
    return np.load(GOLD)


def test_ref_fast_render_matches_reference(gold):
    scene, b = ref_setup(gold)
    np.testing.assert_allclose(scene.nets['rgb_enc'](torch.as_tensor(b['ref'], dtype=torch.float64)).numpy(),
                               gold['z_ref'], **TOL)
    o = O.ref_fast_render(scene, b, relight_probes=True)
    np.testing.assert_allclose(o['rgb'].numpy(), gold['fr_rgb'], **TOL)
    np.testing.assert_allclose(o['rgb_probes'].numpy(), gold['fr_rgb_probes'], **TOL)
    o = O.ref_fast_render(scene, b, relight_probes=True, opt_scale=np.array([0.7, 1.1, 1.3]))
    np.testing.assert_allclose(o['rgb'].numpy(), gold['fr_scaled_rgb'], **TOL)          # raw BRDF: unaffected by the scale
    np.testing.assert_allclose(o['rgb'].numpy(), gold['fr_rgb'], **TOL)
    np.testing.assert_allclose(o['rgb_probes'].numpy(), gold['fr_scaled_rgb_probes'], **TOL)
    n = int(gold['n'])
    em = (np.arange(n) % 4 == 1).astype(np.float32)[:, None].repeat(3, 1)
    o = O.ref_fast_render(scene, b, relight_probes=True, edit_mask=em,
                          edit_material={'diff': [-1.0, 0, 0], 'spec': [0.04, 0.05, 0.06], 'rough': [0.6]})
    np.testing.assert_allclose(o['rgb'].numpy(), gold['fr_edit_rgb'], **TOL)
    np.testing.assert_allclose(o['rgb_probes'].numpy(), gold['fr_edit_rgb_probes'], **TOL)


def test_ref_call_matches_reference(gold):
    scene, b = ref_setup(gold)
    o = O.ref_call(scene, b, 'vali', relight_probes=True)
    for k in ('rgb', 'normal', 'albedo', 'spec', 'rough', 'ks', 'basecolor', 'rgb_spec', 'rgb_diff', 'rgb_probes'):
        np.testing.assert_allclose(o[k].numpy(), gold['vali_' + k], err_msg=k, **TOL)
    np.testing.assert_allclose(o['_rgb_linear'].numpy(), gold['vali_lk_rgb'], **TOL)
    o = O.ref_call(scene, b, 'test', opt_scale=np.array([0.7, 1.1, 1.3]))
    np.testing.assert_allclose(o['rgb'].numpy(), gold['test_scaled_rgb'], **TOL)
