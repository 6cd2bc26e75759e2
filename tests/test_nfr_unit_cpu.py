"""oracle.unit_call against vectors recorded from the reference's own models/nfr_unit.py::Model.call / gen_z
(oracle/gen_golden_nfr_unit.py: executed through the tf_shim stand-in, float64)."""
import os

import numpy as np
import pytest

from oracle import decomp_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'nfr_unit_ref.npz')
TOL = dict(rtol=1e-9, atol=1e-11)


def unit_setup(g):
    seed, n = int(g['seed']), int(g['n'])
    scene = O.synth_scene(seed, bias_scale=float(g['bias_scale']))
    scene.light = scene.light + float(g['light_shift'])         # negative texels: the model's light property clips at 0
    return scene, O.synth_batch(n, seed, fg_frac=float(g['fg_frac']))


@pytest.fixture(scope='module')
def gold():
    return np.load(GOLD)


def test_unit_call_matches_reference(gold):
    scene, b = unit_setup(gold)
    assert (scene.light < 0).any()
    o = O.unit_call(scene, b, 'vali')
    for k in ('rgb', 'normal', 'albedo', 'spec', 'rough', 'ks', 'basecolor', 'xyz', 'rgb_spec', 'rgb_diff'):
        np.testing.assert_allclose(o[k].numpy(), gold['vali_' + k], err_msg=k, **TOL)
    fg = b['alpha'][:, 0] > 0
    np.testing.assert_allclose(o['_rgb_linear'].numpy(), gold['vali_lk_rgb'], **TOL)
    np.testing.assert_allclose(o['spec'].numpy()[fg], gold['vali_lk_spec'], **TOL)       # loss_kwargs: compact rows
    np.testing.assert_allclose(o['rough'].numpy()[fg], gold['vali_lk_rough'], **TOL)
    np.testing.assert_allclose(b['rgb'][fg], gold['vali_lk_gtc'], rtol=1e-6, atol=1e-7)
    o = O.unit_call(scene, b, 'train')
    assert 'rgb_spec' not in o
    np.testing.assert_allclose(o['rgb'].numpy(), gold['train_rgb'], **TOL)
    np.testing.assert_allclose(o['_rgb_linear'].numpy(), gold['train_lk_rgb'], **TOL)


def test_unit_gen_z_matches_reference(gold):
    scene, b = unit_setup(gold)
    o = O.unit_call(scene, b, 'train')
    for k in ('albedo', 'spec', 'rough', 'z_bias'):
        np.testing.assert_allclose(o[k].numpy(), gold['genz_' + k], err_msg=k, **TOL)
