"""ref_nfr.Model mirror (the residual model of nerfactor/test.py:181-197) on the CUDA path: against the float64 oracle and
against the vectors recorded from the reference's own code (tests/golden/ref_nfr_ref.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import decomp_oracle as O
from tests.test_gpu_parity import _close
from tests.test_ref_nfr_cpu import GOLD, ref_setup

pytestmark = pytest.mark.gpu


def _model(scene, dev):
    from vqnerf_release_b200.nerfactor.models.ref_nfr import Model
    nets = {k: (n.weights, n.biases) for k, n in scene.nets.items()}
    return Model({'data_type': 'nerf'}, nets=nets, light=scene.light,
                 novel_probes={'p%d' % i: p for i, p in enumerate(scene.probes)}, device=dev)


def _batch(b, dev):
    t = lambda a: torch.as_tensor(a).to(dev)
    n = b['xyz'].shape[0]
    return ('v', torch.zeros((n, 2), dtype=torch.int32, device=dev), t(b['rayo']), t(b['rayd']), t(b['rgb']), t(b['alpha']),
            t(b['pred_alpha']), t(b['xyz']), t(b['normal']), t(b['ref']), t(b['lvis']))


def test_ref_nfr_vs_reference_code(cuda_dev):
    g = np.load(GOLD)
    scene, b = ref_setup(g)
    m = _model(scene, cuda_dev)
    bt = _batch(b, cuda_dev)
    z_ref = m._pred_ref_at(bt[9])
    _close(z_ref, g['z_ref'], 'rgb_enc', rtol=1e-4, atol=5e-6)
    pred, _, _, _ = m.fast_render(bt, mode='test', relight_probes=True)
    _close(pred['rgb'], g['fr_rgb'], 'rgb', rtol=1e-4, atol=5e-6)
    _close(pred['rgb_probes'], g['fr_rgb_probes'], 'rgb_probes', rtol=1e-4, atol=5e-6)
    pred, _, _, _ = m.fast_render(bt, mode='test', relight_probes=True, opt_scale=np.array([0.7, 1.1, 1.3], np.float32))
    _close(pred['rgb'], g['fr_scaled_rgb'], 'rgb (opt_scale: raw BRDF)', rtol=1e-4, atol=5e-6)
    _close(pred['rgb_probes'], g['fr_scaled_rgb_probes'], 'rgb_probes (opt_scale)', rtol=1e-4, atol=5e-6)
    n = int(g['n'])
    em = torch.as_tensor((np.arange(n) % 4 == 1).astype(np.float32)[:, None].repeat(3, 1)).to(cuda_dev)
    pred, _, _, _ = m.fast_render(bt, mode='test', relight_probes=True, edit_mask=em,
                                  edit_material={'diff': [-1.0, 0, 0], 'spec': [0.04, 0.05, 0.06], 'rough': [0.6]})
    _close(pred['rgb'], g['fr_edit_rgb'], 'edited rgb', rtol=1e-4, atol=5e-6)
    _close(pred['rgb_probes'], g['fr_edit_rgb_probes'], 'edited rgb_probes', rtol=1e-4, atol=5e-6)
    pred, gt, lk, to_vis = m.call(bt, mode='vali', relight_probes=True)
    for k in ('rgb', 'normal', 'albedo', 'spec', 'rough', 'ks', 'basecolor', 'rgb_spec', 'rgb_diff', 'rgb_probes'):
        _close(pred[k], g['vali_' + k], 'call ' + k, rtol=1e-4, atol=5e-6)
    _close(lk['rgb'], g['vali_lk_rgb'], 'loss_kwargs rgb', rtol=1e-4, atol=5e-6)
    assert set(to_vis) >= {'id', 'hw', 'pred_rgb', 'pred_basecolor', 'gt_rgb', 'gt_alpha'}
    pred, _, _, _ = m.call(bt, mode='test', opt_scale=np.array([0.7, 1.1, 1.3], np.float32))
    _close(pred['rgb'], g['test_scaled_rgb'], 'call test opt_scale', rtol=1e-4, atol=5e-6)
    with pytest.raises(ValueError):
        m.call(bt, mode='bogus')


@pytest.mark.parametrize('precision', ['fp32', 'tf32x3'])
def test_ref_nfr_larger_batch_vs_oracle(cuda_dev, precision):
    n = 3000
    scene = O.synth_scene(21, n_probes=3, bias_scale=0.05)
    scene.nets = O.make_ref_nfr_nets(21, 0.05)
    b = O.synth_batch(n, 21, fg_frac=0.7)
    b['ref'] = np.random.RandomState(2).uniform(0, 1, size=(n, 3)).astype(np.float32)
    from vqnerf_release_b200.nerfactor.models.ref_nfr import Model
    nets = {k: (nn.weights, nn.biases) for k, nn in scene.nets.items()}
    m = Model({'data_type': 'nerf', 'precision': precision}, nets=nets, light=scene.light,
              novel_probes={'p%d' % i: p for i, p in enumerate(scene.probes)}, device=cuda_dev)
    pred, _, _, _ = m.fast_render(_batch(b, cuda_dev), mode='test', relight_probes=True)
    o = O.ref_fast_render(scene, b, relight_probes=True)
    _close(pred['rgb'], o['rgb'], 'rgb', rtol=1e-4, atol=5e-6)
    _close(pred['rgb_probes'], o['rgb_probes'], 'rgb_probes', rtol=1e-4, atol=5e-6)
    bg = torch.as_tensor(b['alpha'][:, 0] <= 0).to(cuda_dev)
    assert float(pred['rgb_probes'][bg].abs().max()) == 0.0
