"""nfr_unit.Model mirror (the warm-up model: main branch without a VQ layer) on the CUDA path against the vectors recorded
from the reference's own code (tests/golden/nfr_unit_ref.npz) and against the float64 oracle on a ragged batch."""
import numpy as np
import pytest
import torch

from oracle import decomp_oracle as O
from tests.test_gpu_parity import _close
from tests.test_nfr_unit_cpu import GOLD, unit_setup

pytestmark = pytest.mark.gpu


def _model(scene, dev, precision='tf32x3'):
    from vqnerf_release_b200.nerfactor.models.nfr_unit import Model
    nets = {dst: (scene.nets[src].weights, scene.nets[src].biases)
            for src, dst in (('fine_enc', 'fine_enc'), ('bottleneck', 'bottleneck'), ('diff_main', 'diff_out'),
                             ('spec_main', 'spec_out'), ('rough_main', 'rough_out'))}
    return Model({'data_type': 'nerf', 'precision': precision}, nets=nets, light=scene.light, device=dev)


def _batch(b, dev):
    t = lambda a: torch.as_tensor(a).to(dev)
    n = b['xyz'].shape[0]
    return ('v', torch.zeros((n, 2), dtype=torch.int32, device=dev), t(b['rayo']), t(b['rayd']), t(b['rgb']), t(b['alpha']),
            t(b['pred_alpha']), t(b['xyz']), t(b['normal']), t(b['lvis']))


@pytest.mark.parametrize('precision', ['tf32x3', 'fp32'])
def test_nfr_unit_vs_reference_code(cuda_dev, precision):
    g = np.load(GOLD)
    scene, b = unit_setup(g)
    m = _model(scene, cuda_dev, precision)
    bt = _batch(b, cuda_dev)
    pred, gt, lk, to_vis = m.call(bt, mode='vali')
    for k in ('rgb', 'normal', 'albedo', 'spec', 'rough', 'ks', 'basecolor', 'xyz', 'rgb_spec', 'rgb_diff'):
        _close(pred[k], g['vali_' + k], k, rtol=1e-4, atol=5e-6)
    for k in ('rgb', 'spec', 'rough', 'gtc'):
        _close(lk[k], g['vali_lk_' + k], 'loss_kwargs ' + k, rtol=1e-4, atol=5e-6)
    assert lk['mode'] == 'vali' and lk['pretrain'] is False and to_vis['pred_rgb'] is pred['rgb']
    assert torch.equal(gt['rgb'], bt[4]) and torch.equal(gt['xyz'], bt[7])             # un-masked in this model (:193)
    pred, _, lk, _ = m.call(bt, mode='train')
    assert 'rgb_spec' not in pred
    _close(pred['rgb'], g['train_rgb'], 'train rgb', rtol=1e-4, atol=5e-6)
    _close(lk['rgb'], g['train_lk_rgb'], 'train lk rgb', rtol=1e-4, atol=5e-6)
    tv = m.gen_z(bt, genz=True)
    for k in ('albedo', 'spec', 'rough', 'z_bias'):
        _close(tv[k], g['genz_' + k], 'gen_z ' + k, rtol=1e-4, atol=5e-6)
    assert 'z_bias' not in m.gen_z(bt)
    with pytest.raises(ValueError):
        m.call(bt, mode='nonsense')
    with pytest.raises(AttributeError):
        m.fast_embed(bt)


def test_nfr_unit_ragged_batch_vs_oracle(cuda_dev):
    scene = O.synth_scene(5, bias_scale=0.05)
    b = O.synth_batch(1237, 5, fg_frac=0.4)                   # several tiles, mostly background
    m = _model(scene, cuda_dev)
    pred, _, lk, _ = m.call(_batch(b, cuda_dev), mode='test')
    o = O.unit_call(scene, b, 'test')
    for k in ('rgb', 'albedo', 'spec', 'rough', 'ks', 'basecolor', 'rgb_spec', 'rgb_diff'):
        _close(pred[k], o[k], k, rtol=1e-4, atol=5e-6)
    _close(m._pred_bias_at(torch.as_tensor(b['xyz']).to(cuda_dev)),
           O.pred_enc_at(scene.nets, torch.as_tensor(b['xyz'], dtype=torch.float64)), 'z_bias', rtol=1e-4, atol=5e-6)
