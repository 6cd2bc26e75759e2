"""GPU parity of the native NeuS networks (vqn_sdf_forward on the fused tcgen05 kernel, colour network through
vqn_net_forward) against (a) vectors recorded from the reference's own modules -- tests/golden/neus_fields_ref.npz,
oracle/gen_golden_neus_fields.py -- and (b) the float64 oracle on larger, ragged batches.

Tolerance (north star): 1e-4 relative in the fp32-parity mode (tf32x3); the absolute floors below are 1e-4 of the
tensor scale (sdf ~ 1, |gradient| ~ 1, features ~ 1)."""
import os

import numpy as np
import pytest
import torch

from oracle import neus_oracle as NO

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'neus_fields_ref.npz')


def _nets(dev, precision='tf32x3'):
    from vqnerf_release_b200.neus.fields import RenderingNetwork, SDFNetwork, SingleVarianceNetwork
    st = NO.make_neus_state(0)
    sdf = SDFNetwork(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=(4,), multires=6, bias=0.5, scale=1.0,
                     geometric_init=True, weight_norm=True, device=dev, precision=precision)
    col = RenderingNetwork(d_feature=256, mode='idr', d_in=9, d_out=3, d_hidden=256, n_layers=4, weight_norm=True,
                           multires_view=4, squeeze_out=True, device=dev, precision=precision)
    sdf.load_state_dict(st['sdf'])
    col.load_state_dict(st['color'])
    return st, sdf, col, SingleVarianceNetwork(0.5, device=dev)


def _t(a, dev):
    return torch.as_tensor(np.asarray(a, np.float32)).to(dev)


def test_sdf_network_vs_reference(cuda_dev):
    g = np.load(GOLD)
    _, sdf_net, col_net, _ = _nets(cuda_dev)
    x = _t(g['pts'], cuda_dev)
    out = sdf_net(x).cpu().numpy()
    assert out.shape == (len(g['pts']), 257)
    np.testing.assert_allclose(out[:, 0], g['sdf'], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(out[:96, 1:], g['feat'], rtol=1e-4, atol=2e-5)
    grad = sdf_net.gradient(x)
    assert grad.shape == (len(g['pts']), 1, 3)
    np.testing.assert_allclose(grad[:, 0].cpu().numpy(), g['grad'], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(sdf_net.sdf(x).cpu().numpy()[:, 0], g['sdf'], rtol=1e-4, atol=2e-5)
    col = col_net(x, _t(g['grad'], cuda_dev), _t(g['dirs'], cuda_dev), _t(out[:, 1:], cuda_dev))
    np.testing.assert_allclose(col.cpu().numpy(), g['color'], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize('n', [1, 31, 33, 127, 129, 5003, 40000])
def test_sdf_forward_vs_oracle_ragged(cuda_dev, n):
    st, sdf_net, col_net, _ = _nets(cuda_dev)
    rng = np.random.RandomState(n)
    pts = rng.uniform(-1.2, 1.2, size=(n, 3)).astype(np.float32)
    eo, eg = NO.sdf_forward(st['sdf'], pts)
    x = _t(pts, cuda_dev)
    rows = col_net.alloc_rows(n, cuda_dev)
    rows.fill_(float('nan'))
    sdf, feat, grad = sdf_net.forward_with_gradient(x, feat_out=rows)      # one launch: jets + features into the rows
    np.testing.assert_allclose(sdf.cpu().numpy()[:, 0], eo[:, 0], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(feat.cpu().numpy(), eo[:, 1:], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(grad.cpu().numpy(), eg, rtol=1e-4, atol=1e-4)
    assert torch.isnan(rows[:, 256:]).all()                                  # only the feature columns were written
    # value-only tiles (128 points per tile) agree with the jet tiles (32 points per tile)
    np.testing.assert_allclose(sdf_net.sdf(x).cpu().numpy(), sdf.cpu().numpy(), rtol=0, atol=2e-6)
    # colour network on the rows the SDF kernel filled
    dirs = rng.normal(size=(n, 3)); dirs = (dirs / np.linalg.norm(dirs, axis=1, keepdims=True)).astype(np.float32)
    col = col_net.forward_rows(rows, x, grad, _t(dirs, cuda_dev))
    ec = NO.color_forward(st['color'], pts, eg, dirs, eo[:, 1:])
    np.testing.assert_allclose(col.cpu().numpy(), ec, rtol=1e-4, atol=3e-5)
    assert not torch.isnan(rows).any()


def test_sdf_forward_empty_and_errors(cuda_dev):
    _, sdf_net, col_net, _ = _nets(cuda_dev)
    out = sdf_net(torch.zeros((0, 3), device=cuda_dev))
    assert out.shape == (0, 257)
    from vqnerf_release_b200.neus.fields import SDFNetwork
    with pytest.raises(NotImplementedError):
        SDFNetwork(d_out=257, d_in=3, d_hidden=256, n_layers=8, multires=6, scale=2.0, device=cuda_dev)
    with pytest.raises(NotImplementedError):
        SDFNetwork(d_out=257, d_in=3, d_hidden=256, n_layers=8, multires=0, device=cuda_dev)
    with pytest.raises(KeyError):
        sdf_net.load_state_dict({'lin0.bias': np.zeros(256, np.float32)})
    sdf_net.precision = 'fp32'
    with pytest.raises(ValueError):                                          # tensor-core kernel only
        sdf_net.sdf(torch.zeros((4, 3), device=cuda_dev))


def test_sdf_forward_bf16_mode(cuda_dev):
    st, sdf_net, _, _ = _nets(cuda_dev, precision='bf16')
    rng = np.random.RandomState(3)
    pts = rng.uniform(-1.0, 1.0, size=(4096, 3)).astype(np.float32)
    eo, _ = NO.sdf_forward(st['sdf'], pts)
    out = sdf_net(_t(pts, cuda_dev)).cpu().numpy()
    assert np.abs(out[:, 0] - eo[:, 0]).max() < 3e-2                         # bf16 operands: 1e-2-class budget
    assert np.abs(out[:, 0] - eo[:, 0]).mean() < 5e-3


def test_render_with_native_networks_vs_reference(cuda_dev):
    """NeuSRenderer.render (64 + 4 x 16 samples, up-sampling included) with the native networks against the render
    the reference's own renderer + networks produced on CPU torch."""
    from vqnerf_release_b200.neus.renderer import NeuSRenderer
    g = np.load(GOLD)
    _, sdf_net, col_net, dev_net = _nets(cuda_dev)
    r = NeuSRenderer(None, sdf_net, dev_net, col_net, n_samples=64, n_importance=64, n_outside=0, up_sample_steps=4,
                     perturb=0.0)
    b = g['rays_o'].shape[0]
    out = r.render(_t(g['rays_o'], cuda_dev), _t(g['rays_d'], cuda_dev), torch.full((b, 1), 2.0, device=cuda_dev),
                   torch.full((b, 1), 6.0, device=cuda_dev), 1.0, background_rgb=torch.ones((1, 3), device=cuda_dev),
                   cos_anneal_ratio=1.0)
    np.testing.assert_allclose(out['weight_sum'].cpu().numpy(), g['render_weight_sum'], rtol=0, atol=2e-3)
    np.testing.assert_allclose(out['color_fine'].cpu().numpy(), g['render_color_fine'], rtol=0, atol=2e-3)
    np.testing.assert_allclose(out['surf'].cpu().numpy(), g['render_surf'], rtol=0, atol=3e-3)
    np.testing.assert_allclose(out['depth'].cpu().numpy(), g['render_depth'], rtol=0, atol=3e-3)
    np.testing.assert_allclose(out['weights'].cpu().numpy(), g['render_weights'], rtol=0, atol=5e-3)
    np.testing.assert_allclose(out['s_val'].cpu().numpy(), g['render_s_val'], rtol=1e-5)
    assert abs(float(out['gradient_error']) - float(g['render_gradient_error'])) < 1e-3
