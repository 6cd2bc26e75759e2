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


def _nets(dev, precision='tf32x3', grad_mode='reverse'):
    from vqnerf_release_b200.neus.fields import RenderingNetwork, SDFNetwork, SingleVarianceNetwork
    st = NO.make_neus_state(0)
    sdf = SDFNetwork(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=(4,), multires=6, bias=0.5, scale=1.0,
                     geometric_init=True, weight_norm=True, device=dev, precision=precision, grad_mode=grad_mode)
    col = RenderingNetwork(d_feature=256, mode='idr', d_in=9, d_out=3, d_hidden=256, n_layers=4, weight_norm=True,
                           multires_view=4, squeeze_out=True, device=dev, precision=precision)
    sdf.load_state_dict(st['sdf'])
    col.load_state_dict(st['color'])
    return st, sdf, col, SingleVarianceNetwork(0.5, device=dev)


def _t(a, dev):
    return torch.as_tensor(np.asarray(a, np.float32)).to(dev)


@pytest.mark.parametrize('grad_mode', ['reverse', 'jet'])
def test_sdf_network_vs_reference(cuda_dev, grad_mode):
    g = np.load(GOLD)
    _, sdf_net, col_net, _ = _nets(cuda_dev, grad_mode=grad_mode)
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


@pytest.mark.parametrize('grad_mode', ['reverse', 'jet'])
@pytest.mark.parametrize('n', [1, 31, 33, 127, 129, 5003, 40000])
def test_sdf_forward_vs_oracle_ragged(cuda_dev, n, grad_mode):
    st, sdf_net, col_net, _ = _nets(cuda_dev, grad_mode=grad_mode)
    rng = np.random.RandomState(n)
    pts = rng.uniform(-1.2, 1.2, size=(n, 3)).astype(np.float32)
    eo, eg = NO.sdf_forward(st['sdf'], pts)
    x = _t(pts, cuda_dev)
    rows = col_net.alloc_rows(n, cuda_dev)
    rows.fill_(float('nan'))
    sdf, feat, grad = sdf_net.forward_with_gradient(x, feat_out=rows)      # one launch: value, gradient, features into the rows
    np.testing.assert_allclose(sdf.cpu().numpy()[:, 0], eo[:, 0], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(feat.cpu().numpy(), eo[:, 1:], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(grad.cpu().numpy(), eg, rtol=1e-4, atol=1e-4)
    assert torch.isnan(rows[:, 256:]).all()                                  # only the feature columns were written
    # value-only launch agrees with the gradient launch (jets: 32 points per tile; reverse: 128)
    np.testing.assert_allclose(sdf_net.sdf(x).cpu().numpy(), sdf.cpu().numpy(), rtol=0, atol=2e-6)
    # gradient without the feature layer (what compute_vis asks for)
    g2 = sdf_net.gradient(x)[:, 0]
    np.testing.assert_allclose(g2.cpu().numpy(), eg, rtol=1e-4, atol=1e-4)
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


def test_light_rays_kernel_vs_oracle(cuda_dev):
    from vqnerf_release_b200 import abi
    g = np.load(os.path.join(os.path.dirname(GOLD), 'neus_vis_ref.npz'))
    rng = np.random.RandomState(5)
    n = 37
    surf = rng.uniform(-0.7, 0.7, size=(n, 3)).astype(np.float32)
    nrm = rng.normal(size=(n, 3)); nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    d, front, near, far = NO.light_rays(surf, nrm, g['lxyz'], 1.0)
    for l0, nc in ((0, 512), (100, 7), (511, 1)):
        ro, rd, nr, fr, ft = abi.neus_light_rays(_t(surf, cuda_dev), _t(nrm, cuda_dev), _t(g['lxyz'], cuda_dev), l0, nc, 1.0)
        assert ro.shape == (n * nc, 3) and ft.shape == (n * nc, 1)
        np.testing.assert_array_equal(ro.cpu().numpy().reshape(n, nc, 3), np.broadcast_to(surf[:, None, :], (n, nc, 3)))
        np.testing.assert_allclose(rd.cpu().numpy().reshape(n, nc, 3), d[:, l0:l0 + nc], rtol=0, atol=3e-7)
        np.testing.assert_allclose(fr.cpu().numpy().reshape(n, nc), far[:, l0:l0 + nc], rtol=2e-5, atol=0)     # -b + sqrt(disc) cancels
        np.testing.assert_allclose(nr.cpu().numpy().reshape(n, nc), near[:, l0:l0 + nc], rtol=2e-5, atol=0)
        lcos = np.einsum('ijk,ik->ij', d.astype(np.float64), nrm.astype(np.float64))[:, l0:l0 + nc]
        sure = np.abs(lcos) > 1e-6                                   # the sign of a grazing cosine may round either way
        np.testing.assert_array_equal((ft.cpu().numpy().reshape(n, nc) > 0)[sure], front[:, l0:l0 + nc][sure])
    with pytest.raises(ValueError):
        abi.neus_light_rays(_t(surf, cuda_dev), _t(nrm, cuda_dev), _t(g['lxyz'], cuda_dev), 510, 3, 1.0)


def test_compute_vis_vs_reference(cuda_dev):
    """Light-visibility extraction with the native renderer against the reference's renderer + networks
    driven by the loop of gen_geo.py:202-244 (tests/golden/neus_vis_ref.npz); two chunkings give the same buffer."""
    from vqnerf_release_b200.neus.gen_geo import compute_vis
    from vqnerf_release_b200.neus.renderer import NeuSRenderer
    g = np.load(os.path.join(os.path.dirname(GOLD), 'neus_vis_ref.npz'))
    _, sdf_net, col_net, dev_net = _nets(cuda_dev)
    r = NeuSRenderer(None, sdf_net, dev_net, col_net, n_samples=64, n_importance=64, n_outside=0, up_sample_steps=4,
                     perturb=0.0)
    surf, nrm, lxyz = _t(g['surf'], cuda_dev), _t(g['normal'], cuda_dev), _t(g['lxyz'], cuda_dev)
    lvis = compute_vis(r, surf, nrm, lxyz[None], float(g['max_radius']), cos_anneal_ratio=1.0)
    assert lvis.shape == (5, 512)
    got = lvis.cpu().numpy()
    np.testing.assert_array_equal(got == 0.0, g['lvis'] == 0.0)                # back-lit pairs exactly zero
    np.testing.assert_allclose(got, g['lvis'], rtol=0, atol=2e-3)
    lvis2 = compute_vis(r, surf, nrm, lxyz[None], float(g['max_radius']), cos_anneal_ratio=1.0, batch_size=2,
                        rays_per_render=100)
    np.testing.assert_allclose(lvis2.cpu().numpy(), got, rtol=0, atol=1e-6)    # chunking does not change a pair's ray
