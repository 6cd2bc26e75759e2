"""GPU parity of the NeuS per-ray scan kernels against vectors recorded from the reference's own code
(tests/golden/neus_ref.npz, see oracle/gen_golden_neus.py) and against the NumPy oracle on other shapes."""
import os

import numpy as np
import pytest
import torch

from oracle import neus_oracle as NO

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'neus_ref.npz')


def _t(a, dev):
    return torch.as_tensor(np.asarray(a, np.float32)).to(dev)


def test_up_sample_vs_reference(cuda_dev):
    from vqnerf_release_b200 import abi
    g = np.load(GOLD)
    for i in range(int(g['n_up'])):
        out = abi.neus_up_sample(_t(g['rays_o'], cuda_dev), _t(g['rays_d'], cuda_dev), _t(g['up%d_z_vals' % i], cuda_dev),
                                 _t(g['up%d_sdf' % i], cuda_dev), float(g['up%d_r_limit' % i]),
                                 int(g['up%d_n_importance' % i]), float(g['up%d_inv_s' % i]))
        np.testing.assert_allclose(out.cpu().numpy(), g['up%d_out' % i], rtol=0, atol=5e-5, err_msg='step %d' % i)


def test_cat_z_vals_vs_reference(cuda_dev):
    from vqnerf_release_b200 import abi
    g = np.load(GOLD)
    for i in range(int(g['n_up'])):
        last = bool(g['cat%d_last' % i])
        if last:
            z_out, sdf_out = abi.neus_cat_z_vals(_t(g['cat%d_z_vals' % i], cuda_dev), _t(g['cat%d_new_z' % i], cuda_dev))
            assert sdf_out is None
        else:
            z_out, sdf_out = abi.neus_cat_z_vals(_t(g['cat%d_z_vals' % i], cuda_dev), _t(g['cat%d_new_z' % i], cuda_dev),
                                                 _t(g['cat%d_sdf' % i], cuda_dev), _t(g['cat%d_new_sdf' % i], cuda_dev))
            np.testing.assert_array_equal(sdf_out.cpu().numpy(), g['cat%d_sdf_out' % i])
        np.testing.assert_array_equal(z_out.cpu().numpy(), g['cat%d_z_out' % i])     # sorting is bit-exact
        assert (np.diff(z_out.cpu().numpy(), axis=1) >= 0).all()


def test_composite_vs_reference(cuda_dev):
    from vqnerf_release_b200 import abi
    g = np.load(GOLD)
    o = abi.neus_composite(_t(g['rays_o'], cuda_dev), _t(g['rays_d'], cuda_dev), _t(g['core_z_vals'], cuda_dev),
                           _t(g['core_sdf'], cuda_dev), _t(g['core_gradients'], cuda_dev),
                           _t(g['core_sampled_color'], cuda_dev), float(g['core_inv_s']),
                           float(g['core_cos_anneal_ratio']), float(g['core_sample_dist']), float(g['core_radius']),
                           _t(g['core_background_rgb'], cuda_dev))
    for k in ('color', 'weights', 'surf', 'depth', 'cdf', 'inside_sphere', 'mid_z_vals', 'dists'):
        np.testing.assert_allclose(o[k].cpu().numpy(), g['core_out_' + k], rtol=1e-4, atol=2e-5, err_msg=k)
    ge = o['grad_err_sums'].cpu().numpy()
    assert abs(ge[0] / (ge[1] + 1e-5) - float(g['core_out_gradient_error'])) < 1e-5
    np.testing.assert_allclose(o['weight_sum'].cpu().numpy(), g['render_weight_sum'], atol=2e-5)
    np.testing.assert_allclose(o['weight_max'].cpu().numpy(), g['render_weight_max'], atol=2e-5)
    pts, dirs = abi.neus_mid_points(_t(g['rays_o'], cuda_dev), _t(g['rays_d'], cuda_dev), _t(g['core_z_vals'], cuda_dev),
                                    float(g['core_sample_dist']))
    exp = g['rays_o'][:, None, :] + g['rays_d'][:, None, :] * g['core_out_mid_z_vals'][..., None]
    np.testing.assert_allclose(pts.cpu().numpy(), exp, atol=2e-6)


@pytest.mark.parametrize('b,s,imp', [(1, 2, 1), (513, 64, 16), (512, 112, 16), (100, 128, 64), (7, 37, 5)])
def test_scan_kernels_vs_oracle_random(cuda_dev, b, s, imp):
    from vqnerf_release_b200 import abi
    rng = np.random.RandomState(b + s)
    o = rng.normal(size=(b, 3)); o = (3.0 * o / np.linalg.norm(o, axis=1, keepdims=True)).astype(np.float32)
    d = -o / 3.0 + rng.normal(size=(b, 3)).astype(np.float32) * 0.1
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    z = np.sort(rng.uniform(1.5, 4.5, size=(b, s)).astype(np.float32), axis=1)
    pts = o[:, None, :] + d[:, None, :] * z[..., None]
    sdf = (np.linalg.norm(pts, axis=-1) - 0.6).astype(np.float32)
    got = abi.neus_up_sample(_t(o, cuda_dev), _t(d, cuda_dev), _t(z, cuda_dev), _t(sdf, cuda_dev), 1.0, imp, 64.0)
    exp = NO.up_sample(o, d, z, sdf, 1.0, imp, 64.0)
    np.testing.assert_allclose(got.cpu().numpy(), exp, rtol=0, atol=1e-4)
    new_z = got.cpu().numpy()
    new_sdf = rng.normal(size=new_z.shape).astype(np.float32)
    if s + imp <= 256:
        z_out, sdf_out = abi.neus_cat_z_vals(_t(z, cuda_dev), got, _t(sdf, cuda_dev), _t(new_sdf, cuda_dev))
        ez, es = NO.cat_z_vals(z, new_z, sdf, new_sdf)
        np.testing.assert_array_equal(z_out.cpu().numpy(), ez)
        np.testing.assert_array_equal(sdf_out.cpu().numpy(), es)
    grad = (pts / np.linalg.norm(pts, axis=-1, keepdims=True)).astype(np.float32) * rng.uniform(0.8, 1.2, size=(b, s, 1)).astype(np.float32)
    col = rng.uniform(0, 1, size=(b, s, 3)).astype(np.float32)
    for car, bg in ((0.0, None), (1.0, np.array([0.5, 0.25, 1.0], np.float32))):
        oo = abi.neus_composite(_t(o, cuda_dev), _t(d, cuda_dev), _t(z, cuda_dev), _t(sdf, cuda_dev), _t(grad, cuda_dev),
                                _t(col, cuda_dev), 300.0, car, 2.0 / 64, 1.0, None if bg is None else _t(bg, cuda_dev))
        e = NO.composite(o, d, z, sdf, grad, col, 300.0, car, 2.0 / 64, 1.0, bg)
        for k in ('color', 'weights', 'surf', 'depth', 'cdf', 'inside_sphere', 'mid_z_vals', 'dists', 'weight_sum',
                  'weight_max'):
            np.testing.assert_allclose(oo[k].cpu().numpy(), e[k], rtol=2e-4, atol=5e-5, err_msg=k)
        w = oo['weights'].cpu().numpy()
        assert (w >= 0).all() and (w.sum(1) <= 1.0 + 1e-4).all()     # compositing weights are a sub-partition of unity
    with pytest.raises(ValueError):
        abi.neus_up_sample(_t(o, cuda_dev), _t(d, cuda_dev), torch.zeros((b, 129), device=cuda_dev),
                           torch.zeros((b, 129), device=cuda_dev), 1.0, 4, 64.0)


def test_renderer_mirror_runs_end_to_end(cuda_dev):
    """NeuSRenderer mirror with small torch networks: API surface + consistency with the oracle composite."""
    from vqnerf_release_b200.neus.renderer import NeuSRenderer

    class Sdf(torch.nn.Module):
        def forward(self, x):
            s = x.norm(dim=-1, keepdim=True) - 0.5
            return torch.cat([s, x.repeat(1, 2)], -1)

        def sdf(self, x):
            return self.forward(x)[:, :1]

        def gradient(self, x):
            return (x / x.norm(dim=-1, keepdim=True)).unsqueeze(1)

    dev_net = lambda x: torch.full((x.shape[0], 1), 200.0, device=x.device)
    col_net = lambda pts, g, dirs, f: torch.sigmoid(pts * 3)
    r = NeuSRenderer(None, Sdf(), dev_net, col_net, 64, 64, 0, 4, 0.0)
    b = 300
    gen = torch.Generator().manual_seed(0)
    o = torch.nn.functional.normalize(torch.randn((b, 3), generator=gen), dim=1).mul(4).to(cuda_dev)
    d = torch.nn.functional.normalize(-o + 0.2 * torch.randn((b, 3), generator=gen).to(cuda_dev), dim=1)
    near = torch.full((b, 1), 2.0, device=cuda_dev); far = torch.full((b, 1), 6.0, device=cuda_dev)
    out = r.render(o, d, near, far, 1.0, background_rgb=torch.ones((1, 3), device=cuda_dev), cos_anneal_ratio=1.0)
    assert set(out) == {'color_fine', 's_val', 'cdf_fine', 'weight_sum', 'weight_max', 'gradients', 'weights',
                        'gradient_error', 'inside_sphere', 'surf', 'depth'}
    assert out['weights'].shape == (b, 128) and out['color_fine'].shape == (b, 3)
    hit = out['weight_sum'][:, 0] > 0.99
    assert hit.any()
    surf_r = out['surf'][hit].norm(dim=-1)
    assert float((surf_r - 0.5).abs().max()) < 0.02         # rays that hit the sphere land on its surface
    with pytest.raises(NotImplementedError):
        NeuSRenderer(None, Sdf(), dev_net, col_net, 64, 64, 32, 4, 0.0)


@pytest.mark.parametrize('n_rays,s,i,i2', [(1, 2, 1, 1), (7, 64, 16, 16), (513, 80, 16, 16), (300, 96, 16, 16), (33, 37, 5, 9)])
def test_fused_scan_steps_equal_separate_kernels(cuda_dev, n_rays, s, i, i2):
    """neus_scan_step (cat_z_vals + up_sample [+ last cat_z_vals + mid points] in one launch) == the separate kernels, bit
    for bit -- the separate kernels are the ones pinned against the reference's vectors above."""
    from vqnerf_release_b200 import abi
    g = torch.Generator(device='cpu').manual_seed(n_rays * 1000 + s)
    r = lambda *sh: torch.rand(sh, generator=g)
    rays_o = (r(n_rays, 3) - 0.5).to(cuda_dev)
    rays_d = torch.nn.functional.normalize(r(n_rays, 3) - 0.5, dim=1).to(cuda_dev)
    z = torch.sort(r(n_rays, s) * 2 + 0.5, dim=1)[0]
    z[:, 1::7] = z[:, 0::7][:, :z[:, 1::7].shape[1]]              # exact ties: the stable order matters
    z = torch.sort(z, dim=1)[0].to(cuda_dev)
    new_z = (r(n_rays, i) * 2 + 0.5).to(cuda_dev)
    new_z[:, 0] = z[:, 0]                                          # a tie between the two lists
    sdf, new_sdf = (r(n_rays, s) - 0.4).to(cuda_dev), (r(n_rays, i) - 0.4).to(cuda_dev)
    radius, inv_s, sd = 1.0, 128.0, 0.03
    z_ref, sdf_ref = abi.neus_cat_z_vals(z, new_z, sdf, new_sdf)
    nz_ref = abi.neus_up_sample(rays_o, rays_d, z_ref, sdf_ref, radius, i2, inv_s)
    o = abi.neus_scan_step(rays_o, rays_d, z, new_z, sdf, new_sdf, radius, i2, inv_s)
    assert torch.equal(o['z'], z_ref) and torch.equal(o['sdf'], sdf_ref) and torch.equal(o['new_z'], nz_ref)
    pts_ref = rays_o[:, None, :] + rays_d[:, None, :] * nz_ref[..., None]
    assert float((o['pts'] - pts_ref).abs().max()) < 1e-6           # (fma contraction differs from torch's mul + add)
    nz2, pts2 = abi.neus_up_sample_pts(rays_o, rays_d, z_ref, sdf_ref, radius, i2, inv_s)
    assert torch.equal(nz2, nz_ref) and torch.equal(pts2, o['pts'])
    # last step: + cat_z_vals(last=True) + mid points
    zf_ref, _ = abi.neus_cat_z_vals(z_ref, nz_ref)
    pm_ref, dm_ref = abi.neus_mid_points(rays_o, rays_d, zf_ref, sd)
    f = abi.neus_scan_step(rays_o, rays_d, z, new_z, sdf, new_sdf, radius, i2, inv_s, final_merge=True, sample_dist=sd,
                           want_merged=False)
    assert torch.equal(f['z_final'], zf_ref)
    assert torch.equal(f['mid_pts'], pm_ref) and torch.equal(f['mid_dirs'], dm_ref)
