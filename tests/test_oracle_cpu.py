"""CPU tests of the oracles (no GPU, no product code on the compute path).

* NeuS oracle vs golden vectors recorded from the reference's own code (tests/golden/neus_ref.npz).
* decomp oracle: known-answer checks that do not depend on the reference (solid-angle sum, white-furnace
  Lambertian integral, Sonnet EMA first-step identity, first-index arg-min ties, sRGB round trip) plus the
  committed float64 regression vectors (tests/golden/decomp_oracle.npz).
"""
import math
import os

import numpy as np
import pytest
import torch

from oracle import decomp_oracle as O
from oracle import neus_oracle as NO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def neus_gold():
    return np.load(os.path.join(GOLD, 'neus_ref.npz'))


def test_neus_up_sample_matches_reference(neus_gold):
    g = neus_gold
    for i in range(int(g['n_up'])):
        out = NO.up_sample(g['rays_o'], g['rays_d'], g['up%d_z_vals' % i], g['up%d_sdf' % i],
                           float(g['up%d_r_limit' % i]), int(g['up%d_n_importance' % i]), float(g['up%d_inv_s' % i]))
        np.testing.assert_allclose(out, g['up%d_out' % i], rtol=0, atol=2e-5)


def test_neus_cat_z_vals_matches_reference(neus_gold):
    g = neus_gold
    for i in range(int(g['n_up'])):
        last = bool(g['cat%d_last' % i])
        new_sdf = None if last else g['cat%d_new_sdf' % i]
        z_out, sdf_out = NO.cat_z_vals(g['cat%d_z_vals' % i], g['cat%d_new_z' % i],
                                       None if last else g['cat%d_sdf' % i], new_sdf)
        np.testing.assert_array_equal(z_out, g['cat%d_z_out' % i])
        if not last:
            np.testing.assert_array_equal(sdf_out, g['cat%d_sdf_out' % i])


def test_neus_composite_matches_reference(neus_gold):
    g = neus_gold
    o = NO.composite(g['rays_o'], g['rays_d'], g['core_z_vals'], g['core_sdf'], g['core_gradients'],
                     g['core_sampled_color'], float(g['core_inv_s']), float(g['core_cos_anneal_ratio']),
                     float(g['core_sample_dist']), float(g['core_radius']), g['core_background_rgb'])
    for k in ('color', 'weights', 'surf', 'depth', 'cdf', 'inside_sphere', 'mid_z_vals', 'dists'):
        np.testing.assert_allclose(o[k], g['core_out_' + k], rtol=1e-4, atol=2e-5, err_msg=k)
    assert abs(o['gradient_error'] - float(g['core_out_gradient_error'])) < 1e-5
    np.testing.assert_allclose(o['weight_sum'], g['render_weight_sum'], atol=2e-5)
    np.testing.assert_allclose(o['weight_max'], g['render_weight_max'], atol=2e-5)


# ---------------------------------------------------------------------------- decomp oracle: known answers
def test_light_probe_solid_angles():
    xyz, areas = O.gen_light_xyz(16, 32)
    assert xyz.shape == (16, 32, 3) and areas.shape == (16, 32)
    assert abs(areas.sum() - 4 * math.pi) < 1e-12
    np.testing.assert_allclose(np.linalg.norm(xyz, axis=-1), 100.0, rtol=1e-12)
    assert xyz[0, :, 2].min() > 0 and xyz[-1, :, 2].max() < 0          # first row = north (lat > 0)
    assert np.all(areas > 0)


def test_white_furnace_lambertian():
    """albedo a, no specular, constant radiance L, full visibility, point at the origin looking up:
    rgb = a/pi * sum_{cos>0} cos * L * dOmega ~= a * L  (the 16x32 midpoint rule is within 2 %)."""
    dt = torch.float64
    lxyz, lareas = (torch.as_tensor(a, dtype=dt) for a in O.gen_light_xyz(16, 32))
    pts = torch.zeros((1, 3), dtype=dt)
    l = O.calc_ldir(lxyz, pts)
    n = torch.tensor([[0., 0., 1.]], dtype=dt)
    v = torch.tensor([[0., 0., 1.]], dtype=dt)
    albedo = torch.tensor([[0.3, 0.5, 0.7]], dtype=dt)
    brdf, glossy, diffuse = O.get_brdf(l, v, n, albedo, torch.ones((1, 1), dtype=dt), torch.zeros((1, 3), dtype=dt))
    light = torch.full((16, 32, 3), 0.4, dtype=dt)
    rgb, _ = O.render(diffuse, l, n, lareas, light)
    np.testing.assert_allclose(rgb.numpy(), (albedo * 0.4).numpy(), rtol=0.02)
    assert float(glossy.min()) >= 0.0


def test_sonnet_ema_first_step_is_identity():
    ema = O.SonnetEMA([4], 0.999, torch.float64)
    v = torch.tensor([1.0, 2.0, 0.0, 5.0], dtype=torch.float64)
    np.testing.assert_allclose(ema(v).numpy(), v.numpy(), rtol=1e-12)      # debiased: hidden/(1-decay)
    a2 = ema(2 * v)
    expect = (v * 0.999 * 0.001 + 2 * v * 0.001) / (1 - 0.999 ** 2)
    np.testing.assert_allclose(a2.numpy(), expect.numpy(), rtol=1e-12)


def test_vq_first_min_tie_and_threshold_mask():
    dt = torch.float64
    z = 256
    cb = torch.zeros((z, 3), dtype=dt)
    cb[0, 0] = 1.0
    cb[0, 1] = 1.0       # duplicate of codeword 0 -> tie
    cb[1, 2] = 1.0
    x = torch.zeros((2, z), dtype=dt)
    x[0, 0] = 1.0
    x[1, 1] = 1.0
    vq = O.VectorQuantizerEMA(z, 3, 0.1, dtype=dt)
    out = vq(x, cb, False)
    assert out['encoding_indices'].tolist() == [0, 2]
    # drop codeword 0 (roll < thres): row 0 must fall to its duplicate, index 1
    out = vq(x, cb, False, thres=torch.tensor([[0.5, 0.0, 0.0]], dtype=dt), roll=torch.tensor([[0.1, 0.9, 0.9]], dtype=dt))
    assert out['encoding_indices'].tolist() == [1, 2]
    assert float(out['loss']) == 0.0
    # perplexity of a 50/50 split over two codes
    assert abs(float(out['perplexity']) - 2.0) < 1e-6


def test_srgb_round_trip_and_embed_layout():
    t = torch.linspace(0, 1, 101, dtype=torch.float64)
    np.testing.assert_allclose(O.srgb2linear(O.linear2srgb(t)).numpy(), t.numpy(), atol=1e-6)
    x = torch.tensor([[0.1, -0.2, 0.3]], dtype=torch.float64)
    e = O.embed(x, 10)
    assert e.shape == (1, 63)
    np.testing.assert_allclose(e[0, 3:6].numpy(), np.sin(x[0].numpy()))
    np.testing.assert_allclose(e[0, 6:9].numpy(), np.cos(x[0].numpy()))
    np.testing.assert_allclose(e[0, 57:60].numpy(), np.sin(512 * x[0].numpy()))


def test_l2_normalize_epsilon_on_squared_norm():
    x = torch.tensor([[1e-4, 0.0, 0.0]], dtype=torch.float64)      # squared norm 1e-8 < 1e-6
    y = O.safe_l2_normalize(x, 1)
    assert abs(float(y[0, 0]) - 1e-4 / 1e-3) < 1e-12


def test_mlp_param_counts():
    nets = O.make_vq_nfr_nets(0)
    count = lambda n: sum(w.size + b.size for w, b in zip(n.weights, n.biases))
    assert count(nets['fine_enc']) == 65792 and count(nets['bottleneck']) == 115328
    assert count(nets['diff_main']) == 99843 and count(nets['spec_main']) == 99073
    assert sum(count(n) for n in nets.values()) == 777868


def test_decomp_oracle_matches_committed_vectors():
    g = np.load(os.path.join(GOLD, 'decomp_oracle.npz'))
    scene = O.synth_scene(int(g['seed']), n_probes=int(g['n_probes']), bias_scale=float(g['bias_scale']))
    batch = O.synth_batch(int(g['n']), int(g['seed']), fg_frac=float(g['fg_frac']))
    r = O.fast_render(scene, batch, torch.float64, relight_probes=True, gen_embed=True)
    for k in ('basecolor', 'albedo', 'spec', 'rough', 'rgb', 'rgb_probes', 'embed'):
        np.testing.assert_allclose(r[k].numpy(), g['fr_' + k], rtol=1e-9, atol=1e-12, err_msg=k)
    r32 = O.fast_render(scene, batch, torch.float32, relight_probes=True)
    for k in ('albedo', 'spec', 'rough', 'rgb', 'rgb_probes'):
        np.testing.assert_allclose(r32[k].numpy(), g['fr_' + k], rtol=1e-4, atol=2e-6, err_msg=k + ' fp32')
    vq = O.VectorQuantizerEMA(O.Z_DIM, O.NUM_EMBED, O.COMMITMENT_COST, dtype=torch.float64)
    for step in range(2):
        c = O.call_forward(scene, batch, vq, 'train', thres=g['thres'], roll=g['roll'], dtype=torch.float64)
        scene.codebook = c['update'].numpy().astype(np.float32)
        for k in ('rgb_linear', 'vq_rgb_linear', 'embed_ind', 'update', 'vq_loss', 'perplexity'):
            np.testing.assert_allclose(c[k].numpy(), g['call%d_%s' % (step, k)], rtol=1e-9, atol=1e-12)


def test_neus_networks_match_reference():
    """oracle SDFNetwork / RenderingNetwork restatement vs the reference's own modules (autograd gradient included):
    tests/golden/neus_fields_ref.npz, recorded by oracle/gen_golden_neus_fields.py (float32 torch on CPU)."""
    g = np.load(os.path.join(GOLD, 'neus_fields_ref.npz'))
    st = NO.make_neus_state(0)
    out, grad = NO.sdf_forward(st['sdf'], g['pts'])
    np.testing.assert_allclose(out[:, 0], g['sdf'], rtol=0, atol=3e-6)
    np.testing.assert_allclose(out[:96, 1:], g['feat'], rtol=0, atol=3e-6)
    np.testing.assert_allclose(grad, g['grad'], rtol=0, atol=2e-5)
    col = NO.color_forward(st['color'], g['pts'], g['grad'], g['dirs'], out[:, 1:])
    np.testing.assert_allclose(col, g['color'], rtol=0, atol=3e-6)
    # the forward-mode gradient is the derivative of the oracle's own sdf (central differences in float64)
    x = g['pts'][:16].astype(np.float64)
    for k in range(3):
        e = np.zeros(3); e[k] = 1e-6
        fd = (NO.sdf_forward(st['sdf'], x + e)[0][:, 0] - NO.sdf_forward(st['sdf'], x - e)[0][:, 0]) / 2e-6
        np.testing.assert_allclose(grad[:16, k], fd, rtol=0, atol=1e-6)


def test_neus_compute_vis_matches_reference():
    """oracle light-visibility extraction (ray set-up + render + 1 - weight_sum) vs the loop of gen_geo.py:202-244
    driven around the reference's own renderer and networks (tests/golden/neus_vis_ref.npz)."""
    g = np.load(os.path.join(GOLD, 'neus_vis_ref.npz'))
    st = NO.make_neus_state(0)
    inv_s = float(np.exp(np.float32(g['variance']) * 10.0))
    k = 1                                                     # one probe point (~260 rays) keeps the CPU suite short
    lvis = NO.compute_vis(st, g['surf'][:k], g['normal'][:k], g['lxyz'], float(g['max_radius']), inv_s, 1.0)
    assert lvis.shape == (k, 512)
    np.testing.assert_array_equal(lvis == 0.0, g['lvis'][:k] == 0.0)          # back-lit pairs are exactly zero
    np.testing.assert_allclose(lvis, g['lvis'][:k], rtol=0, atol=2e-3)
    assert 0.05 < (g['lvis'][:k] > 0.5).mean() < 0.95                          # occluded and visible pairs both present
