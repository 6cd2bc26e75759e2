"""Generate golden vectors for the NeuS per-ray scan path FROM THE REFERENCE ITSELF.

Runs in the authoring container only (needs /root/reference): imports
geo/NeuS-ours2/models/{renderer,fields,embedder}.py unmodified (with `mcubes` and `icecream` stubbed,
renderer.py:6-7), builds the networks of confs/nerf.conf:53-86, renders a few rays on CPU torch and records
the inputs/outputs of up_sample, cat_z_vals and render_core.  Output: tests/golden/neus_ref.npz (committed).

    python oracle/gen_golden_neus.py
"""
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference/geo/NeuS-ours2'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'neus_ref.npz')


def main():
    sys.modules.setdefault('mcubes', types.ModuleType('mcubes'))
    ic_mod = types.ModuleType('icecream')
    ic_mod.ic = lambda *a, **k: None
    sys.modules.setdefault('icecream', ic_mod)
    sys.path.insert(0, REF)
    from models.fields import RenderingNetwork, SDFNetwork, SingleVarianceNetwork   # noqa: E402
    from models.renderer import NeuSRenderer                                        # noqa: E402

    torch.manual_seed(0)
    np.random.seed(0)
    sdf_net = SDFNetwork(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=(4,), multires=6, bias=0.5,
                         scale=1.0, geometric_init=True, weight_norm=True)
    color_net = RenderingNetwork(d_feature=256, mode='idr', d_in=9, d_out=3, d_hidden=256, n_layers=4,
                                 weight_norm=True, multires_view=4, squeeze_out=True)
    dev_net = SingleVarianceNetwork(init_val=0.3)
    renderer = NeuSRenderer(None, sdf_net, dev_net, color_net, n_samples=64, n_importance=64, n_outside=0,
                            up_sample_steps=4, perturb=0.0)

    B = 24
    rng = np.random.RandomState(1)
    o = rng.normal(size=(B, 3))
    o = 4.0 * o / np.linalg.norm(o, axis=1, keepdims=True)
    target = rng.uniform(-0.6, 0.6, size=(B, 3))
    d = target - o
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays_o = torch.tensor(o, dtype=torch.float32)
    rays_d = torch.tensor(d, dtype=torch.float32)

    rec = {'rays_o': rays_o.numpy(), 'rays_d': rays_d.numpy()}
    ups, cats = [], []
    orig_up, orig_cat, orig_core = renderer.up_sample, renderer.cat_z_vals, renderer.render_core

    def up_hook(rays_o, rays_d, z_vals, sdf, r_limit, n_importance, inv_s):
        out = orig_up(rays_o, rays_d, z_vals, sdf, r_limit, n_importance, inv_s)
        ups.append(dict(z_vals=z_vals.detach().numpy().copy(), sdf=sdf.detach().numpy().copy(), r_limit=r_limit,
                        n_importance=n_importance, inv_s=inv_s, out=out.detach().numpy().copy()))
        return out

    def cat_hook(rays_o, rays_d, z_vals, new_z_vals, sdf, last=False):
        z_out, sdf_out = orig_cat(rays_o, rays_d, z_vals, new_z_vals, sdf, last=last)
        r = dict(z_vals=z_vals.detach().numpy().copy(), new_z=new_z_vals.detach().numpy().copy(),
                 sdf=sdf.detach().numpy().copy(), last=last, z_out=z_out.detach().numpy().copy(),
                 sdf_out=sdf_out.detach().numpy().copy())
        if not last:
            pts = rays_o[:, None, :] + rays_d[:, None, :] * new_z_vals[..., :, None]
            r['new_sdf'] = sdf_net.sdf(pts.reshape(-1, 3)).reshape(new_z_vals.shape).detach().numpy().copy()
        cats.append(r)
        return z_out, sdf_out

    core = {}

    def core_hook(rays_o, rays_d, z_vals, sample_dist, radius, sdf_network, deviation_network, color_network,
                  **kw):
        captured = {}

        class ColorTap:
            def __call__(self, pts, gradients, dirs, feat):
                out = color_network(pts, gradients, dirs, feat)
                captured['sampled_color'] = out.detach().numpy().copy()
                captured['gradients'] = gradients.detach().numpy().copy()
                return out

        ret = orig_core(rays_o, rays_d, z_vals, sample_dist, radius, sdf_network, deviation_network, ColorTap(),
                        **kw)
        core.update(dict(z_vals=z_vals.detach().numpy().copy(), sample_dist=sample_dist, radius=radius,
                         cos_anneal_ratio=kw.get('cos_anneal_ratio', 0.0),
                         background_rgb=None if kw.get('background_rgb') is None
                         else kw['background_rgb'].detach().numpy().copy(),
                         sdf=ret['sdf'].detach().numpy().copy(), gradients=captured['gradients'],
                         sampled_color=captured['sampled_color'],
                         inv_s=float(1.0 / ret['s_val'].detach().numpy().reshape(-1)[0])))
        for k in ('color', 'dists', 'mid_z_vals', 'weights', 'cdf', 'gradient_error', 'inside_sphere', 'surf',
                  'depth'):
            core['out_' + k] = ret[k].detach().numpy().copy()
        return ret

    renderer.up_sample, renderer.cat_z_vals, renderer.render_core = up_hook, cat_hook, core_hook
    bg = torch.tensor([[1.0, 1.0, 1.0]])
    near_t = torch.full((B, 1), 2.0)
    far_t = torch.full((B, 1), 6.0)
    out = renderer.render(rays_o, rays_d, near=near_t, far=far_t, radius=1.0, background_rgb=bg, cos_anneal_ratio=0.7)
    for i, u in enumerate(ups):
        for k, v in u.items():
            rec['up%d_%s' % (i, k)] = np.asarray(v)
    for i, c in enumerate(cats):
        for k, v in c.items():
            rec['cat%d_%s' % (i, k)] = np.asarray(v)
    for k, v in core.items():
        if v is not None:
            rec['core_' + k] = np.asarray(v)
    for k in ('color_fine', 'weight_sum', 'weight_max', 'surf', 'depth'):
        rec['render_' + k] = out[k].detach().numpy().copy()
    rec['n_up'] = np.asarray(len(ups))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **rec)
    print('wrote', OUT, {k: np.asarray(v).shape for k, v in rec.items() if k.startswith('core_out')})
    print('weight_sum range', float(out['weight_sum'].min()), float(out['weight_sum'].max()))


if __name__ == '__main__':
    main()
