"""Golden vectors for the light-visibility extraction FROM THE REFERENCE'S OWN renderer + networks
(authoring container only; needs /root/reference).

gen_geo.py's Runner cannot be imported (pyhocon, trimesh, datasets), so the loop body of Runner.compute_vis
(gen_geo.py:203-244) is driven here with its own operations -- the direction / front-lit / intersect_circle arithmetic
is executed with torch exactly as written there -- around the UNMODIFIED models.renderer.NeuSRenderer.render and
models.fields networks (parameters: oracle/neus_oracle.py::make_neus_state(0)).  Probe points sit outside the
synthetic object so that the rays towards the lights are partly occluded.  Output: tests/golden/neus_vis_ref.npz.

    python oracle/gen_golden_neus_vis.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..'))
REF = '/root/reference/geo/NeuS-ours2'
OUT = os.path.join(HERE, '..', 'tests', 'golden', 'neus_vis_ref.npz')


def main():
    from oracle import neus_oracle as NO
    sys.modules.setdefault('mcubes', types.ModuleType('mcubes'))
    ic_mod = types.ModuleType('icecream')
    ic_mod.ic = lambda *a, **k: None
    sys.modules.setdefault('icecream', ic_mod)
    sys.path.insert(0, REF)
    from models.fields import RenderingNetwork, SDFNetwork, SingleVarianceNetwork   # noqa: E402
    from models.renderer import NeuSRenderer                                        # noqa: E402

    sdf_net = SDFNetwork(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=(4,), multires=6, bias=0.5,
                         scale=1.0, geometric_init=True, weight_norm=True)
    color_net = RenderingNetwork(d_feature=256, mode='idr', d_in=9, d_out=3, d_hidden=256, n_layers=4,
                                 weight_norm=True, multires_view=4, squeeze_out=True)
    dev_net = SingleVarianceNetwork(init_val=0.5)
    st = NO.make_neus_state(0)
    sdf_net.load_state_dict({k: torch.tensor(v) for k, v in st['sdf'].items()})
    color_net.load_state_dict({k: torch.tensor(v) for k, v in st['color'].items()})
    renderer = NeuSRenderer(None, sdf_net, dev_net, color_net, n_samples=64, n_importance=64, n_outside=0,
                            up_sample_steps=4, perturb=0.0)

    # lights: gen_light_xyz(16, 32) restated by the decomposition oracle (brdf/renderer.py:184-219), radius 100
    from oracle import decomp_oracle as O
    lxyz, _ = O.gen_light_xyz(16, 32)
    lxyz_flat = torch.tensor(lxyz.reshape(1, -1, 3), dtype=torch.float32)
    rng = np.random.RandomState(11)
    n = 5
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    surf = (u * rng.uniform(0.4, 0.8, size=(n, 1))).astype(np.float32)
    nrm = u + 0.7 * rng.normal(size=(n, 3)); nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    max_radius = 1.0
    surf_batch, normal_batch = torch.tensor(surf), torch.tensor(nrm)
    n_lights = lxyz_flat.shape[1]
    lvis_hit = np.zeros((n, n_lights), dtype=np.float32)
    lpix_chunk = 64
    for i in range(0, n_lights, lpix_chunk):                               # gen_geo.py:202-244
        end_i = min(n_lights, i + lpix_chunk)
        lxyz_chunk = lxyz_flat[:, i:end_i, :]
        surf2l = lxyz_chunk - surf_batch[:, None, :]
        surf2l = surf2l / torch.linalg.norm(surf2l, ord=2, dim=-1, keepdim=True)
        surf2l_flat = surf2l.reshape((-1, 3))
        surf_flat = surf_batch[:, None, :].repeat(1, surf2l.shape[1], 1).reshape((-1, 3))
        lcos = torch.einsum('ijk,ik->ij', surf2l, normal_batch)
        front_lit = lcos > 0
        if torch.sum(front_lit) == 0:
            continue
        ff = front_lit.reshape((-1,))
        x, d = surf_flat[ff], surf2l_flat[ff]
        b = 2. * torch.sum(x * d, dim=-1)                                  # intersect_circle, :346-357
        a = torch.sum(d * d, dim=-1)
        c = torch.sum(x * x, dim=-1) - max_radius ** 2
        eps = torch.ones_like(a) * 1e-7
        denom = torch.where(2 * a > eps, 2 * a, eps)
        t1 = (-b + torch.sqrt(torch.square(b) - 4. * a * c)) / denom
        t2 = (-b - torch.sqrt(torch.square(b) - 4. * a * c)) / denom
        far = torch.where(t1 > t2, t1, t2)[:, None]
        n_far, n_near = far / 2., torch.ones_like(far) * 0.1
        near = torch.where(n_near < n_far, n_near, n_far)
        out = renderer.render(x, d, near, far, max_radius, cos_anneal_ratio=1.0, background_rgb=None)
        occu = out['weight_sum'].detach().cpu().numpy()
        full = np.zeros(lvis_hit.shape, dtype=bool)
        full[:, i:end_i] = front_lit.numpy()
        lvis_hit[full] = 1. - occu[:, 0]
        print('lights', i, end_i, 'front-lit rays', int(ff.sum()))
    np.savez_compressed(OUT, surf=surf, normal=nrm, lxyz=lxyz.reshape(-1, 3).astype(np.float32), lvis=lvis_hit,
                        max_radius=np.float32(max_radius), variance=np.float32(0.5))
    print('wrote', OUT, os.path.getsize(OUT), 'bytes; lvis histogram',
          np.histogram(lvis_hit, bins=[-0.01, 0.0, 0.05, 0.5, 0.95, 1.01])[0])


if __name__ == '__main__':
    main()
