"""Golden vectors for the NeuS networks FROM THE REFERENCE ITSELF (authoring container only; needs /root/reference).

Imports geo/NeuS-ours2/models/{fields,embedder,renderer}.py unmodified (`mcubes` / `icecream` stubbed), builds the
networks of confs/nerf.conf:53-86, loads the synthetic parameters of oracle/neus_oracle.py::make_neus_state(0) with
load_state_dict, and records SDFNetwork.forward / .gradient (autograd), RenderingNetwork.forward and one
NeuSRenderer.render of a few rays.  Output: tests/golden/neus_fields_ref.npz (committed; the weights are NOT stored --
tests rebuild them from the same seed).

    python oracle/gen_golden_neus_fields.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..'))
REF = '/root/reference/geo/NeuS-ours2'
OUT = os.path.join(HERE, '..', 'tests', 'golden', 'neus_fields_ref.npz')


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

    rng = np.random.RandomState(7)
    n = 160
    pts = rng.uniform(-1.0, 1.0, size=(n, 3)).astype(np.float32)
    pts[:8] *= 0.05                                        # near the centre (negative sdf)
    x = torch.tensor(pts)
    out = sdf_net(x)
    grad = sdf_net.gradient(torch.tensor(pts)).squeeze(1)
    dirs = rng.normal(size=(n, 3)); dirs = (dirs / np.linalg.norm(dirs, axis=1, keepdims=True)).astype(np.float32)
    col = color_net(x, grad.detach(), torch.tensor(dirs), out[:, 1:].detach())
    rec = {'pts': pts, 'dirs': dirs, 'sdf': out[:, 0].detach().numpy(), 'feat': out[:96, 1:].detach().numpy(),
           'grad': grad.detach().numpy(), 'color': col.detach().numpy()}

    # one full render (up-sampling included) of a few rays with these networks
    B = 16
    o = rng.normal(size=(B, 3)); o = 4.0 * o / np.linalg.norm(o, axis=1, keepdims=True)
    target = rng.uniform(-0.35, 0.35, size=(B, 3))
    d = target - o; d = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays_o, rays_d = torch.tensor(o, dtype=torch.float32), torch.tensor(d, dtype=torch.float32)
    renderer = NeuSRenderer(None, sdf_net, dev_net, color_net, n_samples=64, n_importance=64, n_outside=0,
                            up_sample_steps=4, perturb=0.0)
    ret = renderer.render(rays_o, rays_d, near=torch.full((B, 1), 2.0), far=torch.full((B, 1), 6.0), radius=1.0,
                          background_rgb=torch.ones((1, 3)), cos_anneal_ratio=1.0)
    rec.update({'rays_o': rays_o.numpy(), 'rays_d': rays_d.numpy()})
    for k in ('color_fine', 'weight_sum', 'weight_max', 'surf', 'depth', 'weights', 'gradients', 'gradient_error',
              's_val'):
        rec['render_' + k] = ret[k].detach().numpy()
    np.savez_compressed(OUT, **rec)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')
    print('sdf range', rec['sdf'].min(), rec['sdf'].max(), '|grad|', np.linalg.norm(rec['grad'], axis=1).mean())
    print('weight_sum', rec['render_weight_sum'].ravel())


if __name__ == '__main__':
    main()
