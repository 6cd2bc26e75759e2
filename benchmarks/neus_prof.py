"""One launch each of the NeuS network kernels on 1 M points (for ncu): SDF value + feature + gradient (reverse mode, then
jets), SDF value only, colour network.  `ncu --set full -k regex:mlp_tc_kernel python benchmarks/neus_prof.py`"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from vqnerf_release_b200.neus.fields import RenderingNetwork, SDFNetwork   # noqa: E402


def main():
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    sdf_net = SDFNetwork(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=(4,), multires=6, bias=0.5, scale=1.0,
                         geometric_init=True, weight_norm=True, device=dev)
    col_net = RenderingNetwork(d_feature=256, mode='idr', d_in=9, d_out=3, d_hidden=256, n_layers=4, weight_norm=True,
                               multires_view=4, squeeze_out=True, device=dev)
    n = 1 << 20
    pts = torch.rand((n, 3), device=dev) * 2 - 1
    dirs = torch.nn.functional.normalize(torch.randn((n, 3), device=dev), dim=1)
    rows = col_net.alloc_rows(n, dev)
    for _ in range(2):
        sdf_net.grad_mode = 'reverse'
        sdf, feat, grad = sdf_net.forward_with_gradient(pts, feat_out=rows)     # value + feature + gradient (reverse mode)
        sdf_net.grad_mode = 'jet'
        sdf_net.forward_with_gradient(pts, feat_out=rows)                       # the same by jets
        s = sdf_net.sdf(pts)                                                    # value only
        col = col_net.forward_rows(rows, pts, grad, dirs)                       # colour network
    torch.cuda.synchronize()
    print('ok', float(sdf.mean()), float(s.mean()), float(col.mean()))


if __name__ == '__main__':
    main()
