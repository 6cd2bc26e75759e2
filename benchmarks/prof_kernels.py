"""Single launches of the secondary kernels, for ncu captures (profiles/README.md lists the commands):

    python benchmarks/prof_kernels.py neus_scan [rays]   # up_sample, 3 fused scan steps, composite
    python benchmarks/prof_kernels.py train              # one eager training step at 8192 rays (dense / shade_bwd / loss / EMA / Adam)
    python benchmarks/prof_kernels.py vq K               # indices-only VQ assignment of 4 M latents
    python benchmarks/prof_kernels.py sdf_grad           # SDFNetwork value + gradient of 1 M points, with and without the feature layer
Every mode runs its launches twice (the first pass warms caches and lazy initialisation: give ncu `-s <launches of one pass>`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def neus_scan(b):
    from vqnerf_release_b200 import abi
    dev = torch.device('cuda:0')
    g = torch.Generator(device=dev).manual_seed(b)
    rays_o = torch.randn((b, 3), generator=g, device=dev)
    rays_o = 4.0 * rays_o / rays_o.norm(dim=1, keepdim=True)
    rays_d = -rays_o / 4.0 + 0.05 * torch.randn((b, 3), generator=g, device=dev)
    rays_d = rays_d / rays_d.norm(dim=1, keepdim=True)
    z0 = torch.linspace(2.0, 6.0, 64, device=dev)[None, :].expand(b, 64).contiguous()
    sdf0 = ((rays_o[:, None, :] + rays_d[:, None, :] * z0[..., None]).norm(dim=-1) - 1.0).contiguous()
    new_sdf = [torch.rand((b, 16), generator=g, device=dev) - 0.5 for _ in range(3)]
    grads = torch.randn((b, 128, 3), generator=g, device=dev)
    cols = torch.rand((b, 128, 3), generator=g, device=dev)
    sdf_f = torch.rand((b, 128), generator=g, device=dev) - 0.5
    for _ in range(2):
        z, sdf = z0, sdf0
        nz, _ = abi.neus_up_sample_pts(rays_o, rays_d, z, sdf, 1.0, 16, 64)
        for i in range(3):
            o = abi.neus_scan_step(rays_o, rays_d, z, nz, sdf, new_sdf[i], 1.0, 16, 64 * 2 ** (i + 1), final_merge=(i == 2),
                                   sample_dist=2.0 / 64, want_merged=(i < 2))
            if i < 2:
                z, sdf, nz = o['z'], o['sdf'], o['new_z']
        out = abi.neus_composite(rays_o, rays_d, o['z_final'], sdf_f, grads, cols, 300.0, 1.0, 2.0 / 64, 1.0)
    torch.cuda.synchronize()
    print('neus_scan ok', b, float(out['weight_sum'].mean()))


def train():
    import bench
    from vqnerf_release_b200.nerfactor import train_nfr as T
    from vqnerf_release_b200.nerfactor.models.vq_nfr import Model
    dev = torch.device('cuda:0')
    n = 8192
    model = Model({'data_type': 'nerf', 'random_seed': 2}, device=dev)
    model.assume_all_foreground = True
    host = bench.synth_view(n, 2000, 0)
    d = {k: torch.from_numpy(host[k]).to(dev) for k in host}
    batch = ('synthetic', torch.zeros((n, 2), dtype=torch.int32, device=dev), d['rayo'], d['rayd'], d['rgb'], d['alpha'],
             d['pred_alpha'], d['xyz'], d['normal'], d['lvis'])
    opt = T.Adam(learning_rate=5e-4)
    for _ in range(2):
        loss, _, _ = T.train_iter(model, batch, opt, n // 2, thres=[0.0] * 3 + [0.3] * 12)
    torch.cuda.synchronize()
    print('train ok', float(loss))


def vq(k):
    from vqnerf_release_b200 import abi
    dev = torch.device('cuda:0')
    g = torch.Generator(device=dev).manual_seed(0)
    lat = torch.nn.functional.normalize(torch.randn((1 << 22, 256), generator=g, device=dev), dim=1)
    cb = torch.nn.functional.normalize(torch.randn((256, k), generator=g, device=dev), dim=0)
    for _ in range(2):
        out = abi.vq_assign(lat, cb, want_quantize=False)      # indices only: the tcgen05 kernel for K > 32
    torch.cuda.synchronize()
    print('vq ok', k, int(out['indices'].sum()))


def sdf_grad():
    from vqnerf_release_b200.neus.fields import SDFNetwork
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    net = SDFNetwork(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=(4,), multires=6, bias=0.5, scale=1.0,
                     geometric_init=True, weight_norm=True, device=dev)
    pts = torch.rand((1 << 20, 3), device=dev) * 2 - 1
    feat = torch.empty((1 << 20, 320), device=dev)
    for _ in range(2):
        s1, _, g1 = net.forward_with_gradient(pts, feat_out=feat)
        s2, _, g2 = net.forward_with_gradient(pts, want_feat=False)
    torch.cuda.synchronize()
    print('sdf_grad ok', float(s1.mean()), float((g1 - g2).abs().max()))


if __name__ == '__main__':
    mode = sys.argv[1]
    if mode == 'neus_scan':
        neus_scan(int(sys.argv[2]) if len(sys.argv) > 2 else 65536)
    elif mode == 'train':
        train()
    elif mode == 'vq':
        vq(int(sys.argv[2]))
    else:
        sdf_grad()
