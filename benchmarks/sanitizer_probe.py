"""Small workloads of the tcgen05 kernels for compute-sanitizer (one tool per run):
    compute-sanitizer --tool racecheck python benchmarks/sanitizer_probe.py vq
    compute-sanitizer --tool synccheck python benchmarks/sanitizer_probe.py mlp
vq: indices-only VQ assignment, K = 64 and 128, 8 tiles per CTA (TMEM ping-pong, ring wrap, drain warps all exercised);
mlp: the fused encoder + heads launch on 3 tiles per CTA."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from vqnerf_release_b200 import abi
    dev = torch.device('cuda:0')
    g = torch.Generator(device=dev).manual_seed(0)
    if sys.argv[1] == 'vq':
        n = 148 * 128 * 8
        lat = torch.nn.functional.normalize(torch.randn((n, 256), generator=g, device=dev), dim=1)
        for k in (64, 128):
            cb = torch.nn.functional.normalize(torch.randn((256, k), generator=g, device=dev), dim=0)
            idx = abi.vq_assign(lat, cb, want_quantize=False)['indices']
            ref = torch.argmin((cb * cb).sum(0)[None, :] - 2.0 * (lat.double() @ cb.double()).float(), dim=1)
            print('vq K=%d rows %d mismatches vs fp64-product arg-min: %d' % (k, n, int((idx != ref).sum())))
    else:
        from vqnerf_release_b200.nerfactor.models.vq_nfr import Model
        m = Model({'data_type': 'nerf'}, device=dev)
        n = 148 * 128 * 3
        xyz = torch.rand((n, 3), generator=g, device=dev) * 2 - 1
        nets = [m.net[k].packed for k in ('fine_enc', 'bottleneck', 'diff_main', 'spec_main', 'rough_main')]
        outs = abi.mlp_main(*nets, m.embedder['xyz'].n_freqs, xyz, precision='tf32x3')[1:]
        ref = abi.mlp_main(*nets, m.embedder['xyz'].n_freqs, xyz, precision='fp32')[1:]
        print('mlp_main rows %d max |tc - fp32| = %.2e' % (n, max(float((a - b).abs().max()) for a, b in zip(outs, ref))))
    torch.cuda.synchronize()


if __name__ == '__main__':
    main()
