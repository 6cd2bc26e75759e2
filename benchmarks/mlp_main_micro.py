"""mlp_main (encoder + bottleneck + three heads, ONE launch) alone: CUDA-event time, TFLOP/s and a cross-process output check.

    python benchmarks/mlp_main_micro.py [--points 640000] [--precision tf32x3] [--save out.pt | --compare out.pt]
`--save` stores the three head outputs of a fixed seed; `--compare` asserts that this process (e.g. VQN_TC_CG2=1) produces
the same values as the saved ones (e.g. VQN_TC_CG2=0) and both agree with the fp32 FFMA kernel."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

FLOP_PER_POINT = 953600


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--points', type=int, default=640000)
    ap.add_argument('--precision', default='tf32x3')
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--save')
    ap.add_argument('--compare')
    args = ap.parse_args()
    from vqnerf_release_b200 import abi
    from vqnerf_release_b200.nerfactor.models.vq_nfr import Model
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    m = Model({'data_type': 'nerf', 'precision': args.precision}, device=dev)
    g = torch.Generator(device='cpu').manual_seed(1)
    xyz = (torch.rand((args.points, 3), generator=g) * 2 - 1).to(dev)
    nf = m.embedder['xyz'].n_freqs
    nets = [m.net[k].packed for k in ('fine_enc', 'bottleneck', 'diff_main', 'spec_main', 'rough_main')]

    def run(prec):
        return abi.mlp_main(*nets, nf, xyz, precision=prec)[1:]

    outs = run(args.precision)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(3):
        run(args.precision)
    ev[0].record()
    for _ in range(args.reps):
        run(args.precision)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / args.reps
    print('mlp_main %s cg2=%s: %d points %.4f ms = %.1f TFLOP/s (fp32-equivalent)' % (
        args.precision, os.environ.get('VQN_TC_CG2', 'default'), args.points, ms,
        FLOP_PER_POINT * args.points / ms / 1e9))
    ref = run('fp32')
    for name, a, b in zip(('basecolor', 'ks', 'rough'), outs, ref):
        print('  %-9s max |tc - fp32 FFMA| = %.3e' % (name, float((a - b).abs().max())))
        assert float((a - b).abs().max()) < (1e-2 if args.precision == 'bf16' else 2e-5)
    if args.save:
        torch.save([o.cpu() for o in outs], args.save)
    if args.compare:
        old = torch.load(args.compare)
        for name, a, b in zip(('basecolor', 'ks', 'rough'), outs, old):
            d = float((a.cpu() - b).abs().max())
            print('  %-9s max |this - saved| = %.3e' % (name, d))
            assert d < 2e-6


if __name__ == '__main__':
    main()
