"""Decomp training-step micro-benchmark (BASELINE configs[3]) for profiling: eager train_iter steps on one GPU.

    python benchmarks/train_step.py [--rays 8192] [--steps 5] [--graph]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rays', type=int, default=8192)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--graph', action='store_true')
    args = ap.parse_args()
    from vqnerf_release_b200.nerfactor import train_nfr as T
    from vqnerf_release_b200.nerfactor.models.vq_nfr import Model
    dev = torch.device('cuda:0')
    n = args.rays
    model = Model({'data_type': 'nerf', 'random_seed': 2}, device=dev)
    model.assume_all_foreground = True
    host = bench.synth_view(n, 2000, 0)
    d = {k: torch.from_numpy(host[k]).to(dev) for k in host}
    batch = ('synthetic', torch.zeros((n, 2), dtype=torch.int32, device=dev), d['rayo'], d['rayd'], d['rgb'], d['alpha'],
             d['pred_alpha'], d['xyz'], d['normal'], d['lvis'])
    opt = T.Adam(learning_rate=5e-4)
    thres = [0.0] * 3 + [0.3] * 12
    step = T.GraphedTrainIter(model, opt, n // 2, batch) if args.graph else None
    run = (lambda: step(batch, thres=thres)) if args.graph else (lambda: T.train_iter(model, batch, opt, n // 2, thres=thres))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, _, _ = run()
    e1.record()
    torch.cuda.synchronize()
    print('ms/step %.3f  rays/s %.3e  loss %.5f' % (e0.elapsed_time(e1) / args.steps, n / (e0.elapsed_time(e1) / args.steps) * 1e3, float(loss)))


if __name__ == '__main__':
    main()
