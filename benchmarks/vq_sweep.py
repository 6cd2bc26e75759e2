"""VQ codebook-assignment sweep (BASELINE.json configs[2]): N latents x K codewords, Z = 256.

    python benchmarks/vq_sweep.py [--n 4194304] [--ks 8,15,32,...] [--check]

Prints one JSON line per (N, K): assigns/s, GB/s of algorithmic bytes (N*(4Z+8) + 4ZK) and the fraction of the
roofline max(bytes / HBM peak, flops / FFMA peak).  --check compares the indices of a 256k-row sample with a
float64 torch reference on the GPU (bit-exact except rows whose top-2 gap is < 1e-6 relative)."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=4 * 1024 * 1024)
    ap.add_argument('--ks', default='8,15,16,32,64,128,256,512,1024')
    ap.add_argument('--check', action='store_true')
    ap.add_argument('--reps', type=int, default=5)
    args = ap.parse_args()
    from vqnerf_release_b200 import _lib, abi
    dev = torch.device('cuda:0')
    ctx = _lib.Context.get(dev)
    hbm = 6550.1
    try:
        hbm = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    g = torch.Generator(device=dev).manual_seed(0)
    lat = abi.l2_normalize_rows(torch.rand((args.n, 256), generator=g, device=dev))
    for k in [int(v) for v in args.ks.split(',')]:
        cb = abi.get_codebook(torch.rand((256, k), generator=g, device=dev))
        reps = args.reps if k <= 64 else 2
        for _ in range(2):
            out = abi.vq_assign(lat, cb, want_quantize=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = abi.vq_assign(lat, cb, want_quantize=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        nbytes = args.n * (4 * 256 + 8) + 4 * 256 * k
        flops = 2.0 * args.n * 256 * k + 3.0 * args.n * 256
        rec = {'n': args.n, 'K': k, 'ms': round(ms, 4), 'assigns_per_s': args.n / ms * 1e3,
               'gbs': nbytes / ms / 1e6, 'hbm_frac': nbytes / ms / 1e6 / hbm, 'tflops': flops / ms / 1e9}
        if args.check:
            m = min(args.n, 262144)
            x64, c64 = lat[:m].double(), cb.double()
            d = (x64 * x64).sum(1, keepdim=True) - 2 * x64 @ c64 + (c64 * c64).sum(0, keepdim=True)
            ref = d.argmin(1)
            top2 = torch.topk(d, min(2, k), dim=1, largest=False).values
            gap = (top2[:, -1] - top2[:, 0]) / top2[:, 0].abs().clamp_min(1e-12) if k > 1 else torch.ones(m, device=dev)
            mism = out['indices'][:m] != ref
            rec['mismatch'] = int(mism.sum())
            rec['mismatch_outside_tol'] = int((mism & (gap >= 1e-6)).sum())
        print(json.dumps(rec), flush=True)


if __name__ == '__main__':
    main()
