"""VQ codebook-assignment sweep (BASELINE.json configs[2]): N latents x K codewords, Z = 256.

    python benchmarks/vq_sweep.py [--n 4194304] [--ks 8,15,32,...] [--check]

Prints one JSON line per (N, K): assigns/s, GB/s of algorithmic bytes (N*(4Z+8) + 4ZK) and the fraction of the
roofline max(bytes / HBM peak, flops / 3xTF32 tensor peak).  --check compares EVERY row with the float64 arg-min of the
reference's distance formula, computed on the GPU in chunks (bit-exact except rows whose top-2 gap is < 1e-6 relative),
and runs the kernel a second time to assert bit-identical output.  --n takes a comma-separated list (BASELINE.json
configs[2]: 1 M .. 64 M latents)."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', default=str(4 * 1024 * 1024))
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
    tf32x3 = 1638.4 / 2 / 3
    for n in [int(v) for v in args.n.split(',')]:
        g = torch.Generator(device=dev).manual_seed(0)
        lat = torch.empty((n, 256), device=dev)
        for i in range(0, n, 1 << 22):                      # generated and normalised in slices: 64 M rows are 64 GB
            lat[i:i + (1 << 22)] = abi.l2_normalize_rows(torch.rand((min(1 << 22, n - i), 256), generator=g, device=dev))
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
            nbytes = n * (4 * 256 + 8) + 4 * 256 * k
            flops = 2.0 * n * 256 * k + 3.0 * n * 256
            bound_ms = max(nbytes / hbm / 1e6, flops / tf32x3 / 1e9)
            rec = {'n': n, 'K': k, 'ms': round(ms, 4), 'assigns_per_s': n / ms * 1e3,
                   'gbs': nbytes / ms / 1e6, 'hbm_frac': nbytes / ms / 1e6 / hbm, 'tflops': flops / ms / 1e9,
                   'roofline_frac': bound_ms / ms}
            if args.check:
                idx = out['indices']
                again = abi.vq_assign(lat, cb, want_quantize=False)['indices']
                rec['deterministic'] = bool(torch.equal(idx, again))
                c64 = cb.double()
                c2 = (c64 * c64).sum(0, keepdim=True)
                mism = outside = 0
                for i in range(0, n, 262144):
                    x64 = lat[i:i + 262144].double()
                    d = (x64 * x64).sum(1, keepdim=True) - 2 * x64 @ c64 + c2
                    top2 = torch.topk(d, min(2, k), dim=1, largest=False).values
                    ref = torch.argmax((d == top2[:, :1]).to(torch.int8), dim=1)      # first arg-min
                    gap = (top2[:, -1] - top2[:, 0]) / top2[:, 0].abs().clamp_min(1e-12) if k > 1 else \
                        torch.ones(x64.shape[0], device=dev)
                    mm = idx[i:i + 262144] != ref
                    mism += int(mm.sum())
                    outside += int((mm & (gap >= 1e-6)).sum())
                rec['rows_checked'] = n
                rec['mismatch'] = mism
                rec['mismatch_outside_tol'] = outside
            print(json.dumps(rec), flush=True)
        del lat


if __name__ == '__main__':
    main()
