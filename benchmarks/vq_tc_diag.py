"""Diagnostic for the K <= 128 tcgen05 VQ kernel: every row against the float64 arg-min, repeated, per debug variant.

    python benchmarks/vq_tc_diag.py [--n 4194304] [--ks 64,128] [--flags 0,1,2,4,8,16]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def ref_argmin(lat, cb, chunk=262144):
    idx, gap, d1, d2 = [], [], [], []
    c64 = cb.double()
    c2 = (c64 * c64).sum(0, keepdim=True)
    for i in range(0, lat.shape[0], chunk):
        x = lat[i:i + chunk].double()
        d = (x * x).sum(1, keepdim=True) - 2 * x @ c64 + c2
        t = torch.topk(d, 2, dim=1, largest=False)
        idx.append(t.indices[:, 0])
        gap.append((t.values[:, 1] - t.values[:, 0]) / t.values[:, 0].abs().clamp_min(1e-12))
    return torch.cat(idx), torch.cat(gap)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=4 * 1024 * 1024)
    ap.add_argument('--ks', default='64,128')
    ap.add_argument('--flags', default='0,1,2,4,8,16')
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--detail', type=int, default=24)
    args = ap.parse_args()
    from vqnerf_release_b200 import abi
    dev = torch.device('cuda:0')
    g = torch.Generator(device=dev).manual_seed(0)
    lat = abi.l2_normalize_rows(torch.rand((args.n, 256), generator=g, device=dev))
    for k in [int(v) for v in args.ks.split(',')]:
        cb = abi.get_codebook(torch.rand((256, k), generator=g, device=dev))
        ref, gap = ref_argmin(lat, cb)
        for fl in [int(v) for v in args.flags.split(',')]:
            os.environ['VQN_VQ_TC_FLAGS'] = str(fl)
            dbg = torch.zeros((args.n, 8), device=dev)
            os.environ['VQN_VQ_TC_DBG'] = hex(dbg.data_ptr())
            outs = []
            for rep in range(args.reps):
                out = abi.vq_assign(lat, cb, want_quantize=False)['indices']
                torch.cuda.synchronize()
                outs.append(out.clone())
            os.environ.pop('VQN_VQ_TC_DBG', None)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for rep in range(5):
                abi.vq_assign(lat, cb, want_quantize=False)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            recs = []
            for out in outs:
                mism = out != ref
                recs.append((int(mism.sum()), int((mism & (gap >= 1e-6)).sum())))
            same = all(bool((o == outs[0]).all()) for o in outs[1:])
            print(json.dumps({'n': args.n, 'K': k, 'flags': fl, 'mismatch(all,outside_tol)': recs,
                              'deterministic': same, 'ms': round(ms, 4)}), flush=True)
            if fl == 0 or recs[-1][0] <= 50:
                out = outs[-1]
                bad = torch.nonzero((out != ref) & (gap >= 1e-6))[:, 0]
                c64 = cb.double()
                for row in bad[:args.detail].tolist():
                    x = lat[row].double()
                    d = (c64 * c64).sum(0) - 2 * x @ c64
                    dd = dbg[row].tolist()
                    print('  row %d r=%d cta=%d it=%d ref=%d got=%d gap=%.3g  true d[ref]=%.8f d[got]=%.8f | kernel b=%.8f s=%.8f bi=%d si=%d xs=%.5f near=%d'
                          % (row, row % 128, (row // 128) % 148, (row // 128) // 148, int(ref[row]), int(out[row]),
                             float(gap[row]), float(d[int(ref[row])]), float(d[int(out[row])]), dd[0], dd[1], int(dd[2]),
                             int(dd[3]), dd[4], int(dd[5])), flush=True)
                if len(bad):
                    r = bad % 128
                    print('  rows-in-tile histogram (by warp quarter):', torch.bincount(r // 32, minlength=4).tolist(),
                          ' iteration histogram:', torch.bincount((bad // 128) // 148).tolist(), flush=True)
    os.environ.pop('VQN_VQ_TC_DBG', None)
    os.environ.pop('VQN_VQ_TC_FLAGS', None)


if __name__ == '__main__':
    main()
