"""Per-layer time stamps of the tcgen05 MLP kernel's MMA thread (CTA 0, first 4 tiles) -- diagnostic.

    VQN_EXTRA_NVCC_FLAGS=-DVQN_TC_TRACE python -m vqnerf_release_b200.build --force    # the stamps are compiled out by default
    python benchmarks/tc_trace.py [--precision tf32x3]
Prints, per layer: wait for the previous drain, wait for the first A/W chunk, time to issue all chunks (which
includes waiting for A slots to be refilled), in SM clock cycles."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--precision', default='tf32x3')
    ap.add_argument('--points', type=int, default=640000)
    args = ap.parse_args()
    from vqnerf_release_b200 import _lib, abi
    from vqnerf_release_b200.nerfactor.models.vq_nfr import Model
    dev = torch.device('cuda:0')
    m = Model({'data_type': 'nerf', 'precision': args.precision}, device=dev)
    ctx = _lib.Context.get(dev)
    n = args.points
    xyz = torch.rand((n, 3), device=dev) * 2 - 1
    nf = m.embedder['xyz'].n_freqs
    buf = torch.zeros((320 + 8 * 96,), dtype=torch.int64, device=dev)      # 4 tiles x TC_MAX_LAYERS (20) x 4 stamps, then the producer trace
    ctx.lib.vqn_debug_tc_trace.argtypes = [C.c_void_p]
    for which in ('encoder', 'heads'):
        z = abi.pred_enc_at(m.net['fine_enc'].packed, m.net['bottleneck'].packed, nf, xyz, precision=args.precision)
        torch.cuda.synchronize()
        buf.zero_()
        ctx.lib.vqn_debug_tc_trace(C.c_void_p(buf.data_ptr()))
        if which == 'encoder':
            abi.pred_enc_at(m.net['fine_enc'].packed, m.net['bottleneck'].packed, nf, xyz, precision=args.precision)
        else:
            abi.pred_heads(m.net['diff_main'].packed, m.net['spec_main'].packed, m.net['rough_main'].packed, z, 1.0, 0.0,
                           args.precision)
        torch.cuda.synchronize()
        ctx.lib.vqn_debug_tc_trace(None)
        raw = buf.cpu().numpy()
        t = raw[:320].reshape(4, 20, 4)
        pt = raw[320:].reshape(96, 8)
        print('==', which, args.precision)
        for tile in (1, 2):
            base = t[tile, 0, 0]
            nl = int((t[tile, :, 3] != 0).sum())
            print(' tile %d (starts %d cycles after tile %d started)' % (tile, base - t[tile - 1, 0, 0], tile - 1))
            for l in range(nl):
                a, b, c, d = t[tile, l]
                print('   layer %2d: begin +%6d | drain-wait %5d | first-chunk wait %5d | issue %6d' % (l, a - base, b - a, c - b, d - c))
            print('   tile total: %d cycles' % (t[tile, nl - 1, 3] - base))
        print(' producer thread 0 (group 0), tile 1: per chunk [layer*1000+seg*100+chunk]: load+math | slot wait | split+store | fences | arrive | gap to next')
        rows = [r for r in pt if r[0] != 0]
        for i, r in enumerate(rows):
            gap = rows[i + 1][0] - r[5] if i + 1 < len(rows) else 0
            print('   %5d: %6d %6d %6d %6d %6d | %6d' % (r[6], r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], gap))


if __name__ == '__main__':
    main()
