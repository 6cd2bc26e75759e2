"""Per-layer time stamps of the fused TRAINING launches (vqn_net_forward_train / vqn_net_backward_train) of one head network
at 8192 rows (one 128-row tile per CTA): CTA 0's MMA warp, tile 0 -- diagnostic.

    VQN_EXTRA_NVCC_FLAGS=-DVQN_TC_TRACE python -m vqnerf_release_b200.build --force
    python benchmarks/tc_trace_train.py [--rows 8192]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows', type=int, default=8192)
    args = ap.parse_args()
    from vqnerf_release_b200 import _lib as L, abi
    dev = torch.device('cuda:0')
    ctx = L.Context.get(dev)
    n = args.rows
    g = torch.Generator(device='cpu').manual_seed(0)
    shapes = {'head [256,128,3] skip 1': (256, [256, 128, 3], ['relu', 'relu', 'sigmoid'], 1),
              'bottleneck [128,256,256]': (128, [128, 256, 256], [None, 'relu', 'sigmoid'], None),
              'fine_enc [128]*4 skip 2': (64, [128] * 4, ['relu'] * 4, 2)}
    buf = torch.zeros((320 + 8 * 96,), dtype=torch.int64, device=dev)
    ctx.lib.vqn_debug_tc_trace.argtypes = [C.c_void_p]
    pad4 = lambda v: (v + 3) // 4 * 4
    for name, (in_dim, widths, acts, skip) in shapes.items():
        Ws, bs, d = [], [], in_dim
        for i, w in enumerate(widths):
            Ws.append((torch.randn((d, w), generator=g) / np.sqrt(d)).to(dev))
            bs.append(torch.zeros((w,), device=dev))
            d = w + (in_dim if skip == i else 0)
        net = abi.PackedNet(Ws, bs, acts, skip_at=skip)
        ld = [pad4(w + (in_dim if skip == i else 0)) for i, w in enumerate(widths)]
        y = [torch.zeros((n, l), device=dev) for l in ld]
        dz = [torch.zeros((n, pad4(w)), device=dev) for w in widths]
        lddz = [t.shape[1] for t in dz]
        x = torch.randn((n, in_dim), generator=g).to(dev)
        dz[-1].normal_()
        d_in = torch.zeros((n, in_dim), device=dev)
        head = skip == 1
        net.repack_tc('tf32x3')

        def fwd():
            net.forward_train(x, in_dim, n, y, ld)

        def bwd():
            if skip == 2:
                net.backward_train(dz[-1], lddz[-1], n, y, ld, dz, lddz)
            else:
                net.backward_train(dz[-1], lddz[-1], n, y, ld, dz, lddz, d_in, in_dim, 2 if head else 0)

        for label, fn in (('forward (MODE 3)', fwd), ('backward (MODE 4)', bwd)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            buf.zero_()
            ctx.lib.vqn_debug_tc_trace(C.c_void_p(buf.data_ptr()))
            fn()
            torch.cuda.synchronize()
            ctx.lib.vqn_debug_tc_trace(None)
            t = buf.cpu().numpy()[:320].reshape(4, 20, 4)
            nl = int((t[0, :, 3] != 0).sum())
            print('== %s, %s: %.1f us per launch (%d rows)' % (name, label, us, n))
            if nl == 0:
                print('   (no stamps: build with -DVQN_TC_TRACE)')
                continue
            base = t[0, 0, 0]
            for l in range(nl):
                a, b, c, d = t[0, l]
                print('   layer %2d: begin +%6d | drain-wait %5d | first-chunk wait %5d | issue %6d' % (l, a - base, b - a, c - b, d - c))
            print('   last MMA issued %d cycles after the first layer began (%.1f us at 1.965 GHz)'
                  % (t[0, nl - 1, 3] - base, (t[0, nl - 1, 3] - base) / 1965.0))
            k = t[3, 19]
            if k[0]:
                print('   kernel (CTA 0, thread 0): entry -> prologue done %d | -> first layer begins %d | last MMA issued -> thread 0 '
                      'past its final drain %d | -> all warps done + TMEM freed %d | entry -> exit %d cycles (%.1f us)'
                      % (k[1] - k[0], base - k[1], k[2] - t[0, nl - 1, 3], k[3] - k[2], k[3] - k[0], (k[3] - k[0]) / 1965.0))


if __name__ == '__main__':
    main()
