"""Per-call time of the training Dense kernels (forward / backward-data) at the config #4 sizes.
VQN_DENSE_TC_MIN_M selects the path (default 1024: tcgen05 from 1024 rows; a huge value: warp-level mma.sync)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from vqnerf_release_b200 import abi   # noqa: E402


def main():
    dev = torch.device('cuda:0')
    m = 8192
    for (k, n) in ((63, 128), (128, 128), (191, 128), (128, 256), (256, 256), (384, 3)):
        ldx = (k + 3) // 4 * 4
        X = torch.randn((m, ldx), device=dev)
        W = torch.randn((k, n), device=dev) * 0.1
        b = torch.zeros((n,), device=dev)
        Y = torch.zeros((m, n), device=dev)
        dX = torch.zeros((m, ldx), device=dev)
        res = []
        for name, fn in (('fwd', lambda: abi.dense_forward(X, ldx, W, b, Y, n, m, k, n, 1, 1.0, 0.0)),
                         ('bwd', lambda: abi.dense_backward_data(Y, n, W, dX, ldx, X, ldx, 1, False, m, k, n))):
            for _ in range(5):
                fn()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(20):
                    fn()
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            res.append('%s %.1f us' % (name, e0.elapsed_time(e1) / 200 * 1e3))
        print('m=%d k=%d n=%d: %s' % (m, k, n, ', '.join(res)), flush=True)


if __name__ == '__main__':
    main()
