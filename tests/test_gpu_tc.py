"""tcgen05 primitive self-test: one 128 x N x K tile through the same descriptor / swizzle / TMEM / mbarrier
helpers (csrc/tc_common.cuh) that the tensor-core MLP kernels use, against a CPU GEMM of identically rounded
operands."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tf32(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


@pytest.mark.parametrize('mode,n,k', [(0, 16, 32), (0, 128, 64), (0, 256, 128), (1, 16, 64), (1, 256, 128),
                                     (1, 128, 256), (2, 16, 32), (2, 256, 64), (2, 128, 64)])
def test_tc_selftest_gemm(cuda_dev, mode, n, k):
    from vqnerf_release_b200 import abi
    rng = np.random.RandomState(mode * 1000 + n + k)
    a = rng.normal(size=(128, k)).astype(np.float32)
    b = rng.normal(size=(n, k)).astype(np.float32)
    d = abi.tc_selftest(torch.as_tensor(a).to(cuda_dev), torch.as_tensor(b).to(cuda_dev), mode).cpu().numpy()
    if mode == 0:
        ref = _tf32(a).astype(np.float64) @ _tf32(b).astype(np.float64).T
        tol = 2e-5
    elif mode == 1:
        ar = torch.as_tensor(a).to(torch.bfloat16).double().numpy()
        br = torch.as_tensor(b).to(torch.bfloat16).double().numpy()
        ref = ar @ br.T
        tol = 2e-5
    else:
        ref = a.astype(np.float64) @ b.astype(np.float64).T
        tol = 3e-6            # 3xTF32: fp32-level accuracy (plain tf32 would be ~1e-3)
    scale = np.sqrt(k)
    err = np.abs(d - ref).max() / scale
    assert err < tol, 'mode %d N=%d K=%d: max err / sqrt(K) = %.3e' % (mode, n, k, err)
