"""Mirror of decomp/nerfvq_nfr3/nerfactor/util/math.py:63-64."""
import torch

from ... import abi


def safe_l2_normalize(x: torch.Tensor, axis=None, eps: float = 1e-6) -> torch.Tensor:
    """tf.linalg.l2_normalize(x, axis, epsilon=1e-6) for the shapes the hot path uses:
    rows of an [n,d] tensor (axis in {1,-1}) or columns of a [Z,K] codebook (axis=0, via transpose)."""
    if eps != 1e-6:
        raise ValueError('the kernels hard-code epsilon=1e-6 (util/math.py:63)')
    if x.dim() != 2:
        shp = x.shape
        if axis in (-1, x.dim() - 1):
            return abi.l2_normalize_rows(x.reshape(-1, shp[-1])).reshape(shp)
        raise ValueError('unsupported shape/axis for safe_l2_normalize')
    if axis in (1, -1):
        return abi.l2_normalize_rows(x)
    if axis == 0:
        return abi.l2_normalize_rows(x.t().contiguous()).t().contiguous()
    raise ValueError('axis must be 0, 1 or -1')
