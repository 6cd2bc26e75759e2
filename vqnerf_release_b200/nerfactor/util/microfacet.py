"""Mirror of decomp/nerfvq_nfr3/nerfactor/util/microfacet.py:9-39 (materialising variant; the hot path uses
the fused vqn_shade kernel and never builds [N,512,3] tensors)."""
import torch

from ... import abi


def get_brdf(pts2l, pts2c, normal, albedo=None, rough=None, f0=None):
    n = pts2c.shape[0]
    dev = pts2c.device
    if albedo is None:
        albedo = torch.ones((n, 3), dtype=torch.float32, device=dev)
    if f0 is None:
        f0 = 0.91 * torch.ones((n, 3), dtype=torch.float32, device=dev)
    if rough is None:
        rough = torch.ones((n, 1), dtype=torch.float32, device=dev)
    if pts2l.shape[1] != 512:
        raise ValueError('get_brdf kernel is built for the 16x32 (512-light) probe')
    return abi.eval_brdf(pts2l, pts2c, normal, albedo, f0, rough)
