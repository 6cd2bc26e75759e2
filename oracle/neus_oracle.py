"""CPU oracle for the NeuS per-ray scan path (TEST INFRASTRUCTURE ONLY -- see decomp_oracle.py header).

NumPy restatement of geo/NeuS-ours2/models/renderer.py: sample_pdf :39-69 (det=True), up_sample :131-175,
cat_z_vals :177-191 (sort half) and the compositing half of render_core :193-297 (n_outside == 0).
PINNED: tests/test_oracle_cpu.py checks every function here against tests/golden/neus_ref.npz, which
oracle/gen_golden_neus.py recorded from the reference's own code running on CPU torch.
"""
import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def sample_pdf_det(bins, weights, n_samples, dtype=np.float32):
    """renderer.py:39-69 with det=True"""
    weights = (weights + dtype(1e-5)).astype(dtype)
    pdf = weights / np.sum(weights, -1, keepdims=True)
    cdf = np.cumsum(pdf, -1, dtype=dtype)
    cdf = np.concatenate([np.zeros_like(cdf[..., :1]), cdf], -1)
    u = np.linspace(0. + 0.5 / n_samples, 1. - 0.5 / n_samples, n_samples).astype(dtype)
    u = np.broadcast_to(u, cdf.shape[:-1] + (n_samples,))
    inds = np.stack([np.searchsorted(cdf[i], u[i], side='right') for i in range(cdf.shape[0])])
    below = np.maximum(0, inds - 1)
    above = np.minimum(cdf.shape[-1] - 1, inds)
    cdf_b, cdf_a = np.take_along_axis(cdf, below, 1), np.take_along_axis(cdf, above, 1)
    bin_b, bin_a = np.take_along_axis(bins, below, 1), np.take_along_axis(bins, above, 1)
    denom = cdf_a - cdf_b
    denom = np.where(denom < 1e-5, np.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    return (bin_b + t * (bin_a - bin_b)).astype(dtype)


def up_sample(rays_o, rays_d, z_vals, sdf, r_limit, n_importance, inv_s, dtype=np.float32):
    """renderer.py:131-175"""
    rays_o, rays_d, z_vals, sdf = (np.asarray(a, dtype) for a in (rays_o, rays_d, z_vals, sdf))
    b, s = z_vals.shape
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]
    radius = np.linalg.norm(pts, axis=-1)
    inside = (radius[:, :-1] < r_limit) | (radius[:, 1:] < r_limit)
    sdf = sdf.reshape(b, s)
    prev_sdf, next_sdf = sdf[:, :-1], sdf[:, 1:]
    prev_z, next_z = z_vals[:, :-1], z_vals[:, 1:]
    mid_sdf = (prev_sdf + next_sdf) * dtype(0.5)
    cos_val = (next_sdf - prev_sdf) / (next_z - prev_z + dtype(1e-5))
    prev_cos = np.concatenate([np.zeros((b, 1), dtype), cos_val[:, :-1]], -1)
    cos_val = np.minimum(prev_cos, cos_val)
    cos_val = np.clip(cos_val, -1e3, 0.0).astype(dtype) * inside
    dist = next_z - prev_z
    prev_esti = mid_sdf - cos_val * dist * dtype(0.5)
    next_esti = mid_sdf + cos_val * dist * dtype(0.5)
    prev_cdf = _sigmoid(prev_esti * dtype(inv_s))
    next_cdf = _sigmoid(next_esti * dtype(inv_s))
    alpha = (prev_cdf - next_cdf + dtype(1e-5)) / (prev_cdf + dtype(1e-5))
    trans = np.cumprod(np.concatenate([np.ones((b, 1), dtype), 1. - alpha + dtype(1e-7)], -1), -1, dtype=dtype)[:, :-1]
    weights = alpha * trans
    return sample_pdf_det(z_vals, weights.astype(dtype), n_importance, dtype)


def cat_z_vals(z_vals, new_z, sdf=None, new_sdf=None):
    """renderer.py:177-191 (sort + index-gather of sdf)"""
    z = np.concatenate([z_vals, new_z], -1)
    index = np.argsort(z, axis=-1, kind='stable')
    z_out = np.take_along_axis(z, index, -1)
    sdf_out = None
    if sdf is not None and new_sdf is not None:
        s = np.concatenate([sdf.reshape(z_vals.shape), new_sdf.reshape(new_z.shape)], -1)
        sdf_out = np.take_along_axis(s, index, -1)
    return z_out, sdf_out


def composite(rays_o, rays_d, z_vals, sdf, gradients, sampled_color, inv_s, cos_anneal_ratio, sample_dist, radius,
              background_rgb=None, dtype=np.float32):
    """renderer.py:200-297, compositing half, n_outside == 0"""
    rays_o, rays_d, z_vals = (np.asarray(a, dtype) for a in (rays_o, rays_d, z_vals))
    b, s = z_vals.shape
    sdf = np.asarray(sdf, dtype).reshape(b * s, 1)
    gradients = np.asarray(gradients, dtype).reshape(b * s, 3)
    sampled_color = np.asarray(sampled_color, dtype).reshape(b, s, 3)
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = np.concatenate([dists, np.full((b, 1), sample_dist, dtype)], -1)
    mid_z = z_vals + dists * dtype(0.5)
    pts = rays_o[:, None, :] + rays_d[:, None, :] * mid_z[..., :, None]
    dirs = np.broadcast_to(rays_d[:, None, :], pts.shape).reshape(-1, 3)
    pts = pts.reshape(-1, 3)
    true_cos = np.sum(dirs * gradients, -1, keepdims=True)
    relu = lambda x: np.maximum(x, 0)
    iter_cos = -(relu(-true_cos * dtype(0.5) + dtype(0.5)) * dtype(1.0 - cos_anneal_ratio) +
                 relu(-true_cos) * dtype(cos_anneal_ratio))
    est_next = sdf + iter_cos * dists.reshape(-1, 1) * dtype(0.5)
    est_prev = sdf - iter_cos * dists.reshape(-1, 1) * dtype(0.5)
    prev_cdf = _sigmoid(est_prev * dtype(inv_s))
    next_cdf = _sigmoid(est_next * dtype(inv_s))
    p, c = prev_cdf - next_cdf, prev_cdf
    alpha = np.clip(((p + dtype(1e-5)) / (c + dtype(1e-5))).reshape(b, s), 0.0, 1.0).astype(dtype)
    pts_radius = np.linalg.norm(pts, axis=-1, keepdims=True).reshape(b, s)
    inside = (pts_radius < radius).astype(dtype)
    relax = (pts_radius < radius * 1.1).astype(dtype)
    trans = np.cumprod(np.concatenate([np.ones((b, 1), dtype), 1. - alpha + dtype(1e-7)], -1), -1, dtype=dtype)[:, :-1]
    weights = alpha * trans
    wsum = weights.sum(-1, keepdims=True)
    color = (sampled_color * weights[:, :, None]).sum(1)
    surf = (pts.reshape(b, s, 3) * weights[:, :, None]).sum(1)
    depth = np.linalg.norm(surf - rays_o, axis=-1, keepdims=True)
    if background_rgb is not None:
        color = color + np.asarray(background_rgb, dtype).reshape(1, 3) * (1.0 - wsum)
    gerr = (np.linalg.norm(gradients.reshape(b, s, 3), axis=-1) - 1.0) ** 2
    gerr = (relax * gerr).sum() / (relax.sum() + 1e-5)
    return {'color': color, 'weights': weights, 'surf': surf, 'depth': depth, 'cdf': c.reshape(b, s),
            'inside_sphere': inside, 'mid_z_vals': mid_z, 'dists': dists, 'weight_sum': wsum,
            'weight_max': weights.max(-1, keepdims=True), 'gradient_error': gerr}
