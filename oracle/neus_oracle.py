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


# ---------------------------------------------------------------------------------------------
# SDFNetwork / RenderingNetwork (geo/NeuS-ours2/models/fields.py:9-172, embedder.py:6-50), float64 restatement.
# PINNED: tests/test_oracle_cpu.py compares these with tests/golden/neus_fields_ref.npz, recorded by
# oracle/gen_golden_neus_fields.py from the reference's own modules (autograd gradient included).
# ---------------------------------------------------------------------------------------------
def make_neus_state(seed=0, d_hidden=256, n_layers=8, skip_in=(4,), multires=6, d_feature=256, bias=0.5,
                    color_layers=4, multires_view=4):
    """Random-init parameters under the reference's state_dict names (lin{l}.weight_g / .weight_v / .bias), drawn with
    NumPy so that the golden generator (which loads them into the reference modules) and the tests build identical
    networks without 3 MB of weights in the repository.  SDF net: the geometric initialisation of fields.py:45-61;
    colour net: nn.Linear's default U(-1/sqrt(in), 1/sqrt(in)).  weight_g = |v| * U(0.9, 1.1) (weight_norm starts at
    g = |v|; the factor makes the normalisation observable)."""
    rng = np.random.RandomState(seed)
    d0 = 3 + 6 * multires
    dims = [d0] + [d_hidden] * n_layers + [1 + d_feature]
    sdf = {}
    for l in range(len(dims) - 1):
        out_dim = dims[l + 1] - dims[0] if (l + 1) in skip_in else dims[l + 1]
        if l == len(dims) - 2:
            w = rng.normal(np.sqrt(np.pi) / np.sqrt(dims[l]), 1e-4, size=(out_dim, dims[l]))
            b = np.full(out_dim, -bias)
        elif l == 0:
            w = np.zeros((out_dim, dims[l]))
            w[:, :3] = rng.normal(0.0, np.sqrt(2) / np.sqrt(out_dim), size=(out_dim, 3))
            b = np.zeros(out_dim)
        elif l in skip_in:
            w = rng.normal(0.0, np.sqrt(2) / np.sqrt(out_dim), size=(out_dim, dims[l]))
            w[:, -(dims[0] - 3):] = 0.0
            b = np.zeros(out_dim)
        else:
            w = rng.normal(0.0, np.sqrt(2) / np.sqrt(out_dim), size=(out_dim, dims[l]))
            b = np.zeros(out_dim)
        # the zeroed high-frequency columns would make the embedding a no-op: give them small weights so that the
        # parity tests exercise every column
        w = w + rng.normal(0.0, 0.02 / np.sqrt(out_dim), size=w.shape) * (w == 0.0)
        g = np.linalg.norm(w, axis=1, keepdims=True) * rng.uniform(0.9, 1.1, size=(out_dim, 1))
        sdf['lin%d.weight_v' % l] = w.astype(np.float32)
        sdf['lin%d.weight_g' % l] = g.astype(np.float32)
        sdf['lin%d.bias' % l] = (b + rng.normal(0.0, 0.001, size=b.shape) * (l < len(dims) - 2)).astype(np.float32)
    cd = [9 + d_feature + 6 * multires_view] + [d_hidden] * color_layers + [3]
    col = {}
    for l in range(len(cd) - 1):
        k = 1.0 / np.sqrt(cd[l])
        w = rng.uniform(-k, k, size=(cd[l + 1], cd[l]))
        col['lin%d.weight_v' % l] = w.astype(np.float32)
        col['lin%d.weight_g' % l] = (np.linalg.norm(w, axis=1, keepdims=True)
                                     * rng.uniform(0.9, 1.1, size=(cd[l + 1], 1))).astype(np.float32)
        col['lin%d.bias' % l] = rng.uniform(-k, k, size=cd[l + 1]).astype(np.float32)
    return {'sdf': sdf, 'color': col}


def effective_weight(state, l):
    """nn.utils.weight_norm (fields.py:63-64, dim=0): w = g * v / |v|, norm over the input axis of each output row."""
    if 'lin%d.weight' % l in state:
        return np.asarray(state['lin%d.weight' % l], np.float64)
    v = np.asarray(state['lin%d.weight_v' % l], np.float64)
    g = np.asarray(state['lin%d.weight_g' % l], np.float64)
    return g * v / np.linalg.norm(v, axis=1, keepdims=True)


def n_lin(state):
    return len([k for k in state if k.endswith('.bias')])


def embed_jac(x, multires):
    """embedder.py:11-35: [x, sin(x f0), cos(x f0), ...], f_k = 2^k; also d embed / d x  [n, 3, dim]."""
    x = np.asarray(x, np.float64)
    n = x.shape[0]
    cols, jac = [x], [np.broadcast_to(np.eye(3), (n, 3, 3))]
    for k in range(multires):
        f = 2.0 ** k
        s, c = np.sin(x * f), np.cos(x * f)
        cols += [s, c]
        jac += [np.eye(3)[None] * (f * c)[:, None, :], np.eye(3)[None] * (-f * s)[:, None, :]]
    return np.concatenate(cols, -1), np.concatenate(jac, -1)


def _softplus100(x):
    bt = 100.0 * x
    return np.where(bt > 20.0, x, np.log1p(np.exp(np.minimum(bt, 20.0))) / 100.0)


def sdf_forward(state, x, multires=6, skip_in=(4,), scale=1.0, want_grad=True):
    """SDFNetwork.forward + .gradient (fields.py:74-112): returns (out [n, 1 + d_feature], d sdf / d x [n, 3]).
    The gradient is propagated forward (Jacobian-vector form), which equals the reference's autograd result."""
    x = np.asarray(x, np.float64) * scale
    e, de = embed_jac(x, multires)
    de = de * scale
    h, dh = e, (de if want_grad else None)
    L = n_lin(state)
    for l in range(L):
        if l in skip_in:
            h = np.concatenate([h, e], -1) / np.sqrt(2)
            if want_grad:
                dh = np.concatenate([dh, de], -1) / np.sqrt(2)
        w = effective_weight(state, l)
        b = np.asarray(state['lin%d.bias' % l], np.float64)
        pre = h @ w.T + b
        dpre = dh @ w.T if want_grad else None
        if l < L - 1:
            h = _softplus100(pre)
            if want_grad:
                dh = dpre * _sigmoid(100.0 * pre)[:, None, :]
        else:
            h, dh = pre, dpre
    out = np.concatenate([h[:, :1] / scale, h[:, 1:]], -1)
    return out, (dh[:, :, 0] / scale if want_grad else None)


def color_forward(state, points, normals, view_dirs, feature_vectors, multires_view=4, squeeze_out=True):
    """RenderingNetwork.forward, mode 'idr' (fields.py:147-172)."""
    v, _ = embed_jac(view_dirs, multires_view)
    h = np.concatenate([np.asarray(points, np.float64), v, np.asarray(normals, np.float64),
                        np.asarray(feature_vectors, np.float64)], -1)
    L = n_lin(state)
    for l in range(L):
        h = h @ effective_weight(state, l).T + np.asarray(state['lin%d.bias' % l], np.float64)
        if l < L - 1:
            h = np.maximum(h, 0.0)
    return _sigmoid(h) if squeeze_out else h


# ---------------------------------------------------------------------------------------------
# NeuSRenderer.render (renderer.py:299-401, n_outside == 0, perturb == 0) and Runner.compute_vis
# (gen_geo.py:182-257, intersect_circle :346-357) assembled from the pieces above.
# PINNED by tests/golden/neus_vis_ref.npz (oracle/gen_golden_neus_vis.py: the reference's own renderer + networks).
# ---------------------------------------------------------------------------------------------
def render(state, rays_o, rays_d, near, far, radius, inv_s, cos_anneal_ratio=0.0, background_rgb=None,
           n_samples=64, n_importance=64, up_sample_steps=4, need_color=True):
    f32 = np.float32
    rays_o, rays_d = np.asarray(rays_o, f32), np.asarray(rays_d, f32)
    near, far = np.asarray(near, f32).reshape(-1, 1), np.asarray(far, f32).reshape(-1, 1)
    b = rays_o.shape[0]
    sample_dist = 2 * radius / n_samples
    z = (near + (far - near) * np.linspace(0.0, 1.0, n_samples, dtype=f32)[None, :]).astype(f32)
    sdf_of = lambda zz: sdf_forward(state['sdf'], (rays_o[:, None, :] + rays_d[:, None, :] * zz[..., None])
                                    .reshape(-1, 3), want_grad=False)[0][:, 0].reshape(zz.shape).astype(f32)
    if n_importance > 0:
        sdf = sdf_of(z)
        for i in range(up_sample_steps):
            new_z = up_sample(rays_o, rays_d, z, sdf, radius, n_importance // up_sample_steps, 64 * 2 ** i)
            if i + 1 == up_sample_steps:
                z, _ = cat_z_vals(z, new_z)
            else:
                z, sdf = cat_z_vals(z, new_z, sdf, sdf_of(new_z))
    s = z.shape[1]
    dists = np.concatenate([z[:, 1:] - z[:, :-1], np.full((b, 1), sample_dist, f32)], -1)
    mid = z + dists * f32(0.5)
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * mid[..., None]).reshape(-1, 3)
    out, grad = sdf_forward(state['sdf'], pts)
    if need_color:
        dirs = np.broadcast_to(rays_d[:, None, :], (b, s, 3)).reshape(-1, 3)
        col = color_forward(state['color'], pts, grad, dirs, out[:, 1:])
    else:
        col = np.zeros((b * s, 3))
    return composite(rays_o, rays_d, z, out[:, 0], grad, col, inv_s, cos_anneal_ratio, sample_dist, radius,
                     background_rgb)


def light_rays(surf, normal, lxyz, radius):
    """gen_geo.py:203-228: directions to every light, front-lit mask, near / far of the visibility rays (float32)."""
    f32 = np.float32
    surf, normal, lxyz = np.asarray(surf, f32), np.asarray(normal, f32), np.asarray(lxyz, f32).reshape(-1, 3)
    d = lxyz[None, :, :] - surf[:, None, :]
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    front = np.einsum('ijk,ik->ij', d, normal) > 0
    x = np.broadcast_to(surf[:, None, :], d.shape)
    bq = f32(2.) * np.sum(x * d, -1)
    a = np.sum(d * d, -1)
    c = np.sum(x * x, -1) - f32(radius) ** 2
    denom = np.where(2 * a > 1e-7, 2 * a, f32(1e-7))
    disc = np.sqrt(np.square(bq) - f32(4.) * a * c)
    t1, t2 = (-bq + disc) / denom, (-bq - disc) / denom
    far = np.where(t1 > t2, t1, t2).astype(f32)
    near = np.minimum(f32(0.1), far / f32(2.)).astype(f32)
    return d.astype(f32), front, near, far


def compute_vis(state, surf, normal, lxyz, radius, inv_s, cos_anneal_ratio=1.0):
    """gen_geo.py:182-257 without the dataset / file handling: lvis [N, n_lights] = 1 - weight_sum, 0 where back-lit."""
    d, front, near, far = light_rays(surf, normal, lxyz, radius)
    n, nl = front.shape
    lvis = np.zeros((n, nl), np.float32)
    o = np.broadcast_to(np.asarray(surf, np.float32)[:, None, :], d.shape)
    r = render(state, o[front], d[front], near[front], far[front], radius, inv_s, cos_anneal_ratio, None,
               need_color=False)
    lvis[front] = 1.0 - r['weight_sum'][:, 0]
    return lvis
