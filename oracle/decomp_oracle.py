"""CPU oracle for the VQ-NeRF decomposition-stage shading path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product (`vqnerf_release_b200/`) may
import this module; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` do, and only as the
checker / the CPU baseline.

PARITY PINNED against the reference's own code, executed here: the reference's decomp stage is TensorFlow 2.4.1 +
dm-sonnet 2.0.0 + tensorflow-probability 0.12.1, none of which is installable in this image (no network), and the
reference ships no tests, fixtures or golden vectors for this path -- but its model code
is pure `tf.*` calls, so `oracle/gen_golden_decomp_ref.py` (and `gen_golden_ref_nfr.py`, `gen_golden_nfr_unit.py`,
`gen_golden_shape_unit.py`) import the UNMODIFIED reference modules with the `oracle/tf_shim` stand-in for
tensorflow / tensorflow_probability / sonnet first on sys.path, execute `Model.fast_render`, `Model.call`,
`Model.vq_test`, `Model.compute_loss`, `VectorQuantizerEMA.__call__`, `microfacet.get_brdf`, ... on torch-CPU and record
`tests/golden/{decomp_ref,ref_nfr_ref,nfr_unit_ref,shape_unit_ref}.npz`.  This file (an op-for-op restatement of the
reference's arithmetic in PyTorch-CPU: float64 = "truth", float32 = emulation of the TF fp32 op sequence, including
the [N,512,3] intermediates the reference materialises) is compared with those vectors to 1e-9 in
tests/test_oracle_vs_reference_cpu.py, test_ref_nfr_cpu.py, test_nfr_unit_cpu.py; the CUDA kernels are compared with
the same vectors in the `-m gpu` tests.  What remains a restatement is the semantics of the individual `tf.*` ops
inside the shim (oracle/tf_shim/README.md).  Each function cites the reference file:line it follows (paths relative
to /root/reference/decomp/nerfvq_nfr3/).

Third-party arithmetic restated from published sources (not vendored in the
reference): dm-sonnet 2.0.0 `moving_averages.ExponentialMovingAverage`
(zero-debiased EMA), tfp 0.12.1 `clip_by_value_preserve_gradient` (forward ==
clip, gradient == identity), TF `l2_normalize` (epsilon on the squared norm),
TF `divide_no_nan`, Keras Dense / glorot_uniform.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

# ----------------------------------------------------------------------------
# constants of the shipped config (nerfactor/config/vq_nfr.ini)
# ----------------------------------------------------------------------------
MLP_WIDTH = 128        # vq_nfr.ini:97
Z_DIM = 256            # vq_nfr.ini:100 (conv_width)
N_FREQS_XYZ = 10       # vq_nfr.ini:104
LIGHT_H = 16           # vq_nfr.ini:51
NUM_EMBED = 15         # vq_nfr.ini:113
COMMITMENT_COST = 0.1  # vq_nfr.ini:114

ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2


# ----------------------------------------------------------------------------
# TF primitive restatements
# ----------------------------------------------------------------------------
def safe_l2_normalize(x: torch.Tensor, axis: int, eps: float = 1e-6) -> torch.Tensor:
    """nerfactor/util/math.py:63-64 -> tf.linalg.l2_normalize(x, axis, epsilon):
    x * rsqrt(max(sum(x^2, axis), epsilon)) -- epsilon bounds the SQUARED norm."""
    sq = torch.sum(x * x, dim=axis, keepdim=True)
    return x * torch.rsqrt(torch.clamp(sq, min=eps))


def divide_no_nan(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """tf.math.divide_no_nan: 0 where the denominator is exactly 0."""
    b_safe = torch.where(b == 0, torch.ones_like(b), b)
    out = a / b_safe
    return torch.where(b == 0, torch.zeros_like(out), out)


def clip_preserve_grad(x: torch.Tensor, lo: float, hi: float) -> torch.Tensor:
    """tfp.math.clip_by_value_preserve_gradient: forward = clip, backward = identity."""
    return x + (torch.clamp(x, lo, hi) - x).detach()


# ----------------------------------------------------------------------------
# light probe geometry
# ----------------------------------------------------------------------------
def gen_light_xyz(envmap_h: int, envmap_w: int, envmap_radius: float = 1e2):
    """brdf/renderer.py:184-219 + third_party/xiuminglib/xiuminglib/geometry/sph.py:184-190.
    float64 NumPy, as the reference; the model casts to fp32 (vq_nfr.py:73-75)."""
    lat_step = np.pi / (envmap_h + 2)
    lng_step = 2 * np.pi / (envmap_w + 2)
    lats = np.linspace(np.pi / 2 - lat_step, -np.pi / 2 + lat_step, envmap_h)
    lngs = np.linspace(np.pi - lng_step, -np.pi + lng_step, envmap_w)
    lngs, lats = np.meshgrid(lngs, lats)
    r = envmap_radius * np.ones_like(lats)
    z = r * np.sin(lats)
    x = r * np.cos(lats) * np.cos(lngs)
    y = r * np.cos(lats) * np.sin(lngs)
    xyz = np.stack((x, y, z), axis=-1)
    sin_colat = np.sin(np.pi / 2 - lats)
    areas = 4 * np.pi * sin_colat / np.sum(sin_colat)
    return xyz, areas


# ----------------------------------------------------------------------------
# networks
# ----------------------------------------------------------------------------
def embed(x: torch.Tensor, n_freqs: int = N_FREQS_XYZ) -> torch.Tensor:
    """nerfactor/networks/embedder.py:23-47 with the kwargs of models/shape.py:82-89:
    [x, sin(x f0), cos(x f0), sin(x f1), ...], f_k = 2**linspace(0, n-1, n)."""
    outs = [x]
    for k in range(n_freqs):
        f = float(2.0 ** k)
        outs.append(torch.sin(x * f))
        outs.append(torch.cos(x * f))
    return torch.cat(outs, dim=-1)


@dataclass
class Net:
    """nerfactor/networks/mlp.py:24-50 -- Keras Dense stack, kernel [in,out], bias [out];
    after layer i in skip_at the output becomes concat(y, x_input) (y first)."""
    weights: List[np.ndarray]
    biases: List[np.ndarray]
    acts: List[int]
    skip_at: Optional[int] = None

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        dt = x.dtype
        x_ = x
        y = x
        for i, (w, b, a) in enumerate(zip(self.weights, self.biases, self.acts)):
            y = x_ @ torch.as_tensor(w, dtype=dt) + torch.as_tensor(b, dtype=dt)
            if a == ACT_RELU:
                y = torch.relu(y)
            elif a == ACT_SIGMOID:
                y = torch.sigmoid(y)
            if self.skip_at is not None and i == self.skip_at:
                y = torch.cat((y, x), dim=-1)
            x_ = y
        return y


def glorot_uniform(rng: np.random.RandomState, fan_in: int, fan_out: int) -> np.ndarray:
    """Keras Dense default kernel init: U(-l, l), l = sqrt(6/(fan_in+fan_out)); bias zeros."""
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=(fan_in, fan_out)).astype(np.float32)


def make_net(rng, in_dim: int, widths: Sequence[int], acts: Sequence[int],
             skip_at: Optional[int] = None, bias_scale: float = 0.0) -> Net:
    ws, bs = [], []
    d = in_dim
    for i, w in enumerate(widths):
        ws.append(glorot_uniform(rng, d, w))
        if bias_scale > 0:
            bs.append((rng.uniform(-1, 1, size=(w,)) * bias_scale).astype(np.float32))
        else:
            bs.append(np.zeros((w,), np.float32))
        d = w
        if skip_at is not None and i == skip_at:
            d = w + in_dim
    return Net(ws, bs, list(acts), skip_at)


def make_vq_nfr_nets(seed: int = 0, bias_scale: float = 0.0) -> Dict[str, Net]:
    """The 8 nets of models/vq_nfr.py:135-164 (VQ twins) + models/nfr_unit.py:110-129."""
    rng = np.random.RandomState(seed)
    R, S, N_ = ACT_RELU, ACT_SIGMOID, ACT_NONE
    emb = 3 + 3 * 2 * N_FREQS_XYZ
    nets = {}
    nets['fine_enc'] = make_net(rng, emb, [MLP_WIDTH] * 4, [R] * 4, skip_at=2, bias_scale=bias_scale)
    nets['bottleneck'] = make_net(rng, MLP_WIDTH, [MLP_WIDTH, Z_DIM, Z_DIM], [N_, R, S], bias_scale=bias_scale)
    for name, out in (('diff_main', 3), ('spec_main', 1), ('rough_main', 1),
                      ('diff_vq', 3), ('spec_vq', 3), ('rough_vq', 1)):
        nets[name] = make_net(rng, Z_DIM, [Z_DIM, Z_DIM // 2, out], [R, R, S], skip_at=1,
                              bias_scale=bias_scale)
    return nets


# ----------------------------------------------------------------------------
# model pieces
# ----------------------------------------------------------------------------
def calc_ldir(lxyz: torch.Tensor, pts: torch.Tensor) -> torch.Tensor:
    """models/shape.py:103-110"""
    surf2l = lxyz.reshape(1, -1, 3) - pts[:, None, :]
    return safe_l2_normalize(surf2l, axis=2)


def calc_vdir(cam_loc: torch.Tensor, pts: torch.Tensor) -> torch.Tensor:
    """models/shape.py:112-119"""
    return safe_l2_normalize(cam_loc - pts, axis=1)


def normal_correct(normal: torch.Tensor, surf2c: torch.Tensor) -> torch.Tensor:
    """models/vq_nfr.py:830-833"""
    cos = torch.sum(normal * surf2c, dim=-1, keepdim=True)
    return torch.where(cos >= 0, normal, -normal)


def pred_enc_at(nets: Dict[str, Net], pts: torch.Tensor) -> torch.Tensor:
    """models/vq_nfr.py:771-784 (chunking is arithmetic-neutral)."""
    return nets['bottleneck'](nets['fine_enc'](embed(pts)))


def pred_head(nets, name: str, z: torch.Tensor, slope: float = 1.0, bias: float = 0.0):
    """models/vq_nfr.py:786-828; slope/bias only for the diffuse heads (albedo_slope/bias)."""
    y = nets[name](z)
    if name.startswith('diff'):
        y = slope * y + bias
    return y


def get_codebook(raw_codebook: torch.Tensor) -> torch.Tensor:
    """models/vq_nfr.py:761-769: clip to [0,1] then l2-normalise each column ([Z,K])."""
    c = clip_preserve_grad(raw_codebook, 0.0, 1.0)
    return safe_l2_normalize(c, axis=0)


# ---- microfacet BRDF ---------------------------------------------------------
def _get_gsub(cos_theta, alpha):
    """util/microfacet.py:49-69 (_get_gl and _get_gv share this form)."""
    cos_theta = clip_preserve_grad(cos_theta, 0.0, 1.0)
    cos_theta_sq = cos_theta * cos_theta
    denom_a = torch.abs(alpha ** 2 + (1 - alpha ** 2) * cos_theta_sq)
    denom = cos_theta + torch.sqrt(denom_a)
    return divide_no_nan(2 * cos_theta, denom)


def get_brdf(pts2l, pts2c, normal, albedo, rough, f0):
    """util/microfacet.py:9-39 and helpers :41-89.  Returns (brdf, glossy, diffuse), each [N,L,3]."""
    pts2l = safe_l2_normalize(pts2l, axis=2)
    pts2c = safe_l2_normalize(pts2c, axis=1)
    normal = safe_l2_normalize(normal, axis=1)
    h = pts2l + pts2c[:, None, :]
    h = safe_l2_normalize(h, axis=2)
    # _get_f :82-89
    cos_hv = torch.einsum('ijk,ik->ij', h, pts2c)[:, :, None]
    cos_hv = clip_preserve_grad(cos_hv, 0.0, 1.0)
    f0_ = f0[:, None, :]
    f = f0_ + (1 - f0_) * (1 - cos_hv) ** 5
    alpha = rough ** 2                              # :25
    # _get_d :71-80
    a_ = alpha[:, None, :]
    cos_m = torch.einsum('ijk,ik->ij', h, normal)
    cos_m = clip_preserve_grad(cos_m, 0.0, 1.0)
    cos_m_sq = cos_m * cos_m
    denom_d = math.pi * (cos_m_sq[:, :, None] * (a_ ** 2 - 1) + 1) ** 2
    d = divide_no_nan(a_ ** 2, denom_d)
    # _get_g :41-69
    g_l = _get_gsub(torch.einsum('ijk,ik->ij', pts2l, normal)[:, :, None], a_)
    g_v = _get_gsub(torch.einsum('ij,ij->i', normal, pts2c)[:, None, None], a_)
    g = g_l * g_v
    l_dot_n = torch.einsum('ijk,ik->ij', pts2l, normal)[:, :, None]
    v_dot_n = torch.einsum('ij,ij->i', pts2c, normal)[:, None, None]
    denom = 4 * torch.abs(l_dot_n) * torch.abs(v_dot_n)
    glossy = divide_no_nan(f * g * d, denom)
    diffuse = (albedo / math.pi)[:, None, :].expand_as(glossy)
    return glossy + diffuse, glossy, diffuse


def linear2srgb(t: torch.Tensor) -> torch.Tensor:
    """util/img.py:142-165 (clip to [0,1] first, :62-75)."""
    t = torch.clamp(t, 0.0, 1.0)
    lin = t * 12.92
    nonlin = 1.055 * torch.pow(t, 1 / 2.4) - 0.055
    return torch.where(t <= 0.0031308, lin, nonlin)


def srgb2linear(t: torch.Tensor) -> torch.Tensor:
    """util/img.py:167-186"""
    lin = t / 12.92
    nonlin = torch.pow((t + 0.055) / 1.055, 2.4)
    return torch.where(t <= 0.04045, lin, nonlin)


def render(brdf, l, n, lareas, light, light_vis=None, probes: Optional[torch.Tensor] = None,
           gamma: Optional[Tuple[float, float]] = None):
    """models/vq_nfr.py:694-733 (_render / integrate).  `light` [16,32,3] is already
    clip(_light, 0, inf) (:759); `probes` [P,16,32,3] are the novel probes in dict order.
    gamma=(bias, index) only for data_type != 'nerf' (:715-716)."""
    cos = torch.einsum('ijk,ik->ij', l, n)
    areas = lareas.reshape(1, -1, 1)
    front_lit = (cos > 0).to(cos.dtype)
    lvis = front_lit if light_vis is None else front_lit * light_vis

    def integrate(lt):
        light_flat = lt.reshape(-1, 3)
        lgt = lvis[:, :, None] * light_flat[None, :, :]
        contrib = brdf * lgt * cos[:, :, None] * areas
        rgb = torch.sum(contrib, dim=1)
        if gamma is not None:
            rgb = (rgb * gamma[0]) ** gamma[1]
        return clip_preserve_grad(rgb, 0.0, 1.0)

    rgb = integrate(light)
    rgb_probes = None
    if probes is not None:
        rgb_probes = torch.stack([integrate(p) for p in probes], dim=1)
    return rgb, rgb_probes


# ---- vector quantiser ----------------------------------------------------------
class SonnetEMA:
    """dm-sonnet 2.0.0 sonnet/src/moving_averages.py ExponentialMovingAverage (not vendored;
    restated): update(v): counter += 1; hidden -= (hidden - v) * (1 - decay);
    average = hidden / (1 - decay**counter).  initialize() zero-fills hidden/average."""

    def __init__(self, shape, decay: float, dtype=torch.float32):
        self.decay = decay
        self.hidden = torch.zeros(shape, dtype=dtype)
        self.average = torch.zeros(shape, dtype=dtype)
        self.counter = 0

    def __call__(self, value: torch.Tensor) -> torch.Tensor:
        self.counter += 1
        self.hidden = self.hidden - (self.hidden - value) * (1.0 - self.decay)
        self.average = self.hidden / (1.0 - self.decay ** self.counter)
        return self.average


class VectorQuantizerEMA:
    """nerfactor/networks/vq_layers.py:208-349."""

    def __init__(self, embedding_dim, num_embeddings, commitment_cost, decay=0.999,
                 epsilon=1e-5, dtype=torch.float32):
        self.embedding_dim = embedding_dim
        self.num_embeddings = num_embeddings
        self.commitment_cost = commitment_cost
        self.decay = decay
        self.epsilon = epsilon
        self.ema_cluster_size = SonnetEMA([num_embeddings], decay, dtype)          # :249-251
        self.ema_dw = SonnetEMA([embedding_dim, num_embeddings], decay, dtype)     # :253-255

    def __call__(self, inputs, codebook, is_training, thres=None, roll=None):
        """`roll` replaces tf.random.uniform (:286-288) so the test can inject it."""
        dt = inputs.dtype
        flat = inputs.reshape(-1, self.embedding_dim)
        distances = (torch.sum(flat ** 2, 1, keepdim=True)
                     - 2 * (flat @ codebook)
                     + torch.sum(codebook ** 2, 0, keepdim=True))                  # :279-282
        if thres is not None:
            mask_value = torch.max(distances)
            sel_mask = (roll >= thres).to(dt)
            distances = distances * sel_mask + mask_value * (1.0 - sel_mask)       # :284-290
        idx = torch.argmax(-distances, 1)       # first max of -d == first min of d (:292)
        # torch.argmax does not guarantee first-index ties on all builds: enforce it.
        dmin = distances.min(dim=1, keepdim=True).values
        idx = torch.argmax((distances == dmin).to(torch.int8), dim=1)
        encodings = torch.nn.functional.one_hot(idx, self.num_embeddings).to(dt)   # :293
        quantized = codebook.t()[idx]                                              # :346-349
        e_latent = torch.mean((quantized.detach() - inputs) ** 2)                  # :302
        ret = {}
        if is_training:
            cnt = torch.sum(encodings, dim=0)
            cs = self.ema_cluster_size(cnt)                                        # :305-306
            dw = flat.t() @ encodings                                              # :308
            ema_dw = self.ema_dw(dw)                                               # :309
            n = torch.sum(cs)
            cs = (cs + self.epsilon) / (n + self.num_embeddings * self.epsilon) * n  # :311-313
            w = ema_dw / cs.reshape(1, -1)                                         # :315-316
            used = (cnt > 0).to(dt)                                                # :318
            ret['update'] = w * used[None, :] + codebook * (1.0 - used[None, :])   # :319
        loss = self.commitment_cost * e_latent                                     # :321,324
        quantized_ste = inputs + (quantized - inputs).detach()                     # :327
        avg_probs = torch.mean(encodings, 0)
        perplexity = torch.exp(-torch.sum(avg_probs * torch.log(avg_probs + 1e-10)))  # :328-330
        ret.update({'quantize': quantized_ste, 'loss': loss, 'perplexity': perplexity,
                    'encodings': encodings, 'encoding_indices': idx, 'distances': distances})
        return ret


def top2_gap_rel(distances: torch.Tensor) -> torch.Tensor:
    """Relative gap between the two smallest distances of each row -- the tolerance clause
    of BASELINE.json (indices bit-exact unless this gap < 1e-6)."""
    if distances.shape[1] < 2:
        return torch.full((distances.shape[0],), float('inf'), dtype=distances.dtype)
    v, _ = torch.topk(distances, 2, dim=1, largest=False)
    return (v[:, 1] - v[:, 0]) / torch.clamp(torch.abs(v[:, 0]), min=1e-30)


# ----------------------------------------------------------------------------
# whole-path entry points (forward only)
# ----------------------------------------------------------------------------
@dataclass
class Scene:
    """Replicated state of models/vq_nfr.py Model that the path reads."""
    nets: Dict[str, Net]
    light: np.ndarray                 # [16,32,3]  raw _light variable
    codebook: np.ndarray              # [Z,K]      raw _codebook variable
    probes: Optional[np.ndarray] = None   # [P,16,32,3]
    albedo_slope: float = 1.0
    albedo_bias: float = 0.0
    data_type: str = 'nerf'
    gamma: Tuple[float, float] = (1.0, 1.0)
    lxyz: np.ndarray = field(default_factory=lambda: gen_light_xyz(LIGHT_H, 2 * LIGHT_H)[0])
    lareas: np.ndarray = field(default_factory=lambda: gen_light_xyz(LIGHT_H, 2 * LIGHT_H)[1])


def _t(a, dt):
    if isinstance(a, torch.Tensor):       # autograd leaves of train_step() pass through
        return a.to(dt)
    return torch.as_tensor(np.asarray(a), dtype=dt)


def fast_render(scene: Scene, batch: Dict[str, np.ndarray], dtype=torch.float64,
                relight_probes: bool = False, opt_scale=None, gen_embed: bool = False,
                return_brdf: bool = False, dst_env: Optional[int] = None, edit_mask=None,
                edit_material=None) -> Dict[str, torch.Tensor]:
    """models/vq_nfr.py:262-398.  Returns the `pred` dict entries (full length, zeros at
    background rows) plus the compacted intermediates under '_'-prefixed keys.
    `dst_env` = index into scene.probes of the light the main render uses (`self.novel_probes[dst_env]`, :699-700;
    None = the learned `_light`); `edit_mask [N,>=1]` / `edit_material {'diff','spec','rough'}` = the material edit
    of :293-295, 324-330 (`_update_material`: src * (1 - mask) + mask * update, skipped when update[0] < 0)."""
    dt = dtype
    alpha = _t(batch['alpha'], dt)
    mask = alpha[:, 0] > 0
    rayo, xyz, normal = (_t(batch[k], dt)[mask] for k in ('rayo', 'xyz', 'normal'))
    lvis = _t(batch['lvis'], dt)[mask] if batch.get('lvis') is not None else None
    lxyz, lareas = _t(scene.lxyz, torch.float32).to(dt), _t(scene.lareas, torch.float32).to(dt)
    surf2l = calc_ldir(lxyz, xyz)
    surf2c = calc_vdir(rayo, xyz)
    normal_pred = normal_correct(normal, surf2c)
    z_enc = pred_enc_at(scene.nets, xyz)
    out = {}
    if gen_embed:
        z_norm = safe_l2_normalize(z_enc, axis=1)
        cb = get_codebook(_t(scene.codebook, dt))
        vq = VectorQuantizerEMA(Z_DIM, cb.shape[1], COMMITMENT_COST, dtype=dt)(z_norm, cb, False)
        out['_embed_ind'] = vq['encoding_indices'] + 1
        out['_vq_distances'] = vq['distances']
    rough = pred_head(scene.nets, 'rough_main', z_enc)
    basecolor = pred_head(scene.nets, 'diff_main', z_enc, scene.albedo_slope, scene.albedo_bias)
    ks = pred_head(scene.nets, 'spec_main', z_enc)
    spec = ks * basecolor
    albedo = (1 - ks) * basecolor
    if edit_mask is not None:
        em = (_t(edit_mask, dt)[mask][..., 0:1] > 0).to(dt)                          # :293-295
        upd = lambda src, u: src * (1.0 - em) + em * _t(np.asarray([u], np.float32), dt)   # :258-260
        if not edit_material['diff'][0] < 0:
            albedo = upd(albedo, edit_material['diff'])
        if not edit_material['spec'][0] < 0:
            spec = upd(spec, edit_material['spec'])
        if not edit_material['rough'][0] < 0:
            rough = upd(rough, edit_material['rough'])
    if opt_scale is not None:
        s = _t(opt_scale, dt)
        s_albedo, s_spec = albedo * s, spec * s
    else:
        s_albedo, s_spec = albedo, spec
    brdf, _, _ = get_brdf(surf2l, surf2c, normal_pred, s_albedo, rough, s_spec)
    light = torch.clamp(_t(scene.light, dt), min=0.0)
    if dst_env is not None:
        light = _t(scene.probes[dst_env], dt)                                        # :699-700 (not clipped)
    probes = _t(scene.probes, dt) if (relight_probes and scene.probes is not None) else None
    gamma = None if scene.data_type == 'nerf' else scene.gamma
    rgb_pred, rgb_probes = render(brdf, surf2l, normal_pred, lareas, light, lvis, probes, gamma)
    n = alpha.shape[0]

    def scatter(v):
        full = torch.zeros((n,) + tuple(v.shape[1:]), dtype=v.dtype)
        full[mask] = v
        return full

    out.update({'basecolor': scatter(basecolor), 'albedo': scatter(albedo), 'spec': scatter(spec),
                'rough': scatter(rough), '_rgb_linear': scatter(rgb_pred), '_z_enc': z_enc,
                '_ks': ks})
    rgb_s = linear2srgb(rgb_pred) if scene.data_type == 'nerf' else rgb_pred
    out['rgb'] = scatter(rgb_s)
    if rgb_probes is not None:
        rp = linear2srgb(rgb_probes) if scene.data_type == 'nerf' else rgb_probes
        out['rgb_probes'] = scatter(rp)
    if gen_embed:
        out['embed'] = scatter(out['_embed_ind'][:, None])
    if return_brdf:
        out['_brdf'] = brdf
    return out


def call_forward(scene: Scene, batch: Dict[str, np.ndarray], vq: VectorQuantizerEMA,
                 mode: str = 'train', thres=None, roll=None, dtype=torch.float64):
    """models/vq_nfr.py:534-692 forward (no loss).  Mutates `vq` EMA state and returns the
    codebook update when mode == 'train'."""
    dt = dtype
    alpha = _t(batch['alpha'], dt)
    mask = alpha[:, 0] > 0
    rayo, xyz, normal = (_t(batch[k], dt)[mask] for k in ('rayo', 'xyz', 'normal'))
    lvis = _t(batch['lvis'], dt)[mask] if batch.get('lvis') is not None else None
    lxyz, lareas = _t(scene.lxyz, torch.float32).to(dt), _t(scene.lareas, torch.float32).to(dt)
    surf2l = calc_ldir(lxyz, xyz)
    surf2c = calc_vdir(rayo, xyz)
    normal_pred = normal_correct(normal, surf2c)
    z_enc = pred_enc_at(scene.nets, xyz)
    z_norm = safe_l2_normalize(z_enc, axis=1)
    cb = get_codebook(_t(scene.codebook, dt))
    th = None if thres is None else _t(thres, dt).reshape(1, -1)
    rl = None if roll is None else _t(roll, dt).reshape(1, -1)
    vq_outs = vq(z_norm, cb, is_training=(mode == 'train'), thres=th, roll=rl)
    z_vq = vq_outs['quantize']
    rough = pred_head(scene.nets, 'rough_main', z_enc)
    basecolor = pred_head(scene.nets, 'diff_main', z_enc, scene.albedo_slope, scene.albedo_bias)
    ks = pred_head(scene.nets, 'spec_main', z_enc)
    spec, albedo = ks * basecolor, (1 - ks) * basecolor
    light = clip_preserve_grad(_t(scene.light, dt), 0.0, float('inf'))             # :759
    gamma = None if scene.data_type == 'nerf' else scene.gamma
    brdf, brdf_spec, brdf_diff = get_brdf(surf2l, surf2c, normal_pred, albedo, rough, spec)
    rgb_pred, _ = render(brdf, surf2l, normal_pred, lareas, light, lvis, None, gamma)
    out = {'rgb_linear': rgb_pred, 'albedo': albedo, 'spec': spec, 'rough': rough, 'ks': ks,
           'z_enc': z_enc, 'z_vq': z_vq, 'embed_ind': vq_outs['encoding_indices'] + 1,
           'vq_loss': vq_outs['loss'], 'perplexity': vq_outs['perplexity'],
           'distances': vq_outs['distances'], 'normal': normal_pred}
    if mode != 'train':
        out['rgb_diff'], _ = render(brdf_diff, surf2l, normal_pred, lareas, light, lvis, None, gamma)
        out['rgb_spec'], _ = render(brdf_spec, surf2l, normal_pred, lareas, light, lvis, None, gamma)
    else:
        out['update'] = vq_outs['update']
    vq_rough = pred_head(scene.nets, 'rough_vq', z_vq)
    vq_albedo = pred_head(scene.nets, 'diff_vq', z_vq, scene.albedo_slope, scene.albedo_bias)
    vq_spec = pred_head(scene.nets, 'spec_vq', z_vq)
    vq_brdf, _, _ = get_brdf(surf2l, surf2c, normal_pred, vq_albedo, vq_rough, vq_spec)
    vq_rgb, _ = render(vq_brdf, surf2l, normal_pred, lareas, light, lvis, None, gamma)
    out.update({'vq_rgb_linear': vq_rgb, 'vq_albedo': vq_albedo, 'vq_spec': vq_spec,
                'vq_rough': vq_rough, 'mask': mask})
    return out


# ----------------------------------------------------------------------------
# ref_nfr.Model (models/ref_nfr.py): the residual model test.py renders before the VQ model
# ----------------------------------------------------------------------------
def make_ref_nfr_nets(seed: int = 0, bias_scale: float = 0.0) -> Dict[str, Net]:
    """models/ref_nfr.py:137-160: fine_enc / bottleneck / spec_out come from the VQ-stage model (:141-146),
    rgb_enc = [z, z, z] (none, relu, sigmoid) on the 3-d reference RGB (:148), diff_out / rough_out = [z, z/2, out]
    with skip_at=[1] on the 512-d concat [z_xyz, z_ref] (:149-152)."""
    base = make_vq_nfr_nets(seed, bias_scale)
    rng = np.random.RandomState(seed + 77)
    R, S, N_ = ACT_RELU, ACT_SIGMOID, ACT_NONE
    nets = {'fine_enc': base['fine_enc'], 'bottleneck': base['bottleneck'], 'spec_out': base['spec_main']}
    nets['rgb_enc'] = make_net(rng, 3, [Z_DIM, Z_DIM, Z_DIM], [N_, R, S], bias_scale=bias_scale)
    nets['diff_out'] = make_net(rng, 2 * Z_DIM, [Z_DIM, Z_DIM // 2, 3], [R, R, S], skip_at=1, bias_scale=bias_scale)
    nets['rough_out'] = make_net(rng, 2 * Z_DIM, [Z_DIM, Z_DIM // 2, 1], [R, R, S], skip_at=1, bias_scale=bias_scale)
    return nets


def _ref_materials(scene: Scene, xyz, ref):
    """ref_nfr.py:196-208: z_xyz -> ks; z_bias = concat(z_xyz, rgb_enc(ref)) -> basecolor, rough."""
    nets = scene.nets
    z_xyz = nets['bottleneck'](nets['fine_enc'](embed(xyz)))
    ks = nets['spec_out'](z_xyz)
    z_bias = torch.cat([z_xyz, nets['rgb_enc'](ref)], dim=-1)
    basecolor = scene.albedo_slope * nets['diff_out'](z_bias) + scene.albedo_bias
    rough = nets['rough_out'](z_bias)
    return ks, basecolor, rough, ks * basecolor, (1 - ks) * basecolor


def ref_fast_render(scene: Scene, batch: Dict[str, np.ndarray], dtype=torch.float64, relight_probes: bool = False,
                    opt_scale=None, edit_mask=None, edit_material=None) -> Dict[str, torch.Tensor]:
    """models/ref_nfr.py:306-417: rgb = the RAW BRDF under the model light; rgb_probes = the opt_scale'd BRDF under the
    novel probes; batch['ref'] = the reference RGB."""
    dt = dtype
    alpha = _t(batch['alpha'], dt)
    mask = alpha[:, 0] > 0
    rayo, xyz, normal, ref = (_t(batch[k], dt)[mask] for k in ('rayo', 'xyz', 'normal', 'ref'))
    lvis = _t(batch['lvis'], dt)[mask] if batch.get('lvis') is not None else None
    lxyz, lareas = _t(scene.lxyz, torch.float32).to(dt), _t(scene.lareas, torch.float32).to(dt)
    surf2l, surf2c = calc_ldir(lxyz, xyz), calc_vdir(rayo, xyz)
    normal_pred = normal_correct(normal, surf2c)
    ks, basecolor, rough, spec, albedo = _ref_materials(scene, xyz, ref)
    if edit_mask is not None:
        em = (_t(edit_mask, dt)[mask][..., 0:1] > 0).to(dt)
        upd = lambda src, u: src * (1.0 - em) + em * _t(np.asarray([u], np.float32), dt)
        if not edit_material['diff'][0] < 0:
            albedo = upd(albedo, edit_material['diff'])
        if not edit_material['spec'][0] < 0:
            spec = upd(spec, edit_material['spec'])
        if not edit_material['rough'][0] < 0:
            rough = upd(rough, edit_material['rough'])
    raw_brdf, _, _ = get_brdf(surf2l, surf2c, normal_pred, albedo, rough, spec)
    if opt_scale is not None:
        s = _t(opt_scale, dt)
        albedo, spec = albedo * s, spec * s
    brdf, _, _ = get_brdf(surf2l, surf2c, normal_pred, albedo, rough, spec)
    light = _t(scene.light, dt)                       # ref_nfr.py:88,424: np_light.npy as loaded, not clipped
    gamma = None if scene.data_type == 'nerf' else scene.gamma
    rgb_pred, _ = render(raw_brdf, surf2l, normal_pred, lareas, light, lvis, None, gamma)
    n = alpha.shape[0]

    def scatter(v):
        full = torch.zeros((n,) + tuple(v.shape[1:]), dtype=v.dtype)
        full[mask] = v
        return full
    to_s = linear2srgb if scene.data_type == 'nerf' else (lambda v: v)
    out = {'rgb': scatter(to_s(rgb_pred))}
    if relight_probes and scene.probes is not None:
        _, rgb_probes = render(brdf, surf2l, normal_pred, lareas, light, lvis, _t(scene.probes, dt), gamma)
        out['rgb_probes'] = scatter(to_s(rgb_probes))
    return out


def ref_call(scene: Scene, batch: Dict[str, np.ndarray], mode: str = 'vali', dtype=torch.float64,
             relight_probes: bool = False, opt_scale=None) -> Dict[str, torch.Tensor]:
    """models/ref_nfr.py:176-300 forward: full-length pred dict entries."""
    dt = dtype
    alpha = _t(batch['alpha'], dt)
    mask = alpha[:, 0] > 0
    rayo, xyz, normal, ref = (_t(batch[k], dt)[mask] for k in ('rayo', 'xyz', 'normal', 'ref'))
    lvis = _t(batch['lvis'], dt)[mask] if batch.get('lvis') is not None else None
    lxyz, lareas = _t(scene.lxyz, torch.float32).to(dt), _t(scene.lareas, torch.float32).to(dt)
    surf2l, surf2c = calc_ldir(lxyz, xyz), calc_vdir(rayo, xyz)
    normal_pred = normal_correct(normal, surf2c)
    ks, basecolor, rough, spec, albedo = _ref_materials(scene, xyz, ref)
    if opt_scale is not None and mode == 'test':
        s = _t(opt_scale, dt)
        albedo, spec = albedo * s, spec * s
    brdf, brdf_spec, brdf_diff = get_brdf(surf2l, surf2c, normal_pred, albedo, rough, spec)
    light = _t(scene.light, dt)                       # ref_nfr.py:88,424: np_light.npy as loaded, not clipped
    gamma = None if scene.data_type == 'nerf' else scene.gamma
    probes = _t(scene.probes, dt) if (relight_probes and scene.probes is not None) else None
    rgb_pred, rgb_probes = render(brdf, surf2l, normal_pred, lareas, light, lvis, probes, gamma)
    n = alpha.shape[0]

    def scatter(v):
        full = torch.zeros((n,) + tuple(v.shape[1:]), dtype=v.dtype)
        full[mask] = v
        return full
    to_s = linear2srgb if scene.data_type == 'nerf' else (lambda v: v)
    out = {'rgb': scatter(to_s(rgb_pred)), 'normal': scatter(normal_pred), 'albedo': scatter(albedo),
           'spec': scatter(spec), 'rough': scatter(rough), 'ks': scatter(ks), 'basecolor': scatter(basecolor),
           '_rgb_linear': rgb_pred}
    if mode != 'train':
        out['rgb_diff'] = scatter(render(brdf_diff, surf2l, normal_pred, lareas, light, lvis, None, gamma)[0])
        out['rgb_spec'] = scatter(render(brdf_spec, surf2l, normal_pred, lareas, light, lvis, None, gamma)[0])
    if rgb_probes is not None:
        out['rgb_probes'] = scatter(to_s(rgb_probes))
    return out


def unit_call(scene: Scene, batch: Dict[str, np.ndarray], mode: str = 'vali', dtype=torch.float64) -> Dict[str, torch.Tensor]:
    """models/nfr_unit.py:182-271 forward (the warm-up model: the main branch without a VQ layer; scene.nets carries
    diff_main / spec_main / rough_main as its diff_out / spec_out / rough_out): full-length pred dict entries."""
    dt = dtype
    alpha = _t(batch['alpha'], dt)
    mask = alpha[:, 0] > 0
    rayo, xyz, normal = (_t(batch[k], dt)[mask] for k in ('rayo', 'xyz', 'normal'))
    lvis = _t(batch['lvis'], dt)[mask] if batch.get('lvis') is not None else None
    lxyz, lareas = _t(scene.lxyz, torch.float32).to(dt), _t(scene.lareas, torch.float32).to(dt)
    surf2l, surf2c = calc_ldir(lxyz, xyz), calc_vdir(rayo, xyz)
    normal_pred = normal_correct(normal, surf2c)
    nets = scene.nets
    z_bias = pred_enc_at(nets, xyz)                                                # :208 _pred_bias_at
    basecolor = pred_head(nets, 'diff_main', z_bias, scene.albedo_slope, scene.albedo_bias)
    rough = pred_head(nets, 'rough_main', z_bias)
    ks = pred_head(nets, 'spec_main', z_bias)
    spec, albedo = ks * basecolor, (1 - ks) * basecolor                            # :213-214
    brdf, brdf_spec, brdf_diff = get_brdf(surf2l, surf2c, normal_pred, albedo, rough, spec)
    light = torch.clamp(_t(scene.light, dt), min=0.0)                              # :321-327 the light property clips at 0
    gamma = None if scene.data_type == 'nerf' else scene.gamma
    rgb_pred, _ = render(brdf, surf2l, normal_pred, lareas, light, lvis, None, gamma)
    n = alpha.shape[0]

    def scatter(v):
        full = torch.zeros((n,) + tuple(v.shape[1:]), dtype=v.dtype)
        full[mask] = v
        return full
    to_s = linear2srgb if scene.data_type == 'nerf' else (lambda v: v)
    out = {'rgb': scatter(to_s(rgb_pred)), 'normal': scatter(normal_pred), 'albedo': scatter(albedo),
           'spec': scatter(spec), 'rough': scatter(rough), 'ks': scatter(ks), 'basecolor': scatter(basecolor),
           'xyz': scatter(xyz), 'z_bias': scatter(z_bias), '_rgb_linear': rgb_pred}
    if mode != 'train':
        out['rgb_diff'] = scatter(render(brdf_diff, surf2l, normal_pred, lareas, light, lvis, None, gamma)[0])
        out['rgb_spec'] = scatter(render(brdf_spec, surf2l, normal_pred, lareas, light, lvis, None, gamma)[0])
    return out


# ----------------------------------------------------------------------------
# training step: compute_loss (models/vq_nfr.py:876-986) + train_iter (train_nfr.py:562-576)
# ----------------------------------------------------------------------------
@dataclass
class LossConfig:
    """nerfactor/config/vq_nfr.ini:115-131"""
    vq_loss_weight: float = 1.0
    chr_alpha: float = 60.0
    chr_thres: float = 0.1
    combine_weight: float = 0.2
    mat_sloss_weight: float = 0.05
    chromaticity_loss_weight: float = 1.0
    sim_loss_weight: float = 1e-4
    lambert_weight: float = 1e-3


def rgb2chromaticity(rgb: torch.Tensor) -> torch.Tensor:
    """models/vq_nfr.py:1135-1137"""
    denom = torch.sqrt(torch.sum(rgb * rgb, dim=-1, keepdim=True))
    return divide_no_nan(rgb, denom)


def mse(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """tf.keras.losses.MSE: mean over the last axis -> per-row [n]."""
    return torch.mean((a - b) ** 2, dim=-1)


def sim_loss(codebook_raw: torch.Tensor) -> torch.Tensor:
    """models/vq_nfr.py:958-972: -log(min_{i != j} ||c_i - c_j||) on the normalised codebook.
    DEVIATION (documented in DESIGN.md): tf.sqrt's gradient at the K diagonal zeros is 0 * 0.5/0 = NaN in
    TF (and in torch); the masked product gives those entries zero weight, so the oracle and the product
    define their contribution as 0 (sqrt is only taken off the diagonal)."""
    cb = get_codebook(codebook_raw).t()                                     # [K,Z]
    k = cb.shape[0]
    sq = torch.sum((cb[:, None, :] - cb[None, :, :]) ** 2, dim=-1)
    eye = torch.eye(k, dtype=cb.dtype)
    dist = torch.sqrt(sq + eye) * (1 - eye)                                 # == sqrt(sq) off the diagonal
    max_value = torch.max(dist)
    masked = dist * (1 - eye) + eye * max_value
    return -torch.log(torch.min(masked))


def compute_loss(mode: str, gtc, rgb, vqrgb, vqloss=None, z=None, spec=None, rough=None,
                 codebook_raw=None, cfg: LossConfig = LossConfig(), data_type: str = 'nerf'):
    """models/vq_nfr.py:876-986.  Returns (per-example loss [n], loss_dict).  `codebook_raw` is the
    `_codebook` variable AFTER the EMA assign of the forward pass (get_codebook() is re-evaluated here)."""
    if data_type == 'nerf':
        linear_gt, srgb_pred = srgb2linear(gtc), linear2srgb(rgb)
    else:
        linear_gt, srgb_pred = gtc, rgb
    ld = {}
    if mode != 'train':
        ld['rgb'] = mse(gtc, srgb_pred)
        ld['vqrgb'] = mse(gtc, linear2srgb(vqrgb))
        ld['chromaticity'] = mse(rgb2chromaticity(linear_gt), rgb2chromaticity(vqrgb))
        return ld['rgb'] + ld['vqrgb'] + ld['chromaticity'], ld
    ld['rgb'] = cfg.combine_weight * mse(linear_gt, rgb)
    loss = ld['rgb']
    ld['vqrgb'] = mse(linear_gt, vqrgb)
    loss = loss + ld['vqrgb']
    ld['vqloss'] = cfg.vq_loss_weight * vqloss
    loss = loss + ld['vqloss']
    schr_gt = rgb2chromaticity(gtc)
    if cfg.chromaticity_loss_weight > 0:
        ld['chromaticity'] = cfg.chromaticity_loss_weight * mse(rgb2chromaticity(linear_gt),
                                                                rgb2chromaticity(vqrgb))
        loss = loss + ld['chromaticity']
    if cfg.mat_sloss_weight > 0:
        chr1, chr2 = schr_gt[::2, :], schr_gt[1::2, :]
        chr_e = torch.sqrt(torch.sum((chr1 - chr2) ** 2, dim=-1))
        chr_e = torch.where(chr_e > cfg.chr_thres, chr_e, torch.zeros_like(chr_e))
        w_chr = torch.exp(-cfg.chr_alpha * chr_e)
        mat1, mat2 = z[::2, :], z[1::2, :]
        chr_sl = w_chr * (1.0 - torch.sum(mat1 * mat2, dim=-1))
        chr_sl = torch.stack((chr_sl, chr_sl), dim=-1).reshape(-1)
        ld['chr_smooth'] = cfg.mat_sloss_weight * chr_sl
        loss = loss + ld['chr_smooth']
    if cfg.sim_loss_weight > 0:
        ld['sim_smooth'] = cfg.sim_loss_weight * sim_loss(codebook_raw)
        loss = loss + ld['sim_smooth']
    if cfg.lambert_weight > 0:
        sg_rough = rough.detach()
        sg_rough = torch.where(sg_rough < 0.5, torch.zeros_like(sg_rough), 2 * sg_rough - 1.0)
        ld['lambert'] = cfg.lambert_weight * torch.max(spec, dim=-1).values * sg_rough[:, 0]
        loss = loss + ld['lambert']
    ld['loss'] = loss
    return loss, ld


def train_step(scene: Scene, batch: Dict[str, np.ndarray], vq: VectorQuantizerEMA, thres=None, roll=None,
               cfg: LossConfig = LossConfig(), global_bs: Optional[int] = None, dtype=torch.float64):
    """train_nfr.py:562-576 train_iter, gradients by autograd over the float64 restatement of
    Model.call(mode='train') + compute_loss.  weighted_loss = sum(per_example_loss) / global_bs
    (tf.nn.compute_average_loss).  Returns loss, per-variable gradients, the EMA codebook update and the
    forward outputs.  Mutates `vq` (EMA state) like the reference's forward."""
    dt = dtype
    leaf = lambda a: torch.as_tensor(np.asarray(a), dtype=dt).clone().requires_grad_(True)
    nets_t = {k: Net([leaf(w) for w in n.weights], [leaf(b) for b in n.biases], n.acts, n.skip_at)
              for k, n in scene.nets.items()}
    light_t = leaf(scene.light)
    # non-'nerf' data: the tone parameters are trainable too (vq_nfr.py:736-745: [_gamma_bias, clip_preserve(_gamma_index, 0, 5)])
    g_bias, g_index = leaf(np.asarray(scene.gamma[0], np.float64)), leaf(np.asarray(scene.gamma[1], np.float64))
    gamma_t = (g_bias, clip_preserve_grad(g_index, 0.0, 5.0)) if scene.data_type != 'nerf' else scene.gamma
    sc = Scene(nets=nets_t, light=light_t, codebook=scene.codebook, probes=None,
               albedo_slope=scene.albedo_slope, albedo_bias=scene.albedo_bias, data_type=scene.data_type,
               gamma=gamma_t, lxyz=scene.lxyz, lareas=scene.lareas)
    out = call_forward(sc, batch, vq, 'train', thres, roll, dt)
    cb_after = leaf(out['update'].detach().numpy())                 # _codebook.assign(update) (:582-583)
    mask = out['mask']
    gtc = _t(batch['rgb'], dt)[mask]
    loss, ld = compute_loss('train', gtc, out['rgb_linear'], out['vq_rgb_linear'], out['vq_loss'], out['z_vq'],
                            out['spec'], out['rough'], cb_after, cfg, scene.data_type)
    n_rows = loss.shape[0]
    gbs = float(global_bs if global_bs is not None else n_rows)
    weighted = torch.sum(loss) / gbs
    weighted.backward()
    grads = {k: ([w.grad for w in n.weights], [b.grad for b in n.biases]) for k, n in nets_t.items()}
    return {'loss': weighted.detach(), 'per_example': loss.detach(), 'loss_dict': {k: v.detach() for k, v in ld.items()},
            'grads': grads, 'dlight': light_t.grad, 'dcodebook': cb_after.grad, 'update': out['update'].detach(),
            'dgamma': (g_bias.grad, g_index.grad),
            'out': {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}}


def adam_amsgrad(param, grad, m, v, vhat, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-7):
    """tf.keras.optimizers.Adam(amsgrad=True) dense update (TF 2.4 optimizer_v2/adam.py):
    lr_t = lr sqrt(1-b2^t)/(1-b1^t); m,v EMAs; vhat = max(vhat, v); p -= lr_t m / (sqrt(vhat) + eps).
    `step` is the 1-based iteration count after the increment.  Returns (param, m, v, vhat)."""
    lr_t = lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    m = m + (grad - m) * (1.0 - beta1)
    v = v + (grad * grad - v) * (1.0 - beta2)
    vhat = torch.maximum(vhat, v)
    return param - lr_t * m / (torch.sqrt(vhat) + eps), m, v, vhat


# ----------------------------------------------------------------------------
# synthetic inputs -- oracle-side copy, independent of the product
# ----------------------------------------------------------------------------
def synth_batch(n: int, seed: int = 0, fg_frac: float = 1.0, with_lvis: bool = True):
    rng = np.random.RandomState(seed)
    xyz = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
    rayo = rng.normal(size=(n, 3))
    rayo = (4.0 * rayo / np.linalg.norm(rayo, axis=1, keepdims=True)).astype(np.float32)
    normal = rng.normal(size=(n, 3))
    normal = (normal / np.linalg.norm(normal, axis=1, keepdims=True)).astype(np.float32)
    lvis = rng.uniform(0, 1, size=(n, 512)).astype(np.float32) if with_lvis else None
    alpha = (rng.uniform(0, 1, size=(n, 1)) < fg_frac).astype(np.float32)
    rgb = rng.uniform(0, 1, size=(n, 3)).astype(np.float32)
    return {'xyz': xyz, 'rayo': rayo, 'rayd': -rayo / 4.0, 'normal': normal, 'lvis': lvis,
            'alpha': alpha, 'pred_alpha': alpha.copy(), 'rgb': rgb}


def synth_scene(seed: int = 0, k: int = NUM_EMBED, n_probes: int = 0, bias_scale: float = 0.0,
                data_type: str = 'nerf') -> Scene:
    rng = np.random.RandomState(seed + 1000)
    light = (np.abs(rng.normal(size=(LIGHT_H, 2 * LIGHT_H, 3))) * 0.5).astype(np.float32)
    cb = rng.uniform(0, 1, size=(Z_DIM, k)).astype(np.float32)
    probes = None
    if n_probes > 0:
        probes = (np.abs(rng.normal(size=(n_probes, LIGHT_H, 2 * LIGHT_H, 3))) * 0.5).astype(np.float32)
    return Scene(nets=make_vq_nfr_nets(seed, bias_scale), light=light, codebook=cb, probes=probes,
                 data_type=data_type, gamma=(1.1, 0.9))


# ----------------------------------------------------------------------------
# codebook initialisation: nerfactor/util/torch_kmeans.py:23-94 restated in float64 NumPy
# ----------------------------------------------------------------------------
def kmeans_oracle(x: np.ndarray, num_clusters: int, tol: float = 1e-4, seed: int = 1, max_iter: int = 10000):
    """Lloyd iterations exactly as the reference: initial centres = rows np.random.choice(N, K, replace=False)
    after np.random.seed(seed) (:15-19); assignment = first arg-min of the squared euclidean distance (:58-60);
    centre = member mean (:65-70; an empty cluster keeps its centre -- the reference would produce NaN);
    stop when (sum_k |delta c_k|)^2 < tol (:72-75,91).  Returns (ids, centres, iterations)."""
    x = np.asarray(x, np.float64)
    np.random.seed(seed)
    centers = x[np.random.choice(len(x), num_clusters, replace=False)].copy()
    ids = None
    for it in range(1, max_iter + 1):
        d = (x * x).sum(1, keepdims=True) - 2.0 * x @ centers.T + (centers * centers).sum(1)[None, :]
        ids = d.argmin(1)
        new = centers.copy()
        for k in range(num_clusters):
            sel = x[ids == k]
            if len(sel):
                new[k] = sel.mean(0)
        shift = np.sqrt(((new - centers) ** 2).sum(1)).sum()
        centers = new
        if shift ** 2 < tol:
            break
    return ids, centers, it


# ----------------------------------------------------------------------------
# training-batch assembler: nerfactor/train_nfr.py:380-467 with the product's counter-hash RNG
# ----------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def _rnd_u32(seed: int, stream: int, idx: int) -> int:
    return (_splitmix64((_splitmix64((seed ^ (stream << 56)) & _M64) + idx) & _M64) >> 32) & 0xFFFFFFFF


def outer_sample_rows(alpha_hw: np.ndarray, bs: int, seed: int, alpha_thres: Optional[float] = 0.9) -> np.ndarray:
    """train_nfr.py:401-448 (pure Python, small views only): the [p1, p1_n, p2, p2_n, ...] linear pixel indices.
    tf.random.uniform is replaced by the documented splitmix64 counter hash (TF's streams are not reproducible
    outside TF): stream 1 = the neighbour of interior pixel q, stream 2 = the s-th draw among the valid pairs."""
    h, w = alpha_hw.shape
    jit = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]           # :401-402
    valid, nbs = [], []
    for i in range(1, h - 1):
        for j in range(1, w - 1):
            q = (i - 1) * (w - 2) + (j - 1)
            di, dj = jit[_rnd_u32(seed, 1, q) % 8]
            ok = True
            if alpha_thres is not None:
                ok = alpha_hw[i, j] > alpha_thres and alpha_hw[i + di, j + dj] > alpha_thres   # :428-431
            if ok:
                valid.append(i * w + j)
                nbs.append((i + di) * w + (j + dj))
    rows = np.full((2 * bs,), -1, np.int32)
    if valid:
        for s in range(bs):
            r = _rnd_u32(seed, 2, s) % len(valid)                                              # :438-439
            rows[2 * s], rows[2 * s + 1] = valid[r], nbs[r]
    return rows
