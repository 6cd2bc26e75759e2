"""Torch-tensor wrappers, one per C-ABI entry point (include/vqnerf_b200.h).

Each wrapper validates device/dtype/contiguity, allocates outputs with torch (device memory only)
and launches on torch's current stream.  Nothing here computes on the host or in PyTorch.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

F32 = torch.float32


def _ctx(t: torch.Tensor) -> L.Context:
    return L.Context.get(t.device)


def _f(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != F32:
        t = t.to(F32)
    return t.contiguous()


# ---------------------------------------------------------------------------------------------
def gen_light_xyz(envmap_h: int, envmap_w: int, envmap_radius: float = 1e2) -> Tuple[np.ndarray, np.ndarray]:
    """brdf/renderer.py:184-219 (float64 host arrays)."""
    lib = L.load()
    xyz = np.zeros((envmap_h, envmap_w, 3), np.float64)
    areas = np.zeros((envmap_h, envmap_w), np.float64)
    L.check(lib.vqn_gen_light_xyz(envmap_h, envmap_w, float(envmap_radius),
                                  xyz.ctypes.data_as(C.POINTER(C.c_double)),
                                  areas.ctypes.data_as(C.POINTER(C.c_double))))
    return xyz, areas


class PackedNet:
    """vqn_net handle for one mlp.Network; holds references to the caller's weight tensors."""

    def __init__(self, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], acts: Sequence,
                 skip_at: Optional[int] = None):
        assert len(weights) == len(biases) == len(acts)
        self.weights = [_f(w) for w in weights]
        self.biases = [_f(b) for b in biases]
        self.acts = [L.act_code(a) for a in acts]
        self.skip_at = -1 if skip_at is None else int(skip_at)
        self.ctx = _ctx(self.weights[0])
        self.in_dim = int(self.weights[0].shape[0])
        self.out_dim = int(self.weights[-1].shape[1])
        self.handle = C.c_void_p()
        d = self._desc()
        L.check(self.ctx.lib.vqn_net_create(self.ctx.handle, C.byref(d), C.byref(self.handle),
                                            L.stream_ptr(self.weights[0].device)))

    def _desc(self) -> L.NetDesc:
        d = L.NetDesc()
        n = len(self.weights)
        if n > L.VQN_MAX_LAYERS:
            raise ValueError('at most %d layers' % L.VQN_MAX_LAYERS)
        d.n_layers, d.in_dim, d.skip_at = n, self.in_dim, self.skip_at
        prev = self.in_dim
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            exp_in = prev + (self.in_dim if (self.skip_at >= 0 and i == self.skip_at + 1) else 0)
            if w.dim() != 2 or w.shape[0] != exp_in or b.shape != (w.shape[1],):
                raise ValueError('layer %d: kernel %s / bias %s do not chain (expected in=%d)'
                                 % (i, tuple(w.shape), tuple(b.shape), exp_in))
            d.widths[i], d.acts[i] = int(w.shape[1]), self.acts[i]
            d.w[i], d.b[i] = w.data_ptr(), b.data_ptr()
            prev = int(w.shape[1])
        return d

    def repack(self):
        d = self._desc()
        L.check(self.ctx.lib.vqn_net_repack(self.handle, C.byref(d), L.stream_ptr(self.weights[0].device)))

    def repack_tc(self, precision='tf32x3'):
        """Refresh only the tensor-core weight images of `precision` (after an optimizer step)."""
        L.check(self.ctx.lib.vqn_net_repack_tc(self.handle, L.precision_code(precision),
                                               L.stream_ptr(self.weights[0].device)))

    def backward_train(self, dz_last: torch.Tensor, lddz_last: int, n: int, ys: Sequence[torch.Tensor],
                       ldys: Sequence[int], dzs: Sequence[torch.Tensor], lddzs: Sequence[int],
                       d_input: Optional[torch.Tensor] = None, ld_din: int = 0, din_mode: int = 0,
                       din_y: Optional[torch.Tensor] = None, ld_din_y: int = 0, din_act: int = 0) -> None:
        """vqn_net_backward_train: the backward-data chain of the whole network in one launch (tf32x3).  din_y: the
        network's input is the activated output of another layer -- d_input becomes that layer's dz (times act'(din_y))."""
        nl = len(self.weights)
        yp = (C.c_void_p * nl)(*[y.data_ptr() for y in ys])
        lp = (C.c_int64 * nl)(*[int(v) for v in ldys])
        zp = (C.c_void_p * nl)(*[z.data_ptr() for z in dzs])
        zl = (C.c_int64 * nl)(*[int(v) for v in lddzs])
        L.check(self.ctx.lib.vqn_net_backward_train(self.ctx.handle, self.handle, dz_last.data_ptr(), int(lddz_last), int(n),
                                                    yp, lp, zp, zl, None if d_input is None else d_input.data_ptr(),
                                                    int(ld_din), int(din_mode),
                                                    None if din_y is None else din_y.data_ptr(), int(ld_din_y), int(din_act),
                                                    L.stream_ptr(dz_last.device)))

    def forward_train(self, x: torch.Tensor, ldx: int, n: int, ys: Sequence[torch.Tensor], lds: Sequence[int],
                      out_scale: float = 1.0, out_bias: float = 0.0, precision='tf32x3') -> None:
        """vqn_net_forward_train: the whole network in one launch, every layer's output stored into ys[i] (ld lds[i])."""
        nl = len(self.weights)
        if len(ys) != nl or len(lds) != nl:
            raise ValueError('one activation buffer per layer')
        yp = (C.c_void_p * nl)(*[y.data_ptr() for y in ys])
        lp = (C.c_int64 * nl)(*[int(v) for v in lds])
        L.check(self.ctx.lib.vqn_net_forward_train(self.ctx.handle, self.handle, L.ptr(x, F32), int(ldx), int(n), yp, lp,
                                                   float(out_scale), float(out_bias), L.precision_code(precision),
                                                   L.stream_ptr(x.device)))

    def forward(self, x: torch.Tensor, precision='fp32') -> torch.Tensor:
        x = _f(x)
        if x.dim() != 2 or x.shape[1] != self.in_dim:
            raise ValueError('expected input [n,%d], got %s' % (self.in_dim, tuple(x.shape)))
        y = torch.empty((x.shape[0], self.out_dim), dtype=F32, device=x.device)
        L.check(self.ctx.lib.vqn_net_forward(self.handle, L.ptr(x), x.shape[0], L.ptr(y),
                                             L.precision_code(precision), L.stream_ptr(x.device)))
        return y

    def __del__(self):
        try:
            if getattr(self, 'handle', None) and self.handle.value:
                self.ctx.lib.vqn_net_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


def embed(x: torch.Tensor, n_freqs: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Embedder.__call__ (embedder.py:23-47).  out: an existing [n, ld >= 3 + 6 n_freqs] buffer to fill in place (columns
    beyond the embedding are left untouched)."""
    x = _f(x)
    c = _ctx(x)
    if out is None:
        out = torch.empty((x.shape[0], 3 + 6 * n_freqs), dtype=F32, device=x.device)
    elif out.dtype != F32 or out.dim() != 2 or out.shape[0] != x.shape[0] or not out.is_contiguous():
        raise ValueError('embed: out must be a contiguous float32 [n, ld] tensor')
    L.check(c.lib.vqn_embed_ld(c.handle, L.ptr(x), x.shape[0], n_freqs, L.ptr(out), out.shape[1], L.stream_ptr(x.device)))
    return out


def pred_enc_at(fine_enc: PackedNet, bottleneck: PackedNet, n_freqs: int, pts: torch.Tensor,
                row_idx: Optional[torch.Tensor] = None, n: Optional[int] = None, precision='fp32',
                out: Optional[torch.Tensor] = None, n_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    pts = _f(pts)
    n = pts.shape[0] if n is None else n
    z = out if out is not None else torch.empty((n, bottleneck.out_dim), dtype=F32, device=pts.device)
    c = _ctx(pts)
    L.check(c.lib.vqn_pred_enc_at(c.handle, fine_enc.handle, bottleneck.handle, n_freqs, L.ptr(pts),
                                  L.ptr(row_idx, torch.int32), L.ptr(n_dev, torch.int32), n, L.ptr(z),
                                  L.precision_code(precision),
                                  L.stream_ptr(pts.device)))
    return z


def pred_heads(diff: Optional[PackedNet], spec: Optional[PackedNet], rough: Optional[PackedNet],
               z: torch.Tensor, slope: float = 1.0, bias: float = 0.0, precision='fp32',
               n_dev: Optional[torch.Tensor] = None):
    z = _f(z)
    n = z.shape[0]
    outs = [torch.empty((n, h.out_dim), dtype=F32, device=z.device) if h is not None else None
            for h in (diff, spec, rough)]
    c = _ctx(z)
    L.check(c.lib.vqn_pred_heads(c.handle, *(h.handle if h is not None else None for h in (diff, spec, rough)),
                                 L.ptr(z), L.ptr(n_dev, torch.int32), n, float(slope), float(bias),
                                 *(L.ptr(o) for o in outs),
                                 L.precision_code(precision), L.stream_ptr(z.device)))
    return tuple(outs)


def mlp_main(fine_enc: PackedNet, bottleneck: PackedNet, diff: PackedNet, spec: PackedNet, rough: PackedNet,
             n_freqs: int, pts: torch.Tensor, row_idx=None, n_dev=None, n: Optional[int] = None, slope: float = 1.0,
             bias: float = 0.0, want_z: bool = False, precision='fp32'):
    """Fused encoder + three main heads (one launch).  Returns (z or None, basecolor, ks, rough), all compact."""
    pts = _f(pts)
    n = pts.shape[0] if n is None else n
    dev = pts.device
    # the latent stays on chip (fp32: shared memory; tensor-core modes: a per-CTA L2-resident scratch tile) unless the
    # caller wants it; network shapes the fused tensor-core program does not cover fall back to two launches staged in z
    outs = [torch.empty((n, h.out_dim), dtype=F32, device=dev) for h in (diff, spec, rough)]
    c = _ctx(pts)

    def launch(z):
        L.check(c.lib.vqn_mlp_main(c.handle, fine_enc.handle, bottleneck.handle, diff.handle, spec.handle, rough.handle,
                                   n_freqs, L.ptr(pts), L.ptr(row_idx, torch.int32), L.ptr(n_dev, torch.int32), n,
                                   float(slope), float(bias), L.ptr(z), *(L.ptr(o) for o in outs),
                                   L.precision_code(precision), L.stream_ptr(dev)))

    z = torch.empty((n, bottleneck.out_dim), dtype=F32, device=dev) if want_z else None
    try:
        launch(z)
    except NotImplementedError as e:
        if z is not None or 'z_out buffer is required' not in str(e):
            raise
        z = torch.empty((n, bottleneck.out_dim), dtype=F32, device=dev)
        launch(z)
        z = None
    return (z,) + tuple(outs)


def get_codebook(raw: torch.Tensor) -> torch.Tensor:
    raw = _f(raw)
    out = torch.empty_like(raw)
    c = _ctx(raw)
    L.check(c.lib.vqn_get_codebook(c.handle, L.ptr(raw), raw.shape[0], raw.shape[1], L.ptr(out),
                                   L.stream_ptr(raw.device)))
    return out


def l2_normalize_rows(x: torch.Tensor) -> torch.Tensor:
    x = _f(x)
    out = torch.empty_like(x)
    c = _ctx(x)
    L.check(c.lib.vqn_l2_normalize_rows(c.handle, L.ptr(x), x.shape[0], x.shape[1], L.ptr(out),
                                        L.stream_ptr(x.device)))
    return out


def vq_stats_size(z_dim: int, k: int) -> int:
    return int(L.load().vqn_vq_stats_size(int(z_dim), int(k)))


def vq_assign(inputs: torch.Tensor, codebook: torch.Tensor, sel_mask: Optional[torch.Tensor] = None,
              normalize_inputs: bool = False, want_quantize: bool = True, want_distances: bool = False,
              want_znorm: bool = False, stats: Optional[torch.Tensor] = None, want_dw: bool = False):
    """Returns dict(indices int64 [n], quantize, distances, z_norm) -- absent outputs are None."""
    inputs, codebook = _f(inputs), _f(codebook)
    n, zd = inputs.shape
    k = codebook.shape[1]
    if codebook.shape[0] != zd:
        raise ValueError('codebook must be [%d,K]' % zd)
    dev = inputs.device
    idx = torch.empty((n,), dtype=torch.int64, device=dev)
    quant = torch.empty_like(inputs) if want_quantize else None
    dist = torch.empty((n, k), dtype=F32, device=dev) if want_distances else None
    zn = torch.empty_like(inputs) if want_znorm else None
    if sel_mask is not None:
        sel_mask = _f(sel_mask).reshape(-1)
        if sel_mask.numel() != k:
            raise ValueError('sel_mask must have K entries')
    if stats is not None and (stats.dtype != torch.float64 or stats.numel() != vq_stats_size(zd, k)):
        raise ValueError('stats must be float64 [K+2+Z*K]')
    c = _ctx(inputs)
    L.check(c.lib.vqn_vq_assign(c.handle, L.ptr(inputs), n, zd, L.ptr(codebook), k, L.ptr(sel_mask),
                                int(normalize_inputs), L.ptr(idx), L.ptr(quant), L.ptr(dist), L.ptr(zn),
                                L.ptr(stats), int(want_dw), L.stream_ptr(dev)))
    return {'indices': idx, 'quantize': quant, 'distances': dist, 'z_norm': zn}


def vq_ema_update(stats: torch.Tensor, codebook: torch.Tensor, decay: float, epsilon: float,
                  commitment_cost: float, is_training: bool, state: Optional[dict]):
    """state: dict(cs_hidden, cs_average, dw_hidden, dw_average, counters) caller-owned tensors."""
    codebook = _f(codebook)
    zd, k = codebook.shape
    dev = codebook.device
    loss = torch.empty((1,), dtype=F32, device=dev)
    perp = torch.empty((1,), dtype=F32, device=dev)
    update = torch.empty_like(codebook) if is_training else None
    c = _ctx(codebook)
    st = state or {}
    L.check(c.lib.vqn_vq_ema_update(
        c.handle, L.ptr(stats, torch.float64), zd, k, L.ptr(codebook), float(decay), float(epsilon),
        float(commitment_cost), int(is_training), L.ptr(st.get('cs_hidden')), L.ptr(st.get('cs_average')),
        L.ptr(st.get('dw_hidden')), L.ptr(st.get('dw_average')), L.ptr(st.get('counters'), torch.int64),
        L.ptr(update), L.ptr(loss), L.ptr(perp), L.stream_ptr(dev)))
    return update, loss, perp


LVIS_FORMATS = {torch.float32: 0, torch.float16: 1, torch.uint8: 2}


def compress_lvis(lvis: torch.Tensor, fmt: str) -> torch.Tensor:
    """Compact light-visibility formats of the host-buffer path (explicit opt-in; the reference stores float32):
    'f16' = IEEE half (relative error 2^-11 per light: ~1e-5 on a shaded radiance), 'u8' = round(255 v) (absolute error
    <= 1/510 per light: ~5e-5 typical, up to ~4e-4 on a radiance dominated by a few lights).  Works on host or device
    tensors; `Model.fast_render` / `fast_render_host` accept the result in place of the float32 tensor."""
    if fmt in ('f32', None):
        return lvis.to(F32)
    if fmt == 'f16':
        return lvis.to(torch.float16)
    if fmt == 'u8':
        return (lvis.to(F32).clamp(0, 1) * 255.0 + 0.5).to(torch.uint8)
    raise ValueError("lvis format must be 'f32', 'f16' or 'u8'")


def shade(xyz, rayo, normal, lvis, albedo, spec, rough, lxyz, lareas, lights, *, row_idx=None, n_dev=None,
          n: Optional[int] = None, n_total: Optional[int] = None, to_srgb=False, gamma=None, clip_light0=True,
          want_split=False, want_normal=False, out_rgb: Optional[torch.Tensor] = None,
          peer_ptrs: Optional[Sequence[int]] = None, peer_row0: int = 0, no_clip: bool = False):
    """Fused _calc_ldir/_calc_vdir/_normal_correct/_eval_brdf_at/_render.  lights [1+P,512,3].
    no_clip: the raw, un-clipped integral (training of non-'nerf' data applies gamma afterwards: gamma_forward).
    Returns dict(rgb [n_total,1+P,3], rgb_diff, rgb_spec, normal) (full-length when row_idx is given)."""
    xyz, rayo, normal = _f(xyz), _f(rayo), _f(normal)
    albedo, spec, rough = _f(albedo), _f(spec), _f(rough)
    lights = _f(lights)
    if lights.dim() == 2:
        lights = lights[None]
    npb = lights.shape[0]
    dev = xyz.device
    n = albedo.shape[0] if n is None else n
    n_total = xyz.shape[0] if n_total is None else n_total
    alloc = torch.zeros if row_idx is not None else torch.empty
    rgb = out_rgb if out_rgb is not None else alloc((n_total, npb, 3), dtype=F32, device=dev)
    a = L.ShadeArgs()
    a.xyz, a.rayo, a.normal = xyz.data_ptr(), rayo.data_ptr(), normal.data_ptr()
    if lvis is not None:
        # float32 = the reference's lvis.npy; float16 / uint8 (v = q / 255) = the compact opt-in formats (LVIS_FORMATS)
        if lvis.dtype not in LVIS_FORMATS:
            lvis = lvis.to(F32)
        if not lvis.is_cuda or not lvis.is_contiguous():
            raise ValueError('lvis must be a contiguous CUDA tensor')
        a.lvis = lvis.data_ptr()
        a.lvis_format = LVIS_FORMATS[lvis.dtype]
    a.albedo, a.spec, a.rough = albedo.data_ptr(), spec.data_ptr(), rough.data_ptr()
    if row_idx is not None:
        a.row_idx = L.ptr(row_idx, torch.int32).value
    if n_dev is not None:
        a.n_dev = L.ptr(n_dev, torch.int32).value
    a.n = n
    lxyz, lareas = _f(lxyz).reshape(-1, 3), _f(lareas).reshape(-1)
    a.lxyz, a.lareas, a.lights = lxyz.data_ptr(), lareas.data_ptr(), lights.data_ptr()
    a.n_probes, a.to_srgb, a.clip_light0 = npb, int(to_srgb), int(clip_light0)
    if gamma is not None:
        a.use_gamma, a.gamma_bias, a.gamma_index = 1, float(gamma[0]), float(gamma[1])
    a.rgb = rgb.data_ptr()
    a.no_clip = int(bool(no_clip))
    if peer_ptrs:
        if len(peer_ptrs) > 8:
            raise ValueError('at most 8 peers (one NVSwitch box)')
        for q, pp in enumerate(peer_ptrs):
            a.peer_rgb[q] = int(pp)
        a.n_peers, a.peer_row0 = len(peer_ptrs), int(peer_row0)
    out = {'rgb': rgb, 'rgb_diff': None, 'rgb_spec': None, 'normal': None}
    if want_split:
        out['rgb_diff'] = alloc((n_total, 3), dtype=F32, device=dev)
        out['rgb_spec'] = alloc((n_total, 3), dtype=F32, device=dev)
        a.rgb_diff, a.rgb_spec = out['rgb_diff'].data_ptr(), out['rgb_spec'].data_ptr()
    if want_normal:
        out['normal'] = alloc((n_total, 3), dtype=F32, device=dev)
        a.normal_out = out['normal'].data_ptr()
    c = _ctx(xyz)
    L.check(c.lib.vqn_shade(c.handle, C.byref(a), L.stream_ptr(dev)))
    return out


def eval_brdf(pts2l, pts2c, normal, albedo, spec, rough):
    pts2l, pts2c, normal, albedo, spec, rough = map(_f, (pts2l, pts2c, normal, albedo, spec, rough))
    n = pts2c.shape[0]
    outs = [torch.empty((n, 512, 3), dtype=F32, device=pts2c.device) for _ in range(3)]
    c = _ctx(pts2c)
    L.check(c.lib.vqn_eval_brdf(c.handle, L.ptr(pts2l), L.ptr(pts2c), L.ptr(normal), L.ptr(albedo), L.ptr(spec),
                                L.ptr(rough), n, *(L.ptr(o) for o in outs), L.stream_ptr(pts2c.device)))
    return tuple(outs)


def render(brdf, l, normal, lvis, lareas, light, gamma=None):
    brdf, l, normal, lareas, light = map(_f, (brdf, l, normal, lareas, light))
    lvis = _f(lvis) if lvis is not None else None
    n = normal.shape[0]
    rgb = torch.empty((n, 3), dtype=F32, device=normal.device)
    c = _ctx(normal)
    g = gamma or (1.0, 1.0)
    L.check(c.lib.vqn_render(c.handle, L.ptr(brdf), L.ptr(l), L.ptr(normal), L.ptr(lvis), L.ptr(lareas.reshape(-1)),
                             L.ptr(light.reshape(-1, 3)), n, int(gamma is not None), float(g[0]), float(g[1]),
                             L.ptr(rgb), L.stream_ptr(normal.device)))
    return rgb


def material_combine(basecolor, ks, opt_scale=None, n_dev=None, want_scaled=True):
    basecolor, ks = _f(basecolor), _f(ks)
    n = basecolor.shape[0]
    albedo, spec = torch.empty_like(basecolor), torch.empty_like(basecolor)
    a_s = torch.empty_like(basecolor) if (want_scaled and opt_scale is not None) else None
    s_s = torch.empty_like(basecolor) if (want_scaled and opt_scale is not None) else None
    if opt_scale is not None:
        opt_scale = _f(opt_scale).reshape(-1)
    c = _ctx(basecolor)
    L.check(c.lib.vqn_material_combine(c.handle, L.ptr(basecolor), L.ptr(ks), L.ptr(opt_scale),
                                       L.ptr(n_dev, torch.int32), n, L.ptr(albedo), L.ptr(spec), L.ptr(a_s),
                                       L.ptr(s_s), L.stream_ptr(basecolor.device)))
    return albedo, spec, (a_s if a_s is not None else albedo), (s_s if s_s is not None else spec)


def peer_clear_background(alpha, peer_ptrs, peer_row0, width):
    """Zeros for the background rows of this rank's shard in the peers' image buffers (fused gather, dist.PeerImage)."""
    import ctypes as C
    alpha = _f(alpha)
    if alpha.dim() == 1:
        alpha = alpha[:, None]
    arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
    c = _ctx(alpha)
    L.check(c.lib.vqn_peer_clear_background(c.handle, L.ptr(alpha), alpha.shape[1], alpha.shape[0], int(peer_row0),
                                            int(width), C.cast(arr, C.c_void_p), len(peer_ptrs),
                                            L.stream_ptr(alpha.device)))


def peer_frame_sync(counter, local_flags, peer_flag_ptrs, world, rank, dst):
    """Frame hand-shake of the fused gather on one destination rank (vqn_peer_frame_sync; dist.PeerImage.barrier)."""
    import ctypes as C
    arr = (C.c_void_p * len(peer_flag_ptrs))(*[int(p) for p in peer_flag_ptrs])
    c = _ctx(counter)
    L.check(c.lib.vqn_peer_frame_sync(c.handle, counter.data_ptr(), local_flags.data_ptr(), C.cast(arr, C.c_void_p),
                                      int(world), int(rank), int(dst), L.stream_ptr(counter.device)))


def material_edit(edit_mask, edit_material, row_idx, n_dev, albedo, spec, rough, opt_scale=None, albedo_s=None,
                  spec_s=None):
    """fast_render's `_update_material` (models/vq_nfr.py:258-260, 324-330), in place on the compact tensors.  A channel
    list whose first entry is negative is left alone, as in the reference."""
    import ctypes as C
    edit_mask = _f(edit_mask)
    if edit_mask.dim() == 1:
        edit_mask = edit_mask[:, None]
    stride = edit_mask.shape[1]

    def host(vals, k):
        vals = list(vals)
        if vals[0] < 0:
            return None
        if len(vals) != k:
            raise ValueError('edit_material: expected %d values, got %d' % (k, len(vals)))
        return (C.c_float * k)(*[float(v) for v in vals])
    d3, s3, r1 = host(edit_material['diff'], 3), host(edit_material['spec'], 3), host(edit_material['rough'], 1)
    cptr = lambda a: C.cast(a, C.c_void_p) if a is not None else None
    if opt_scale is not None:
        opt_scale = _f(opt_scale).reshape(-1)
    c = _ctx(albedo)
    L.check(c.lib.vqn_material_edit(c.handle, L.ptr(edit_mask), stride, L.ptr(row_idx, torch.int32),
                                    L.ptr(n_dev, torch.int32), albedo.shape[0], cptr(d3), cptr(s3), cptr(r1),
                                    L.ptr(opt_scale), L.ptr(albedo), L.ptr(spec), L.ptr(rough),
                                    L.ptr(albedo_s if albedo_s is not albedo else None),
                                    L.ptr(spec_s if spec_s is not spec else None), L.stream_ptr(albedo.device)))


def linear2srgb(x):
    x = _f(x)
    out = torch.empty_like(x)
    c = _ctx(x)
    L.check(c.lib.vqn_linear2srgb(c.handle, L.ptr(x), x.numel(), L.ptr(out), L.stream_ptr(x.device)))
    return out


def srgb2linear(x):
    x = _f(x)
    out = torch.empty_like(x)
    c = _ctx(x)
    L.check(c.lib.vqn_srgb2linear(c.handle, L.ptr(x), x.numel(), L.ptr(out), L.stream_ptr(x.device)))
    return out


def compact_mask(alpha: torch.Tensor):
    """ind = where(alpha[:,0] > 0): returns (row_idx int32 [n_total], n_active int32 [1]) on device."""
    a = _f(alpha).reshape(-1)
    n = a.shape[0]
    row_idx = torch.empty((max(n, 1),), dtype=torch.int32, device=a.device)
    n_active = torch.zeros((1,), dtype=torch.int32, device=a.device)
    work = torch.empty((n // 1024 + 2,), dtype=torch.int32, device=a.device)   # per-call: safe across streams
    c = _ctx(a)
    L.check(c.lib.vqn_compact_mask(c.handle, L.ptr(a), n, L.ptr(row_idx), L.ptr(n_active), L.ptr(work),
                                   L.stream_ptr(a.device)))
    return row_idx, n_active


def scatter_rows(compact: torch.Tensor, row_idx: torch.Tensor, n_total: int, n_dev=None,
                 n: Optional[int] = None) -> torch.Tensor:
    compact = _f(compact)
    c_ = int(np.prod(compact.shape[1:])) if compact.dim() > 1 else 1
    n = compact.shape[0] if n is None else n
    out = torch.zeros((n_total,) + tuple(compact.shape[1:]), dtype=F32, device=compact.device)
    c = _ctx(compact)
    L.check(c.lib.vqn_scatter_rows(c.handle, L.ptr(compact), L.ptr(row_idx, torch.int32), L.ptr(n_dev, torch.int32),
                                   n, c_, L.ptr(out), L.stream_ptr(compact.device)))
    return out


def scatter_rows_multi(compacts, row_idx: torch.Tensor, n_total: int, n_dev=None, n: Optional[int] = None):
    """scatter_rows for up to 4 compact [n, c_q] tensors that share the row list: ONE zero fill + ONE launch."""
    compacts = [_f(c) for c in compacts]
    widths = [int(np.prod(c.shape[1:])) if c.dim() > 1 else 1 for c in compacts]
    n = compacts[0].shape[0] if n is None else n
    dev = compacts[0].device
    flat = torch.zeros((n_total * sum(widths),), dtype=F32, device=dev)
    outs, off = [], 0
    for c, w in zip(compacts, widths):
        outs.append(flat[off:off + n_total * w].view((n_total,) + tuple(c.shape[1:])))
        off += n_total * w
    k = len(compacts)
    src = (C.c_void_p * k)(*[c.data_ptr() for c in compacts])
    dst = (C.c_void_p * k)(*[o.data_ptr() for o in outs])
    wid = (C.c_int32 * k)(*widths)
    ctx = _ctx(compacts[0])
    L.check(ctx.lib.vqn_scatter_rows_multi(ctx.handle, src, wid, k, L.ptr(row_idx, torch.int32), L.ptr(n_dev, torch.int32),
                                           n, dst, L.stream_ptr(dev)))
    return outs


# ---- NeuS -------------------------------------------------------------------------------------
def neus_up_sample(rays_o, rays_d, z_vals, sdf, r_limit, n_importance, inv_s):
    rays_o, rays_d, z_vals, sdf = map(_f, (rays_o, rays_d, z_vals, sdf))
    b, s = z_vals.shape
    sdf = sdf.reshape(b, s)
    out = torch.empty((b, n_importance), dtype=F32, device=z_vals.device)
    c = _ctx(z_vals)
    L.check(c.lib.vqn_neus_up_sample(c.handle, L.ptr(rays_o), L.ptr(rays_d), L.ptr(z_vals), L.ptr(sdf), b, s,
                                     float(r_limit), int(n_importance), float(inv_s), L.ptr(out),
                                     L.stream_ptr(z_vals.device)))
    return out


def neus_up_sample_pts(rays_o, rays_d, z_vals, sdf, r_limit, n_importance, inv_s):
    """up_sample + the positions o + d z of the new samples (the next SDF call's input): (new_z [B,I], pts [B,I,3])."""
    rays_o, rays_d, z_vals, sdf = map(_f, (rays_o, rays_d, z_vals, sdf))
    b, s = z_vals.shape
    sdf = sdf.reshape(b, s)
    out = torch.empty((b, n_importance), dtype=F32, device=z_vals.device)
    pts = torch.empty((b, n_importance, 3), dtype=F32, device=z_vals.device)
    c = _ctx(z_vals)
    L.check(c.lib.vqn_neus_up_sample_pts(c.handle, L.ptr(rays_o), L.ptr(rays_d), L.ptr(z_vals), L.ptr(sdf), b, s,
                                         float(r_limit), int(n_importance), float(inv_s), L.ptr(out), L.ptr(pts),
                                         L.stream_ptr(z_vals.device)))
    return out, pts


def neus_scan_step(rays_o, rays_d, z_vals, new_z, sdf, new_sdf, r_limit, n_importance, inv_s, final_merge=False,
                   sample_dist=0.0, want_merged=True, want_dirs=True):
    """One hierarchical-sampling step in one launch: cat_z_vals(z_vals, new_z, sdf, new_sdf) -> up_sample of the next
    step on the merged row [-> the last cat_z_vals and render_core's mid points when final_merge].  Returns a dict with
    'z', 'sdf' (merged, when want_merged), 'new_z', 'pts' (not final) or 'z_final', 'mid_pts', 'mid_dirs' (final)."""
    rays_o, rays_d, z_vals, new_z = map(_f, (rays_o, rays_d, z_vals, new_z))
    b, s = z_vals.shape
    i = new_z.shape[1]
    dev = z_vals.device
    sdf, new_sdf = _f(sdf).reshape(b, s), _f(new_sdf).reshape(b, i)
    a = L.NeusStepArgs()
    a.rays_o, a.rays_d, a.z_vals, a.new_z = rays_o.data_ptr(), rays_d.data_ptr(), z_vals.data_ptr(), new_z.data_ptr()
    a.sdf, a.new_sdf = sdf.data_ptr(), new_sdf.data_ptr()
    a.n_rays, a.n_samples, a.n_new, a.n_importance, a.final_merge = b, s, i, int(n_importance), int(bool(final_merge))
    a.r_limit, a.inv_s, a.sample_dist = float(r_limit), float(inv_s), float(sample_dist)
    out = {}
    e = lambda *shape: torch.empty(shape, dtype=F32, device=dev)
    if want_merged:
        out['z'], out['sdf'] = e(b, s + i), e(b, s + i)
        a.z_out, a.sdf_out = out['z'].data_ptr(), out['sdf'].data_ptr()
    if final_merge:
        t = s + i + int(n_importance)
        out['z_final'], out['mid_pts'] = e(b, t), e(b, t, 3)
        a.z_final, a.mid_pts = out['z_final'].data_ptr(), out['mid_pts'].data_ptr()
        if want_dirs:
            out['mid_dirs'] = e(b, t, 3)
            a.mid_dirs = out['mid_dirs'].data_ptr()
    else:
        out['new_z'], out['pts'] = e(b, int(n_importance)), e(b, int(n_importance), 3)
        a.new_z_out, a.pts_out = out['new_z'].data_ptr(), out['pts'].data_ptr()
    c = _ctx(z_vals)
    L.check(c.lib.vqn_neus_scan_step(c.handle, C.byref(a), L.stream_ptr(dev)))
    return out


def neus_cat_z_vals(z_vals, new_z, sdf=None, new_sdf=None):
    z_vals, new_z = _f(z_vals), _f(new_z)
    b, s = z_vals.shape
    i = new_z.shape[1]
    sdf = _f(sdf).reshape(b, s) if sdf is not None else None
    new_sdf = _f(new_sdf).reshape(b, i) if new_sdf is not None else None
    z_out = torch.empty((b, s + i), dtype=F32, device=z_vals.device)
    sdf_out = torch.empty((b, s + i), dtype=F32, device=z_vals.device) if (sdf is not None and new_sdf is not None) \
        else None
    c = _ctx(z_vals)
    L.check(c.lib.vqn_neus_cat_z_vals(c.handle, L.ptr(z_vals), L.ptr(new_z), L.ptr(sdf), L.ptr(new_sdf), b, s, i,
                                      L.ptr(z_out), L.ptr(sdf_out), L.stream_ptr(z_vals.device)))
    return z_out, sdf_out


def neus_mid_points(rays_o, rays_d, z_vals, sample_dist):
    rays_o, rays_d, z_vals = map(_f, (rays_o, rays_d, z_vals))
    b, s = z_vals.shape
    pts = torch.empty((b, s, 3), dtype=F32, device=z_vals.device)
    dirs = torch.empty((b, s, 3), dtype=F32, device=z_vals.device)
    c = _ctx(z_vals)
    L.check(c.lib.vqn_neus_mid_points(c.handle, L.ptr(rays_o), L.ptr(rays_d), L.ptr(z_vals), b, s,
                                      float(sample_dist), L.ptr(pts), L.ptr(dirs), L.stream_ptr(z_vals.device)))
    return pts, dirs


def neus_composite(rays_o, rays_d, z_vals, sdf, gradients, sampled_color, inv_s, cos_anneal_ratio, sample_dist,
                   radius, background_rgb=None, want_grad_err=True):
    rays_o, rays_d, z_vals = map(_f, (rays_o, rays_d, z_vals))
    b, s = z_vals.shape
    dev = z_vals.device
    sdf = _f(sdf).reshape(b, s)
    gradients = _f(gradients).reshape(b, s, 3)
    sampled_color = _f(sampled_color).reshape(b, s, 3)
    a = L.NeusCompositeArgs()
    a.rays_o, a.rays_d, a.z_vals = rays_o.data_ptr(), rays_d.data_ptr(), z_vals.data_ptr()
    a.sdf, a.gradients, a.sampled_color = sdf.data_ptr(), gradients.data_ptr(), sampled_color.data_ptr()
    a.n_rays, a.n_samples = b, s
    a.inv_s, a.cos_anneal_ratio, a.sample_dist, a.radius = float(inv_s), float(cos_anneal_ratio), \
        float(sample_dist), float(radius)
    bg = None
    if background_rgb is not None:
        bg = _f(background_rgb).reshape(-1)
        a.background_rgb = bg.data_ptr()
    o = {k: torch.empty(shape, dtype=F32, device=dev) for k, shape in (
        ('color', (b, 3)), ('weights', (b, s)), ('surf', (b, 3)), ('depth', (b, 1)), ('cdf', (b, s)),
        ('inside_sphere', (b, s)), ('mid_z_vals', (b, s)), ('dists', (b, s)), ('weight_sum', (b, 1)),
        ('weight_max', (b, 1)))}
    for k, v in o.items():
        setattr(a, k, v.data_ptr())
    ge = torch.zeros((2,), dtype=torch.float64, device=dev) if want_grad_err else None
    if ge is not None:
        a.grad_err_sums = ge.data_ptr()
    c = _ctx(z_vals)
    L.check(c.lib.vqn_neus_composite(c.handle, C.byref(a), L.stream_ptr(dev)))
    o['grad_err_sums'] = ge
    return o


def sdf_forward(trunk: PackedNet, w_sdf: torch.Tensor, b_sdf: torch.Tensor, feat: Optional[PackedNet], n_freqs: int,
                pts: torch.Tensor, want_grad: bool = False, feat_out: Optional[torch.Tensor] = None,
                precision='tf32x3', grad_mode: str = 'reverse'):
    """vqn_sdf_forward (fields.py:74-112).  feat_out: None (sdf only) or a [n, stride >= d_feature] row buffer that
    receives the feature vector in its first columns.  grad_mode: 'reverse' (act' stash + transposed layers) or 'jet'
    (forward-mode tangent rows).  Returns (sdf [n,1], grad [n,3] or None)."""
    if grad_mode not in ('reverse', 'jet'):
        raise ValueError("grad_mode must be 'reverse' or 'jet'")
    pts = _f(pts)
    n = pts.shape[0]
    if pts.dim() != 2 or pts.shape[1] != 3:
        raise ValueError('expected points [n,3], got %s' % (tuple(pts.shape),))
    sdf = torch.empty((n, 1), dtype=F32, device=pts.device)
    grad = torch.empty((n, 3), dtype=F32, device=pts.device) if want_grad else None
    if n == 0:
        L.precision_code(precision)
        return sdf, grad
    stride = 0
    if feat_out is not None:
        if feat is None or feat_out.dim() != 2 or feat_out.shape[0] != n or feat_out.dtype != F32 \
                or feat_out.stride(1) != 1 or feat_out.stride(0) < feat.out_dim:
            raise ValueError('feat_out must be a float32 [n, >= d_feature] row buffer')
        stride = feat_out.stride(0)
    c = _ctx(pts)
    L.check(c.lib.vqn_sdf_forward(c.handle, trunk.handle, L.ptr(w_sdf, F32), L.ptr(b_sdf, F32),
                                  feat.handle if feat_out is not None else None, n_freqs, L.ptr(pts), n, L.ptr(sdf),
                                  C.c_void_p(feat_out.data_ptr()) if feat_out is not None else None, stride,
                                  L.ptr(grad), 1 if grad_mode == 'reverse' else 0, L.precision_code(precision),
                                  L.stream_ptr(pts.device)))
    return sdf, grad


def neus_color_input(pts, dirs, normals, multires_view: int, rows: torch.Tensor, col_off: int, width: int):
    """vqn_neus_color_input: columns [col_off, col_off + width) of rows <- [pts, embed(dirs), normals, 0...]."""
    pts, dirs, normals = map(_f, (pts, dirs, normals))
    n = pts.shape[0]
    if rows.dtype != F32 or rows.dim() != 2 or rows.shape[0] != n or rows.stride(1) != 1:
        raise ValueError('rows must be a float32 [n, stride] buffer')
    c = _ctx(pts)
    L.check(c.lib.vqn_neus_color_input(c.handle, L.ptr(pts), L.ptr(dirs), L.ptr(normals), n, int(multires_view),
                                       C.c_void_p(rows.data_ptr()), rows.stride(0), int(col_off), int(width),
                                       L.stream_ptr(pts.device)))


def neus_light_rays(surf, normal, lxyz, l0: int, n_chunk: int, radius: float):
    """vqn_neus_light_rays: rays from every surface point to lights [l0, l0 + n_chunk).  Returns rays_o [M,3],
    rays_d [M,3], near [M,1], far [M,1], front [M,1] (1.0 where front-lit), M = n_pts * n_chunk."""
    surf, normal, lxyz = map(_f, (surf, normal, lxyz))
    lxyz = lxyz.reshape(-1, 3)
    n = surf.shape[0]
    if l0 < 0 or n_chunk < 1 or l0 + n_chunk > lxyz.shape[0]:
        raise ValueError('light chunk [%d, %d) outside the %d lights' % (l0, l0 + n_chunk, lxyz.shape[0]))
    m = n * n_chunk
    dev = surf.device
    rays_o = torch.empty((m, 3), dtype=F32, device=dev)
    rays_d = torch.empty((m, 3), dtype=F32, device=dev)
    near, far, front = (torch.empty((m, 1), dtype=F32, device=dev) for _ in range(3))
    if m:
        c = _ctx(surf)
        L.check(c.lib.vqn_neus_light_rays(c.handle, L.ptr(surf), L.ptr(normal), L.ptr(lxyz), n, int(l0), int(n_chunk),
                                          float(radius), L.ptr(rays_o), L.ptr(rays_d), L.ptr(near), L.ptr(far),
                                          L.ptr(front), L.stream_ptr(dev)))
    return rays_o, rays_d, near, far, front


def neus_lvis_scatter(weight_sum, row_idx, n: int, l0: int, n_chunk: int, lvis: torch.Tensor):
    """lvis[p, l0 + j] = 1 - weight_sum[i] for the compacted pair row_idx[i] = p * n_chunk + j (first n rows)."""
    weight_sum = _f(weight_sum)
    if lvis.dtype != F32 or not lvis.is_contiguous() or lvis.dim() != 2:
        raise ValueError('lvis must be a contiguous float32 [n_pts, n_lights] tensor')
    c = _ctx(lvis)
    L.check(c.lib.vqn_neus_lvis_scatter(c.handle, L.ptr(weight_sum), L.ptr(row_idx, torch.int32), None, int(n), int(l0),
                                        int(n_chunk), int(lvis.shape[1]), L.ptr(lvis), L.stream_ptr(lvis.device)))


# ---- tensor-core primitive self-test ------------------------------------------------------------
def tc_selftest(a: torch.Tensor, b: torch.Tensor, mode: int) -> torch.Tensor:
    """d[128,n] = a[128,k] @ b[n,k]^T on tcgen05 (mode 0 tf32, 1 bf16, 2 3xTF32, 3 shipped tf32+bf16 scheme with A in TMEM)."""
    a, b = _f(a), _f(b)
    if a.shape[0] != 128 or a.shape[1] != b.shape[1]:
        raise ValueError('a must be [128,k], b [n,k]')
    d = torch.empty((128, b.shape[0]), dtype=F32, device=a.device)
    c = _ctx(a)
    L.check(c.lib.vqn_tc_selftest(c.handle, int(mode), b.shape[0], a.shape[1], L.ptr(a), L.ptr(b), L.ptr(d),
                                  L.stream_ptr(a.device)))
    return d


# ---------------------------------------------------------------------------------------------
# training step (include/vqnerf_b200.h "training step"): raw-pointer wrappers, no allocation --
# the caller (nerfactor/train_nfr.py) owns every buffer so that a step can be replayed from a CUDA graph
# ---------------------------------------------------------------------------------------------
def _p(t, off_elems: int = 0):
    """device pointer of tensor `t` advanced by `off_elems` elements (views into concat / flat buffers)"""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr() + off_elems * t.element_size())


def dense_forward(x, ldx, w, b, y, ldy, m, k, n, act, out_scale=1.0, out_bias=0.0, x_off=0, y_off=0):
    c = _ctx(w)
    L.check(c.lib.vqn_dense_forward(c.handle, _p(x, x_off), ldx, _p(w), _p(b), _p(y, y_off), ldy, m, k, n, act,
                                    float(out_scale), float(out_bias), L.stream_ptr(w.device)))


def dense_backward_data(dz, lddz, w, dx, lddx, yprev, ldyp, act_prev, accumulate, m, k, n, w_row0=0, dx_off=0,
                        yprev_off=0):
    """dx[:, dx_off:dx_off+k] (+)= (dz . w[w_row0:w_row0+k, :]^T) * act_prev'(yprev[:, yprev_off:...])"""
    c = _ctx(w)
    L.check(c.lib.vqn_dense_backward_data(c.handle, _p(dz), lddz, _p(w, w_row0 * n), _p(dx, dx_off), lddx,
                                          _p(yprev, yprev_off), ldyp, act_prev, int(accumulate), m, k, n,
                                          L.stream_ptr(w.device)))


def dense_backward_weights(x, ldx, dz, lddz, dw, db, m, k, n, x_off=0):
    c = _ctx(dw)
    L.check(c.lib.vqn_dense_backward_weights(c.handle, _p(x, x_off), ldx, _p(dz), lddz, _p(dw), _p(db), m, k, n,
                                             L.stream_ptr(dw.device)))


def _ptr_int(t, off_elems=0):
    return None if t is None else t.data_ptr() + 4 * int(off_elems)


def bwd_data_problem(dz, lddz, w, dx, lddx, yprev, ldyp, act_prev, accumulate, m, k, n, w_row0=0, dx_off=0):
    """One problem of dense_backward_data_batched (same meaning as dense_backward_data; accumulate 2 = atomic add)."""
    q = L.DenseProblem()
    q.a, q.lda, q.w, q.ldw = _ptr_int(dz), lddz, _ptr_int(w, w_row0 * n), n
    q.out, q.ldo, q.yprev, q.ldy = _ptr_int(dx, dx_off), lddx, _ptr_int(yprev), ldyp
    q.m, q.k, q.n, q.act_prev, q.accumulate = m, k, n, act_prev, int(accumulate)
    return q


def bwd_weights_problem(x, ldx, dz, lddz, dw, db, m, k, n, dw_off=0):
    """dW[dw_off:][k, n] += x[m, k]^T dz[m, n]; db[n] += column sums of dz (db may be None)."""
    q = L.DenseProblem()
    q.a, q.lda, q.w, q.ldw = _ptr_int(x), ldx, _ptr_int(dz), lddz
    q.out, q.ldo, q.colsum = _ptr_int(dw, dw_off), n, _ptr_int(db)
    q.m, q.k, q.n = m, k, n
    return q


def _dense_batched(fn_name, problems, dev):
    c = L.Context.get(dev)
    fn = getattr(c.lib, fn_name)
    for i in range(0, len(problems), 32):
        chunk = problems[i:i + 32]
        arr = (L.DenseProblem * len(chunk))(*chunk)
        L.check(fn(c.handle, arr, len(chunk), L.stream_ptr(dev)))


def dense_backward_data_batched(problems, dev):
    """Independent backward-data GEMMs (bwd_data_problem) in ONE launch per 32 problems."""
    _dense_batched('vqn_dense_backward_data_batched', problems, dev)


def dense_backward_weights_batched(problems, dev):
    """Every weight-gradient GEMM of a step (bwd_weights_problem) in ONE launch per 32 problems."""
    _dense_batched('vqn_dense_backward_weights_batched', problems, dev)


def act_backward(dy, lddy, y, ldy, m, n, act, scale, out_scale, out_bias, dz, lddz):
    c = _ctx(dy)
    L.check(c.lib.vqn_act_backward(c.handle, _p(dy), lddy, _p(y), ldy, m, n, act, float(scale), float(out_scale),
                                   float(out_bias), _p(dz), lddz, L.stream_ptr(dy.device)))


def act_backward_batched(jobs, dev):
    """jobs: list of the argument tuples of act_backward (<= 8) -- ONE launch."""
    if not jobs:
        return
    c = L.Context.get(dev)
    arr = (L.ActJob * len(jobs))()
    for q, (dy, lddy, y, ldy, m, n, act, scale, out_scale, out_bias, dz, lddz) in zip(arr, jobs):
        q.dy, q.y, q.dz = _ptr_int(dy), (_ptr_int(y) if y is not None else None), _ptr_int(dz)
        q.lddy, q.ldy, q.lddz, q.m, q.n, q.act = lddy, ldy, lddz, m, n, act
        q.scale, q.out_scale, q.out_bias = float(scale), float(out_scale), float(out_bias)
    L.check(c.lib.vqn_act_backward_batched(c.handle, arr, len(jobs), L.stream_ptr(dev)))


def train_scalars(sums, vq_loss, sim_loss, inv_gbs, vq_w, sim_w, out):
    """out[4] = [weighted loss, vq_w*vq_loss, sim_w*sim_loss, rows/global_bs] on the device (train_nfr.py:571)."""
    c = _ctx(sums)
    L.check(c.lib.vqn_train_scalars(c.handle, _p(sums), _p(vq_loss), _p(sim_loss) if sim_loss is not None else None,
                                    float(inv_gbs), float(vq_w), float(sim_w), _p(out), L.stream_ptr(sums.device)))
    return out


def copy_cols(src, lds, dst, ldd, m, w, dst_off=0):
    c = _ctx(src)
    L.check(c.lib.vqn_copy_cols(c.handle, _p(src), lds, _p(dst, dst_off), ldd, m, w, L.stream_ptr(src.device)))


def gamma_forward(lin, gpar, out=None):
    """out = clip((lin * gpar[0]) ^ clip(gpar[1], 0, 5), 0, 1) with gpar = [_gamma_bias, _gamma_index] on the device."""
    lin = _f(lin)
    out = torch.empty_like(lin) if out is None else out
    c = _ctx(lin)
    L.check(c.lib.vqn_gamma_forward(c.handle, _p(lin), _p(gpar), _p(out), lin.numel(), L.stream_ptr(lin.device)))
    return out


def gamma_backward(lin, gpar, d_out, d_lin, d_gpar):
    """d_lin = d_out * d out / d lin;  d_gpar[0..1] += gradients of the two tone parameters (see gamma_forward)."""
    c = _ctx(lin)
    L.check(c.lib.vqn_gamma_backward(c.handle, _p(lin), _p(gpar), _p(d_out), _p(d_lin), _p(d_gpar), lin.numel(),
                                     L.stream_ptr(lin.device)))


def copy_cols_batched(jobs, dev):
    """jobs: list of (src, lds, dst, ldd, m, w, dst_off) -- the arguments of copy_cols -- in ONE launch per 16 jobs."""
    c = L.Context.get(dev)
    for i in range(0, len(jobs), 16):
        chunk = jobs[i:i + 16]
        arr = (L.CopyJob * len(chunk))()
        for q, (src, lds, dst, ldd, m, w, dst_off) in zip(arr, chunk):
            q.src, q.dst, q.lds, q.ldd, q.m, q.w = _ptr_int(src), _ptr_int(dst, dst_off), lds, ldd, m, w
        L.check(c.lib.vqn_copy_cols_batched(c.handle, arr, len(chunk), L.stream_ptr(dev)))


def nets_repack_tc(packed_nets, precision='tf32x3'):
    """Refresh the pre-split tensor-core weight images of several PackedNets in ONE launch (after an optimizer step)."""
    if not packed_nets:
        return
    arr = (C.c_void_p * len(packed_nets))(*[p.handle.value for p in packed_nets])
    c = packed_nets[0].ctx
    L.check(c.lib.vqn_nets_repack_tc(arr, len(packed_nets), L.precision_code(precision),
                                     L.stream_ptr(packed_nets[0].weights[0].device)))


def shade_backward(xyz, rayo, normal, lvis, albedo, spec, rough, lxyz, lareas, light, d_rgb, d_albedo, d_spec,
                   d_rough, d_light, clip_light0=True, row_idx=None):
    c = _ctx(xyz)
    n = albedo.shape[0]
    L.check(c.lib.vqn_shade_backward(c.handle, _p(xyz), _p(rayo), _p(normal), _p(lvis), _p(row_idx), n, _p(albedo),
                                     _p(spec), _p(rough), _p(lxyz), _p(lareas), _p(light), int(clip_light0),
                                     _p(d_rgb), _p(d_albedo), _p(d_spec), _p(d_rough), _p(d_light),
                                     L.stream_ptr(xyz.device)))


def loss_train(gtc, rgb, vqrgb, z_vq, spec, rough, data_is_nerf, combine_weight, chromaticity_weight,
               mat_sloss_weight, lambert_weight, chr_alpha, chr_thres, inv_global_bs, loss_rows, d_rgb, d_vqrgb,
               d_z, d_spec, sums):
    c = _ctx(gtc)
    n, zd = z_vq.shape
    L.check(c.lib.vqn_loss_train(c.handle, _p(gtc), _p(rgb), _p(vqrgb), _p(z_vq), _p(spec), _p(rough), n, zd,
                                 int(data_is_nerf), float(combine_weight), float(chromaticity_weight),
                                 float(mat_sloss_weight), float(lambert_weight), float(chr_alpha), float(chr_thres),
                                 float(inv_global_bs), _p(loss_rows), _p(d_rgb), _p(d_vqrgb), _p(d_z), _p(d_spec),
                                 _p(sums), L.stream_ptr(gtc.device)))


def vq_backward(z_enc, indices, codebook, d_zvq, commit_coef, d_zenc, accumulate=True, act=0, dz_out=None):
    """dz_out (optional, [n, ld]): also d_zenc * act'(z_enc) -- the dz of the layer whose activated output z_enc is."""
    c = _ctx(z_enc)
    L.check(c.lib.vqn_vq_backward_act(c.handle, _p(z_enc), _p(indices), _p(codebook), codebook.shape[1], _p(d_zvq),
                                      float(commit_coef), z_enc.shape[0], int(accumulate), _p(d_zenc), int(act),
                                      None if dz_out is None else _p(dz_out), 0 if dz_out is None else dz_out.shape[1],
                                      L.stream_ptr(z_enc.device)))


def zero_batched(tensors, dev):
    """Clear up to 8 contiguous device tensors in ONE launch."""
    c = L.Context.get(dev)
    for i in range(0, len(tensors), 8):
        chunk = tensors[i:i + 8]
        ptrs = (C.c_void_p * len(chunk))(*[t.data_ptr() for t in chunk])
        sizes = (C.c_int64 * len(chunk))(*[t.numel() * t.element_size() for t in chunk])
        L.check(c.lib.vqn_zero_batched(c.handle, ptrs, sizes, len(chunk), L.stream_ptr(dev)))


def train_pack_stats(stats64, stats32, rows_slot, rows):
    """stats32 = float(stats64); rows_slot[0] += rows (this rank's active-row count, part of the all-reduce buffer)."""
    c = _ctx(stats64)
    L.check(c.lib.vqn_train_pack_stats(c.handle, _p(stats64), _p(stats32), stats64.numel(), _p(rows_slot), float(rows),
                                       L.stream_ptr(stats64.device)))


def material_combine_backward(basecolor, ks, d_albedo, d_spec, d_spec_extra, d_basecolor, d_ks):
    c = _ctx(basecolor)
    L.check(c.lib.vqn_material_combine_backward(c.handle, _p(basecolor), _p(ks), _p(d_albedo), _p(d_spec),
                                                _p(d_spec_extra), basecolor.shape[0], _p(d_basecolor), _p(d_ks),
                                                L.stream_ptr(basecolor.device)))


def codebook_sim_loss(raw_codebook, grad_scale, loss_out, d_raw, accumulate=False):
    c = _ctx(raw_codebook)
    zd, k = raw_codebook.shape
    L.check(c.lib.vqn_codebook_sim_loss(c.handle, _p(raw_codebook), zd, k, float(grad_scale), _p(loss_out),
                                        _p(d_raw), int(accumulate), L.stream_ptr(raw_codebook.device)))


def adam_amsgrad(param, grad, m, v, vhat, lr_t, beta1, beta2, epsilon, lr_t_dev=None):
    c = _ctx(param)
    L.check(c.lib.vqn_adam_amsgrad(c.handle, _p(param), _p(grad), _p(m), _p(v), _p(vhat), param.numel(),
                                   float(lr_t), _p(lr_t_dev), float(beta1), float(beta2), float(epsilon),
                                   L.stream_ptr(param.device)))


def cast_f64_f32(src, dst):
    c = _ctx(src)
    L.check(c.lib.vqn_cast_f64_f32(c.handle, _p(src), _p(dst), src.numel(), L.stream_ptr(src.device)))


def cast_f32_f64(src, dst):
    c = _ctx(src)
    L.check(c.lib.vqn_cast_f32_f64(c.handle, _p(src), _p(dst), src.numel(), L.stream_ptr(src.device)))


# ---------------------------------------------------------------------------------------------
# training-batch assembler (outer_sample, nerfactor/train_nfr.py:380-467)
# ---------------------------------------------------------------------------------------------
def sample_pairs(alpha: torch.Tensor, h: int, w: int, bs: int, seed: int, alpha_thres: Optional[float] = 0.9):
    """rows int32 [2*bs] = [p1, p1_n, p2, p2_n, ...] and n_valid int32 [1] (device)."""
    a = _f(alpha).reshape(-1)
    if a.numel() != h * w:
        raise ValueError('alpha must hold one H x W view')
    dev = a.device
    total = (h - 2) * (w - 2)
    flag = torch.empty((total,), dtype=F32, device=dev)
    nb = torch.empty((total,), dtype=torch.int32, device=dev)
    valid = torch.empty((total,), dtype=torch.int32, device=dev)
    work = torch.empty((total // 1024 + 2,), dtype=torch.int32, device=dev)
    n_valid = torch.zeros((1,), dtype=torch.int32, device=dev)
    rows = torch.empty((2 * bs,), dtype=torch.int32, device=dev)
    c = _ctx(a)
    L.check(c.lib.vqn_sample_pairs(c.handle, L.ptr(a), h, w, int(alpha_thres is not None),
                                   float(alpha_thres if alpha_thres is not None else 0.0), bs, int(seed) & (2 ** 64 - 1),
                                   L.ptr(flag), L.ptr(nb), L.ptr(valid), L.ptr(work), L.ptr(n_valid), L.ptr(rows),
                                   L.stream_ptr(dev)))
    return rows, n_valid


def gather_rows(src: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
    src = _f(src)
    flat = src.reshape(src.shape[0], -1)
    out = torch.empty((rows.shape[0], flat.shape[1]), dtype=F32, device=src.device)
    c = _ctx(src)
    L.check(c.lib.vqn_gather_rows(c.handle, L.ptr(flat), L.ptr(rows, torch.int32), rows.shape[0], flat.shape[1],
                                  L.ptr(out), L.stream_ptr(src.device)))
    return out.reshape((rows.shape[0],) + tuple(src.shape[1:]))
