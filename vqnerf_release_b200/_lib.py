"""ctypes binding of libvqnerf_b200.so (the C ABI declared in include/vqnerf_b200.h).

There is NO CPU fallback: importing this module without the built library, or creating a
context without an sm_100 GPU, raises.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libvqnerf_b200.so')

VQN_MAX_LAYERS = 8
PREC_FP32, PREC_BF16, PREC_TF32X3 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SOFTPLUS100 = 0, 1, 2, 3
_ACT_BY_NAME = {None: ACT_NONE, 'none': ACT_NONE, 'linear': ACT_NONE, 'relu': ACT_RELU, 'sigmoid': ACT_SIGMOID,
                'softplus100': ACT_SOFTPLUS100}
_PREC_BY_NAME = {'fp32': PREC_FP32, 'bf16': PREC_BF16, 'tf32x3': PREC_TF32X3}

(OK, ERR_INVALID_ARG, ERR_CUDA, ERR_NONFINITE, ERR_UNSUPPORTED, ERR_ZERO_NORM) = range(6)


class NonFiniteError(FloatingPointError):
    """tf.errors.InvalidArgumentError from tf.debugging.check_numerics in the reference."""


class NetDesc(C.Structure):
    _fields_ = [('n_layers', C.c_int32), ('in_dim', C.c_int32), ('skip_at', C.c_int32), ('reserved', C.c_int32),
                ('widths', C.c_int32 * VQN_MAX_LAYERS), ('acts', C.c_int32 * VQN_MAX_LAYERS),
                ('w', C.c_void_p * VQN_MAX_LAYERS), ('b', C.c_void_p * VQN_MAX_LAYERS)]


class ShadeArgs(C.Structure):
    _fields_ = [('xyz', C.c_void_p), ('rayo', C.c_void_p), ('normal', C.c_void_p), ('lvis', C.c_void_p),
                ('albedo', C.c_void_p), ('spec', C.c_void_p), ('rough', C.c_void_p), ('row_idx', C.c_void_p),
                ('n_dev', C.c_void_p), ('n', C.c_int64),
                ('lxyz', C.c_void_p), ('lareas', C.c_void_p), ('lights', C.c_void_p),
                ('n_probes', C.c_int32), ('clip_light0', C.c_int32), ('to_srgb', C.c_int32), ('use_gamma', C.c_int32),
                ('gamma_bias', C.c_float), ('gamma_index', C.c_float),
                ('rgb', C.c_void_p), ('rgb_diff', C.c_void_p), ('rgb_spec', C.c_void_p), ('normal_out', C.c_void_p),
                ('peer_rgb', C.c_void_p * 8), ('n_peers', C.c_int32), ('lvis_format', C.c_int32),
                ('peer_row0', C.c_int64), ('no_clip', C.c_int32), ('reserved0', C.c_int32)]


class NeusCompositeArgs(C.Structure):
    _fields_ = [('rays_o', C.c_void_p), ('rays_d', C.c_void_p), ('z_vals', C.c_void_p), ('sdf', C.c_void_p),
                ('gradients', C.c_void_p), ('sampled_color', C.c_void_p),
                ('n_rays', C.c_int64), ('n_samples', C.c_int32), ('reserved', C.c_int32),
                ('inv_s', C.c_float), ('cos_anneal_ratio', C.c_float), ('sample_dist', C.c_float),
                ('radius', C.c_float), ('background_rgb', C.c_void_p),
                ('color', C.c_void_p), ('weights', C.c_void_p), ('surf', C.c_void_p), ('depth', C.c_void_p),
                ('cdf', C.c_void_p), ('inside_sphere', C.c_void_p), ('mid_z_vals', C.c_void_p),
                ('dists', C.c_void_p), ('weight_sum', C.c_void_p), ('weight_max', C.c_void_p),
                ('grad_err_sums', C.c_void_p)]


class NeusStepArgs(C.Structure):
    _fields_ = [('rays_o', C.c_void_p), ('rays_d', C.c_void_p), ('z_vals', C.c_void_p), ('new_z', C.c_void_p),
                ('sdf', C.c_void_p), ('new_sdf', C.c_void_p),
                ('n_rays', C.c_int64), ('n_samples', C.c_int32), ('n_new', C.c_int32), ('n_importance', C.c_int32),
                ('final_merge', C.c_int32),
                ('r_limit', C.c_float), ('inv_s', C.c_float), ('sample_dist', C.c_float), ('reserved', C.c_float),
                ('z_out', C.c_void_p), ('sdf_out', C.c_void_p), ('new_z_out', C.c_void_p), ('pts_out', C.c_void_p),
                ('z_final', C.c_void_p), ('mid_pts', C.c_void_p), ('mid_dirs', C.c_void_p)]


class CopyJob(C.Structure):
    _fields_ = [('src', C.c_void_p), ('dst', C.c_void_p), ('lds', C.c_int64), ('ldd', C.c_int64), ('m', C.c_int64),
                ('w', C.c_int32), ('reserved', C.c_int32)]


class ActJob(C.Structure):
    _fields_ = [('dy', C.c_void_p), ('y', C.c_void_p), ('dz', C.c_void_p), ('lddy', C.c_int64), ('ldy', C.c_int64),
                ('lddz', C.c_int64), ('m', C.c_int64), ('n', C.c_int32), ('act', C.c_int32), ('scale', C.c_float),
                ('out_scale', C.c_float), ('out_bias', C.c_float), ('reserved', C.c_float)]


class DenseProblem(C.Structure):
    _fields_ = [('a', C.c_void_p), ('lda', C.c_int64), ('w', C.c_void_p), ('ldw', C.c_int64),
                ('out', C.c_void_p), ('ldo', C.c_int64), ('yprev', C.c_void_p), ('ldy', C.c_int64),
                ('colsum', C.c_void_p), ('m', C.c_int64), ('k', C.c_int32), ('n', C.c_int32),
                ('act_prev', C.c_int32), ('accumulate', C.c_int32)]


_P, _I, _L, _F, _D = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double

# name -> (restype, argtypes); every symbol declared in include/vqnerf_b200.h
SIGNATURES = {
    'vqn_abi_version': (_I, []),
    'vqn_status_str': (C.c_char_p, [_I]),
    'vqn_last_error': (C.c_char_p, []),
    'vqn_ctx_create': (_I, [_I, C.POINTER(_P)]),
    'vqn_ctx_destroy': (_I, [_P]),
    'vqn_ctx_launch_count': (_L, [_P]),
    'vqn_ctx_check_numerics': (_I, [_P, _P]),
    'vqn_vq_stats_size': (_L, [_I, _I]),
    'vqn_compact_workspace_size': (_L, [_L]),
    'vqn_sample_pairs_workspace_size': (_L, [_I, _I]),
    'vqn_gen_light_xyz': (_I, [_I, _I, _D, C.POINTER(_D), C.POINTER(_D)]),
    'vqn_net_create': (_I, [_P, C.POINTER(NetDesc), C.POINTER(_P), _P]),
    'vqn_net_repack': (_I, [_P, C.POINTER(NetDesc), _P]),
    'vqn_net_destroy': (_I, [_P]),
    'vqn_net_out_dim': (_I, [_P]),
    'vqn_net_forward': (_I, [_P, _P, _L, _P, _I, _P]),
    'vqn_embed': (_I, [_P, _P, _L, _I, _P, _P]),
    'vqn_embed_ld': (_I, [_P, _P, _L, _I, _P, _L, _P]),
    'vqn_pred_enc_at': (_I, [_P, _P, _P, _I, _P, _P, _P, _L, _P, _I, _P]),
    'vqn_pred_heads': (_I, [_P, _P, _P, _P, _P, _P, _L, _F, _F, _P, _P, _P, _I, _P]),
    'vqn_mlp_main': (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _L, _F, _F, _P, _P, _P, _P, _I, _P]),
    'vqn_get_codebook': (_I, [_P, _P, _I, _I, _P, _P]),
    'vqn_l2_normalize_rows': (_I, [_P, _P, _L, _I, _P, _P]),
    'vqn_vq_assign': (_I, [_P, _P, _L, _I, _P, _I, _P, _I, _P, _P, _P, _P, _P, _I, _P]),
    'vqn_vq_ema_update': (_I, [_P, _P, _I, _I, _P, _F, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'vqn_shade': (_I, [_P, C.POINTER(ShadeArgs), _P]),
    'vqn_eval_brdf': (_I, [_P, _P, _P, _P, _P, _P, _P, _L, _P, _P, _P, _P]),
    'vqn_render': (_I, [_P, _P, _P, _P, _P, _P, _P, _L, _I, _F, _F, _P, _P]),
    'vqn_material_combine': (_I, [_P, _P, _P, _P, _P, _L, _P, _P, _P, _P, _P]),
    'vqn_peer_clear_background': (_I, [_P, _P, _I, _L, _L, _I, _P, _I, _P]),
    'vqn_peer_frame_sync': (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    'vqn_material_edit': (_I, [_P, _P, _I, _P, _P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'vqn_linear2srgb': (_I, [_P, _P, _L, _P, _P]),
    'vqn_srgb2linear': (_I, [_P, _P, _L, _P, _P]),
    'vqn_compact_mask': (_I, [_P, _P, _L, _P, _P, _P, _P]),
    'vqn_scatter_rows': (_I, [_P, _P, _P, _P, _L, _I, _P, _P]),
    'vqn_scatter_rows_multi': (_I, [_P, _P, _P, _I, _P, _P, _L, _P, _P]),
    'vqn_neus_up_sample': (_I, [_P, _P, _P, _P, _P, _L, _I, _F, _I, _F, _P, _P]),
    'vqn_neus_up_sample_pts': (_I, [_P, _P, _P, _P, _P, _L, _I, _F, _I, _F, _P, _P, _P]),
    'vqn_neus_scan_step': (_I, [_P, C.POINTER(NeusStepArgs), _P]),
    'vqn_neus_cat_z_vals': (_I, [_P, _P, _P, _P, _P, _L, _I, _I, _P, _P, _P]),
    'vqn_neus_composite': (_I, [_P, C.POINTER(NeusCompositeArgs), _P]),
    'vqn_neus_mid_points': (_I, [_P, _P, _P, _P, _L, _I, _F, _P, _P, _P]),
    'vqn_sdf_forward': (_I, [_P, _P, _P, _P, _P, _I, _P, _L, _P, _P, _L, _P, _I, _I, _P]),
    'vqn_neus_color_input': (_I, [_P, _P, _P, _P, _L, _I, _P, _L, _I, _I, _P]),
    'vqn_neus_light_rays': (_I, [_P, _P, _P, _P, _L, _I, _I, _F, _P, _P, _P, _P, _P, _P]),
    'vqn_neus_lvis_scatter': (_I, [_P, _P, _P, _P, _L, _I, _I, _I, _P, _P]),
    'vqn_dense_forward': (_I, [_P, _P, _L, _P, _P, _P, _L, _L, _I, _I, _I, _F, _F, _P]),
    'vqn_net_forward_train': (_I, [_P, _P, _P, _L, _L, C.POINTER(_P), C.POINTER(_L), _F, _F, _I, _P]),
    'vqn_net_repack_tc': (_I, [_P, _I, _P]),
    'vqn_net_backward_train': (_I, [_P, _P, _P, _L, _L, C.POINTER(_P), C.POINTER(_L), C.POINTER(_P), C.POINTER(_L), _P, _L,
                                    _I, _P, _L, _I, _P]),
    'vqn_nets_repack_tc': (_I, [C.POINTER(_P), _I, _I, _P]),
    'vqn_copy_cols_batched': (_I, [_P, C.POINTER(CopyJob), _I, _P]),
    'vqn_dense_backward_data': (_I, [_P, _P, _L, _P, _P, _L, _P, _L, _I, _I, _L, _I, _I, _P]),
    'vqn_dense_backward_weights': (_I, [_P, _P, _L, _P, _L, _P, _P, _L, _I, _I, _P]),
    'vqn_gamma_forward': (_I, [_P, _P, _P, _P, _L, _P]),
    'vqn_gamma_backward': (_I, [_P, _P, _P, _P, _P, _P, _L, _P]),
    'vqn_dense_backward_data_batched': (_I, [_P, C.POINTER(DenseProblem), _I, _P]),
    'vqn_dense_backward_weights_batched': (_I, [_P, C.POINTER(DenseProblem), _I, _P]),
    'vqn_act_backward': (_I, [_P, _P, _L, _P, _L, _L, _I, _I, _F, _F, _F, _P, _L, _P]),
    'vqn_copy_cols': (_I, [_P, _P, _L, _P, _L, _L, _I, _P]),
    'vqn_shade_backward': (_I, [_P] * 6 + [_L] + [_P] * 6 + [_I] + [_P] * 6),
    'vqn_loss_train': (_I, [_P] * 7 + [_L, _I, _I] + [_F] * 7 + [_P] * 7),
    'vqn_vq_backward': (_I, [_P, _P, _P, _P, _I, _P, _F, _L, _I, _P, _P]),
    'vqn_material_combine_backward': (_I, [_P, _P, _P, _P, _P, _P, _L, _P, _P, _P]),
    'vqn_codebook_sim_loss': (_I, [_P, _P, _I, _I, _F, _P, _P, _I, _P]),
    'vqn_adam_amsgrad': (_I, [_P, _P, _P, _P, _P, _P, _L, _F, _P, _F, _F, _F, _P]),
    'vqn_cast_f64_f32': (_I, [_P, _P, _P, _L, _P]),
    'vqn_cast_f32_f64': (_I, [_P, _P, _P, _L, _P]),
    'vqn_zero_batched': (_I, [_P, C.POINTER(_P), C.POINTER(_L), _I, _P]),
    'vqn_train_pack_stats': (_I, [_P, _P, _P, _L, _P, _F, _P]),
    'vqn_vq_backward_act': (_I, [_P, _P, _P, _P, _I, _P, _F, _L, _I, _P, _I, _P, _L, _P]),
    'vqn_train_scalars': (_I, [_P, _P, _P, _P, _F, _F, _F, _P, _P]),
    'vqn_act_backward_batched': (_I, [_P, _P, _I, _P]),
    'vqn_sample_pairs': (_I, [_P, _P, _I, _I, _I, _F, _I, C.c_uint64, _P, _P, _P, _P, _P, _P, _P]),
    'vqn_gather_rows': (_I, [_P, _P, _P, _L, _I, _P, _P]),
    'vqn_microbench_fma': (_I, [_P, _I, _I, C.POINTER(_D)]),
    'vqn_tc_selftest': (_I, [_P, _I, _I, _I, _P, _P, _P, _P]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """dlopen the in-tree library; raise if it was not built (no fallback of any kind)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    'libvqnerf_b200.so is not built: run `python -m vqnerf_release_b200.build` '
                    '(or __graft_entry__.build()); this package has no CPU / PyTorch fallback')
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)      # AttributeError if the symbol is missing
                fn.restype, fn.argtypes = res, args
            if lib.vqn_abi_version() != 1:
                raise ImportError('libvqnerf_b200.so ABI version mismatch')
            _lib = lib
    return _lib


def check(status: int) -> None:
    """Turn a vqn_status into the exception type the reference would raise."""
    if status == OK:
        return
    lib = load()
    detail = lib.vqn_last_error().decode(errors='replace')
    msg = '%s: %s' % (lib.vqn_status_str(status).decode(), detail)
    if status == ERR_INVALID_ARG:
        raise ValueError(msg)
    if status in (ERR_NONFINITE, ERR_ZERO_NORM):
        raise NonFiniteError(msg)
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def act_code(a) -> int:
    if isinstance(a, str):
        a = a.lower()
    if a not in _ACT_BY_NAME:
        raise ValueError('unsupported activation %r' % (a,))
    return _ACT_BY_NAME[a]


def precision_code(p) -> int:
    if isinstance(p, int):
        return p
    if p not in _PREC_BY_NAME:
        raise ValueError('precision must be one of %s' % sorted(_PREC_BY_NAME))
    return _PREC_BY_NAME[p]


class Context:
    """One vqn_ctx per device (cached)."""
    _cache = {}

    def __init__(self, device_index: int):
        lib = load()
        h = C.c_void_p()
        check(lib.vqn_ctx_create(int(device_index), C.byref(h)))
        self.handle = h
        self.device_index = int(device_index)
        self.lib = lib

    @classmethod
    def get(cls, device) -> 'Context':
        import torch
        dev = torch.device(device)
        if dev.type != 'cuda':
            raise RuntimeError('vqnerf_release_b200 runs on CUDA (sm_100a) only; got device %s' % (dev,))
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if idx not in cls._cache:
            cls._cache[idx] = cls(idx)
        return cls._cache[idx]

    def launch_count(self) -> int:
        return int(self.lib.vqn_ctx_launch_count(self.handle))

    def check_numerics(self, stream_ptr) -> None:
        check(self.lib.vqn_ctx_check_numerics(self.handle, stream_ptr))


def stream_ptr(device=None):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t, dtype=None, allow_none=True):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    import torch
    if t is None:
        if allow_none:
            return None
        raise ValueError('tensor required')
    if not t.is_cuda:
        raise ValueError('expected a CUDA tensor')
    if not t.is_contiguous():
        raise ValueError('expected a contiguous tensor')
    if dtype is not None and t.dtype != dtype:
        raise ValueError('expected dtype %s, got %s' % (dtype, t.dtype))
    return C.c_void_p(t.data_ptr())
