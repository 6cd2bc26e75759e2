"""Mirror of decomp/nerfvq_nfr3/nerfactor/datasets/shape_unit.py::Dataset -- the on-disk buffer format either side of the
shading path (second half).

A scene is a directory tree

    <data_root>/{train,val}_???/metadata.json          camera: cam_transform_mat, cam_angle_x, imh, imw (, cx, cy)
    <data_root>/{train,val}_???/rgba.png                ground-truth RGBA
    <data_nerf_root>/<view id>/xyz.npy     [H,W,3]      surface points        (geo stage, gen_geo.py)
    <data_nerf_root>/<view id>/normal.npy  [H,W,3]      surface normals
    <data_nerf_root>/<view id>/alpha.png   [H,W]        predicted alpha
    <data_nerf_root>/<view id>/lvis.npy    [H,W,512]    light visibility      (data_type == 'nerf' only)

and a view becomes the batch tuple the model consumes (shape_unit.py:93-110):
`(id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal[, lvis])`, every tensor flattened to [H*W, c].

Same method names and arithmetic as the reference (`_glob`, `_load_data`, `_gen_rays`, `_sample_rays`,
`decompose_projection_matrix`, `_parse_id`, `get_n_views`); what differs is the container: no tf.data pipeline -- `view(i)`
returns the tuple as PINNED host tensors, ready for `Model.fast_render_host` (copies overlap the kernels) or `.to(device)`.
`lvis_format` ('f32' default, 'f16', 'u8') applies abi.compress_lvis to the visibility rows at load time (opt-in).
`write_view` stores geo-stage outputs in the same format (the producer side, gen_geo.py:246-257).
"""
from __future__ import annotations

import glob
import json
import os
from os.path import basename, dirname, join

import numpy as np
import torch


def _read_img(path):
    from PIL import Image                                   # xm.io.img.read: "whatever Pillow supports"
    with Image.open(path) as im:
        return np.array(im)


def _normalize_uint(arr):
    """xm.img.normalize_uint: uint8 / uint16 -> float64 in [0, 1]"""
    if arr.dtype not in (np.uint8, np.uint16):
        raise TypeError(arr.dtype)
    return arr.astype(float) / np.iinfo(arr.dtype).max


def _resize(img, new_h):
    """The reference resizes with TF's antialiased bilinear filter (util/img.py:100-135 through xm.img.resize); buffers are
    normally stored at the working resolution (imh), so this is the rare path: cv2 area/linear interpolation."""
    import cv2
    h, w = img.shape[:2]
    new_w = int(round(w * new_h / h))
    interp = cv2.INTER_AREA if new_h < h else cv2.INTER_LINEAR
    out = cv2.resize(img, (new_w, new_h), interpolation=interp)
    return out.reshape((new_h, new_w) + img.shape[2:])


class Dataset:
    MODES = ('train', 'vali', 'test', 'render')

    def __init__(self, config, mode, debug=False, lvis_format='f32'):
        if mode not in self.MODES:
            raise AssertionError("Accepted dataset modes: 'train', 'vali', 'test', 'render', but input is %s" % mode)
        self.config, self.mode, self.debug = config, mode, debug
        self.lvis_format = lvis_format
        self.meta2buf = {}
        self.files = self._glob()
        assert self.files, "No file to process into a dataset"

    def _cfg(self, key, kind=str, fallback=None):
        c = self.config
        if hasattr(c, 'has_option'):
            if not c.has_option('DEFAULT', key):
                if fallback is None:
                    raise KeyError(key)
                return fallback
            v = c.get('DEFAULT', key)
        else:
            if key not in c:
                if fallback is None:
                    raise KeyError(key)
                return fallback
            v = c[key]
        if kind is bool:
            return str(v).strip().lower() in ('1', 'true', 'yes', 'on')
        return kind(v)

    # ------------------------------------------------------------------ shape_unit.py:46-90
    def _glob(self):
        root = self._cfg('data_root')
        nerf_root = self._cfg('data_nerf_root')
        self.data_type = self._cfg('data_type')
        mode_str = 'train' if self.mode in ('train', 'render') else 'val'
        metadata_dir = join(root, '%s_002' % mode_str) if self.debug else join(root, '%s_???' % mode_str)
        metadata_paths, incomplete = [], []
        for metadata_path in sorted(glob.glob(join(metadata_dir, 'metadata.json'))):
            id_ = self._parse_id(metadata_path)
            paths = {'xyz': join(nerf_root, id_, 'xyz.npy'), 'normal': join(nerf_root, id_, 'normal.npy'),
                     'alpha': join(nerf_root, id_, 'alpha.png'), 'rgba': join(dirname(metadata_path), 'rgba.png')}
            if self.data_type == 'nerf':
                paths['lvis'] = join(nerf_root, id_, 'lvis.npy')
            if all(os.path.exists(p) for p in paths.values()):
                metadata_paths.append(metadata_path)
                self.meta2buf[metadata_path] = paths
            else:
                incomplete.append(metadata_path)        # skipped: at least one paired buffer is missing (:82-86)
        self.incomplete = incomplete
        return metadata_paths

    @staticmethod
    def _parse_id(metadata_path):
        return basename(dirname(metadata_path))

    def get_n_views(self):
        return len(self.files)

    def _get_batch_size(self):
        """shape_unit.py:323-335: training batches are n_rays_per_step pairs, the others one whole view."""
        if self.mode == 'train':
            return self._cfg('n_rays_per_step', int)
        ret = self._load_data(self.files[0])
        return int(ret[2].shape[0] * ret[2].shape[1])

    # ------------------------------------------------------------------ shape_unit.py:149-262
    def _load_data(self, metadata_path):
        imh = self._cfg('imh', int)
        white_bg = self._cfg('white_bg', bool)
        id_ = self._parse_id(metadata_path)
        with open(metadata_path) as fh:
            metadata = json.load(fh)
        if self.data_type == 'dtu':
            k = imh / metadata['imh']
            imw = int(k * metadata['imw'])
            scaled_projection = (np.array(metadata['world_mat']) @ np.array(metadata['scale_mat']))[0:3, 0:4]
            intrinsic, cam_to_world = self.decompose_projection_matrix(scaled_projection)
            intrinsic[:2, :3] = intrinsic[:2, :3] * k
            rayo, rayd = self._gen_rays(cam_to_world, np.linalg.inv(intrinsic), imh, imw)
        else:
            imw = int(metadata['imw'] * imh / metadata['imh'])
            cam_to_world = np.array([float(x) for x in metadata['cam_transform_mat'].split(',')]).reshape(4, 4)
            cx = cy = None
            if 'cx' in metadata:
                cx, cy = imh / metadata['imh'] * metadata['cx'], imh / metadata['imh'] * metadata['cy']
            rayo, rayd = self._gen_rays(cam_to_world, metadata['cam_angle_x'], imh, imw, cx, cy)
        rayo, rayd = rayo.astype(np.float32), rayd.astype(np.float32)
        paths = self.meta2buf[metadata_path]
        xyz = np.load(paths['xyz'])
        normal = np.load(paths['normal'])
        pred_alpha = _normalize_uint(_read_img(paths['alpha']))
        rgba = _read_img(paths['rgba'])
        assert rgba.ndim == 3 and rgba.shape[2] == 4, "Input image is not RGBA"
        rgba = _normalize_uint(rgba)
        rgb = rgba[:, :, :3]
        alpha = pred_alpha if self.mode == 'test' else rgba[:, :, 3]       # test views: gt_alpha = pred_alpha (:197)
        if imh != xyz.shape[0]:
            xyz = _resize(xyz, imh)
        if imh != normal.shape[0]:
            normal = _resize(normal, imh)
        if imh != alpha.shape[0]:
            alpha = _resize(alpha, imh)
        if imh != pred_alpha.shape[0]:
            pred_alpha = _resize(pred_alpha, imh)
        if imh != rgb.shape[0]:
            rgb = _resize(rgb, imh)
        # collapsed point and camera (occupancy accumulating to 0): push the point 0.1 along the ray (:236-238)
        zero_bg = np.linalg.norm(xyz - rayo, axis=-1) == 0.
        xyz[zero_bg] = rayo[zero_bg] + rayd[zero_bg] * 0.1
        # all-zero normals -> +y, then re-normalise (:240-243)
        zero_bg = np.mean(normal, axis=-1) == 0.
        normal[zero_bg] = np.array([0., 1., 0.])
        normal = normal / np.linalg.norm(normal, axis=2, keepdims=True)
        # composite onto white / black (:245-248, util/img.py:78-97)
        bg = np.ones_like(rgb) if white_bg else np.zeros_like(rgb)
        a3 = np.tile(alpha.reshape(alpha.shape + (1,)), (1, 1, rgb.shape[2])) if alpha.ndim == 2 else alpha
        rgb = (np.multiply(rgb, a3) + np.multiply(bg, 1. - a3)).astype(np.float32)
        out = [id_, rayo, rayd, rgb, alpha.astype(np.float32), pred_alpha.astype(np.float32), xyz.astype(np.float32),
               normal.astype(np.float32)]
        if self.data_type == 'nerf':
            lvis = np.load(paths['lvis'])
            if imh != lvis.shape[0]:
                lvis = _resize(lvis, imh)
            out.append(np.clip(lvis, 0, 1).astype(np.float32))
        return tuple(out)

    # ------------------------------------------------------------------ shape_unit.py:265-296
    def _gen_rays(self, to_world, intrinsic, imh, imw, cx=None, cy=None):
        cam_loc = to_world[:3, 3]
        rayo = np.tile(cam_loc[None, None, :], (imh, imw, 1))
        xs = np.linspace(0, imw, imw, endpoint=False)
        ys = np.linspace(0, imh, imh, endpoint=False)
        xs, ys = np.meshgrid(xs, ys)
        if self.data_type == 'dtu':
            p = np.stack((xs, ys, np.ones_like(xs)), axis=-1)
            p = (intrinsic[None, None, :3, :3] @ p[..., None])[..., 0]
            rayd = p / np.linalg.norm(p, ord=2, axis=-1, keepdims=True)
            rayd = (to_world[None, None, :3, :3] @ rayd[..., None])[..., 0]
        else:
            fl = .5 * imw / np.tan(.5 * intrinsic)
            if cx is None:
                cx = .5 * imw
            if cy is None:
                cy = .5 * imh
            rayd = np.stack(((xs - cx) / fl, -(ys - cy) / fl, -np.ones_like(xs)), axis=-1)   # camera frame
            rayd = np.sum(rayd[:, :, np.newaxis, :] * to_world[:3, :3], axis=-1)             # world frame
        return rayo, rayd

    def decompose_projection_matrix(self, P):
        """shape_unit.py:298-315 (DTU): intrinsics and camera-to-world pose of a 3x4 projection matrix."""
        import cv2
        out = cv2.decomposeProjectionMatrix(P)
        K, R, t = out[0], out[1], out[2]
        K = K / K[2, 2]
        intrinsics = np.eye(4)
        intrinsics[:3, :3] = K
        pose = np.eye(4, dtype=np.float32)
        pose[:3, :3] = R.transpose()
        pose[:3, 3] = (t[:3] / t[3])[:, 0]
        return intrinsics, pose

    # ------------------------------------------------------------------ shape_unit.py:93-128
    def _sample_rays(self, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, lvis=None):
        flat = lambda a, c: a.reshape(-1, c)
        out = [flat(rayo, 3), flat(rayd, 3), flat(rgb, 3), flat(alpha, 1), flat(pred_alpha, 1), flat(xyz, 3),
               flat(normal, 3)]
        if self.data_type == 'nerf':
            if lvis is None:
                raise ValueError('NeRF data requires lvis')
            out.append(lvis.reshape(-1, lvis.shape[2]))
        return tuple(out)

    def view(self, i, device=None, pin=True):
        """One whole view as the model's batch tuple (`_process_example_precache` + `_process_example_postcache`)."""
        data = self._load_data(self.files[i])
        id_, hw = data[0], data[3].shape[:2]
        rows = self._sample_rays(*data[1:])
        n = rows[0].shape[0]
        ts = [torch.from_numpy(np.ascontiguousarray(r)) for r in rows]
        if self.data_type == 'nerf' and self.lvis_format not in (None, 'f32'):
            from ... import abi
            ts[-1] = abi.compress_lvis(ts[-1], self.lvis_format)
        hw_t = torch.tensor([list(hw)], dtype=torch.int32).repeat(n, 1)            # (:107-108)
        if device is not None:
            ts = [t.to(device) for t in ts]
            hw_t = hw_t.to(device)
        elif pin and torch.cuda.is_available():
            ts = [t.pin_memory() for t in ts]
        return (id_, hw_t) + tuple(ts)

    def __len__(self):
        return len(self.files)

    def __getitem__(self, i):
        return self.view(i)


def write_view(data_root, data_nerf_root, view_id, metadata, rgba_u8, xyz, normal, alpha_u8, lvis=None):
    """The producer side of the format (geo stage, gen_geo.py:246-257 + the renderer's metadata.json / rgba.png)."""
    from PIL import Image
    vdir, ndir = join(data_root, view_id), join(data_nerf_root, view_id)
    os.makedirs(vdir, exist_ok=True)
    os.makedirs(ndir, exist_ok=True)
    with open(join(vdir, 'metadata.json'), 'w') as fh:
        json.dump(metadata, fh)
    Image.fromarray(np.asarray(rgba_u8, np.uint8), 'RGBA').save(join(vdir, 'rgba.png'))
    np.save(join(ndir, 'xyz.npy'), np.asarray(xyz, np.float32))
    np.save(join(ndir, 'normal.npy'), np.asarray(normal, np.float32))
    Image.fromarray(np.asarray(alpha_u8, np.uint8), 'L').save(join(ndir, 'alpha.png'))
    if lvis is not None:
        np.save(join(ndir, 'lvis.npy'), np.asarray(lvis, np.float32))
