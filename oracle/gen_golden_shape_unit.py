"""Golden vectors for the on-disk buffer loader, produced by the REFERENCE'S OWN CODE.

Writes a small synthetic scene (two 6x8 views: one 'nerf' Blender-style camera, exercised in train and test mode) in
the reference's directory layout, then runs the unmodified `nerfactor/datasets/shape_unit.py::Dataset._load_data` /
`_gen_rays` / `_sample_rays` (through the `oracle/tf_shim` TensorFlow stand-in; the constructor's tf.data pipeline is
bypassed with `Dataset.__new__`) and records the scene's raw arrays together with the reference's outputs.

    python oracle/gen_golden_shape_unit.py      -> tests/golden/shape_unit_ref.npz       (needs /root/reference)
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('VQN_REFERENCE', '/root/reference') + '/decomp/nerfvq_nfr3'
OUT = os.path.join(HERE, '..', 'tests', 'golden', 'shape_unit_ref.npz')
sys.dont_write_bytecode = True
sys.path[:0] = [os.path.join(HERE, 'tf_shim'), REF, os.path.join(REF, 'nerfactor'), os.path.join(HERE, '..')]


def synth_scene(seed=0, h=6, w=8):
    rng = np.random.RandomState(seed)
    views = {}
    for vid in ('train_000', 'val_000'):
        ang = rng.uniform(0, 2 * np.pi)
        c2w = np.eye(4)
        c2w[:3, :3] = np.linalg.qr(rng.standard_normal((3, 3)))[0]
        c2w[:3, 3] = 4.0 * np.array([np.cos(ang), np.sin(ang), 0.3])
        xyz = rng.uniform(-1, 1, size=(h, w, 3)).astype(np.float32)
        normal = rng.standard_normal((h, w, 3)).astype(np.float32)
        normal[0, :3] = 0.0                                        # all-zero normals -> +y
        alpha = (rng.uniform(0, 1, size=(h, w)) * 255).astype(np.uint8)
        alpha[:, :4] = 0
        rgba = (rng.uniform(0, 1, size=(h, w, 4)) * 255).astype(np.uint8)
        lvis = rng.uniform(-0.1, 1.1, size=(h, w, 512)).astype(np.float32)       # clipped to [0, 1] by the loader
        views[vid] = dict(metadata={'cam_transform_mat': ','.join('%.17g' % v for v in c2w.reshape(-1)),
                                    'cam_angle_x': 0.6911, 'imh': h, 'imw': w}, rgba=rgba, xyz=xyz, normal=normal,
                          alpha=alpha, lvis=lvis, c2w=c2w)
    return views


def main():
    import configparser
    import tensorflow as tf  # noqa: F401  (the shim)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        from nerfactor.datasets.shape_unit import Dataset
    from vqnerf_release_b200.nerfactor.datasets.shape_unit import write_view
    views = synth_scene()
    rec = {}
    with tempfile.TemporaryDirectory() as tmp:
        root, nroot = os.path.join(tmp, 'data'), os.path.join(tmp, 'surf')
        for vid, v in views.items():
            write_view(root, nroot, vid, v['metadata'], v['rgba'], v['xyz'], v['normal'], v['alpha'], v['lvis'])
            # the collapsed-point rule needs xyz == rayo at one pixel: camera location at pixel (2, 5)
            x = np.load(os.path.join(nroot, vid, 'xyz.npy'))
            x[2, 5] = v['c2w'][:3, 3].astype(np.float32)
            np.save(os.path.join(nroot, vid, 'xyz.npy'), x)
            v['xyz'] = x
            for k in ('rgba', 'xyz', 'normal', 'alpha', 'lvis'):
                rec['%s_in_%s' % (vid, k)] = v[k]
            rec['%s_in_metadata' % vid] = np.array(str(v['metadata']))
            rec['%s_in_cam' % vid] = v['c2w']
        cfg = configparser.ConfigParser()
        cfg['DEFAULT'] = {'data_root': root, 'data_nerf_root': nroot, 'data_type': 'nerf', 'model': 'vq_nfr', 'imh': '6',
                          'white_bg': 'True', 'use_nerf_alpha': 'False', 'random_seed': '2', 'n_rays_per_step': '1024'}
        for mode, vid in (('train', 'train_000'), ('test', 'val_000'), ('vali', 'val_000')):
            ds = Dataset.__new__(Dataset)
            ds.config, ds.mode, ds.debug, ds.meta2buf = cfg, mode, False, {}
            files = ds._glob()
            assert len(files) == 1 and ds._parse_id(files[0]) == vid, files
            out = ds._load_data(tf.convert_to_tensor(np.array(files[0])) if False else _Str(files[0]))
            names = ('id', 'rayo', 'rayd', 'rgb', 'alpha', 'pred_alpha', 'xyz', 'normal', 'lvis')
            for nm, a in zip(names[1:], out[1:]):
                rec['%s_%s' % (mode, nm)] = np.asarray(a, np.float32)          # tf.py_function casts to float32 (:136-139)
            flat = ds._sample_rays(*[np.asarray(a, np.float32) for a in out[1:]])
            rec['%s_flat_shapes' % mode] = np.array([list(np.asarray(f).shape) for f in flat])
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **rec)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


class _Str(str):
    """tutil.eager_tensor_to_str(path) calls .numpy().decode() on the py_function argument"""

    def numpy(self):
        return self.encode()


if __name__ == '__main__':
    main()
