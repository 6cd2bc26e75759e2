"""On-disk buffer loader (second half) against vectors recorded from the reference's own
`datasets/shape_unit.py::Dataset._load_data / _gen_rays / _sample_rays` (oracle/gen_golden_shape_unit.py)."""
import ast
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'shape_unit_ref.npz')


@pytest.fixture(scope='module')
def scene(tmp_path_factory):
    from vqnerf_release_b200.nerfactor.datasets.shape_unit import write_view
    g = np.load(GOLD)
    tmp = tmp_path_factory.mktemp('scene')
    root, nroot = str(tmp / 'data'), str(tmp / 'surf')
    for vid in ('train_000', 'val_000'):
        meta = ast.literal_eval(str(g[vid + '_in_metadata']))
        write_view(root, nroot, vid, meta, g[vid + '_in_rgba'], g[vid + '_in_xyz'], g[vid + '_in_normal'],
                   g[vid + '_in_alpha'], g[vid + '_in_lvis'])
    cfg = {'data_root': root, 'data_nerf_root': nroot, 'data_type': 'nerf', 'imh': 6, 'white_bg': True,
           'n_rays_per_step': 1024}
    return g, cfg


@pytest.mark.parametrize('mode', ['train', 'test', 'vali'])
def test_load_data_matches_reference(scene, mode):
    from vqnerf_release_b200.nerfactor.datasets.shape_unit import Dataset
    g, cfg = scene
    ds = Dataset(cfg, mode)
    assert ds.get_n_views() == 1
    out = ds._load_data(ds.files[0])
    assert out[0] == ('train_000' if mode == 'train' else 'val_000')
    for nm, a in zip(('rayo', 'rayd', 'rgb', 'alpha', 'pred_alpha', 'xyz', 'normal', 'lvis'), out[1:]):
        assert a.dtype == np.float32, nm
        np.testing.assert_allclose(a, g['%s_%s' % (mode, nm)], rtol=1e-6, atol=1e-7, err_msg=nm)
    # the special cases the reference handles: collapsed point (xyz == camera) moved 0.1 along the ray, zero normals
    # -> +y, visibility clipped to [0, 1], test views use the predicted alpha as ground truth
    assert np.allclose(out[6][2, 5], out[1][2, 5] + 0.1 * out[2][2, 5])
    assert np.allclose(out[7][0, :3], [0., 1., 0.])
    assert out[8].min() >= 0 and out[8].max() <= 1
    if mode == 'test':
        np.testing.assert_array_equal(out[4], out[5])
    flat = ds._sample_rays(*out[1:])
    np.testing.assert_array_equal(np.array([list(f.shape) for f in flat]), g[mode + '_flat_shapes'])


def test_view_is_the_models_batch_tuple(scene):
    from vqnerf_release_b200.nerfactor.datasets.shape_unit import Dataset
    g, cfg = scene
    ds = Dataset(cfg, 'test')
    b = ds.view(0, pin=False)
    assert len(b) == 10 and b[0] == 'val_000'
    n = 6 * 8
    assert b[1].shape == (n, 2) and b[1].dtype == torch.int32 and b[1][0].tolist() == [6, 8]
    for t, c in zip(b[2:], (3, 3, 3, 1, 1, 3, 3, 512)):
        assert t.shape == (n, c) and t.dtype == torch.float32 and t.is_contiguous()
    np.testing.assert_allclose(b[9].numpy().reshape(6, 8, 512), g['test_lvis'], rtol=1e-6, atol=1e-7)
    assert ds._get_batch_size() == n
    assert Dataset(cfg, 'train')._get_batch_size() == 1024
    # opt-in compact visibility rows
    b8 = Dataset(cfg, 'test', lvis_format='u8').view(0, pin=False)
    assert b8[9].dtype == torch.uint8 and float((b8[9].float() / 255 - b[9]).abs().max()) <= 0.5 / 255 + 1e-6


def test_incomplete_views_are_skipped(scene, tmp_path):
    from vqnerf_release_b200.nerfactor.datasets.shape_unit import Dataset
    g, cfg = scene
    os.makedirs(os.path.join(cfg['data_root'], 'val_001'), exist_ok=True)
    with open(os.path.join(cfg['data_root'], 'val_001', 'metadata.json'), 'w') as fh:
        fh.write('{}')
    ds = Dataset(cfg, 'vali')
    assert ds.get_n_views() == 1 and len(ds.incomplete) == 1
    with pytest.raises(AssertionError):
        Dataset(dict(cfg, data_root=str(tmp_path)), 'vali')
