"""Host-side logic on CPU: config mirror, row sharding, and the world_size-2 gloo path of the single
render gather / training all-reduce."""
import configparser
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vqnerf_release_b200 import dist as vdist


def test_cfg_accepts_configparser_and_dict(built_lib):
    from vqnerf_release_b200.nerfactor.models.vq_nfr import _Cfg
    cp = configparser.ConfigParser()
    cp.read_string('[DEFAULT]\nnum_embed = 7\ndata_type = nerf\nno_brdf_chunk = False\nalbedo_slope = 0.77\n')
    c = _Cfg(cp)
    assert c.getint('num_embed') == 7 and c.get('data_type') == 'nerf' and not c.getboolean('no_brdf_chunk')
    assert abs(c.getfloat('albedo_slope') - 0.77) < 1e-12
    assert c.getint('conv_width') == 256                   # default of vq_nfr.ini:100
    assert c.getint('brdf_chunk_size', fallback=123) == 123
    d = _Cfg({'num_embed': 31})
    assert d.getint('num_embed') == 31 and d.getint('light_h') == 16


@pytest.mark.parametrize('n,w,align', [(640000, 8, 1), (10, 3, 1), (7, 8, 1), (65536, 8, 2), (11, 4, 2), (0, 2, 1)])
def test_shard_rows_partition(n, w, align):
    spans = [vdist.shard_rows(n, r, w, align) for r in range(w)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans[:-1], spans[1:]):
        assert b == c and a <= b
    sizes = [b - a for a, b in spans]
    assert sum(sizes) == n and max(sizes) - min(sizes) <= 2 * align - 1   # last unit may be partial
    if align == 2:
        assert all(a % 2 == 0 for a, _ in spans)           # (pixel, neighbour) pairs stay together
    with pytest.raises(ValueError):
        vdist.shard_rows(n, w, w)


def _worker(rank, world, port, n_total, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        full = torch.arange(n_total * 3, dtype=torch.float32).reshape(n_total, 3)
        a, b = vdist.shard_rows(n_total, rank, world)
        got = vdist.gather_rows(full[a:b].clone(), n_total)
        ok = torch.equal(got, full)
        got_dst = vdist.gather_rows(full[a:b].clone(), n_total, dst=0)
        ok = ok and ((got_dst is None) == (rank != 0))
        # one flat all-reduce over mixed dtypes (grads fp32 | VQ stats fp64)
        g = torch.full((5,), float(rank + 1))
        s = torch.full((4,), float(rank + 1), dtype=torch.float64)
        vdist.allreduce_flat_([g, s])
        tot = world * (world + 1) / 2
        ok = ok and bool((g == tot).all()) and bool((s == tot).all())
        hook = vdist.make_stats_allreduce()
        st = torch.ones((3,), dtype=torch.float64) * (rank + 1)
        hook(st)
        ok = ok and bool((st == tot).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_total', [10, 9])     # even and ragged blocks
def test_gather_and_allreduce_world2_gloo(n_total):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def test_single_process_paths_are_identity():
    t = torch.ones((4, 3))
    assert vdist.gather_rows(t, 4) is t
    vdist.allreduce_flat_([t])
    assert bool((t == 1).all())


def test_adam_schedule_matches_keras(built_lib):
    """train_nfr.Adam.lr_t == tf.keras Adam's lr_t with ExponentialDecay (host arithmetic, no GPU):
    lr * rate ** (step / decay_steps) * sqrt(1 - b2^t) / (1 - b1^t), the schedule seeing the pre-increment step."""
    import math
    from vqnerf_release_b200.nerfactor import train_nfr as T
    opt = T.Adam(learning_rate=5e-4, decay_steps=500_000, decay_rate=0.1)
    for t in (1, 2, 10, 1000, 500_001):
        opt.iterations = t
        want = 5e-4 * 0.1 ** ((t - 1) / 500_000) * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        assert abs(opt.lr_t() - want) <= 1e-12 * want
    assert T.Adam(learning_rate=1e-3).lr_t.__self__.decay_steps == -1
    with pytest.raises(NotImplementedError):
        T.Adam(amsgrad=False)
    assert T._pad4(63) == 64 and T._pad4(64) == 64
    assert T.NET_ORDER[0] == 'fine_enc' and len(T.NET_ORDER) == 8
