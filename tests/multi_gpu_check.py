"""2-rank data-parallel training step == single-device global batch.  Run on a 2-GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tests/multi_gpu_check.py

Every rank first runs the FULL 512-row batch alone (no process group yet), then the ranks shard the same batch
(pairs kept together, dist.shard_rows(align=2)), run the step with the single NCCL all-reduce of
[gradients | VQ statistics | loss sums], and compare parameters, EMA state and loss with the full-batch run.
Also checks the pixel-sharded render + one all-gather against the single-device image."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import decomp_oracle as O  # noqa: E402  (synthetic inputs only)


def build(dev):
    from vqnerf_release_b200.nerfactor.models.vq_nfr import Model
    scene = O.synth_scene(0, bias_scale=0.05, n_probes=2)
    nets = {k: (n.weights, n.biases) for k, n in scene.nets.items()}
    m = Model({'data_type': 'nerf'}, nets=nets, light=scene.light, codebook=scene.codebook.T.copy(),
              novel_probes={'p%d' % i: p for i, p in enumerate(scene.probes)}, device=dev)
    m.assume_all_foreground = True
    return m


def batch_tuple(b, dev, lo, hi):
    t = lambda a: torch.as_tensor(a[lo:hi]).to(dev).contiguous()
    n = hi - lo
    return ('s', torch.zeros((n, 2), dtype=torch.int32, device=dev), t(b['rayo']), t(b['rayd']), t(b['rgb']),
            t(b['alpha']), t(b['pred_alpha']), t(b['xyz']), t(b['normal']), t(b['lvis']))


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
    torch.cuda.set_device(dev)
    from vqnerf_release_b200 import dist as vdist
    from vqnerf_release_b200.nerfactor import train_nfr as T
    n, gbs, k = 512, 256, 15
    batch = O.synth_batch(n, 0)
    thres = np.array([0.0] * 3 + [0.4] * 12)
    # ---- full batch on one device (no process group) ----
    m1 = build(dev)
    o1 = T.Adam(learning_rate=5e-4)
    img1 = m1.fast_render(batch_tuple(batch, dev, 0, n), mode='test', relight_probes=True)[0]['rgb_probes'].clone()
    losses1 = []
    for it in range(2):
        roll = np.random.RandomState(it).uniform(0, 1, size=(1, k))
        l, _, _ = T.train_iter(m1, batch_tuple(batch, dev, 0, n), o1, gbs, thres=thres, roll=roll)
        losses1.append(float(l))
    torch.cuda.synchronize()
    # ---- sharded ----
    dist.init_process_group('nccl', device_id=dev)
    lo, hi = vdist.shard_rows(n, rank, world, align=2)
    m2 = build(dev)
    o2 = T.Adam(learning_rate=5e-4)
    img_local = m2.fast_render(batch_tuple(batch, dev, lo, hi), mode='test', relight_probes=True)[0]['rgb_probes']
    img2 = vdist.gather_rows(img_local, n)
    assert torch.equal(img1, img2), 'sharded render + all-gather differs from the single-device image'
    # fused gather: the shading kernel stores its rows into every rank's symmetric-memory image (P2P over NVLink)
    peer = vdist.PeerImage(n, (3, 3), dev)
    peer.begin_frame()
    pred_p = m2.fast_render(batch_tuple(batch, dev, lo, hi), mode='test', relight_probes=True, peer_image=peer)[0]
    peer.barrier()
    torch.cuda.synchronize()
    # m1 has been trained for two steps meanwhile; compare against a fresh single-device model (which takes the
    # warp-per-point kernel at this size: same arithmetic, fp32 sums in a different order)
    m3 = build(dev)
    ref_all = m3.fast_render(batch_tuple(batch, dev, 0, n), mode='test', relight_probes=True)[0]['rgb_probes']
    assert torch.allclose(peer.tensor[:, 1:, :], ref_all, rtol=5e-5, atol=2e-6), 'fused P2P gather differs from the single-device image'
    assert torch.equal(pred_p['rgb_probes'], peer.tensor[lo:hi, 1:, :]), 'local rows differ from the rows stored into the image'
    print('rank %d: fused P2P gather OK (max |diff| vs single device %.2e)' % (rank, (peer.tensor[:, 1:, :] - ref_all).abs().max().item()), flush=True)
    img_full = peer.tensor.clone()
    # gather semantics: only rank 0's buffer receives the rows
    peer0 = vdist.PeerImage(n, (3, 3), dev, dst=0)
    peer0.begin_frame()
    m2.fast_render(batch_tuple(batch, dev, lo, hi), mode='test', relight_probes=True, peer_image=peer0)
    peer0.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        assert torch.equal(peer0.tensor, img_full), 'gather-to-rank-0 image differs from the all-gather image'
    else:
        assert float(peer0.tensor.abs().max()) == 0.0, 'a non-destination rank received rows'
    print('rank %d: fused gather to rank 0 OK' % rank, flush=True)
    # frames with a CHANGING foreground mask: background rows must read zero in every frame (no stale pixels of the frame
    # two buffers ago), and five back-to-back frames without host synchronisation must each equal the single-device image
    rng = np.random.RandomState(11)
    ref_frames, frames = [], []
    for f in range(5):
        bf = dict(batch)
        bf['alpha'] = (rng.uniform(0, 1, size=(n, 1)) < 0.6).astype(np.float32)
        bf['pred_alpha'] = bf['alpha']
        m3.assume_all_foreground = False
        ref_frames.append(m3.fast_render(batch_tuple(bf, dev, 0, n), mode='test', relight_probes=True)[0]['rgb_probes'].clone())
        frames.append(batch_tuple(bf, dev, lo, hi))
    m2.assume_all_foreground = False
    got = []
    for f in range(5):
        peer0.begin_frame()
        m2.fast_render(frames[f], mode='test', relight_probes=True, peer_image=peer0)
        peer0.barrier()
        if rank == 0:
            got.append(peer0.tensor[:, 1:, :].clone())       # stream-ordered read of frame f before frame f+1 begins
    torch.cuda.synchronize()
    if rank == 0:
        for f in range(5):
            assert torch.allclose(got[f], ref_frames[f], rtol=5e-5, atol=2e-6), 'frame %d differs (stale background rows?)' % f
            bgrows = (frames[f][5] <= 0) if world == 1 else None
        print('rank 0: 5 frames with changing masks, background rows zero in every frame: OK', flush=True)
    # the same through the captured graphs (GraphedFastRender: two graphs, one per frame buffer)
    from vqnerf_release_b200.nerfactor.models.vq_nfr import GraphedFastRender
    gr = GraphedFastRender(m2, frames[0], peer_image=peer0, mode='test', relight_probes=True)
    for f in range(5):
        pred_g, img_g = gr(frames[f])
        if rank == 0:
            assert torch.allclose(img_g[:, 1:, :], ref_frames[f], rtol=5e-5, atol=2e-6), 'graphed frame %d differs' % f
    torch.cuda.synchronize()
    print('rank %d: graph-replayed fused gather OK' % rank, flush=True)
    m2.assume_all_foreground = True
    losses2 = []
    for it in range(2):
        roll = np.random.RandomState(it).uniform(0, 1, size=(1, k))
        l, _, _ = T.train_iter(m2, batch_tuple(batch, dev, lo, hi), o2, gbs, thres=thres, roll=roll)
        losses2.append(float(l))
    torch.cuda.synchronize()
    p1, p2 = m1._train_state.params, m2._train_state.params
    err = (p1 - p2).abs().max().item()
    cs = (m1.vq_layer.state['cs_hidden'] - m2.vq_layer.state['cs_hidden']).abs().max().item()
    dw = (m1.vq_layer.state['dw_hidden'] - m2.vq_layer.state['dw_hidden']).abs().max().item()
    ok = err < 2e-5 and cs == 0.0 and dw < 1e-5 and np.allclose(losses1, losses2, rtol=1e-5)
    print('rank %d: rows [%d,%d) max|param diff| %.3e  cs_hidden diff %.1e  dw_hidden diff %.1e  loss %s vs %s  -> %s'
          % (rank, lo, hi, err, cs, dw, losses1, losses2, 'OK' if ok else 'MISMATCH'), flush=True)
    # ---- NeuS geo stage: ray-sharded render and point-sharded light visibility == single device ----
    from oracle import neus_oracle as NO
    from vqnerf_release_b200 import abi
    from vqnerf_release_b200.neus.fields import RenderingNetwork, SDFNetwork, SingleVarianceNetwork
    from vqnerf_release_b200.neus.gen_geo import compute_vis, compute_vis_sharded
    from vqnerf_release_b200.neus.renderer import NeuSRenderer, render_sharded
    st = NO.make_neus_state(0)
    sdf_net = SDFNetwork(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=(4,), multires=6, device=dev)
    col_net = RenderingNetwork(d_feature=256, mode='idr', d_in=9, d_out=3, d_hidden=256, n_layers=4, multires_view=4,
                               device=dev)
    sdf_net.load_state_dict(st['sdf']); col_net.load_state_dict(st['color'])
    r = NeuSRenderer(None, sdf_net, SingleVarianceNetwork(0.5, device=dev), col_net, 64, 64, 0, 4, 0.0)
    g = torch.Generator().manual_seed(5)
    nb = 301
    o = torch.nn.functional.normalize(torch.randn((nb, 3), generator=g), dim=1).mul(4).to(dev)
    d = torch.nn.functional.normalize(-o + 0.3 * torch.randn((nb, 3), generator=g).to(dev), dim=1)
    near, far = torch.full((nb, 1), 2.0, device=dev), torch.full((nb, 1), 6.0, device=dev)
    bg = torch.ones((1, 3), device=dev)
    one = r.render(o, d, near, far, 1.0, background_rgb=bg, cos_anneal_ratio=1.0)
    sh = render_sharded(r, o, d, near, far, 1.0, background_rgb=bg, cos_anneal_ratio=1.0)
    for kk in ('color_fine', 'weights', 'weight_sum', 'surf', 'depth'):
        assert torch.equal(one[kk], sh[kk]), 'ray-sharded NeuS render differs in %s' % kk
    surf = 0.5 * torch.nn.functional.normalize(torch.randn((9, 3), generator=g), dim=1).to(dev)
    nrm = torch.nn.functional.normalize(surf + 0.3 * torch.randn((9, 3), generator=g).to(dev), dim=1)
    lx, _ = abi.gen_light_xyz(16, 32)
    lxyz = torch.as_tensor(lx.reshape(1, -1, 3), dtype=torch.float32).to(dev)
    lv1 = compute_vis(r, surf, nrm, lxyz, 1.0)
    lv2 = compute_vis_sharded(r, surf, nrm, lxyz, 1.0)
    assert torch.equal(lv1, lv2), 'point-sharded light visibility differs'
    print('rank %d: NeuS ray-sharded render and point-sharded compute_vis bit-identical to one device' % rank, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)


if __name__ == '__main__':
    main()
