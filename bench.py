#!/usr/bin/env python
"""bench.py -- headline benchmark of the VQ-NeRF decomp shading path on B200.

A "step" is one pass of the hot path over one batch of synthetic input: the full-image relighting call
`Model.fast_render(batch, mode='test', relight_probes=True)` (reference: nerfactor/test.py:254-266 ->
models/vq_nfr.py:262-398) on ONE 800x800 NeRF-Blender-shaped view = 640 000 surface points, 16x32 (512-light)
probe, P novel probes, random-init weights (BASELINE.json configs[1]).  At N GPUs the SAME view is pixel-sharded:
contiguous row blocks of 640 000 / N points per rank (STRONG scaling, BASELINE configs[1] "pixel-sharded at 2/4/8"),
and the single gather of the shaded pixels to rank 0 is inside the timed step -- fused into the shading kernel as P2P
stores into rank 0's symmetric-memory image (`--gather nccl`: a separate NCCL all-gather).  The step is replayed from
one CUDA graph per rank (nerfactor/models/vq_nfr.py::GraphedFastRender).  `weak_scaling` (extra key, N > 1): N whole
views per step, one per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--probes P] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` = shaded points/s of the whole job with inputs resident in HBM; `e2e` = the same
job with HOST (pinned) input buffers: H2D copies of every rank's shard, the gather and the D2H read of the shaded image
inside the timed region.  `--impl reference` times the CPU restatement of the reference (oracle/decomp_oracle.py, kind
"port": the TensorFlow reference cannot run in this image) on the host cores, `--steps` steps of a bounded sample of
the same workload after `--warmup` untimed ones.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 800
N_POINTS = H * W
# algorithmic work per point (BASELINE.md section 4 / SURVEY.md 8d)
MLP_ENC_FLOP = 2 * 179968
MLP_HEADS_FLOP = 2 * 296832
SHADE_FLOP_PER_LIGHT = 110.0
# NeuS networks of confs/nerf.conf: MACs per sample
SDF_TRUNK_MACS = 39 * 256 + 2 * 256 * 256 + 256 * 217 + 4 * 256 * 256      # lin0 .. lin7
SDF_MACS = SDF_TRUNK_MACS + 256 * 257                                       # = 524 544
COLOR_MACS = 289 * 256 + 3 * 256 * 256 + 256 * 3                            # = 271 360
CPU_SAMPLE_POINTS = 4096      # chunk size of the CPU restatement (config #1 size; bounds the [N,512,3] intermediates)
CPU_BUDGET_S = 12.0           # keep adding chunks until this much CPU time has been spent


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--probes', type=int, default=8, help='novel relight probes P (plus the model light)')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--points', type=int, default=N_POINTS, help='points per GPU per step (default 800x800)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--gather', default='p2p', choices=['p2p', 'nccl'],
                    help='multi-GPU image gather: p2p = fused into the shading kernel (peer stores), nccl = all_gather')
    ap.add_argument('--precision', default='tf32x3', choices=['fp32', 'bf16', 'tf32x3'],
                    help='MLP arithmetic: tf32x3 = tcgen05 3xTF32 split, fp32 accumulate (fp32 parity, default)')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def synth_view(n, seed, probes):
    """Synthetic view of SURVEY.md 8d config #2 (all foreground = worst case)."""
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
    rayo = rng.normal(size=(n, 3)).astype(np.float32)
    rayo = 4.0 * rayo / np.linalg.norm(rayo, axis=1, keepdims=True)
    normal = rng.normal(size=(n, 3)).astype(np.float32)
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    lvis = rng.random(size=(n, 512), dtype=np.float32)
    alpha = np.ones((n, 1), np.float32)
    rgb = rng.random(size=(n, 3), dtype=np.float32)
    return {'xyz': xyz, 'rayo': rayo.astype(np.float32), 'rayd': (-rayo / 4.0).astype(np.float32),
            'normal': normal.astype(np.float32), 'lvis': lvis, 'alpha': alpha, 'pred_alpha': alpha.copy(), 'rgb': rgb}


def synth_probes(p, seed=123):
    rng = np.random.default_rng(seed)
    return {'probe%02d' % i: (np.abs(rng.normal(size=(16, 32, 3))) * 0.5).astype(np.float32) for i in range(p)}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': float(max(mx)) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def ncu_traffic(n, probes, precision):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r2c_ncu_traffic.json) -- only valid for the workload it was captured on, else None."""
    p = os.path.join(ROOT, 'profiles', 'r2c_ncu_traffic.json')
    try:
        d = json.load(open(p))
        w = d['workload']
        if (w['points'], w['probes'], w['precision']) != (n, probes, precision):
            return None
        f = d['per_launch_dram_bytes']['mlp_tc_kernel<tf32x3> encoder+heads, ONE launch (640000 points)']
        return {'bytes': f['read'] + f['write'], 'tensor_pct': f['sm__pipe_tensor_cycles_active_pct'],
                'note': 'dram__bytes_read.sum + dram__bytes_write.sum of the single encoder+heads launch: the latent z '
                        '[128,256] tile of a CTA goes through a per-CTA scratch that stays in L2; reads = xyz + weights, the '
                        '17.9 MB of outputs were still in L2 when the kernel ended; algorithmic bytes: 40 B/point = 25.6 MB',
                'source': 'profiles/r2c_ncu_traffic.json, profiles/r2c_ncu_full_mlp_tc_main.txt'}
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops', 1590.0), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1590.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------------------------
def cpu_baseline(probes, chunk=CPU_SAMPLE_POINTS, budget_s=CPU_BUDGET_S):
    """The reference's CPU path: the op-for-op PyTorch-CPU restatement (incl. the [N,512,3] intermediates) on all
    host cores, on a bounded sample of the same workload: chunks of `chunk` points of the 800x800 view are shaded
    one after the other (the reference's own brdf_chunk loop, vq_nfr.py:841-874) until `budget_s` is spent."""
    import torch
    from oracle import decomp_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    scene = O.synth_scene(0, n_probes=probes)
    O.fast_render(scene, O.synth_batch(256, 1), torch.float32, relight_probes=probes > 0)   # warm-up
    batches = [O.synth_batch(chunk, s) for s in range(4)]
    done, t0 = 0, time.perf_counter()
    while True:
        O.fast_render(scene, batches[(done // chunk) % 4], torch.float32, relight_probes=probes > 0)
        done += chunk
        dt = time.perf_counter() - t0
        if dt >= budget_s or done >= N_POINTS:
            break
    return {'value': done / dt, 'unit': 'points/s', 'cores': cores, 'kind': 'port',
            'sample': '%s in chunks of %d, %d relight probes, fp32 torch-CPU restatement of vq_nfr.fast_render incl. '
                      '[N,512,3] intermediates; %.1f s of CPU time'
                      % ('one whole 800x800 view (%d chunks = %d points)' % (done // chunk, done) if done >= N_POINTS
                         else '%d of %d points of the 800x800 view' % (done, N_POINTS), chunk, probes, dt)}, dt, done


def run_reference(args):
    """--impl reference: `--warmup` untimed + `--steps` timed steps; a step = the CPU restatement shading a bounded sample
    of the 800x800 view (whole 4096-point chunks, sized so that the run ends within a couple of minutes)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    from oracle import decomp_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    scene = O.synth_scene(0, n_probes=args.probes)
    chunk = CPU_SAMPLE_POINTS
    batches = [O.synth_batch(chunk, s) for s in range(8)]
    t0 = time.perf_counter()
    O.fast_render(scene, batches[0], torch.float32, relight_probes=args.probes > 0)
    one = time.perf_counter() - t0                       # one chunk (cold): sizes the per-step sample
    budget = 90.0 / (steps + warmup)                     # whole run ~1.5 min of CPU time
    chunks_per_step = int(max(1, min(N_POINTS // chunk, budget / max(one, 1e-3))))
    dts, k = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for _ in range(chunks_per_step):
            O.fast_render(scene, batches[k % len(batches)], torch.float32, relight_probes=args.probes > 0)
            k += 1
        if it >= warmup:
            dts.append(time.perf_counter() - t0)
    pts = chunks_per_step * chunk
    ms = float(np.mean(dts)) * 1e3
    val = pts / (ms * 1e-3)
    cb = {'value': val, 'unit': 'points/s', 'cores': cores, 'kind': 'port',
          'sample': '%d points per step (%d chunks of %d) of the 640 000-point view, %d relight probes, fp32 torch-CPU '
                    'restatement of vq_nfr.fast_render incl. the [N,512,3] intermediates' % (pts, chunks_per_step, chunk,
                                                                                              args.probes)}
    line = {'metric': 'shaded surface points/sec', 'value': val, 'unit': 'points/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': warmup, 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': 'vq_nfr.fast_render full-image relight (800x800 view, 512 lights, P=%d probes): '
                                   'bounded sample of %d points per step, CPU restatement of the reference'
                                   % (args.probes, pts)},
            'cpu_baseline': cb,
            'e2e': {'value': val, 'unit': 'points/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
TRAIN_RAYS_PER_GPU = 8192     # BASELINE configs[3]: 64k-ray batch over 8 GPUs
TRAIN_FLOP_PER_RAY = 3 * (MLP_ENC_FLOP + MLP_HEADS_FLOP + 2 * 297600)   # fwd + 2x bwd of the 8 MLPs


def bench_train_step(dev, rank, world, args):
    """nerfactor/train_nfr.py train_iter: Model.call(mode='train') + compute_loss + backward + Sonnet-EMA codebook
    update + Adam(amsgrad), data-parallel over `world` GPUs with ONE NCCL all-reduce of
    [gradients | VQ statistics | loss sums] per step.  Timed with CUDA events, max over ranks."""
    import torch
    import torch.distributed as dist
    from vqnerf_release_b200 import _lib
    from vqnerf_release_b200.nerfactor import train_nfr as T
    from vqnerf_release_b200.nerfactor.models.vq_nfr import Model
    n = TRAIN_RAYS_PER_GPU
    model = Model({'data_type': 'nerf', 'random_seed': 2},
                  light=(np.abs(np.random.default_rng(5).normal(size=(16, 32, 3))) * 0.5).astype(np.float32), device=dev)
    model.assume_all_foreground = True            # the sampler only draws foreground pixels (train_nfr.py:380-467)
    host = synth_view(n, 2000 + rank, 0)
    keys = ('rayo', 'rayd', 'rgb', 'alpha', 'pred_alpha', 'xyz', 'normal', 'lvis')
    d = {k: torch.from_numpy(host[k]).to(dev) for k in keys}
    batch = ('synthetic', torch.zeros((n, 2), dtype=torch.int32, device=dev), d['rayo'], d['rayd'], d['rgb'], d['alpha'],
             d['pred_alpha'], d['xyz'], d['normal'], d['lvis'])
    opt = T.Adam(learning_rate=5e-4, decay_steps=500_000, decay_rate=0.1)
    gbs = n * world // 2                          # n_rays_per_step counts (pixel, neighbour) pairs
    thres = [0.0] * 3 + [0.3] * 12                # codeword dropout as in the VQ stage (train_nfr.py:185-195)
    ctx = _lib.Context.get(dev)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # eager step (every kernel launched through the C ABI from Python): launch-bound at this batch size
    for _ in range(3):
        T.train_iter(model, batch, opt, gbs, thres=thres)
    sync()
    steps = max(args.steps, 10)
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        T.train_iter(model, batch, opt, gbs, thres=thres)
    e1.record()
    sync()
    launches = (ctx.launch_count() - l0) // steps
    eager_ms = e0.elapsed_time(e1) / steps
    # the same step captured once into a CUDA graph and replayed (the product path of the training loop)
    graphed = T.GraphedTrainIter(model, opt, gbs, batch)
    for _ in range(3):
        graphed(batch, thres=thres)
    sync()
    steps = max(args.steps, 20)
    e0.record()
    for _ in range(steps):
        loss, _, _ = graphed(batch, thres=thres)
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    # the collective alone (same flat buffer), for the "all-reduce time reported separately" clause
    ar_ms = 0.0
    st = model._train_state
    if world > 1:
        sync()
        e0.record()
        for _ in range(10):
            dist.all_reduce(st.gflat, op=dist.ReduceOp.SUM)
        e1.record()
        sync()
        ar_ms = e0.elapsed_time(e1) / 10
    return {'rays_per_s': n * world / (ms * 1e-3), 'ms_per_step': ms, 'rays_per_gpu': n, 'global_rays': n * world,
            'allreduce_ms': ar_ms, 'allreduce_bytes': int(st.gflat.numel() * 4), 'kernel_launches_per_step': int(launches),
            'mlp_tflops': TRAIN_FLOP_PER_RAY * n / (ms * 1e-3) / 1e12, 'loss': float(loss), 'eager_ms_per_step': eager_ms,
            'note': 'whole step (fwd, loss, bwd, all-reduce, EMA, Adam) replayed from one CUDA graph; forward = one fused '
                    'tcgen05 launch per network with saved activations (the six heads forked over three streams inside the '
                    'graph), backward = one fused tcgen05 launch per network for the backward-data chain (transposed weight '
                    'images + the saved activations; head chains forked over the same streams) and ONE tcgen05 launch for all '
                    'weight gradients; '
                    'eager_ms_per_step = the same kernels launched one by one from Python'}


def bench_neus_scan(dev, hbm_peak):
    """NeuS geo stage (BASELINE configs[4], secondary path): hierarchical up_sample (4 x 16 importance samples,
    sample_pdf + merge) and render_core compositing as warp-per-ray scan kernels; the SDF / colour networks are the
    caller's (synthetic sdf / gradients / colours here, so this times exactly the scan kernels).  512 rays per batch
    as in the reference (launch-bound), and 65 536 rays (one 256 x 256 tile) for the HBM fraction."""
    import torch
    from vqnerf_release_b200 import abi
    out = {}
    for b in (512, 65536):
        g = torch.Generator(device=dev).manual_seed(b)
        rays_o = torch.randn((b, 3), generator=g, device=dev)
        rays_o = 4.0 * rays_o / rays_o.norm(dim=1, keepdim=True)
        rays_d = -rays_o / 4.0 + 0.05 * torch.randn((b, 3), generator=g, device=dev)
        rays_d = rays_d / rays_d.norm(dim=1, keepdim=True)
        z0 = torch.linspace(2.0, 6.0, 64, device=dev)[None, :].expand(b, 64).contiguous()
        sdf_of = lambda z: ((rays_o[:, None, :] + rays_d[:, None, :] * z[..., None]).norm(dim=-1) - 1.0).contiguous()
        sdf0 = sdf_of(z0)
        new_sdf = [torch.rand((b, 16), generator=g, device=dev) - 0.5 for _ in range(4)]
        grads = torch.randn((b, 128, 3), generator=g, device=dev)
        cols = torch.rand((b, 128, 3), generator=g, device=dev)
        sdf_f = torch.rand((b, 128), generator=g, device=dev) - 0.5

        def run():
            # the launch sequence of NeuSRenderer.render (neus/renderer.py): up_sample, three fused steps (cat_z_vals +
            # up_sample; the last one + the final cat_z_vals + mid points), compositing
            z, sdf = z0, sdf0
            nz, _ = abi.neus_up_sample_pts(rays_o, rays_d, z, sdf, 1.0, 16, 64)
            for i in range(3):
                o = abi.neus_scan_step(rays_o, rays_d, z, nz, sdf, new_sdf[i], 1.0, 16, 64 * 2 ** (i + 1),
                                       final_merge=(i == 2), sample_dist=2.0 / 64, want_merged=(i < 2))
                if i < 2:
                    z, sdf, nz = o['z'], o['sdf'], o['new_z']
            return abi.neus_composite(rays_o, rays_d, o['z_final'], sdf_f, grads, cols, 300.0, 1.0, 2.0 / 64, 1.0)

        ctx = abi._ctx(rays_o)
        l0 = ctx.launch_count()
        run()
        n_launch = ctx.launch_count() - l0

        for _ in range(3):
            run()
        torch.cuda.synchronize(dev)
        reps = 20 if b == 512 else 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / reps
        # algorithmic bytes: composite B*S*(4+4+12+12) in + B*(3+3+1+S)*4 out; every sampling step reads z, sdf
        # (+ the 16 new z / sdf) and writes the merged z, sdf, 16 new z and their 48 B positions; the last one writes the
        # final z and the 2 x 12 B mid points / directions per sample instead of a merged row
        comp = b * 128 * 32 + b * (7 + 128) * 4
        ups = b * (64 * 8 + 16 * 16) + sum(b * ((s + 16) * 8 + (s + 32) * 8 + 16 * 16) for s in (64, 80)) + \
            b * ((96 + 16) * 8 + 128 * 4 + 128 * 24)
        out['rays_%d' % b] = {'ms': ms, 'rays_per_s': b / (ms * 1e-3), 'samples_per_s': b * 128 / (ms * 1e-3),
                              'gbs': (comp + ups) / (ms * 1e-3) / 1e9, 'hbm_frac': (comp + ups) / (ms * 1e-3) / 1e9 / hbm_peak,
                              'launches': int(n_launch)}
    return out


def bench_neus_render(dev):
    """BASELINE configs[4] end to end with the NATIVE networks (neus/fields.py on the fused tcgen05 kernel):
    NeuSRenderer.render = SDF net on 64 coarse samples, 4 x (up_sample + SDF net on 16 new samples + merge),
    render_core (SDF value + feature + gradient jets in one launch, colour net, compositing); nerf.conf network sizes
    (SDF 39 -> 256 x 8 -> 257, colour 289 -> 256 x 4 -> 3), geometric random init.  Also the SDF kernel alone."""
    import torch
    from vqnerf_release_b200.neus.fields import RenderingNetwork, SDFNetwork, SingleVarianceNetwork
    from vqnerf_release_b200.neus.renderer import NeuSRenderer
    torch.manual_seed(0)
    sdf_net = SDFNetwork(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=(4,), multires=6, bias=0.5, scale=1.0,
                         geometric_init=True, weight_norm=True, device=dev)
    col_net = RenderingNetwork(d_feature=256, mode='idr', d_in=9, d_out=3, d_hidden=256, n_layers=4, weight_norm=True,
                               multires_view=4, squeeze_out=True, device=dev)
    dev_net = SingleVarianceNetwork(0.5, device=dev)
    r = NeuSRenderer(None, sdf_net, dev_net, col_net, 64, 64, 0, 4, 0.0)
    out = {}

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    for b in (512, 65536):
        g = torch.Generator(device=dev).manual_seed(b)
        rays_o = torch.randn((b, 3), generator=g, device=dev)
        rays_o = 4.0 * rays_o / rays_o.norm(dim=1, keepdim=True)
        rays_d = -rays_o / 4.0 + 0.05 * torch.randn((b, 3), generator=g, device=dev)
        rays_d = rays_d / rays_d.norm(dim=1, keepdim=True)
        near = torch.full((b, 1), 2.0, device=dev)
        far = torch.full((b, 1), 6.0, device=dev)
        bg = torch.ones((1, 3), device=dev)
        ms = timed(lambda: r.render(rays_o, rays_d, near, far, 1.0, background_rgb=bg, cos_anneal_ratio=1.0),
                   10 if b == 512 else 3)
        # network evaluations per ray: SDF value on 64 + 3 x 16 samples; SDF value + feature + gradient (reverse mode:
        # forward + transposed trunk) and the colour network on 128 samples
        flop = b * (112 * 2 * (SDF_TRUNK_MACS + 256) + 128 * (2 * (SDF_MACS + SDF_TRUNK_MACS) + 2 * COLOR_MACS))
        out['rays_%d' % b] = {'ms': ms, 'rays_per_s': b / (ms * 1e-3), 'samples_per_s': b * 128 / (ms * 1e-3),
                              'tflops_executed': flop / (ms * 1e-3) / 1e12}
    # light-visibility extraction (gen_geo.compute_vis): 256 surface points x 512 lights
    from vqnerf_release_b200 import abi
    from vqnerf_release_b200.neus.gen_geo import compute_vis
    g = torch.Generator(device=dev).manual_seed(3)
    u = torch.randn((256, 3), generator=g, device=dev)
    u = u / u.norm(dim=1, keepdim=True)
    surf = 0.45 * u
    lx, _ = abi.gen_light_xyz(16, 32)
    lxyz = torch.as_tensor(lx.reshape(1, -1, 3), dtype=torch.float32).to(dev)
    ms = timed(lambda: compute_vis(r, surf, u, lxyz, 1.0, cos_anneal_ratio=1.0), 2)
    lv = compute_vis(r, surf, u, lxyz, 1.0, cos_anneal_ratio=1.0)
    n_rays = int((lv != 0).sum().item())
    out['compute_vis_256pts_x_512lights'] = {
        'ms': ms, 'front_lit_rays': n_rays, 'rays_per_s': n_rays / (ms * 1e-3),
        'lvis_entries_per_s': 256 * 512 / (ms * 1e-3), 'mean_lvis': float(lv.mean().item()),
        'tflops_executed': n_rays * (112 * 2 * (SDF_TRUNK_MACS + 256) + 128 * 2 * 2 * (SDF_TRUNK_MACS + 256)) / (ms * 1e-3) / 1e12,
        'note': 'every front-lit (point, light) pair is one NeuS render of 64 + 4 x 16 samples; colour network skipped'}
    n = 1 << 20
    pts = torch.rand((n, 3), device=dev) * 2 - 1
    rows = col_net.alloc_rows(n, dev)
    ms_rev = timed(lambda: sdf_net.forward_with_gradient(pts, feat_out=rows), 3)
    sdf_net.grad_mode = 'jet'
    ms_jet = timed(lambda: sdf_net.forward_with_gradient(pts, feat_out=rows), 3)
    sdf_net.grad_mode = 'reverse'
    ms_val = timed(lambda: sdf_net.sdf(pts), 3)
    out['sdf_kernel_1M_points'] = {
        'value_feature_gradient_ms': ms_rev, 'points_per_s': n / (ms_rev * 1e-3),
        'tflops': n * 2 * (SDF_MACS + SDF_TRUNK_MACS) / (ms_rev * 1e-3) / 1e12,
        'jet_mode_ms': ms_jet, 'jet_mode_tflops_executed': n * 4 * 2 * SDF_MACS / (ms_jet * 1e-3) / 1e12,
        'sdf_only_ms': ms_val, 'sdf_only_tflops': n * 2 * (SDF_TRUNK_MACS + 256) / (ms_val * 1e-3) / 1e12,
        'note': 'gradient in reverse mode inside the launch: forward on value rows with act\' stashed per CTA in L2, then '
                'the transposed trunk layers (2 row-passes per point); jet mode = 4 tile rows per point (value + 3 '
                'tangents), nothing stored; 3xTF32 tensor roofline 273 TFLOP/s'}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    from vqnerf_release_b200 import _lib, abi, dist as vdist
    from vqnerf_release_b200.nerfactor.models.vq_nfr import GraphedFastRender, Model

    n_view = args.points                                  # ONE view, pixel-sharded over the ranks (strong scaling)
    row_a, row_b = vdist.shard_rows(n_view, rank, world)
    n = row_b - row_a
    P = args.probes
    model = Model({'data_type': 'nerf', 'random_seed': 2, 'precision': args.precision},
                  light=(np.abs(np.random.default_rng(5).normal(size=(16, 32, 3))) * 0.5).astype(np.float32),
                  novel_probes=synth_probes(P), device=dev)
    host = synth_view(n, 1000 + rank, P)                  # this rank's rows of the synthetic view
    keys = ('rayo', 'rayd', 'rgb', 'alpha', 'pred_alpha', 'xyz', 'normal', 'lvis')
    pinned = {k: torch.from_numpy(host[k]).pin_memory() for k in keys}
    devt = {k: pinned[k].to(dev, non_blocking=True) for k in keys}
    hw = torch.zeros((n, 2), dtype=torch.int32, device=dev)
    id_ = 'synthetic'

    def batch_of(d):
        return (id_, hw, d['rayo'], d['rayd'], d['rgb'], d['alpha'], d['pred_alpha'], d['xyz'], d['normal'], d['lvis'])

    ctx = _lib.Context.get(dev)

    # multi-GPU: the single gather of the pixel-sharded render is FUSED into the shading kernel (P2P stores of every
    # shaded row into rank 0's symmetric-memory image over NVLink); --gather nccl selects a separate all-gather
    peer_img, gather_mode = None, 'none'
    if world > 1:
        gather_mode = args.gather
        if gather_mode == 'p2p':
            try:
                peer_img = vdist.PeerImage(n_view, (1 + P, 3), dev, dst=0)      # the image is gathered on rank 0
            except Exception as e:                       # no symmetric memory on this box: fall back to NCCL
                if rank == 0:
                    print('bench.py: symmetric memory unavailable (%s); using the NCCL all-gather' % (e,), file=sys.stderr)
                gather_mode = 'nccl'
        ok = torch.tensor([1 if gather_mode == 'p2p' else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            peer_img, gather_mode = None, 'nccl'

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        """`steps` calls of fn between barrier + synchronize on both sides, CUDA events, max over ranks -> ms per step."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    # ---- device-resident step: fast_render of the shard + gather to rank 0, replayed from ONE CUDA graph --------
    graphed = GraphedFastRender(model, batch_of(devt), peer_image=peer_img, mode='test', relight_probes=True)

    def step():
        pred, img = graphed()
        if peer_img is None and world > 1:
            img = vdist.gather_rows(pred['rgb_probes'] if P > 0 else pred['albedo'], n_view)   # --gather nccl
        return img, pred

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    # launches inside the graph are counted at capture time: one replay = the launches of one captured step
    ms_step = timed(step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = n_view / (ms_step * 1e-3)

    # ---- the same step launched kernel by kernel from Python, with per-stage CUDA events (and the launch count) ----
    def eager_step():
        if peer_img is not None:
            peer_img.begin_frame()
        pred, _, _, _ = model.fast_render(batch_of(devt), mode='test', relight_probes=True, peer_image=peer_img)
        if peer_img is not None:
            peer_img.barrier()
            model._mark('gather_barrier', dev)
        elif world > 1:
            vdist.gather_rows(pred['rgb_probes'] if P > 0 else pred['albedo'], n_view)
            model._mark('gather_nccl', dev)

    for _ in range(3):
        eager_step()
    barrier()
    k_eager = max(3, min(args.steps, 10))
    model.stage_events = []
    l0 = ctx.launch_count()
    eager_ms = timed(eager_step, k_eager)
    launches = (ctx.launch_count() - l0) // k_eager
    events = model.stage_events
    model.stage_events = None
    stage_ms = {}
    per = len(events) // k_eager
    for si in range(k_eager):
        chunk = events[si * per:(si + 1) * per]
        for (na, ea), (nb, eb) in zip(chunk[:-1], chunk[1:]):
            stage_ms.setdefault(nb, []).append(ea.elapsed_time(eb))
    stage_ms = {k: float(np.mean(v)) for k, v in stage_ms.items()}
    if world > 1:                                         # per-stage max over the ranks (the slowest shard sets the step)
        names = sorted(stage_ms)
        t = torch.tensor([stage_ms[k] for k in names], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        stage_ms = {k: float(v) for k, v in zip(names, t.tolist())}

    # ---- end-to-end: host buffers in, gathered image out -----------------------------------------------------
    # Every rank: H2D of the inputs of its shard from pinned host memory (rayo, alpha, xyz, normal, lvis), the kernels,
    # D2H of its rows of every predicted map into pinned host memory -- in 8 chunks on two streams so that copies
    # overlap kernels (Model.fast_render_host).  N > 1: the shaded image is gathered on rank 0 by the fused P2P stores
    # (or the NCCL all-gather) and rank 0 reads the WHOLE image back to its host: all inside the timed region.
    e2e_keys = ('rayo', 'alpha', 'xyz', 'normal', 'lvis')
    h2d = sum(pinned[k].numel() * pinned[k].element_size() for k in e2e_keys)
    host_batch = batch_of(pinned)
    e2e_out = {}
    img_host = None
    if world > 1 and rank == 0:
        img_host = torch.empty((n_view, 1 + P, 3), dtype=torch.float32).pin_memory()

    def e2e_step():
        if peer_img is not None:
            peer_img.begin_frame()
        out = model.fast_render_host(host_batch, n_chunks=8, out=e2e_out, mode='test', relight_probes=True,
                                     peer_image=peer_img)
        if peer_img is not None:
            peer_img.barrier()
            if rank == 0:
                img_host.copy_(peer_img.tensor, non_blocking=True)
        elif world > 1:
            dev_rows = out['rgb_probes'].to(dev, non_blocking=True) if P > 0 else out['albedo'].to(dev, non_blocking=True)
            img = vdist.gather_rows(dev_rows, n_view)
            if rank == 0:
                img_host.copy_(img, non_blocking=True)
        return out

    for _ in range(2):
        e2e_step()
    k_e2e = max(2, min(args.steps, 5))
    e2e_ms = timed(e2e_step, k_e2e)
    d2h = sum(v.numel() * v.element_size() for k, v in e2e_out.items() if k != 'alpha')
    tot = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    h2d_job, d2h_job = int(tot[0].item()), int(tot[1].item()) + (img_host.numel() * 4 if img_host is not None else 0)
    if world > 1:
        g = torch.tensor([float(d2h_job)], dtype=torch.float64, device=dev)
        dist.broadcast(g, src=0)
        d2h_job = int(g.item())

    # ---- the same end-to-end job with the opt-in compact light-visibility host formats (abi.compress_lvis) --------
    # float32 (above) is the reference's lvis.npy format and stays the headline; f16 is inside the 1e-4 parity budget
    # (~1e-5 on a radiance), u8 is not (a few 1e-4) -- both are explicit choices of the caller.
    e2e_compact = {}
    for fmt in ('f16', 'u8'):
        pl = dict(pinned)
        pl['lvis'] = abi.compress_lvis(pinned['lvis'], fmt).pin_memory()
        hb = batch_of(pl)
        outc = {}

        def cstep():
            if peer_img is not None:
                peer_img.begin_frame()
            model.fast_render_host(hb, n_chunks=8, out=outc, mode='test', relight_probes=True, peer_image=peer_img)
            if peer_img is not None:
                peer_img.barrier()
                if rank == 0:
                    img_host.copy_(peer_img.tensor, non_blocking=True)

        for _ in range(2):
            cstep()
        cms = timed(cstep, k_e2e)
        hb_bytes = sum(pl[k].numel() * pl[k].element_size() for k in e2e_keys)
        e2e_compact[fmt] = {'value': n_view / (cms * 1e-3), 'unit': 'points/s', 'ms_per_step': cms,
                            'h2d_bytes_per_step_per_gpu': hb_bytes, 'h2d_gbs_per_gpu': hb_bytes / (cms * 1e-3) / 1e9}
        del pl, hb, outc

    # ---- weak scaling (extra key): N whole views per step, one per GPU, gathered on rank 0 ------------------------
    weak = None
    if world > 1:
        del graphed
        hostw = synth_view(n_view, 3000 + rank, P)
        dw = {k: torch.from_numpy(hostw[k]).to(dev) for k in keys}
        hww = torch.zeros((n_view, 2), dtype=torch.int32, device=dev)
        bw = (id_, hww, dw['rayo'], dw['rayd'], dw['rgb'], dw['alpha'], dw['pred_alpha'], dw['xyz'], dw['normal'], dw['lvis'])
        peer_w = None
        if peer_img is not None:
            peer_w = vdist.PeerImage(n_view * world, (1 + P, 3), dev, dst=0)
        gw = GraphedFastRender(model, bw, peer_image=peer_w, mode='test', relight_probes=True)

        def weak_step():
            pred, img = gw()
            if peer_w is None:
                vdist.gather_rows(pred['rgb_probes'] if P > 0 else pred['albedo'], n_view * world)

        for _ in range(3):
            weak_step()
        wms = timed(weak_step, args.steps)
        weak = {'value': n_view * world / (wms * 1e-3), 'unit': 'points/s', 'ms_per_step': wms,
                'points_per_gpu': n_view, 'note': 'one whole 800x800 view per GPU per step, images gathered on rank 0'}
        del gw, dw, bw, peer_w

    # ---- VQ assignment (the second half of BASELINE.json's metric): 4 M latents x K=15, indices only -------------
    vq = None
    if rank == 0:
        nv, kq = 4 * 1024 * 1024, 15
        g = torch.Generator(device=dev).manual_seed(0)
        lat = torch.rand((nv, 256), generator=g, device=dev)
        lat = abi.l2_normalize_rows(lat)
        cbk = abi.get_codebook(torch.rand((256, kq), generator=g, device=dev))
        for _ in range(3):
            abi.vq_assign(lat, cbk, want_quantize=False)
        torch.cuda.synchronize(dev)
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v0.record()
        for _ in range(5):
            abi.vq_assign(lat, cbk, want_quantize=False)
        v1.record()
        torch.cuda.synchronize(dev)
        vms = v0.elapsed_time(v1) / 5
        vq = {'assigns_per_s': nv / (vms * 1e-3), 'ms': vms, 'latents': nv, 'K': kq,
              'gbs': nv * 1032 / (vms * 1e-3) / 1e9}
        del lat

    # ---- decomp training step (BASELINE configs[3]): 8192 rays/GPU, fwd + bwd + EMA + ONE all-reduce + Adam ------
    train = bench_train_step(dev, rank, world, args)

    if rank == 0:
        hbm_peak, bf16_peak, peak_src = measured_peaks()
        # FP32-FMA peak of this GPU, measured live (not in MEASURED_PEAKS.json)
        import ctypes as C
        tf = C.c_double()
        _lib.check(ctx.lib.vqn_microbench_fma(ctx.handle, 0, 2000, C.byref(tf)))
        fp32_peak = float(tf.value)
        tf2 = C.c_double()
        _lib.check(ctx.lib.vqn_microbench_fma(ctx.handle, 1, 2000, C.byref(tf2)))
        mlp_ms = stage_ms.get('mlp_main', float('nan'))
        shade_ms = stage_ms.get('shade', float('nan'))
        MLP_FLOP = MLP_ENC_FLOP + MLP_HEADS_FLOP
        if args.precision == 'fp32':
            # dominant kernel: mlp_simt_kernel (heads launch); FFMA-bound, so the roofline is the fp32-FMA peak
            ach = MLP_FLOP * n / (mlp_ms * 1e-3) / 1e12
            roof = {'bound': 'fp32_fma', 'kernel': 'mlp_simt_kernel (encoder + 3 main heads fused, %d points/launch, '
                                                   '953 600 FLOP/point)' % n,
                    'achieved': ach, 'peak': fp32_peak, 'unit': 'TFLOP/s', 'frac': ach / fp32_peak,
                    'peak_source': 'FFMA peak measured live by vqn_microbench_fma on this GPU '
                                   '(FFMA2 packed: %.1f TFLOP/s); tensor peak for reference: %.0f bf16 TFLOP/s %s'
                                   % (float(tf2.value), bf16_peak, peak_src),
                    'traffic': None}
        else:
            # tf32 tensor rate = bf16 / 2; three MMAs per product in the 3xTF32 split
            peak = bf16_peak if args.precision == 'bf16' else bf16_peak / 2 / 3
            ach = MLP_FLOP * n / (mlp_ms * 1e-3) / 1e12
            roof = {'bound': 'tensor', 'kernel': 'mlp_tc_kernel<%s> (encoder + bottleneck + 3 heads in one launch, 953 600 FLOP/point)' % args.precision,
                    'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s (fp32-equivalent algorithmic FLOPs)', 'frac': ach / peak,
                    'peak_source': 'bf16 cuBLAS peak %.0f TFLOP/s %s%s; FFMA peak measured live: %.1f TFLOP/s' % (
                        bf16_peak, peak_src, '' if args.precision == 'bf16' else ' / 2 (tf32 rate) / 3 (hi/lo split MMAs)', fp32_peak),
                    'traffic': None}
            tr = ncu_traffic(n, P, args.precision)
            if tr is not None:
                roof['traffic'] = tr['bytes']
                roof['traffic_note'] = tr['note'] + ' [' + tr['source'] + ']'
            roof['binding_resource'] = ('the producer warps (drain -> bias / activation -> hi/lo split -> swizzled store -> fence -> '
                                        'arrive): a dependent chain of ~2000 cycles per 32-K chunk and group at 96 registers / 16 '
                                        'warps; issue slots 44 %, tensor pipe 41 % (profiles/r2c_ncu_source_mlp_tc_main.txt); the '
                                        'MMA issue path (per-instruction election loops) was the bound until round 2 -- see '
                                        'DESIGN.md 4.1')
            if args.precision == 'tf32x3':
                # the kernel evaluates the two correction products as ONE bf16 MMA of twice the K, i.e. 4 instead of 6
                # bf16-equivalent MMA units per fp32-equivalent product: its own instruction stream has a higher ceiling
                roof['frac_of_executed_scheme_peak'] = ach / (bf16_peak / 4)
                roof['executed_scheme_note'] = ('peak above = three kind::tf32 MMAs per product (the plain 3xTF32 split); the '
                                                'kernel issues 1 tf32 + 1 bf16(K=16) MMA per 8 K-values = bf16 peak / 4 = '
                                                '%.0f TFLOP/s fp32-equivalent; ncu sm__pipe_tensor_cycles_active of the fused '
                                                'launch: %s %% (profiles/r2c_ncu_full_mlp_tc_main.txt)'
                                                % (bf16_peak / 4, ('%.0f' % tr['tensor_pct']) if tr is not None else '41'))
        shade_bytes = n * (2048 + 36 + 28 + 12 * (1 + P))
        kernels = {
            'mlp_main': {'ms': mlp_ms, 'tflops': MLP_FLOP * n / (mlp_ms * 1e-3) / 1e12},
            'shade': {'ms': shade_ms, 'gbs': shade_bytes / (shade_ms * 1e-3) / 1e9,
                      'tflops': n * 512 * (SHADE_FLOP_PER_LIGHT + 6 * P) / (shade_ms * 1e-3) / 1e12,   # 512 (110 + 6 P)
                      'fp32_fma_frac': n * 512 * (SHADE_FLOP_PER_LIGHT + 6 * P) / (shade_ms * 1e-3) / 1e12 / fp32_peak,
                      'hbm_frac': shade_bytes / (shade_ms * 1e-3) / 1e9 / hbm_peak},
            'other_ms': {k: v for k, v in stage_ms.items() if k not in ('mlp_main', 'shade')},
        }
        cb = None
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only
            cb, _, _ = cpu_baseline(P)
        line = {
            'metric': 'shaded surface points/sec', 'value': value, 'unit': 'points/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_step, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': {'fp32': 'f32', 'bf16': 'bf16', 'tf32x3': 'f32 (3xTF32 tensor-core split, fp32 accumulate)'}[args.precision],
            'data': 'synthetic',
            'config': {'workload': 'vq_nfr.fast_render full-image relight of ONE 800x800 view = %d points (all foreground), '
                                   '512-light probe + P=%d novel probes, random-init MLPs, K=15 codebook; %d points on '
                                   'each of %d GPU(s)' % (n_view, P, n, world),
                       'points_per_view': n_view, 'points_per_gpu': n, 'probes': P, 'precision': args.precision,
                       'parallelism': ('ONE view, pixel rows sharded x%d (strong scaling), image gathered on rank 0: %s' % (world, {'p2p': 'fused into the shading kernel (P2P stores over NVLink into rank 0\'s symmetric-memory image, double-buffered, one device-side barrier per frame)', 'nccl': 'one NCCL all-gather'}[gather_mode])) if world > 1 else 'single GPU',
                       'launch': 'the step is one CUDA-graph replay per rank (GraphedFastRender); eager_ms_per_step = the same kernels launched one by one from Python',
                       'l2': 'inputs (%.2f GB lvis per GPU per step) larger than the 126 MB L2, no flush needed' % (n * 2048 / 1e9)},
            'eager_ms_per_step': eager_ms,
            'e2e': {'value': n_view / (e2e_ms * 1e-3), 'unit': 'points/s', 'ms_per_step': e2e_ms,
                    'h2d_bytes_per_step': h2d_job, 'd2h_bytes_per_step': d2h_job,
                    'h2d_gbs_per_gpu': h2d / (e2e_ms * 1e-3) / 1e9,
                    'includes': 'H2D of every rank\'s shard inputs from pinned host memory, kernels, D2H of every predicted '
                                'map' + (', the gather of the shaded image on rank 0 and its D2H read (%s)' % gather_mode
                                         if world > 1 else ''),
                    'bound': 'host-to-device link: the job moves %.2f GB of fp32 light visibility per view from pinned '
                             'host memory, overlapped with the kernels (fast_render_host); the device-resident step is '
                             '%.1fx shorter' % (h2d_job / 1e9, e2e_ms / ms_step)},
            'e2e_compact_lvis': dict(e2e_compact, note='same job, light visibility sent as float16 / uint8 (q/255) instead of float32: '
                                                           'opt-in host formats (abi.compress_lvis); error on the shaded radiance vs float32 visibility: '
                                                           'f16 <= 3e-5 (inside the 1e-4 parity budget), u8 <= 6e-4 (outside it) -- tests/test_gpu_parity.py'),
            'weak_scaling': weak,
            'gpu_launches': int(launches) * args.steps,
            'gpu_launches_note': '%d launches of this library\'s kernels per step (counted in the eager pass; the timed region replays the same kernels from the captured graph) x %d steps' % (int(launches), args.steps),
            'clocks': clocks,
            'roofline': roof,
            'kernels': kernels,
            'kernels_note': 'per-stage CUDA-event times of the eager pass' + (', max over ranks' if world > 1 else ''),
            'cpu_baseline': cb,
            'vq_assign': dict(vq, hbm_frac=vq['gbs'] / hbm_peak, bound='hbm', algorithmic_bytes_per_latent=1032),
            'train_step': train,
            'neus_scan': bench_neus_scan(dev, hbm_peak),
            'neus_render': bench_neus_render(dev),
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
