"""shade_pt_kernel (+ its prep kernel) against the shard size: whole rounds of 2368 persistent warps and split last rounds."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vqnerf_release_b200 import abi
from vqnerf_release_b200.nerfactor.models.vq_nfr import Model
dev = torch.device('cuda:0')
m = Model({'data_type': 'nerf'}, device=dev)
for k in range(8):
    m.novel_probes['p%d' % k] = torch.rand((16, 32, 3), device=dev)
lights = m._lights(True, None)
sizes = [int(v) for v in sys.argv[1].split(',')] if len(sys.argv) > 1 else (75776, 80000, 151552, 160000, 640000)
for n in sizes:
    host = bench.synth_view(n, 1, 8)
    d = {k: torch.from_numpy(host[k]).to(dev) for k in ('xyz', 'rayo', 'normal', 'lvis')}
    alb, spec, rough = torch.rand((n, 3), device=dev), torch.rand((n, 3), device=dev) * 0.1, torch.rand((n, 1), device=dev)
    out = torch.empty((n, 9, 3), device=dev)
    run = lambda: abi.shade(d['xyz'], d['rayo'], d['normal'], d['lvis'], alb, spec, rough, m.lxyz, m.lareas, lights, n=n, n_total=n, to_srgb=True, out_rgb=out)
    st = torch.cuda.Stream(device=dev)              # the library's work buffers are per stream: warm up on the capture stream
    st.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(st):
        for _ in range(3): run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    print('n=%7d tiles=%6d rounds=%.2f  shade (prep + kernel, graph replay) %.4f ms' % (n, (n + 31) // 32, (n + 31) // 32 / 2368, e0.elapsed_time(e1) / 20))
