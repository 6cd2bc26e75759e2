"""Mirror of the light-visibility extraction of geo/NeuS-ours2/gen_geo.py: Runner.compute_vis (:182-257) and
Runner.intersect_circle (:346-357) -- SURVEY 8(f) N1, the producer of the shading path's `lvis [N,512]` input.

For every surface point and every light of the 16 x 32 probe that is on the front side of the surface
(direction . normal > 0) a NeuS ray is rendered from the point towards the light, from near = min(0.1, far/2) to the
bounding sphere; lvis = 1 - weight_sum (accumulated opacity), 0 for back-lit pairs.  Ray set-up, front-lit compaction,
the render (native SDF network + scan kernels; the colour network is skipped because only the weights are read) and the
scatter into the [N, n_lights] buffer all run on the device; the only host round trip is the compacted ray count per
chunk (the reference's boolean indexing does the same).  The dataset / mask / file handling around it (gen_rays_at,
lvis.npy / lvis.png writes) stays with the caller.
"""
from __future__ import annotations

import torch

from .. import abi


def compute_vis(renderer, surf, normal, lxyz_flat, max_radius, cos_anneal_ratio=1.0, use_white_bkgd=False,
                batch_size=512, rays_per_render=1 << 16):
    """lvis [N, n_lights] for surface points `surf [N,3]` with normals `normal [N,3]` (already alpha-masked, as
    gen_geo.py:191-192), lights `lxyz_flat [1, n_lights, 3]` (gen_light_xyz), bounding-sphere radius `max_radius`.

    batch_size: surface points per batch (the reference's self.batch_size split, :193-194); the light loop handles
    rays_per_render // batch rays per NeuS render instead of lpix_chunk = 1 light -- the result per pair does not
    depend on the chunking."""
    surf = surf.reshape(-1, 3).to(torch.float32).contiguous()
    normal = normal.reshape(-1, 3).to(torch.float32).contiguous()
    lxyz = lxyz_flat.reshape(-1, 3).to(device=surf.device, dtype=torch.float32).contiguous()
    n_lights = lxyz.shape[0]
    n_pts = surf.shape[0]
    lvis = torch.zeros((n_pts, n_lights), dtype=torch.float32, device=surf.device)
    background_rgb = torch.ones([1, 3], device=surf.device) if use_white_bkgd else None
    for p0 in range(0, n_pts, batch_size):
        sb, nb = surf[p0:p0 + batch_size], normal[p0:p0 + batch_size]
        lv = lvis[p0:p0 + batch_size]                                   # contiguous row block (view)
        lchunk = max(1, min(n_lights, rays_per_render // sb.shape[0]))
        for l0 in range(0, n_lights, lchunk):
            nc = min(lchunk, n_lights - l0)
            rays_o, rays_d, near, far, front = abi.neus_light_rays(sb, nb, lxyz, l0, nc, max_radius)
            row_idx, n_active = abi.compact_mask(front)
            n = int(n_active.item())                                    # one D2H count per chunk
            if n == 0:
                continue                                                # :218-221
            rows = row_idx[:n]
            ro, rd = abi.gather_rows(rays_o, rows), abi.gather_rows(rays_d, rows)
            nr, fr = abi.gather_rows(near, rows), abi.gather_rows(far, rows)
            out = renderer.render(ro, rd, nr, fr, max_radius, cos_anneal_ratio=cos_anneal_ratio,
                                  background_rgb=background_rgb, need_color=False)
            abi.neus_lvis_scatter(out['weight_sum'], rows, n, l0, nc, lv)
    return lvis


def compute_vis_sharded(renderer, surf, normal, lxyz_flat, max_radius, group=None, dst=None, **kw):
    """Surface points sharded over the ranks of `group` (contiguous blocks, dist.shard_rows), one gather of the
    lvis rows at the end (the reference shards whole views over independent processes, gen_geo.py:141-146)."""
    import torch.distributed as dist
    from .. import dist as vdist
    n = surf.reshape(-1, 3).shape[0]
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return compute_vis(renderer, surf, normal, lxyz_flat, max_radius, **kw)
    lo, hi = vdist.shard_rows(n, dist.get_rank(group), dist.get_world_size(group))
    local = compute_vis(renderer, surf.reshape(-1, 3)[lo:hi], normal.reshape(-1, 3)[lo:hi], lxyz_flat, max_radius, **kw)
    return vdist.gather_rows(local, n, group=group, dst=dst)
