"""Mirror of geo/NeuS-ours2/models/renderer.py::NeuSRenderer (secondary path, n_outside == 0).

The per-ray work -- up_sample (+sample_pdf det=True), the sort/merge of cat_z_vals and the SDF-to-alpha
compositing of render_core -- runs in the warp-per-ray kernels of csrc/neus.cu.  The SDF / colour /
variance networks are passed in exactly as the reference passes them; with the native networks of
neus/fields.py (fused tcgen05 kernel) render_core takes the fused route: ONE launch gives sdf, feature vector and
SDF gradient, the feature vector lands directly in the colour network's input rows.  Any other callable (a torch
module) is called the way the reference calls it.
"""
from __future__ import annotations

import torch

from .. import abi


class NeuSRenderer:
    def __init__(self, nerf, sdf_network, deviation_network, color_network, n_samples, n_importance, n_outside,
                 up_sample_steps, perturb):
        if n_outside != 0:
            raise NotImplementedError('n_outside > 0 (render_core_outside) is not on the path: every shipped conf '
                                      'uses n_outside = 0 (confs/nerf.conf:85)')
        self.nerf = nerf
        self.sdf_network = sdf_network
        self.deviation_network = deviation_network
        self.color_network = color_network
        self.n_samples = n_samples
        self.n_importance = n_importance
        self.n_outside = n_outside
        self.up_sample_steps = up_sample_steps
        self.perturb = perturb

    # renderer.py:131-175
    def up_sample(self, rays_o, rays_d, z_vals, sdf, r_limit, n_importance, inv_s):
        return abi.neus_up_sample(rays_o, rays_d, z_vals, sdf, r_limit, n_importance, inv_s)

    # renderer.py:177-191
    def cat_z_vals(self, rays_o, rays_d, z_vals, new_z_vals, sdf, last=False):
        batch_size, n_samples = z_vals.shape
        _, n_importance = new_z_vals.shape
        if last:
            z_out, _ = abi.neus_cat_z_vals(z_vals, new_z_vals)
            return z_out, sdf
        pts = rays_o[:, None, :] + rays_d[:, None, :] * new_z_vals[..., :, None]
        new_sdf = self.sdf_network.sdf(pts.reshape(-1, 3)).reshape(batch_size, n_importance)
        return abi.neus_cat_z_vals(z_vals, new_z_vals, sdf.reshape(batch_size, n_samples), new_sdf)

    # renderer.py:193-297
    def render_core(self, rays_o, rays_d, z_vals, sample_dist, radius, sdf_network, deviation_network,
                    color_network, background_alpha=None, background_sampled_color=None, background_rgb=None,
                    cos_anneal_ratio=0.0, to_light=False, need_color=True, _mid=None):
        # need_color=False (not a reference argument): callers that only read the weights (compute_vis) skip the
        # colour network; every other output is unchanged and 'color' is the composited background only
        if background_alpha is not None or to_light:
            raise NotImplementedError('background model / to_light marching are outside the path (n_outside = 0)')
        batch_size, n_samples = z_vals.shape
        # _mid (not a reference argument): the mid-point positions / directions already produced by the last fused
        # sampling step of render()
        pts, dirs = _mid if _mid is not None else abi.neus_mid_points(rays_o, rays_d, z_vals, float(sample_dist))
        pts = pts.reshape(-1, 3)
        dirs = dirs.reshape(-1, 3)
        if not need_color and hasattr(sdf_network, 'forward_with_gradient'):
            sdf, _, gradients = sdf_network.forward_with_gradient(pts, want_feat=False)
            sampled_color = torch.zeros((batch_size, n_samples, 3), dtype=torch.float32, device=pts.device)
        elif hasattr(sdf_network, 'forward_with_gradient') and hasattr(color_network, 'forward_rows'):
            rows = color_network.alloc_rows(pts.shape[0], pts.device)
            sdf, _, gradients = sdf_network.forward_with_gradient(pts, feat_out=rows)
            sampled_color = color_network.forward_rows(rows, pts, gradients, dirs).reshape(batch_size, n_samples, 3)
        else:
            sdf_nn_output = sdf_network(pts)
            sdf = sdf_nn_output[:, :1]
            feature_vector = sdf_nn_output[:, 1:]
            gradients = sdf_network.gradient(pts).squeeze()
            sampled_color = color_network(pts, gradients, dirs, feature_vector).reshape(batch_size, n_samples, 3)
        if hasattr(deviation_network, 'inv_s'):
            inv_s_f = min(max(deviation_network.inv_s(), 1e-6), 1e6)          # host value: no device sync
            inv_s = torch.full((1, 1), inv_s_f, dtype=torch.float32, device=pts.device)
        else:
            inv_s = deviation_network(torch.zeros([1, 3], device=pts.device))[:, :1].clip(1e-6, 1e6)
            inv_s_f = float(inv_s.reshape(-1)[0])
        o = abi.neus_composite(rays_o, rays_d, z_vals, sdf.detach(), gradients.detach(), sampled_color.detach(),
                               inv_s_f, cos_anneal_ratio, float(sample_dist), float(radius), background_rgb)
        ge = o['grad_err_sums']
        gradient_error = (ge[0] / (ge[1] + 1e-5)).to(torch.float32)
        return {
            'color': o['color'], 'sdf': sdf, 'dists': o['dists'],
            'gradients': gradients.reshape(batch_size, n_samples, 3),
            's_val': 1.0 / inv_s.expand(batch_size * n_samples, 1), 'mid_z_vals': o['mid_z_vals'],
            'weights': o['weights'], 'cdf': o['cdf'], 'gradient_error': gradient_error,
            'inside_sphere': o['inside_sphere'], 'surf': o['surf'], 'depth': o['depth'],
            'weight_sum': o['weight_sum'], 'weight_max': o['weight_max'],
        }

    # renderer.py:299-401
    def render(self, rays_o, rays_d, near, far, radius, perturb_overwrite=-1, background_rgb=None,
               cos_anneal_ratio=0.0, to_light=False, need_color=True):
        if to_light:
            raise NotImplementedError('to_light marching (a per-ray sample_dist, renderer.py:211,302) has no caller in the '
                                      'reference -- gen_geo.compute_vis renders with the default -- and is not built')
        batch_size = len(rays_o)
        dev = rays_o.device
        sample_dist = 2 * radius / self.n_samples
        z_vals = torch.linspace(0.0, 1.0, self.n_samples, device=dev)
        z_vals = (near + (far - near) * z_vals[None, :]).expand(batch_size, self.n_samples).contiguous()
        n_samples = self.n_samples
        perturb = self.perturb
        if perturb_overwrite >= 0:
            perturb = perturb_overwrite
        if perturb > 0:
            t_rand = torch.rand([batch_size, 1], device=dev) - 0.5
            z_vals = z_vals + t_rand * 2. * radius / self.n_samples
        mid = None
        steps = self.up_sample_steps
        n_imp = self.n_importance // steps if steps > 0 else 0
        if (self.n_importance > 0 and steps >= 2 and self.n_samples + steps * n_imp <= 128 and
                type(self).up_sample is NeuSRenderer.up_sample and type(self).cat_z_vals is NeuSRenderer.cat_z_vals):
            # Fused sampling (csrc/neus.cu::neus_scan_step_kernel): one launch per step does cat_z_vals of step i,
            # up_sample of step i + 1 and the positions of its new samples; the last launch also does the final
            # cat_z_vals(last=True) and render_core's mid points.  4 launches instead of 4 + 4 + 1, same values bit for
            # bit (tests/test_gpu_neus.py::test_fused_scan_steps_equal_separate_kernels).
            with torch.no_grad():
                pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]
                sdf = self.sdf_network.sdf(pts.reshape(-1, 3)).reshape(batch_size, self.n_samples)
                new_z, pts = abi.neus_up_sample_pts(rays_o, rays_d, z_vals, sdf, radius, n_imp, 64)
                for i in range(steps - 1):
                    new_sdf = self.sdf_network.sdf(pts.reshape(-1, 3)).reshape(batch_size, n_imp)
                    last = (i + 2 == steps)
                    o = abi.neus_scan_step(rays_o, rays_d, z_vals, new_z, sdf, new_sdf, radius, n_imp, 64 * 2 ** (i + 1),
                                           final_merge=last, sample_dist=sample_dist, want_merged=not last)
                    if last:
                        z_vals, mid = o['z_final'], (o['mid_pts'], o['mid_dirs'])
                    else:
                        z_vals, sdf, new_z, pts = o['z'], o['sdf'], o['new_z'], o['pts']
            n_samples = self.n_samples + self.n_importance
        elif self.n_importance > 0:
            with torch.no_grad():
                pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]
                sdf = self.sdf_network.sdf(pts.reshape(-1, 3)).reshape(batch_size, self.n_samples)
                for i in range(self.up_sample_steps):
                    new_z_vals = self.up_sample(rays_o, rays_d, z_vals, sdf, radius,
                                                self.n_importance // self.up_sample_steps, 64 * 2 ** i)
                    z_vals, sdf = self.cat_z_vals(rays_o, rays_d, z_vals, new_z_vals, sdf,
                                                  last=(i + 1 == self.up_sample_steps))
            n_samples = self.n_samples + self.n_importance
        ret_fine = self.render_core(rays_o, rays_d, z_vals, sample_dist, radius, self.sdf_network,
                                    self.deviation_network, self.color_network, background_rgb=background_rgb,
                                    cos_anneal_ratio=cos_anneal_ratio, need_color=need_color, _mid=mid)
        s_val = ret_fine['s_val'].reshape(batch_size, n_samples).mean(dim=-1, keepdim=True)
        return {
            'color_fine': ret_fine['color'], 's_val': s_val, 'cdf_fine': ret_fine['cdf'],
            'weight_sum': ret_fine['weight_sum'], 'weight_max': ret_fine['weight_max'],
            'gradients': ret_fine['gradients'], 'weights': ret_fine['weights'],
            'gradient_error': ret_fine['gradient_error'], 'inside_sphere': ret_fine['inside_sphere'],
            'surf': ret_fine['surf'], 'depth': ret_fine['depth'],
        }


def render_sharded(renderer: NeuSRenderer, rays_o, rays_d, near, far, radius, group=None, dst=None, **kw):
    """Ray-sharded render (rays are independent given replicated networks): every rank renders the
    contiguous block dist.shard_rows(n_rays, rank, world) of the SAME ray list and the per-ray outputs are gathered
    with ONE collective per output tensor (dist.gather_rows; none at world size 1).  Returns the dict of
    NeuSRenderer.render for all rays (None on ranks != dst when dst is given); 'gradient_error' stays the local
    shard's scalar."""
    import torch.distributed as dist
    from .. import dist as vdist
    n = rays_o.shape[0]
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return renderer.render(rays_o, rays_d, near, far, radius, **kw)
    lo, hi = vdist.shard_rows(n, dist.get_rank(group), dist.get_world_size(group))
    out = renderer.render(rays_o[lo:hi].contiguous(), rays_d[lo:hi].contiguous(), near[lo:hi].contiguous(),
                          far[lo:hi].contiguous(), radius, **kw)
    full = {}
    for k, v in out.items():
        full[k] = v if v.dim() == 0 else vdist.gather_rows(v.contiguous(), n, group=group, dst=dst)
    if dst is not None and dist.get_rank(group) != dst:
        return None
    return full
