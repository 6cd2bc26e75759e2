"""Multi-GPU plumbing for the shading path: one process per GPU, torch.distributed (NCCL over NVLink on
the GPU box, gloo in the CPU tests).

The path shards by surface point: every point is independent given replicated weights,
codebook and probes, so rendering needs exactly ONE collective -- a gather of the per-pixel outputs --
and a training step needs exactly one all-reduce of [gradients | VQ EMA statistics].  Nothing else crosses
GPUs.  The reference's only collective call site is tf.distribute.MirroredStrategy
(decomp/nerfvq_nfr3/nerfactor/trainvali.py:436-446,471,520); its geo stage shards views across independent
processes (geo/NeuS-ours2/gen_geo.py:141-146).
"""
from __future__ import annotations

import os

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_rows(n_total: int, rank: int, world_size: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous pixel-row block [start, stop) of rank `rank` (keeps lvis reads contiguous).  Blocks differ
    by at most `align` rows; `align` = 2 keeps the (pixel, neighbour) pairs of a training batch
    (train_nfr.py:447-448, vq_nfr.py:945-954) on one GPU."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError('bad rank/world_size')
    units = (n_total + align - 1) // align
    base, rem = divmod(units, world_size)
    start_u = rank * base + min(rank, rem)
    stop_u = start_u + base + (1 if rank < rem else 0)
    return min(start_u * align, n_total), min(stop_u * align, n_total)


def shard_sizes(n_total: int, world_size: int, align: int = 1) -> List[int]:
    return [b - a for a, b in (shard_rows(n_total, r, world_size, align) for r in range(world_size))]


def gather_rows(local: torch.Tensor, n_total: int, group=None, align: int = 1,
                dst: Optional[int] = None) -> Optional[torch.Tensor]:
    """The single collective of a pixel-sharded render: concatenates every rank's [n_r, ...] block in rank
    order into [n_total, ...].  dst=None -> all ranks get the image (all_gather); dst=r -> only rank r."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_total, world, align)
    if local.shape[0] != sizes[rank]:
        raise ValueError('rank %d holds %d rows, expected %d' % (rank, local.shape[0], sizes[rank]))
    tail = tuple(local.shape[1:])
    if len(set(sizes)) == 1:
        out = torch.empty((n_total,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out if (dst is None or dst == rank) else None
    # ragged blocks: pad to the largest block, one all_gather, then trim
    mx = max(sizes)
    padded = torch.zeros((mx,) + tail, dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    buf = torch.empty((world * mx,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    if dst is not None and dst != rank:
        return None
    return torch.cat([buf[r * mx:r * mx + sizes[r]] for r in range(world)], dim=0)


def allreduce_flat_(tensors: List[torch.Tensor], group=None) -> None:
    """ONE all-reduce(sum) for a training step: packs [MLP grads | light grad | codebook grad | VQ counts | dw |
    e_latent partial sums] into a flat buffer per dtype, reduces, and scatters back in place."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for dt, ts in by_dtype.items():
        flat = torch.cat([t.reshape(-1) for t in ts])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for t in ts:
            t.copy_(flat[off:off + t.numel()].reshape(t.shape))
            off += t.numel()


def make_stats_allreduce(group=None):
    """Hook for VectorQuantizerEMA.stats_allreduce: global-batch one-hot counts / dw / e_latent sums so that the
    EMA, the `used` mask and the commitment mean equal the single-device global batch."""
    def hook(stats: torch.Tensor) -> None:
        allreduce_flat_([stats], group=group)
    return hook


class PeerImage:
    """Fused gather of a pixel-sharded render: every rank owns image buffers [n_total, c...] in torch symmetric memory
    (P2P-mapped over NVLink / NVSwitch); the shading kernel of rank r stores its rows straight into the destination
    ranks' buffers (vqn_shade_args.peer_rgb), so the "single gather" of SURVEY 8e overlaps the light integral tile by
    tile instead of running as a separate NCCL collective.

    Frame protocol (every rank, on its current stream, per frame):

        img.begin_frame()                      # flips to the other of TWO buffers
        model.fast_render(..., peer_image=img) # live rows -> peers' buffers; background rows of the shard -> zeros
        img.barrier()                          # device-side: all peers' stores have landed
        ... img.tensor is the complete frame on the destination rank(s) until the frame after next begins ...

    Two buffers + one barrier per frame are enough: a peer starts writing frame t+2 into the buffer of frame t only
    after it passed barrier(t+1), which the destination joined -- in stream order -- after its reads of frame t.
    Background (alpha <= 0) rows are written as zeros by their owner (vqn_peer_clear_background), as the reference's
    scatter_nd leaves them, so a frame never shows pixels of an earlier one; both buffers start zero-filled."""

    def __init__(self, n_total: int, tail, device, group=None, dst: Optional[int] = None):
        """dst=None: every rank ends up with the full image (all-gather semantics, world x the NVLink traffic);
        dst=r: only rank r does (gather semantics -- what writing an image needs: the senders push 1/world of the
        image each and rank r's ingress overlaps its own kernels)."""
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.n_total = int(n_total)
        self.tail = tuple(tail)
        self.width = 1
        for t in self.tail:
            self.width *= int(t)
        self.dst = dst
        self._bufs, self._handles, self._ptrs = [], [], []
        for _ in range(2):
            t = symm_mem.empty((self.n_total,) + self.tail, dtype=torch.float32, device=device)
            t.zero_()
            h = symm_mem.rendezvous(t, self.group)
            ptrs = [int(p) for p in h.buffer_ptrs]
            if len(ptrs) != self.world:
                raise RuntimeError('symmetric memory rendezvous returned %d peers for world %d' % (len(ptrs), self.world))
            self._bufs.append(t); self._handles.append(h)
            self._ptrs.append(ptrs if dst is None else [ptrs[int(dst)]])
        self.row0, self.row1 = shard_rows(self.n_total, self.rank, self.world)
        self._frame = 0
        # gather on one rank: per-frame flag hand-shake instead of a barrier over all ranks (vqn_peer_frame_sync)
        self._flags = None
        if dst is not None and self.world > 1 and os.environ.get('VQN_PEER_FLAG_SYNC', '1') != '0':
            self._flags = symm_mem.empty((self.world + 1,), dtype=torch.int32, device=device)
            self._flags.zero_()
            fh = symm_mem.rendezvous(self._flags, self.group)
            self._flag_ptrs = [int(p) for p in fh.buffer_ptrs]
            self._flag_handle = fh
            self._counter = torch.zeros((1,), dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        self._handles[0].barrier()                       # every rank's zero-fill is done before anyone stores

    # -- current frame
    @property
    def tensor(self) -> torch.Tensor:
        return self._bufs[self._frame & 1]

    @property
    def handle(self):
        return self._handles[self._frame & 1]

    @property
    def peer_ptrs(self):
        return self._ptrs[self._frame & 1]

    def begin_frame(self) -> None:
        self._frame += 1

    def clear_background(self, alpha_local: torch.Tensor, row_off: int = 0) -> None:
        """Zeros for this shard's background rows in the destination buffers (called by fast_render; `row_off` = first
        row of `alpha_local` inside the shard, for chunked calls)."""
        from . import abi
        abi.peer_clear_background(alpha_local, self.peer_ptrs, self.row0 + int(row_off), self.width)

    def barrier(self) -> None:
        if self._flags is not None:
            from . import abi
            abi.peer_frame_sync(self._counter, self._flags, self._flag_ptrs, self.world, self.rank, int(self.dst))
        else:
            self.handle.barrier()
