"""Mirror of decomp/nerfvq_nfr3/nerfactor/networks/vq_layers.py:174-349 (VectorQuantizerEMA).

Same constructor arguments, call signature and returned dict keys as the reference.  The Sonnet
ExponentialMovingAverage state (hidden, average, counter for `ema_cluster_size` and `ema_dw`) lives in
torch tensors owned by this object, so `state_dict()` / `load_state_dict()` checkpoint it the way
`tf.train.Checkpoint(net=model)` does in the reference.  No host synchronisation happens in `__call__`
(the reference's `.numpy()` at :318 is not reproduced).
"""
from __future__ import annotations

from typing import Optional

import torch

from ... import abi


class VectorQuantizerEMA:
    def __init__(self, embedding_dim, num_embeddings, commitment_cost, seed=0, decay=0.999, epsilon=1e-5,
                 dtype=torch.float32, name='vector_quantizer_ema', device='cuda'):
        if not 0 <= decay <= 1:
            raise ValueError('decay must be in range [0, 1]')          # vq_layers.py:236-237
        if dtype != torch.float32:
            raise ValueError('only float32 is built')
        self.embedding_dim = int(embedding_dim)
        self.num_embeddings = int(num_embeddings)
        self.decay = float(decay)
        self.commitment_cost = float(commitment_cost)
        self.epsilon = float(epsilon)
        self.name = name
        self.device = torch.device(device)
        self._gen = torch.Generator(device='cpu')
        self._gen.manual_seed(int(seed))
        z, k = self.embedding_dim, self.num_embeddings
        f32 = dict(dtype=torch.float32, device=self.device)
        # ema_cluster_size / ema_dw: initialize(zeros) (:249-255)
        self.state = {
            'cs_hidden': torch.zeros((k,), **f32), 'cs_average': torch.zeros((k,), **f32),
            'dw_hidden': torch.zeros((z, k), **f32), 'dw_average': torch.zeros((z, k), **f32),
            'counters': torch.zeros((2,), dtype=torch.int64, device=self.device),
        }
        # set by the multi-GPU training path: callable(stats float64 tensor) -> None (in-place all-reduce)
        self.stats_allreduce = None

    # -- checkpointing ---------------------------------------------------------------------------
    def state_dict(self):
        return {k: v.clone() for k, v in self.state.items()}

    def load_state_dict(self, sd):
        for k in self.state:
            self.state[k].copy_(sd[k])

    # -- vq_layers.py:257-344 -----------------------------------------------------------------------
    def __call__(self, inputs, codebook, is_training, thres=None, individual=True, roll=None,
                 return_encodings=True, return_distances=True):
        """`roll` (optional, [1,K] or [1,1]) replaces tf.random.uniform(:286-288) for reproducible tests;
        by default it is drawn from this layer's seeded generator."""
        z, k = self.embedding_dim, self.num_embeddings
        if inputs.shape[-1] != z:
            raise ValueError('final dimension of inputs must be embedding_dim=%d' % z)
        flat = inputs.reshape(-1, z)
        sel_mask = None
        if thres is not None:
            thres_t = torch.as_tensor(thres, dtype=torch.float32).reshape(1, -1)
            if roll is None:
                shape = (1, k) if individual else (1, 1)
                roll = torch.rand(shape, generator=self._gen, dtype=torch.float32)
            roll = torch.as_tensor(roll, dtype=torch.float32).reshape(1, -1)
            sel_mask = (roll >= thres_t).to(torch.float32).expand(1, k).reshape(-1).to(flat.device)  # :289
        stats = torch.zeros((abi.vq_stats_size(z, k),), dtype=torch.float64, device=flat.device)
        out = abi.vq_assign(flat, codebook, sel_mask=sel_mask, want_quantize=True,
                            want_distances=return_distances, stats=stats, want_dw=bool(is_training))
        if self.stats_allreduce is not None:
            self.stats_allreduce(stats)       # global-batch statistics before the EMA
        update, loss, perplexity = abi.vq_ema_update(
            stats, codebook, self.decay, self.epsilon, self.commitment_cost, bool(is_training),
            self.state if is_training else None)
        idx = out['indices'].reshape(inputs.shape[:-1])
        ret = {
            'quantize': out['quantize'].reshape(inputs.shape),
            'loss': loss[0],
            'perplexity': perplexity[0],
            'encoding_indices': idx,
            'distances': out['distances'],
        }
        if return_encodings:
            # tf.one_hot(encoding_indices, K) (:293) -- index plumbing, built lazily only when asked for
            ret['encodings'] = torch.nn.functional.one_hot(out['indices'], k).to(torch.float32)
        if is_training:
            ret['update'] = update
        return ret

    def quantize(self, codebook, encoding_indices):
        """vq_layers.py:346-349: embedding_lookup(codebook^T, indices)."""
        return codebook.t()[encoding_indices]
