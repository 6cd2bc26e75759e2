"""Mirror of the training-step half of decomp/nerfvq_nfr3/nerfactor/train_nfr.py (:121-139 optimizer,
:562-576 train_iter): one VQ-stage training iteration = Model.call(mode='train') + compute_loss under a
gradient tape + Adam(amsgrad) apply, here as hand-written forward / backward kernels behind the C ABI.

    optimizer = Adam(learning_rate=5e-4, amsgrad=True, decay_steps=500_000, decay_rate=0.1)
    weighted_loss, partial_to_vis, loss_dict = train_iter(model, batch, optimizer, global_bs, thres=thres)

Data flow of a step (all device-side, no host synchronisation when every row is foreground):
  embed -> fine_enc -> bottleneck -> z_enc -> {VQ assign (+ one-hot / dw statistics), main heads -> shade,
  VQ heads(z_vq) -> shade} -> loss -> backward of each in reverse -> ONE all-reduce of
  [all gradients | VQ statistics | loss sums] (NCCL over NVLink; single GPU: none) -> Sonnet EMA codebook update
  with the GLOBAL statistics -> codebook separation loss on the updated codebook -> Adam on the flat buffer.
All trainable tensors (8 MLPs, _light, _codebook) are views into ONE flat fp32 buffer, and so are their
gradients, so the collective and the optimizer are each a single launch.
"""
from __future__ import annotations

import os

import math
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from .. import _lib as L
from .. import abi

F32 = torch.float32
NET_ORDER = ('fine_enc', 'bottleneck', 'diff_main', 'spec_main', 'rough_main', 'diff_vq', 'spec_vq', 'rough_vq')


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


class Adam:
    """tf.keras.optimizers.Adam(learning_rate, amsgrad=True) (train_nfr.py:121-139) with the optional
    ExponentialDecay schedule lr * rate ** (step / decay_steps) (:123-127)."""

    def __init__(self, learning_rate: float = 5e-4, beta_1: float = 0.9, beta_2: float = 0.999,
                 epsilon: float = 1e-7, amsgrad: bool = True, decay_steps: int = -1, decay_rate: float = 0.1):
        if not amsgrad:
            raise NotImplementedError('the reference trains with amsgrad=True (train_nfr.py:129)')
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self.decay_steps, self.decay_rate = decay_steps, decay_rate
        self.iterations = 0
        self.m = self.v = self.vhat = None

    def lr_t(self) -> float:
        """called after `iterations` was incremented for this step"""
        t = self.iterations
        lr = self.learning_rate
        if self.decay_steps > 0:
            lr = lr * self.decay_rate ** ((t - 1) / self.decay_steps)   # schedule sees the pre-increment step
        return lr * math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)

    def apply_gradients(self, params: torch.Tensor, grads: torch.Tensor, lr_t_dev=None) -> None:
        """lr_t_dev: device scalar holding lr_t (graph replay: the caller advances `iterations` and refreshes it)"""
        self.ensure_state(params)
        if lr_t_dev is None:
            self.iterations += 1
        abi.adam_amsgrad(params, grads, self.m, self.v, self.vhat, self.lr_t() if lr_t_dev is None else 0.0,
                         self.beta_1, self.beta_2, self.epsilon, lr_t_dev=lr_t_dev)

    def ensure_state(self, params: torch.Tensor) -> None:
        if self.m is None:
            self.m, self.v, self.vhat = (torch.zeros_like(params) for _ in range(3))

    def state_dict(self):
        return {'iterations': self.iterations, 'm': self.m, 'v': self.v, 'vhat': self.vhat}


# Training forward: one fused tensor-core launch per network (default) or one Dense kernel per layer
# (VQN_TRAIN_FUSED_FORWARD=0; the per-layer kernels remain the backward path either way)
FUSED_FORWARD = os.environ.get('VQN_TRAIN_FUSED_FORWARD', '1') != '0'
# backward GEMMs batched: the six heads level by level and all weight gradients in one launch (VQN_TRAIN_BATCHED_BACKWARD=0:
# one launch per layer, the per-network order of round 1)
BATCHED_BACKWARD = os.environ.get('VQN_TRAIN_BATCHED_BACKWARD', '1') != '0'
CONCURRENT_HEADS = os.environ.get('VQN_TRAIN_CONCURRENT_HEADS', '1') != '0'
# backward-data chain of a network as ONE launch of the fused tcgen05 kernel (vqn_net_backward_train); 0: batched per-level GEMMs
FUSED_BACKWARD = os.environ.get('VQN_TRAIN_FUSED_BACKWARD', '1') != '0'


class _NetTrain:
    """Activations and gradient scratch of one mlp.Network for a fixed row count (networks/mlp.py:39-50)."""

    def __init__(self, net, in_dim: int, n: int, out_scale: float = 1.0, out_bias: float = 0.0):
        self.net, self.in_dim, self.n = net, in_dim, n
        self.out_scale, self.out_bias = out_scale, out_bias
        self.acts = [L.act_code(a) for a in net.act]
        self.skip = None if net.skip_at is None else int(net.skip_at[0])
        dev = net.kernels[0].device
        self.widths = list(net.widths)
        self.k_in: List[int] = []          # input width of layer i
        self.ld: List[int] = []            # leading dim of the output buffer of layer i
        d = in_dim
        for i, w in enumerate(self.widths):
            self.k_in.append(d)
            d = w + (in_dim if self.skip == i else 0)
            self.ld.append(_pad4(d))
        self.y = [torch.zeros((n, ld), dtype=F32, device=dev) for ld in self.ld]
        self.dz = [torch.empty((n, _pad4(w)), dtype=F32, device=dev) for w in self.widths]
        self.x = None
        self.ldx = None

    @property
    def out(self) -> torch.Tensor:
        return self.y[-1]

    def concat_job(self, x: torch.Tensor, ldx: int):
        """The x half of concat(y, x) (mlp.py:47-48) as a job of abi.copy_cols_batched (None without a skip connection)."""
        if self.skip is None:
            return None
        return (x, ldx, self.y[self.skip], self.ld[self.skip], self.n, self.in_dim, self.widths[self.skip])

    def forward(self, x: torch.Tensor, ldx: int, prepared: bool = False) -> torch.Tensor:
        """prepared: the caller has already refreshed the weight images (abi.nets_repack_tc) and copied the x half of the
        skip concat (concat_job) for several networks in single launches."""
        self.x, self.ldx = x, ldx
        if FUSED_FORWARD and ldx % 4 == 0:
            # the whole network in ONE launch of the fused tcgen05 kernel (every layer's output stored for the backward
            # pass); the pre-split weight images are refreshed first because the optimizer has moved the weights
            packed = self.net.packed
            if not prepared:
                packed.repack_tc('tf32x3')
            packed.forward_train(x, ldx, self.n, self.y, self.ld, self.out_scale, self.out_bias, 'tf32x3')
            if self.skip is not None and not prepared:                    # concat(y, x) (mlp.py:47-48)
                abi.copy_cols(x, ldx, self.y[self.skip], self.ld[self.skip], self.n, self.in_dim,
                              dst_off=self.widths[self.skip])
            return self.y[-1]
        cur, ld = x, ldx
        n_layers = len(self.widths)
        for i, w in enumerate(self.widths):
            last = i == n_layers - 1
            abi.dense_forward(cur, ld, self.net.kernels[i], self.net.biases[i], self.y[i], self.ld[i], self.n,
                              self.k_in[i], w, self.acts[i], self.out_scale if last else 1.0,
                              self.out_bias if last else 0.0)
            if self.skip == i:                                            # concat(y, x) (mlp.py:47-48)
                abi.copy_cols(x, ldx, self.y[i], self.ld[i], self.n, self.in_dim, dst_off=w)
            cur, ld = self.y[i], self.ld[i]
        return self.y[-1]

    def weight_problems(self, dW: List[torch.Tensor], dB: List[torch.Tensor], split_skip: bool = False) -> list:
        """The weight-gradient GEMMs of every layer (dW_i += x_i^T dz_i, dB_i += colsum dz_i) as problems of ONE batched
        launch; valid once the backward-data chain has filled every dz."""
        out = []
        for i, w in enumerate(self.widths):
            xin, ldin = (self.x, self.ldx) if i == 0 else (self.y[i - 1], self.ld[i - 1])
            if split_skip and self.skip is not None and i == self.skip + 1:
                # the layer after the skip concat as TWO problems -- rows [0, w_prev) of dW from y, the rest from the network
                # input itself: nobody has to copy x behind y (three heads of a branch share one x: it stays in L2)
                wp = self.widths[i - 1]
                out.append(abi.bwd_weights_problem(xin, ldin, self.dz[i], self.dz[i].shape[1], dW[i], dB[i], self.n, wp, w))
                out.append(abi.bwd_weights_problem(self.x, self.ldx, self.dz[i], self.dz[i].shape[1], dW[i], None, self.n,
                                                   self.in_dim, w, dw_off=wp * w))
                continue
            out.append(abi.bwd_weights_problem(xin, ldin, self.dz[i], self.dz[i].shape[1], dW[i], dB[i], self.n,
                                               self.k_in[i], w))
        return out

    def data_problems(self, i: int, d_input: Optional[torch.Tensor], ld_din: int, atomic: bool) -> list:
        """Backward-data GEMMs out of layer i's dz (dz[i] must be complete): into dz[i-1] through the previous activation,
        and the x part of a skip concat / the network input into d_input (added; atomically when `atomic`)."""
        w = self.widths[i]
        dz, lddz = self.dz[i], self.dz[i].shape[1]
        acc = 2 if atomic else 1
        out = []
        if i > 0:
            wp = self.widths[i - 1]
            out.append(abi.bwd_data_problem(dz, lddz, self.net.kernels[i], self.dz[i - 1], self.dz[i - 1].shape[1],
                                            self.y[i - 1], self.ld[i - 1], self.acts[i - 1], 0, self.n, wp, w))
            if self.skip == i - 1 and d_input is not None:
                out.append(abi.bwd_data_problem(dz, lddz, self.net.kernels[i], d_input, ld_din, None, 0, L.ACT_NONE, acc,
                                                self.n, self.in_dim, w, w_row0=wp))
        elif d_input is not None:
            out.append(abi.bwd_data_problem(dz, lddz, self.net.kernels[0], d_input, ld_din, None, 0, L.ACT_NONE, acc,
                                            self.n, self.in_dim, w))
        return out

    def act_job(self, dy: torch.Tensor, lddy: int) -> tuple:
        """The arguments of abi.act_backward for the last layer (dz[last] = dy * act'(y[last]) * albedo_slope)."""
        last = len(self.widths) - 1
        return (dy, lddy, self.y[last], self.ld[last], self.n, self.widths[last], self.acts[last],
                self.out_scale, self.out_scale, self.out_bias, self.dz[last], self.dz[last].shape[1])

    def act_backward_last(self, dy: torch.Tensor, lddy: int) -> None:
        abi.act_backward(*self.act_job(dy, lddy))

    def backward_fused(self, dy: torch.Tensor, lddy: int, d_input: Optional[torch.Tensor], ld_din: int, din_mode: int,
                       act_done: bool = False, din_y: Optional[torch.Tensor] = None, ld_din_y: int = 0,
                       din_act: int = 0) -> None:
        """The backward-DATA chain of the network as ONE launch of the fused tcgen05 kernel (vqn_net_backward_train): fills
        every dz[i]; the caller batches the weight-gradient GEMMs (weight_problems).  act_done: dz[last] was already formed
        (the heads' last-layer activation gradients share one launch)."""
        last = len(self.widths) - 1
        if not act_done:
            self.act_backward_last(dy, lddy)
        self.net.packed.backward_train(self.dz[last], self.dz[last].shape[1], self.n, self.y, self.ld, self.dz,
                                       [t.shape[1] for t in self.dz], d_input, ld_din, din_mode, din_y, ld_din_y, din_act)

    def backward(self, dy: torch.Tensor, lddy: int, dW: List[torch.Tensor], dB: List[torch.Tensor],
                 d_input: Optional[torch.Tensor] = None, ld_din: int = 0, weight_list: Optional[list] = None) -> None:
        """dy: gradient w.r.t. the net output [n, out].  Adds weight gradients into dW/dB and (optionally)
        ACCUMULATES the gradient w.r.t. the net input into d_input.  weight_list: defer the weight-gradient GEMMs --
        their problems are appended to the list and the caller launches them in one batch."""
        nl = len(self.widths)
        last = nl - 1
        self.act_backward_last(dy, lddy)
        if weight_list is not None:
            weight_list += self.weight_problems(dW, dB)
        for i in range(last, -1, -1):
            w = self.widths[i]
            dz, lddz = self.dz[i], self.dz[i].shape[1]
            xin, ldin = (self.x, self.ldx) if i == 0 else (self.y[i - 1], self.ld[i - 1])
            if weight_list is None:
                abi.dense_backward_weights(xin, ldin, dz, lddz, dW[i], dB[i], self.n, self.k_in[i], w)
            if i > 0:
                wp = self.widths[i - 1]
                # y part of the input: through the previous layer's activation -> dz[i-1]
                abi.dense_backward_data(dz, lddz, self.net.kernels[i], self.dz[i - 1], self.dz[i - 1].shape[1],
                                        self.y[i - 1], self.ld[i - 1], self.acts[i - 1], False, self.n, wp, w)
                if self.skip == i - 1 and d_input is not None:            # x part of concat(y, x)
                    abi.dense_backward_data(dz, lddz, self.net.kernels[i], d_input, ld_din, None, 0, L.ACT_NONE,
                                            True, self.n, self.in_dim, w, w_row0=wp)
            elif d_input is not None:
                abi.dense_backward_data(dz, lddz, self.net.kernels[0], d_input, ld_din, None, 0, L.ACT_NONE, True,
                                        self.n, self.in_dim, w)


class TrainState:
    """Flat parameter / gradient buffers (views re-pointed into the model) + per-row-count activation sets."""

    def __init__(self, model):
        self.model = model
        dev = model.device
        tensors = []
        for name in NET_ORDER:
            net = model.net[name]
            for k, b in zip(net.kernels, net.biases):
                tensors += [k, b]
        tensors += [model._light, model._codebook]
        self.has_gamma = model.data_type != 'nerf'               # learnable tone scaling [_gamma_bias, _gamma_index] (:736-745)
        if self.has_gamma:
            tensors += [torch.cat([model._gamma_bias.reshape(1), model._gamma_index.reshape(1)]).to(dev)]
        sizes = [_pad4(t.numel()) for t in tensors]              # 16-byte aligned views
        self.n_param = sum(sizes)
        z, k = model.z_dim, model.num_embed
        self.n_stats = abi.vq_stats_size(z, k)
        self.n_sums = 8                                           # loss sums [6] + active rows + spare
        # flat = [params]; gflat = [grads | stats (fp32) | sums]
        self.params = torch.zeros((self.n_param,), dtype=F32, device=dev)
        self.gflat = torch.zeros((self.n_param + _pad4(self.n_stats) + self.n_sums,), dtype=F32, device=dev)
        self.grads = self.gflat[:self.n_param]
        self.stats32 = self.gflat[self.n_param:self.n_param + self.n_stats]
        self.sums = self.gflat[self.n_param + _pad4(self.n_stats):]
        self.stats64 = torch.zeros((self.n_stats,), dtype=torch.float64, device=dev)
        off = 0
        views, gviews = [], []
        for t, sz in zip(tensors, sizes):
            v = self.params[off:off + t.numel()].view(t.shape)
            v.copy_(t)
            views.append(v)
            gviews.append(self.grads[off:off + t.numel()].view(t.shape))
            off += sz
        it, git = iter(views), iter(gviews)
        self.dW: Dict[str, List[torch.Tensor]] = {}
        self.dB: Dict[str, List[torch.Tensor]] = {}
        for name in NET_ORDER:
            net = model.net[name]
            self.dW[name], self.dB[name] = [], []
            for i in range(len(net.kernels)):
                net.kernels[i] = next(it)
                net.biases[i] = next(it)
                self.dW[name].append(next(git))
                self.dB[name].append(next(git))
            net._pack()                                           # the inference handles follow the new storage
        model._light = next(it)
        model._codebook = next(it)
        self.d_light = next(git)
        self.d_codebook = next(git)
        if self.has_gamma:
            self.gamma_par = next(it)                             # device copy the kernels read; the model's views follow
            self.d_gamma = next(git)
            model._gamma_bias, model._gamma_index = self.gamma_par[0:1], self.gamma_par[1:2]
        self.sim_loss = torch.zeros((1,), dtype=F32, device=dev)
        self.scalars = torch.zeros((4,), dtype=F32, device=dev)
        self.acts: Dict[int, dict] = {}
        self.dirty = False
        self.side_streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]   # forked head forwards

    def buffers(self, n: int) -> dict:
        if n in self.acts:
            return self.acts[n]
        m = self.model
        dev = m.device
        z = m.z_dim
        emb = m.embedder['xyz'].out_dims
        e = lambda *s: torch.empty(s, dtype=F32, device=dev)
        b = {
            'embed': torch.zeros((n, _pad4(emb)), dtype=F32, device=dev),
            'nets': {
                'fine_enc': _NetTrain(m.net['fine_enc'], emb, n),
                'bottleneck': _NetTrain(m.net['bottleneck'], m.net['fine_enc'].widths[-1], n),
                'diff_main': _NetTrain(m.net['diff_main'], z, n, m.albedo_slope, m.albedo_bias),
                'spec_main': _NetTrain(m.net['spec_main'], z, n),
                'rough_main': _NetTrain(m.net['rough_main'], z, n),
                'diff_vq': _NetTrain(m.net['diff_vq'], z, n, m.albedo_slope, m.albedo_bias),
                'spec_vq': _NetTrain(m.net['spec_vq'], z, n),
                'rough_vq': _NetTrain(m.net['rough_vq'], z, n),
            },
            'albedo': e(n, 3), 'spec': e(n, 3), 'base_c': e(n, 3), 'ks_c': e(n, 1), 'rough_c': e(n, 1),
            'vq_albedo_c': e(n, 3), 'vq_spec_c': e(n, 3), 'vq_rough_c': e(n, 1),
            'loss_rows': e(n), 'd_rgb': e(n, 3), 'd_vqrgb': e(n, 3), 'd_zvq': e(n, z), 'd_spec_l': e(n, 3),
            'd_albedo': e(n, 3), 'd_spec': e(n, 3), 'd_rough': e(n, 1), 'd_base': e(n, 3), 'd_ks': e(n, 1),
            'd_vq_albedo': e(n, 3), 'd_vq_spec': e(n, 3), 'd_vq_rough': e(n, 1),
            'd_zenc': e(n, z), 'd_h': e(n, m.net['fine_enc'].widths[-1]),
            'rgb_g': e(n, 3), 'vqrgb_g': e(n, 3), 'd_rgb_lin': e(n, 3), 'd_vqrgb_lin': e(n, 3),   # non-'nerf' data: gamma
        }
        self.acts[n] = b
        return b


def _train_state(model) -> TrainState:
    st = getattr(model, '_train_state', None)
    if st is None:
        st = TrainState(model)
        model._train_state = st
    return st


LOSS_DEFAULTS = dict(vq_loss_weight=1.0, chr_alpha=60.0, chr_thres=0.1, combine_weight=0.2, mat_sloss_weight=0.05,
                     chromaticity_loss_weight=1.0, sim_loss_weight=1e-4, lambert_weight=1e-3)   # vq_nfr.ini:115-131


def _compact_col(src: torch.Tensor, ld: int, w: int, dst: torch.Tensor):
    abi.copy_cols(src, ld, dst, w, src.shape[0], w)
    return dst


def train_iter(model, batch, optimizer: Adam, global_bs: int, thres=None, roll=None, group=None,
               apply: bool = True, _sel_mask_dev=None, _lr_t_dev=None):
    """train_nfr.py:562-576.  Returns (weighted_loss [device scalar], partial_to_vis, loss_dict of per-term
    batch means).  `group`: torch.distributed process group for the data-parallel all-reduce (None: default
    group when initialised, else single GPU).  Rows must be (pixel, neighbour) pairs (train_nfr.py:447-448)."""
    m = model
    st = _train_state(m)
    cfg = m.config
    lw = {k: (cfg.getfloat(k, fallback=v) if hasattr(cfg, 'getfloat') else v) for k, v in LOSS_DEFAULTS.items()}
    id_, hw, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, lvis = m._unpack(batch, False)
    n_total = alpha.shape[0]
    full = getattr(m, 'assume_all_foreground', False)
    if full:
        n = n_total
    else:
        row_idx, n_act = abi.compact_mask(alpha)
        n = int(n_act.item())
    if n != n_total:                     # the sampler only draws foreground pixels (train_nfr.py:380-467); rare path
        idx = row_idx[:n].long()
        rayo, rgb, xyz, normal = (t.index_select(0, idx) for t in (rayo, rgb, xyz, normal))
        lvis = lvis.index_select(0, idx) if lvis is not None else None
    if n % 2:
        raise ValueError('training rows must come in (pixel, neighbour) pairs')
    B = st.buffers(n)
    nets = B['nets']
    z = m.z_dim
    K = m.num_embed
    inv_gbs = 1.0 / float(global_bs)
    fused_bwd = FUSED_FORWARD and BATCHED_BACKWARD and FUSED_BACKWARD
    prep = FUSED_FORWARD and BATCHED_BACKWARD
    pack_done = None
    if prep and CONCURRENT_HEADS:
        # ONE refresh of all weight images (17 us), on a side stream under the clears / the embedding / the input copies
        cur = torch.cuda.current_stream(m.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        st.side_streams[0].wait_event(ev)
        with torch.cuda.stream(st.side_streams[0]):
            abi.nets_repack_tc([nets[name].net.packed for name in NET_ORDER], 'tf32x3')
            pack_done = torch.cuda.Event()
            pack_done.record(st.side_streams[0])
    # everything the step accumulates into, cleared in one launch (d_h is stored, not accumulated, by the fused backward)
    abi.zero_batched([st.gflat, st.stats64, B['d_zenc']] + ([] if fused_bwd else [B['d_h']]), m.device)

    # ---- forward ------------------------------------------------------------------------------------------
    emb = m.embedder['xyz']
    E = B['embed']
    c = L.Context.get(m.device)
    xyz_c = xyz.contiguous()
    e_tmp = None if fused_bwd else abi.embed(xyz_c, emb.n_freqs)
    # launches shared by several networks (the step is launch-latency-bound at 8192 rows): ONE refresh of all weight images,
    # the x halves of the skip concats of the networks fed from the same input in one copy launch
    if prep:
        if pack_done is None:
            abi.nets_repack_tc([nets[name].net.packed for name in NET_ORDER], 'tf32x3')
        if fused_bwd:
            # straight into the padded input buffer; the fused backward reads x itself (weight_problems(split_skip=True)),
            # so nobody needs the copy of x behind fine_enc's skip layer
            abi.embed(xyz_c, emb.n_freqs, out=E)
        else:
            abi.copy_cols_batched([(e_tmp, emb.out_dims, E, E.shape[1], n, emb.out_dims, 0),
                                   (e_tmp, emb.out_dims) + nets['fine_enc'].concat_job(E, E.shape[1])[2:]], m.device)
        if pack_done is not None:
            torch.cuda.current_stream(m.device).wait_event(pack_done)
    else:
        abi.copy_cols(e_tmp, emb.out_dims, E, E.shape[1], n, emb.out_dims)
    h = nets['fine_enc'].forward(E, E.shape[1], prepared=prep)
    z_enc = nets['bottleneck'].forward(h, nets['fine_enc'].ld[-1], prepared=prep)
    codebook = m.get_codebook()                                                   # :576, before the EMA assign
    sel_mask = _sel_mask_dev           # graph replay: a static device mask refreshed by the caller
    th = m._thres_mask(thres) if _sel_mask_dev is None else None
    if th is not None:
        vq = m.vq_layer
        if roll is None:
            roll = torch.rand((1, K), generator=vq._gen, dtype=F32)
        roll = torch.as_tensor(roll, dtype=F32).reshape(1, -1)
        sel_mask = (roll >= torch.as_tensor(th, dtype=F32).reshape(1, -1)).to(F32).expand(1, K).reshape(-1).to(m.device)
    vq_out = abi.vq_assign(z_enc, codebook, sel_mask=sel_mask, normalize_inputs=True, want_quantize=True,
                           stats=st.stats64, want_dw=True)                        # :575-577 (l2_normalize fused)
    z_vq, idx = vq_out['quantize'], vq_out['indices']
    if prep and not fused_bwd:
        abi.copy_cols_batched([nets[k].concat_job(z_enc, z) for k in ('diff_main', 'spec_main', 'rough_main')] +
                              [nets[k].concat_job(z_vq, z) for k in ('diff_vq', 'spec_vq', 'rough_vq')], m.device)
    if prep and CONCURRENT_HEADS:
        # The six head networks are independent given z_enc / z_vq, and one fused forward launch occupies only 64 of the 148
        # SMs at 8192 rows (one 128-row tile per CTA): fork them over three streams (the fork / join events become edges of
        # the captured graph).  Nothing is allocated on the side streams -- every buffer is a preallocated one.
        cur = torch.cuda.current_stream(m.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        outs = {}
        groups = (('diff_main', 'rough_vq'), ('spec_main', 'diff_vq'), ('rough_main', 'spec_vq'))
        for si, names in enumerate(groups):
            side = cur if si == 0 else st.side_streams[si - 1]
            if si > 0:
                side.wait_event(ev)
            with torch.cuda.stream(side):
                for k in names:
                    outs[k] = nets[k].forward(z_vq if k.endswith('_vq') else z_enc, z, prepared=True)
            if si > 0:
                done = torch.cuda.Event()
                done.record(side)
                cur.wait_event(done)
        base, ks, rough = outs['diff_main'], outs['spec_main'], outs['rough_main']
    else:
        base = nets['diff_main'].forward(z_enc, z, prepared=prep)
        ks = nets['spec_main'].forward(z_enc, z, prepared=prep)
        rough = nets['rough_main'].forward(z_enc, z, prepared=prep)
    all_heads_done = prep and CONCURRENT_HEADS
    if prep:
        base_c, ks_c, rough_c = B['base_c'], B['ks_c'], B['rough_c']
        jobs = [(base, nets['diff_main'].ld[-1], base_c, 3, n, 3, 0), (ks, nets['spec_main'].ld[-1], ks_c, 1, n, 1, 0),
                (rough, nets['rough_main'].ld[-1], rough_c, 1, n, 1, 0)]
        if all_heads_done:               # the VQ branch's outputs exist already: compact all six in this launch
            va, vs, vr = outs['diff_vq'], outs['spec_vq'], outs['rough_vq']
            va_c, vs_c, vr_c = B['vq_albedo_c'], B['vq_spec_c'], B['vq_rough_c']
            jobs += [(va, nets['diff_vq'].ld[-1], va_c, 3, n, 3, 0), (vs, nets['spec_vq'].ld[-1], vs_c, 3, n, 3, 0),
                     (vr, nets['rough_vq'].ld[-1], vr_c, 1, n, 1, 0)]
        abi.copy_cols_batched(jobs, m.device)
    else:
        base_c = _compact_col(base, nets['diff_main'].ld[-1], 3, B['base_c'])
        ks_c = _compact_col(ks, nets['spec_main'].ld[-1], 1, B['ks_c'])
        rough_c = _compact_col(rough, nets['rough_main'].ld[-1], 1, B['rough_c'])
    albedo, spec, _, _ = abi.material_combine(base_c, ks_c, want_scaled=False)    # :590-591
    lights = m._light.reshape(1, 512, 3)
    is_nerf = m.data_type == 'nerf'
    # non-'nerf' data: rgb = clip((rgb * gamma_bias) ^ gamma_index) with TRAINABLE gamma (:715-718): the integral comes out
    # raw and the tone scaling is its own pair of kernels reading the parameters from the flat device buffer
    sh = abi.shade(xyz, rayo, normal, lvis, albedo, spec, rough_c, m.lxyz, m.lareas, lights, n=n, n_total=n,
                   no_clip=not is_nerf)
    rgb_lin = sh['rgb'].reshape(n, 3)
    rgb_pred = rgb_lin if is_nerf else abi.gamma_forward(rgb_lin, st.gamma_par, B['rgb_g'])
    if not all_heads_done:
        va = nets['diff_vq'].forward(z_vq, z, prepared=prep)
        vs = nets['spec_vq'].forward(z_vq, z, prepared=prep)
        vr = nets['rough_vq'].forward(z_vq, z, prepared=prep)
    if all_heads_done:
        pass
    elif prep:
        va_c, vs_c, vr_c = B['vq_albedo_c'], B['vq_spec_c'], B['vq_rough_c']
        abi.copy_cols_batched([(va, nets['diff_vq'].ld[-1], va_c, 3, n, 3, 0),
                               (vs, nets['spec_vq'].ld[-1], vs_c, 3, n, 3, 0),
                               (vr, nets['rough_vq'].ld[-1], vr_c, 1, n, 1, 0)], m.device)
    else:
        va_c = _compact_col(va, nets['diff_vq'].ld[-1], 3, B['vq_albedo_c'])
        vs_c = _compact_col(vs, nets['spec_vq'].ld[-1], 3, B['vq_spec_c'])
        vr_c = _compact_col(vr, nets['rough_vq'].ld[-1], 1, B['vq_rough_c'])
    sh_vq = abi.shade(xyz, rayo, normal, lvis, va_c, vs_c, vr_c, m.lxyz, m.lareas, lights, n=n, n_total=n,
                      no_clip=not is_nerf)
    vq_rgb_lin = sh_vq['rgb'].reshape(n, 3)
    vq_rgb = vq_rgb_lin if is_nerf else abi.gamma_forward(vq_rgb_lin, st.gamma_par, B['vqrgb_g'])

    # ---- loss + backward ----------------------------------------------------------------------------------
    abi.loss_train(rgb.contiguous(), rgb_pred, vq_rgb, z_vq, spec, rough_c, is_nerf, lw['combine_weight'],
                   lw['chromaticity_loss_weight'], lw['mat_sloss_weight'], lw['lambert_weight'], lw['chr_alpha'],
                   lw['chr_thres'], inv_gbs, B['loss_rows'], B['d_rgb'], B['d_vqrgb'], B['d_zvq'], B['d_spec_l'],
                   st.sums)
    d_rgb, d_vqrgb = B['d_rgb'], B['d_vqrgb']
    if not is_nerf:                      # back through the tone scaling: d rgb_lin, and the gradients of the two parameters
        abi.gamma_backward(rgb_lin, st.gamma_par, B['d_rgb'], B['d_rgb_lin'], st.d_gamma)
        abi.gamma_backward(vq_rgb_lin, st.gamma_par, B['d_vqrgb'], B['d_vqrgb_lin'], st.d_gamma)
        d_rgb, d_vqrgb = B['d_rgb_lin'], B['d_vqrgb_lin']
    # main branch
    abi.shade_backward(xyz, rayo, normal, lvis, albedo, spec, rough_c, m.lxyz, m.lareas, m._light, d_rgb,
                       B['d_albedo'], B['d_spec'], B['d_rough'], st.d_light)
    abi.material_combine_backward(base_c, ks_c, B['d_albedo'], B['d_spec'], B['d_spec_l'], B['d_base'], B['d_ks'])
    d_zenc = B['d_zenc']
    # VQ branch (d_zvq already holds the pair-smoothness gradient)
    abi.shade_backward(xyz, rayo, normal, lvis, va_c, vs_c, vr_c, m.lxyz, m.lareas, m._light, d_vqrgb,
                       B['d_vq_albedo'], B['d_vq_spec'], B['d_vq_rough'], st.d_light)
    d_zvq = B['d_zvq']
    heads = [('diff_main', B['d_base'], 3, d_zenc), ('spec_main', B['d_ks'], 1, d_zenc),
             ('rough_main', B['d_rough'], 1, d_zenc), ('diff_vq', B['d_vq_albedo'], 3, d_zvq),
             ('spec_vq', B['d_vq_spec'], 3, d_zvq), ('rough_vq', B['d_vq_rough'], 1, d_zvq)]
    wlist: list = []                       # every weight-gradient GEMM of the step: ONE batched launch at the end
    if BATCHED_BACKWARD and FUSED_BACKWARD and prep:
        # every head's backward-data chain is ONE launch of the fused tcgen05 kernel (transposed weight images, dz_i stored for
        # the weight gradients, the narrow last layer backwards on the CUDA cores); the six launches are forked over three
        # streams like the forwards; the three heads of a branch add into the same d_z atomically
        abi.act_backward_batched([nets[name].act_job(dy, lddy) for name, dy, lddy, _ in heads], m.device)
        cur = torch.cuda.current_stream(m.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        groups = (heads[0::3], heads[1::3], heads[2::3])
        for si, grp_ in enumerate(groups):
            side = cur if si == 0 else st.side_streams[si - 1]
            if si > 0:
                side.wait_event(ev)
            with torch.cuda.stream(side):
                for name, dy, lddy, dzin in grp_:
                    nets[name].backward_fused(dy, lddy, dzin, z, 2, act_done=True)
            if si > 0:
                done = torch.cuda.Event()
                done.record(side)
                cur.wait_event(done)
    elif BATCHED_BACKWARD:
        # the six heads level by level: the same-level backward-data GEMMs of all heads are independent -> one launch per
        # level (3 launches instead of 24); the three heads of a branch add into the same d_z atomically
        for name, dy, lddy, _ in heads:
            nets[name].act_backward_last(dy, lddy)
        for level in (2, 1, 0):
            probs = []
            for name, _, _, dzin in heads:
                probs += nets[name].data_problems(level, dzin, z, atomic=True)
            abi.dense_backward_data_batched(probs, m.device)
    if BATCHED_BACKWARD:
        head_w = []
        for name, _, _, _ in heads:
            head_w += nets[name].weight_problems(st.dW[name], st.dB[name], split_skip=fused_bwd)
        if CONCURRENT_HEADS and prep:
            # the heads' 18 weight-gradient GEMMs (one full wave of CTAs) run on a side stream UNDER the encoder's backward
            # chain, whose per-layer launches fill less than half of the SMs
            cur = torch.cuda.current_stream(m.device)
            ev = torch.cuda.Event()
            ev.record(cur)
            side = st.side_streams[0]
            side.wait_event(ev)
            with torch.cuda.stream(side):
                abi.dense_backward_weights_batched(head_w, m.device)
                heads_w_done = torch.cuda.Event()
                heads_w_done.record(side)
        else:
            wlist += head_w
            heads_w_done = None
    else:
        heads_w_done = None
        for name, dy, lddy, dzin in heads:
            nets[name].backward(dy, lddy, st.dW[name], st.dB[name], dzin, z)
    commit_coef = lw['vq_loss_weight'] * m.vq_layer.commitment_cost * 2.0 * inv_gbs / z
    bn, fe = nets['bottleneck'], nets['fine_enc']
    if fused_bwd:
        # ... and, z_enc being the bottleneck's activated output, the dz of its last layer in the same pass
        abi.vq_backward(z_enc, idx, codebook, d_zvq, commit_coef, d_zenc, accumulate=True, act=bn.acts[-1], dz_out=bn.dz[-1])
    else:
        abi.vq_backward(z_enc, idx, codebook, d_zvq, commit_coef, d_zenc, accumulate=True)
    d_h = B['d_h']
    wl = wlist if BATCHED_BACKWARD else None
    if fused_bwd:
        # the bottleneck's chain ends in fine_enc's last dz (its input IS fine_enc's activated output): no d_h round trip
        bn.backward_fused(d_zenc, z, fe.dz[-1], fe.dz[-1].shape[1], 0, act_done=True, din_y=fe.y[-1], ld_din_y=fe.ld[-1],
                          din_act=fe.acts[-1])
        fe.backward_fused(None, 0, None, 0, 0, act_done=True)
        wlist += bn.weight_problems(st.dW['bottleneck'], st.dB['bottleneck'], split_skip=True)
        wlist += fe.weight_problems(st.dW['fine_enc'], st.dB['fine_enc'], split_skip=True)
    else:
        bn.backward(d_zenc, z, st.dW['bottleneck'], st.dB['bottleneck'], d_h, d_h.shape[1], weight_list=wl)
        fe.backward(d_h, d_h.shape[1], st.dW['fine_enc'], st.dB['fine_enc'], weight_list=wl)
    if wlist:
        abi.dense_backward_weights_batched(wlist, m.device)
    if heads_w_done is not None:
        torch.cuda.current_stream(m.device).wait_event(heads_w_done)

    # ---- the single collective: [gradients | VQ statistics | loss sums] --------------------------------------
    abi.train_pack_stats(st.stats64, st.stats32, st.sums[6:7], float(n))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(st.gflat, op=dist.ReduceOp.SUM, group=group)
    abi.cast_f32_f64(st.stats32, st.stats64)

    # ---- EMA codebook update with the global statistics (:582-583), separation loss on the updated codebook ----
    vq = m.vq_layer
    update, vq_loss, _ = abi.vq_ema_update(st.stats64, codebook, vq.decay, vq.epsilon, vq.commitment_cost, True,
                                           vq.state)
    m._codebook.copy_(update)
    sim_w = lw['sim_loss_weight']
    if sim_w > 0:
        # gradient scale = sim_w * rows / gbs (scalar broadcast to every row, vq_nfr.py:971-972); rows == known
        n_glob = n * (dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1) \
            if full else None
        scale = sim_w * (float(n_glob) * inv_gbs if n_glob is not None else float((st.sums[6] * inv_gbs).item()))
        abi.codebook_sim_loss(m._codebook, scale, st.sim_loss, st.d_codebook, accumulate=False)
    # weighted = sums[5]/gbs + (rows/gbs) * (vq_w * vq_loss + sim_w * sim_loss): the two broadcast scalars weigh by the share
    # of active rows; one launch for the scalar arithmetic
    sc = abi.train_scalars(st.sums, vq_loss, st.sim_loss if sim_w > 0 else None, inv_gbs, lw['vq_loss_weight'], sim_w,
                           st.scalars)
    weighted = sc[0]
    loss_dict = {'rgb': st.sums[0], 'vqrgb': st.sums[1], 'chromaticity': st.sums[2], 'chr_smooth': st.sums[3],
                 'lambert': st.sums[4], 'vqloss': sc[1], 'sim_smooth': sc[2], 'rows': st.sums[6]}
    if apply:
        optimizer.apply_gradients(st.params, st.grads, lr_t_dev=_lr_t_dev)
        st.dirty = True
    partial_to_vis = {'id': id_, 'hw': hw, 'pred_rgb_linear': rgb_pred, 'pred_vq_rgb_linear': vq_rgb,
                      'pred_albedo': albedo, 'pred_spec': spec, 'pred_rough': rough_c, 'embed_ind': idx + 1,
                      'z_enc': z_enc, 'z_vq': z_vq, 'loss_rows': B['loss_rows']}
    return weighted, partial_to_vis, loss_dict


class GraphedTrainIter:
    """train_iter captured ONCE into a CUDA graph and replayed: the ~150 kernel launches of a step (plus the NCCL
    all-reduce) become one graph launch, which is what makes the 8192-rays-per-GPU step GPU-bound instead of
    launch-bound.  Inputs are copied into static buffers; the codeword-dropout mask and Adam's step-dependent
    rate live in device memory so that a replay needs no re-capture.  Requires an all-foreground batch of a fixed
    row count (what the reference's sampler produces, train_nfr.py:380-467)."""

    def __init__(self, model, optimizer: Adam, global_bs: int, example_batch, use_thres: bool = True, group=None):
        self.model, self.opt, self.gbs, self.group = model, optimizer, int(global_bs), group
        dev = model.device
        model.assume_all_foreground = True
        self.static = [t.clone() if torch.is_tensor(t) else t for t in example_batch]
        self.K = model.num_embed
        self.sel_mask = torch.ones((self.K,), dtype=F32, device=dev) if use_thres else None
        self.lr_t = torch.zeros((1,), dtype=F32, device=dev)
        # per-step host values (Adam's rate, the dropout mask) go through a RING of pinned staging slots, each guarded by
        # an event recorded after its H2D copy: the host may run several replays ahead of the GPU, and rewriting a pinned
        # buffer whose copy is still queued would hand a step the NEXT step's rate / a torn mask
        self._ring = [(torch.zeros((1,), dtype=F32).pin_memory(), torch.ones((self.K,), dtype=F32).pin_memory(),
                       torch.cuda.Event()) for _ in range(4)]
        self._ring_i = 0
        # warm-up on a side stream (allocates the activation set, Adam state, NCCL buffers), then capture; the
        # learned state touched by the warm-up steps is snapshotted and restored so that construction has no effect
        st = _train_state(model)
        optimizer.ensure_state(st.params)
        vq = model.vq_layer
        snap = (st.params.clone(), {k: v.clone() for k, v in vq.state.items()}, optimizer.iterations,
                optimizer.m.clone(), optimizer.v.clone(), optimizer.vhat.clone(), vq._gen.get_state())
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(2):
                self._refresh(None, None)
                self.out = train_iter(model, tuple(self.static), optimizer, self.gbs, group=group,
                                      _sel_mask_dev=self.sel_mask, _lr_t_dev=self.lr_t)
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        st.params.copy_(snap[0])
        for k, v in snap[1].items():
            vq.state[k].copy_(v)
        optimizer.iterations = snap[2]
        optimizer.m.copy_(snap[3]); optimizer.v.copy_(snap[4]); optimizer.vhat.copy_(snap[5])
        vq._gen.set_state(snap[6])
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode='thread_local'):
            self.out = train_iter(model, tuple(self.static), optimizer, self.gbs, group=group,
                                  _sel_mask_dev=self.sel_mask, _lr_t_dev=self.lr_t)

    def _refresh(self, thres, roll):
        self.opt.iterations += 1
        lr_host, mask_host, ev = self._ring[self._ring_i]
        self._ring_i = (self._ring_i + 1) % len(self._ring)
        ev.synchronize()                                  # the copy that last read this slot has completed (no-op if unused)
        lr_host[0] = self.opt.lr_t()
        self.lr_t.copy_(lr_host, non_blocking=True)
        if self.sel_mask is not None:
            if thres is None:
                mask_host.fill_(1.0)
            else:
                th = torch.as_tensor(thres, dtype=F32).reshape(-1)
                if roll is None:
                    roll = torch.rand((1, self.K), generator=self.model.vq_layer._gen, dtype=F32)
                mask_host.copy_((torch.as_tensor(roll, dtype=F32).reshape(-1) >= th).to(F32))
            self.sel_mask.copy_(mask_host, non_blocking=True)
        ev.record(torch.cuda.current_stream(self.model.device))

    def __call__(self, batch, thres=None, roll=None):
        for dst, src in zip(self.static, batch):
            if torch.is_tensor(dst) and src is not dst:
                dst.copy_(src, non_blocking=True)
        self._refresh(thres, roll)
        self.graph.replay()
        return self.out


def outer_sample(batch, config, data_type, alpha_thres=0.9, seed=0, hw=None):
    """train_nfr.py:380-467: n_rays_per_step (pixel, random 8-neighbour) pairs of one H x W view, both above
    alpha_thres, as the 2*bs-row batch [p1, p1_n, p2, p2_n, ...] of every per-pixel tensor.  `hw`: (H, W) if known
    (avoids the reference's hw[0,:].numpy() host sync); `seed`: counter-hash seed of this draw (pass the step number).
    `id_` (strings in the reference) is passed through unchanged."""
    cfg = config
    bs = int(cfg.getint('DEFAULT', 'n_rays_per_step')) if hasattr(cfg, 'getint') and hasattr(cfg, 'has_option') \
        else int(cfg.get('n_rays_per_step', 1024))
    if data_type == 'nerf':
        id_, hw_t, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal, lvis = batch
    else:
        id_, hw_t, rayo, rayd, rgb, alpha, pred_alpha, xyz, normal = batch
        lvis = None
    if hw is None:
        hw = tuple(int(v) for v in hw_t[0, :].tolist())
    h, w = hw
    rows, n_valid = abi.sample_pairs(alpha, h, w, bs, seed, alpha_thres)
    g = lambda t: abi.gather_rows(t, rows)
    hw_out = hw_t.index_select(0, rows.clamp_min(0).long())          # int32 metadata: index plumbing
    out = [id_, hw_out, g(rayo), g(rayd), g(rgb), g(alpha).reshape(-1, 1), g(pred_alpha).reshape(-1, 1), g(xyz),
           g(normal)]
    if data_type == 'nerf':
        if lvis is None:
            raise ValueError('NeRF data requires lvis')
        out.append(g(lvis))
    return tuple(out)


def sync_inference_weights(model) -> None:
    """Re-pack the inference-side weight images (tensor-core swizzled copies) after optimizer steps."""
    st = getattr(model, '_train_state', None)
    if st is not None and st.dirty:
        for name in NET_ORDER:
            model.net[name].weights_updated()
        st.dirty = False
