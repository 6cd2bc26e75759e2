"""Golden vectors for the TRAINING step on non-'nerf' data (no light visibility, learnable tone scaling
`rgb = (rgb * gamma_bias) ** clip(gamma_index, 0, 5)`, models/vq_nfr.py:707, 715-718, 736-745), produced by the REFERENCE'S
OWN CODE through the tf_shim stand-in (see oracle/gen_golden_decomp_ref.py): one `call(mode='train')` + `compute_loss` +
gradients of every trainable variable including `_gamma_bias` / `_gamma_index`.

    python oracle/gen_golden_decomp_real.py       -> tests/golden/decomp_real_ref.npz       (needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden_decomp_ref as G  # noqa: E402

tf, O = G.tf, G.O
OUT = os.path.join(HERE, '..', 'tests', 'golden', 'decomp_real_ref.npz')
GAMMA = (1.15, 0.85)


def main():
    R = G.import_reference()
    n, seed = 98, 7
    rec = {'n': n, 'seed': seed, 'bias_scale': 0.05, 'fg_frac': 0.8, 'gamma': np.array(GAMMA)}
    np_ = lambda v: v.detach().double().numpy()
    tf.set_float(torch.float64)
    try:
        scene = O.synth_scene(seed, bias_scale=0.05, data_type='real')
        scene.gamma = GAMMA
        b = O.synth_batch(n, seed, fg_frac=0.8, with_lvis=False)
        m, _, _ = G.build_model(R, scene, data_type='real')
        batch = G.make_batch(b, data_type='real')
        thres = np.array([0.0] * 3 + [0.3] * 12, np.float32)
        roll = np.random.RandomState(seed + 1).uniform(0, 1, size=(15,)).astype(np.float32)
        rec['thres'], rec['roll'] = thres, roll
        n_fg = int((b['alpha'][:, 0] > 0).sum())
        assert n_fg % 2 == 0
        global_bs = n_fg // 2
        rec['global_bs'] = global_bs
        tf.random.queue.append(roll.reshape(1, -1))
        pred, gt, lk, _ = m.call(batch, mode='train', thres=thres)
        rec['train_rgb'], rec['train_vqrgb'] = np_(pred['rgb']), np_(lk['vqrgb'])
        per_example, ld = m.compute_loss(pred, gt, **dict(lk, keep_batch=True))
        rec['train_per_example'] = np_(per_example)
        weighted = torch.sum(per_example) / global_bs
        rec['train_loss'] = np_(weighted)
        variables = [('light', m._light), ('gamma_bias', m._gamma_bias), ('gamma_index', m._gamma_index)]
        for name in sorted(m.net):
            for li, layer in enumerate(m.net[name].layers):
                variables += [('%s_w%d' % (name, li), layer.kernel), ('%s_b%d' % (name, li), layer.bias)]
        grads = torch.autograd.grad(weighted, [v for _, v in variables], allow_unused=True)
        for (name, _), g_ in zip(variables, grads):
            G.project(name, g_, rec, 'train_d_')
    finally:
        tf.set_float(torch.float32)
    np.savez_compressed(OUT, **rec)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes,', len(rec), 'arrays; d gamma =', rec['train_d_gamma_bias'], rec['train_d_gamma_index'])


if __name__ == '__main__':
    main()
