"""tcgen05 primitive self-test: one 128 x N x K tile through the same descriptor / swizzle / TMEM / mbarrier
helpers (csrc/tc_common.cuh) that the tensor-core MLP kernels use, against a CPU GEMM of identically rounded
operands."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tf32(x):
    """cvt.rna.tf32.f32: round to nearest (ties away) on the 13 dropped mantissa bits"""
    u = x.view(np.uint32).astype(np.uint64) + np.uint64(0x1000)
    return (u & np.uint64(0xFFFFE000)).astype(np.uint32).view(np.float32)


@pytest.mark.parametrize('mode,n,k', [(0, 16, 32), (0, 128, 64), (0, 256, 128), (1, 16, 64), (1, 256, 128),
                                     (1, 128, 256), (2, 16, 32), (2, 256, 64), (2, 128, 64),
                                     (3, 16, 32), (3, 128, 64), (3, 256, 64), (3, 128, 128)])    # 3: A operand from tensor memory
def test_tc_selftest_gemm(cuda_dev, mode, n, k):
    from vqnerf_release_b200 import abi
    rng = np.random.RandomState(mode * 1000 + n + k)
    a = rng.normal(size=(128, k)).astype(np.float32)
    b = rng.normal(size=(n, k)).astype(np.float32)
    d = abi.tc_selftest(torch.as_tensor(a).to(cuda_dev), torch.as_tensor(b).to(cuda_dev), mode).cpu().numpy()
    if mode == 0:
        ref = _tf32(a).astype(np.float64) @ _tf32(b).astype(np.float64).T
        tol = 2e-5
    elif mode == 1:
        ar = torch.as_tensor(a).to(torch.bfloat16).double().numpy()
        br = torch.as_tensor(b).to(torch.bfloat16).double().numpy()
        ref = ar @ br.T
        tol = 2e-5
    else:
        ref = a.astype(np.float64) @ b.astype(np.float64).T
        tol = 3e-6 if mode == 2 else 8e-6   # 3xTF32 with round-to-nearest hi/lo: fp32-level accuracy (plain tf32: ~5e-4)
    scale = np.sqrt(k)
    err = np.abs(d - ref).max() / scale
    print('tc_selftest mode %d N=%d K=%d err/sqrt(K)=%.3e' % (mode, n, k, err))
    assert err < tol, 'mode %d N=%d K=%d: max err / sqrt(K) = %.3e' % (mode, n, k, err)


# ------------------------------------------------------------------------------------------------
# tensor-core MLP modes against the float64 oracle
# ------------------------------------------------------------------------------------------------
def _close(a, b, name, rtol, atol):
    a = a.detach().cpu().double().numpy()
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    err = np.abs(a - b)
    bad = err > atol + rtol * np.abs(b)
    print('%s: max abs err %.3e' % (name, err.max()))
    assert not bad.any(), '%s: %d/%d out of tolerance, max abs err %.3e' % (name, bad.sum(), bad.size, err.max())


@pytest.mark.parametrize('precision,rtol,atol', [('tf32x3', 1e-4, 2e-6), ('bf16', 1e-2, 2e-3)])
def test_tc_mlp_matches_oracle(cuda_dev, precision, rtol, atol):
    from oracle import decomp_oracle as O
    from tests.test_gpu_parity import _model_from_scene
    scene = O.synth_scene(1, bias_scale=0.05)
    m = _model_from_scene(scene, cuda_dev, precision=precision)
    rng = np.random.RandomState(5)
    for n in (1, 127, 128, 129, 1000, 20000):
        pts = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
        z = m._pred_enc_at(torch.as_tensor(pts).to(cuda_dev))
        zo = O.pred_enc_at(scene.nets, torch.as_tensor(pts, dtype=torch.float64))
        _close(z, zo, '%s z_enc n=%d' % (precision, n), rtol, atol)
        # heads are checked on the oracle's latent so that each stage is held to the tolerance on its own
        zin = zo.to(torch.float32).to(cuda_dev)
        for vq in (False, True):
            sfx = '_vq' if vq else '_main'
            _close(m._pred_diff_at(zin, vq), O.pred_head(scene.nets, 'diff' + sfx, zo), 'diff' + sfx, rtol, atol)
            _close(m._pred_spec_at(zin, vq), O.pred_head(scene.nets, 'spec' + sfx, zo), 'spec' + sfx, rtol, atol)
            _close(m._pred_rough_at(zin, vq), O.pred_head(scene.nets, 'rough' + sfx, zo), 'rough' + sfx, rtol, atol)


@pytest.mark.parametrize('precision,rtol,atol', [('tf32x3', 1e-4, 5e-6), ('bf16', 1e-2, 2e-3)])
def test_tc_fast_render_matches_oracle(cuda_dev, precision, rtol, atol):
    from oracle import decomp_oracle as O
    from tests.test_gpu_parity import _batch_tuple, _model_from_scene
    n = 5000
    scene = O.synth_scene(11, n_probes=2, bias_scale=0.05)
    batch = O.synth_batch(n, 11, fg_frac=0.7)
    m = _model_from_scene(scene, cuda_dev, precision=precision)
    pred, _, _, _ = m.fast_render(_batch_tuple(batch, cuda_dev), mode='test', relight_probes=True)
    o = O.fast_render(scene, batch, torch.float64, relight_probes=True)
    for k in ('basecolor', 'albedo', 'spec', 'rough', 'rgb_probes'):
        _close(pred[k], o[k], '%s %s' % (precision, k), rtol, atol)


def test_tc_generic_net_forward(cuda_dev):
    """mlp.Network.__call__ on the tensor cores, incl. a skip connection fed from global rows."""
    from oracle import decomp_oracle as O
    from vqnerf_release_b200.nerfactor.networks.mlp import Network
    nets = O.make_vq_nfr_nets(3, bias_scale=0.1)
    rng = np.random.RandomState(0)
    for name, inp in (('bottleneck', rng.normal(size=(300, 128)).astype(np.float32)),
                      ('diff_main', rng.uniform(0, 1, size=(257, 256)).astype(np.float32))):
        on = nets[name]
        act = [{0: None, 1: 'relu', 2: 'sigmoid'}[a] for a in on.acts]
        net = Network.from_arrays(on.weights, on.biases, act, skip_at=None if on.skip_at is None else [on.skip_at],
                                  device=cuda_dev)
        ref = on(torch.as_tensor(inp, dtype=torch.float64))
        _close(net(torch.as_tensor(inp).to(cuda_dev), precision='tf32x3'), ref, name + ' tf32x3', 1e-4, 2e-6)
        _close(net(torch.as_tensor(inp).to(cuda_dev), precision='bf16'), ref, name + ' bf16', 1e-2, 2e-3)
