"""BASELINE.json's full-size configuration (C2: 1024 molecules x 29 atoms, 5 layers, hidden 128) on the default
fp32-accurate tensor-core path, through size-independent properties: the oracle needs minutes per step at this size,
so parity is anchored on (1) a seeded sub-batch checked against the oracle, (2) molecules being independent --
every per-molecule result of the big batch must equal the same molecule processed in a small batch, (3) the exact
inverse, (4) bitwise determinism of loss and gradients."""
import numpy as np
import pytest
import torch

from gpu_util import gpu_batch, rel_err, to_np
from oracle import enflow_oracle as orc

pytestmark = pytest.mark.gpu
B, NF, L = 1024, 5, 5


def _model(precision='fp32_tc'):
    from enflow_b200.data import synthetic as syn
    from gpu_util import build_model
    return build_model(syn.make_weights(NF, 128, L, seed=0), NF, L, precision=precision)


def _take(arrs, lo, hi):
    off = np.concatenate([[0], np.cumsum(arrs['N'])])
    sl = slice(off[lo], off[hi])
    out = {k: arrs[k][sl] for k in ('h', 'g', 'pos', 'vel', 'box')}
    out['N'], out['r_cut'] = arrs['N'][lo:hi], arrs['r_cut'][lo:hi]
    return out, sl


def test_full_size_c2_properties():
    from enflow_b200.data import synthetic as syn
    from enflow_b200.flow.loss import Alchemical_NLL
    arrs = syn.make_batch('c2', B, ragged=True)
    n_atoms = int(arrs['N'].sum())
    eps = syn.make_noise(n_atoms, NF)
    model = _model()
    nll = Alchemical_NLL(kBT=syn.TRAIN_KBT, softening=syn.TRAIN_SOFTENING)
    runs = []
    for _ in range(2):
        model.zero_grad(set_to_none=True)
        out, ldj = model(gpu_batch(arrs), eps=torch.as_tensor(eps))
        loss = nll(out, ldj)
        loss.backward()
        runs.append((loss.detach().clone(), out.pos.detach().clone(), out.g.detach().clone(), out.ldj_mol.detach().clone(),
                     model.flat_grads.clone()))
    # (4) determinism at full size
    for a, b in zip(runs[0], runs[1]):
        assert torch.equal(a, b)
    assert torch.isfinite(runs[0][4]).all() and float(runs[0][4].abs().max()) > 0
    # (2) independence: molecules 100..116 alone give the same latents and per-molecule log-det
    sub, sl = _take(arrs, 100, 116)
    with torch.no_grad():
        o2, _ = model(gpu_batch(sub), eps=torch.as_tensor(eps[sl]))
    assert rel_err(to_np(o2.pos), to_np(runs[0][1][sl])) < 2e-6
    assert rel_err(to_np(o2.g), to_np(runs[0][2][sl])) < 2e-6
    assert rel_err(to_np(o2.ldj_mol), to_np(runs[0][3][100:116])) < 2e-6
    # (1) the same sub-batch against the oracle (fp64 restatement of the reference)
    p = orc.params_to_torch({k: v for k, v in syn.make_weights(NF, 128, L, seed=0).items()})
    ref, _, ref_ldj_mol = orc.lf_forward(p, L, orc.to_torch(sub), syn.TRAIN_DT, torch.as_tensor(eps[sl], dtype=torch.float64))
    for k in ('pos', 'vel', 'h', 'g'):
        assert rel_err(to_np(getattr(o2, k)), ref[k].numpy()) < 1e-5, k
    assert rel_err(to_np(o2.ldj_mol), ref_ldj_mol.numpy()) < 1e-5
    # (3) exact inverse at full size
    with torch.no_grad():
        out, _ = model(gpu_batch(arrs), eps=torch.as_tensor(eps))
        fwd_ldj = out.ldj_mol.clone()
        back = model.reverse(out, quantize=True)
    assert rel_err(to_np(back.pos), arrs['pos']) < 1e-5
    assert rel_err(to_np(back.vel), arrs['vel']) < 1e-5
    assert np.array_equal(to_np(back.h), arrs['h'])
    assert rel_err(to_np(back.neg_ldj_mol), -to_np(fwd_ldj)) < 1e-5


def test_empty_batch_is_a_no_op():
    arrs = {'h': np.zeros((0, NF)), 'g': np.zeros((0, NF)), 'pos': np.zeros((0, 3)), 'vel': np.zeros((0, 3)),
            'box': np.zeros((0, 3)), 'N': np.zeros(0, dtype=np.int64), 'r_cut': np.zeros(0, dtype=np.float32)}
    model = _model()
    with torch.no_grad():
        out, ldj = model(gpu_batch(arrs), eps=torch.zeros(0, NF))
    assert out.pos.shape == (0, 3) and out.h.shape == (0, NF)
    assert float(ldj) == 0.0


@pytest.mark.parametrize('B', [37, 75])
def test_mid_size_train_step_vs_oracle(B):
    """A few tiles per CTA (the software-pipelined kernels rotate three tile-info buffers and two accumulators):
    37 molecules ~ 2 tiles of the forward and 3 of the backward kernel per CTA, 75 ~ 3-4 and 6-7; loss, latents and
    every gradient against the oracle."""
    from enflow_b200.data import synthetic as syn
    from enflow_b200.flow.loss import Alchemical_NLL
    arrs = syn.make_batch('c2', B, seed=77 + B)
    eps = syn.make_noise(int(arrs['N'].sum()), NF, seed=5)
    sd = syn.make_weights(NF, 128, L, seed=0)
    model = _model()
    out, ldj = model(gpu_batch(arrs), eps=torch.as_tensor(eps))
    loss = Alchemical_NLL(kBT=syn.TRAIN_KBT, softening=syn.TRAIN_SOFTENING)(out, ldj)
    loss.backward()
    ref_loss, ref_grads, ref_state, _, _ = orc.train_step(sd, L, arrs, syn.TRAIN_DT, eps, syn.TRAIN_KBT, syn.TRAIN_SOFTENING)
    assert abs(loss.item() - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    for k in ('pos', 'vel', 'h', 'g'):
        assert rel_err(to_np(getattr(out, k)), ref_state[k].detach().numpy()) < 1e-5, k
    worst = max(np.linalg.norm(to_np(p.grad) - ref_grads[k].numpy()) / max(np.linalg.norm(ref_grads[k].numpy()), 1e-300)
                for k, p in model.named_parameters())
    assert worst < 1e-4, worst


# ---- the other BASELINE.json configurations at bench size: sub-batch vs oracle + independence of molecules --------
def _model_for(nf, precision, seed=0):
    from enflow_b200.data import synthetic as syn
    from gpu_util import build_model
    return build_model(syn.make_weights(nf, 128, L, seed=seed, coord_gain=0.5), nf, L, precision=precision), \
        syn.make_weights(nf, 128, L, seed=seed, coord_gain=0.5)


def _sub_batch_and_independence(config, B, nf, precision, tol, lo, hi, kw=None, reverse=False):
    """Run the full-size batch once; molecules [lo, hi) processed alone must give the same per-molecule results
    (they never interact: data/base.py:129-142, loss.py:13), and that sub-batch must match the fp64 oracle."""
    from enflow_b200.data import synthetic as syn
    arrs = syn.make_batch(config, B, **(kw or {}))
    eps = syn.make_noise(int(arrs['N'].sum()), nf)
    model, sd = _model_for(nf, precision)
    sub, sl = _take(arrs, lo, hi)
    p = orc.params_to_torch(sd)
    if not reverse:
        with torch.no_grad():
            big, _ = model(gpu_batch(arrs), eps=torch.as_tensor(eps))
            small, _ = model(gpu_batch(sub), eps=torch.as_tensor(eps[sl]))
        for k in ('pos', 'vel', 'h', 'g'):
            assert rel_err(to_np(getattr(small, k)), to_np(getattr(big, k)[sl])) < max(2e-6, tol / 5), k
        assert rel_err(to_np(small.ldj_mol), to_np(big.ldj_mol[lo:hi])) < max(2e-6, tol / 5)
        ref, _, ref_ldj_mol = orc.lf_forward(p, L, orc.to_torch(sub), syn.TRAIN_DT, torch.as_tensor(eps[sl], dtype=torch.float64))
        for k in ('pos', 'vel', 'h', 'g'):
            assert rel_err(to_np(getattr(small, k)), ref[k].numpy()) < tol, k
        assert rel_err(to_np(small.ldj_mol), ref_ldj_mol.numpy()) < tol
    else:
        with torch.no_grad():
            big = model.reverse(gpu_batch(arrs), quantize=False)
            small = model.reverse(gpu_batch(sub), quantize=False)
        for k in ('pos', 'vel', 'h', 'g'):
            assert rel_err(to_np(getattr(small, k)), to_np(getattr(big, k)[sl])) < max(2e-6, tol / 5), k
        assert rel_err(to_np(small.neg_ldj_mol), to_np(big.neg_ldj_mol[lo:hi])) < max(2e-6, tol / 5)
        ref, ref_neg = orc.lf_reverse(p, L, orc.to_torch(sub), syn.TRAIN_DT, quantize=False)
        for k in ('pos', 'vel', 'h', 'g'):
            assert rel_err(to_np(getattr(small, k)), ref[k].numpy()) < 2 * tol, k
        assert rel_err(to_np(small.neg_ldj_mol), ref_neg.numpy()) < 2 * tol


def test_full_size_c1_train_yaml_shape():
    """C1: example/train.yaml shape, 64 x 22 atoms, radius graph in the PBC-quirk regime, fp32-accurate mode."""
    _sub_batch_and_independence('c1', 64, 4, 'fp32_tc', 1e-5, 20, 28)


def test_full_size_c3_lj55():
    """C3: 1024 LJ-55 clusters, fully connected, one feature."""
    _sub_batch_and_independence('c3', 1024, 1, 'fp32_tc', 1e-5, 500, 504)


def test_full_size_c5_bf16_radius_graph():
    """C5: 32 x 500-atom fragments, radius-cutoff graph, bf16 edge MLP (tolerance 1e-2)."""
    _sub_batch_and_independence('c5', 32, 5, 'bf16', 1e-2, 7, 8, kw={'n_atoms': 500})


def test_full_size_c4_inverse_pass():
    """C4: the generate.yaml inverse pass on 16 384 x 22-atom latents (PBC regime), product mode, against the oracle's
    inverse on a sub-batch, with the per-molecule -sum(Q)."""
    _sub_batch_and_independence('c4', 16384, 4, 'fp32_tc', 1e-5, 9000, 9008, reverse=True)
