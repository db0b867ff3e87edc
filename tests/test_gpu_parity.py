"""Parity of the CUDA path (through the C ABI) with the reference: golden vectors + CPU oracle.

Every test runs in the three arithmetic modes of the edge MLP (gpu_util.MODES): 'fp32_tc' (the product default:
tcgen05, bf16x3 operand split), 'fp32' (CUDA-core FFMA cross-check) and 'bf16' (tcgen05, one MMA per GEMM).
Tolerances (BASELINE.json north star): log-likelihood and latents within 1e-5 relative in the fp32-accurate modes,
1e-2 under the bf16 MLP; edge/index construction bit-exact.  Gradients are held to 1e-4 (5e-2 in bf16 mode)
relative, L2 per tensor.  `check_close` asserts the max-norm figure and prints the elementwise one (1 % floor) next
to it (`pytest -rP`).
"""
import numpy as np
import pytest
import torch

from golden_util import CASES, load_case
from gpu_util import DEV, MODES, build_model, check_close, gpu_batch, rel_err, to_np
from oracle import enflow_oracle as orc

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-5
modes = pytest.mark.parametrize('mode,tol,gtol', MODES)


def oracle_trace(c):
    p = orc.params_to_torch(c['sd'])
    b = orc.to_torch(c['batch'])
    trace = []
    with torch.no_grad():
        state, ldj, ldj_mol = orc.lf_forward(p, c['L'], b, c['dt'], torch.as_tensor(c['eps']), trace=trace)
        z0, log_q = orc.argmax_forward(p, b['h'], torch.as_tensor(c['eps']))
    return trace, state, ldj, ldj_mol, z0, log_q


def grad_errors(model, ref_grads):
    return sorted(((np.linalg.norm(to_np(p.grad) - ref_grads[k].numpy()) / max(np.linalg.norm(ref_grads[k].numpy()), 1e-300), k)
                   for k, p in model.named_parameters()), reverse=True)


@pytest.mark.parametrize('name', list(CASES))
def test_edges_bit_exact_vs_golden(name):
    """K0 on the reference's own fp64 positions reproduces Data.edges exactly, in the reference order."""
    c = load_case(name)
    trace = oracle_trace(c)[0]
    arrs = dict(c['batch'])
    for i in range(c['L']):
        if i > 0:
            arrs['pos'] = trace[i - 1]['pos'].numpy()          # fp64 positions entering layer i
        e = gpu_batch(arrs).edges
        assert np.array_equal(e.row.cpu().numpy(), c['gold'][f'row{i}'].astype(np.int64)), f'layer {i} row'
        assert np.array_equal(e.col.cpu().numpy(), c['gold'][f'col{i}'].astype(np.int64)), f'layer {i} col'
        r32, c32, rowptr, e_dev = e.csr
        assert bool((r32[1:] >= r32[:-1]).all()), 'CSR form must be row-grouped'
        assert int(rowptr[-1]) == r32.numel()


@pytest.mark.parametrize('name', ['c1_pbc', 'c5_small'])
def test_edges_fp32_inputs_match_oracle(name):
    """fp32 positions (the hot path's state) upcast exactly: same edges as the oracle on those values."""
    c = load_case(name)
    arrs = dict(c['batch'])
    rs = np.random.RandomState(5)
    arrs['pos'] = (arrs['pos'] + rs.normal(0, 0.05, arrs['pos'].shape)).astype(np.float32).astype(np.float64)
    b = orc.to_torch(arrs)
    row, col, _ = orc.build_edges(b['pos'], b['box'], b['N'], b['r_cut'])
    e = gpu_batch(arrs, dtype=torch.float32).edges
    assert np.array_equal(e.row.cpu().numpy(), row.numpy())
    assert np.array_equal(e.col.cpu().numpy(), col.numpy())


@pytest.mark.parametrize('name', list(CASES))
@modes
def test_egcl_layers_vs_golden(name, mode, tol, gtol):
    """EGCL.forward(h, edges) per layer, fed the oracle's layer inputs: Q, F, G against the reference."""
    c = load_case(name)
    trace, _, _, _, z0, _ = oracle_trace(c)
    model = build_model(c['sd'], c['nf'], c['L'], precision=mode)
    arrs = dict(c['batch'])
    h = z0.numpy()
    for i in range(c['L']):
        if i > 0:
            arrs['pos'], h = trace[i - 1]['pos'].numpy(), trace[i - 1]['h'].numpy()
        data = gpu_batch(arrs)
        with torch.no_grad():
            Q, F, G = model.networks[i](torch.as_tensor(h, device=DEV), data.edges)
        for k, v in (('Q', Q), ('G', G)):
            check_close(to_np(v), c['gold'][f'{k}{i}'], tol, f'{mode} layer {i} {k}')
        # F is a mean over edges of terms that cancel: its error is bounded against the scale of the summands
        # (what the latents inherit through vel += F dt), not against the cancelled mean
        err = rel_err(to_np(F), c['gold'][f'F{i}'])
        assert err < 5 * tol, f'{mode} layer {i} F: rel err {err:.3e}'


@pytest.mark.parametrize('name', list(CASES))
@modes
def test_flow_forward_and_loss_vs_golden(name, mode, tol, gtol):
    from enflow_b200.flow.loss import Alchemical_NLL
    c = load_case(name)
    model = build_model(c['sd'], c['nf'], c['L'], precision=mode)
    data = gpu_batch(c['batch'])
    with torch.no_grad():
        out, ldj = model(data, eps=torch.as_tensor(c['eps']))
        loss = Alchemical_NLL(kBT=c['kBT'], softening=c['softening'])(out, ldj)
    for k in ('h', 'g', 'pos', 'vel'):
        check_close(to_np(getattr(out, k)), c['gold'][f'out_{k}'], tol, f'{mode} {k}')
    gl = float(c['gold']['ldj'])
    assert abs(ldj.item() - gl) <= tol * max(abs(gl), 1.0), f'{mode} ldj {ldj.item()} vs {gl}'
    assert abs(loss.item() - float(c['gold']['loss'])) <= tol * abs(float(c['gold']['loss']))


@pytest.mark.parametrize('name', list(CASES))
@modes
def test_flow_train_step_vs_oracle(name, mode, tol, gtol):
    from enflow_b200.flow.loss import Alchemical_NLL
    c = load_case(name)
    model = build_model(c['sd'], c['nf'], c['L'], precision=mode)
    out, ldj = model(gpu_batch(c['batch']), eps=torch.as_tensor(c['eps']))
    loss = Alchemical_NLL(kBT=c['kBT'], softening=c['softening'])(out, ldj)
    loss.backward()
    ref_loss, ref_grads, _, _, _ = orc.train_step(c['sd'], c['L'], c['batch'], c['dt'], c['eps'], c['kBT'], c['softening'])
    assert abs(loss.item() - ref_loss.item()) <= tol * abs(ref_loss.item())
    for k, p in model.named_parameters():
        assert p.grad is not None, k
    worst = grad_errors(model, ref_grads)
    assert worst[0][0] < gtol, f'{mode} worst gradient errors: {worst[:5]}'


@pytest.mark.parametrize('name', list(CASES))
@modes
def test_reverse_vs_golden(name, mode, tol, gtol):
    """LFIntegrator.reverse on the reference's own latents against the reference's inverse (dynamics.py:25-37)."""
    c = load_case(name)
    model = build_model(c['sd'], c['nf'], c['L'], precision=mode)
    arrs = dict(c['batch'])
    for k in ('h', 'g', 'pos', 'vel'):
        arrs[k] = c['gold'][f'out_{k}']
    back = model.reverse(gpu_batch(arrs))
    for k in ('g', 'pos', 'vel'):
        check_close(to_np(getattr(back, k)), c['gold'][f'rev_{k}'], 2 * tol, f'{mode} {k}')
    if mode != 'bf16':      # one-hot after ArgMax.reverse; a bf16-sized perturbation may flip a near-tie
        assert np.array_equal(to_np(back.h), c['gold']['rev_h'])
    else:
        assert (to_np(back.h) != c['gold']['rev_h']).any(axis=1).mean() < 0.02


@modes
def test_round_trip_and_neg_ldj(mode, tol, gtol):
    """reverse(forward(x)) == x (the reference's own self-check, enflow/main.py:275-278) and
    the per-molecule -sum(Q) of the inverse equals minus the forward's per-molecule log-det.  The inverse
    evaluates the same kernels on the same inputs as the forward pass, so it is exact in every mode."""
    c = load_case('c2_ragged')
    model = build_model(c['sd'], c['nf'], c['L'], precision=mode)
    data = gpu_batch(c['batch'])
    with torch.no_grad():
        out, ldj = model(data, eps=torch.as_tensor(c['eps']))
    fwd_ldj_mol = out.ldj_mol.clone()
    back = model.reverse(out, quantize=True)
    check_close(to_np(back.pos), c['batch']['pos'], FWD_TOL, 'pos')
    check_close(to_np(back.vel), c['batch']['vel'], FWD_TOL, 'vel')
    check_close(to_np(back.g), c['batch']['g'], FWD_TOL, 'g')
    assert np.array_equal(to_np(back.h), c['batch']['h'])
    assert rel_err(to_np(back.neg_ldj_mol), -to_np(fwd_ldj_mol)) < FWD_TOL


@pytest.mark.parametrize('name', ['c1_pbc', 'c2_ragged'])
@modes
def test_deterministic_bitwise(name, mode, tol, gtol):
    from enflow_b200.flow.loss import Alchemical_NLL
    c = load_case(name)
    runs = []
    for _ in range(2):
        model = build_model(c['sd'], c['nf'], c['L'], precision=mode)
        out, ldj = model(gpu_batch(c['batch']), eps=torch.as_tensor(c['eps']))
        loss = Alchemical_NLL(kBT=c['kBT'], softening=c['softening'])(out, ldj)
        loss.backward()
        runs.append((loss.detach().clone(), out.pos.detach().clone(), model.flat_grads.clone()))
    assert torch.equal(runs[0][0], runs[1][0])
    assert torch.equal(runs[0][1], runs[1][1])
    assert torch.equal(runs[0][2], runs[1][2]), 'gradients must be bit-identical run to run'


def _random_rotation(rs):
    q, r = np.linalg.qr(rs.normal(size=(3, 3)))
    q = q * np.sign(np.diag(r))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


@modes
def test_equivariance_fc_regime(mode, tol, gtol):
    """Rotation / translation / per-molecule permutation in the well-defined (no-PBC) regime, SURVEY Q6.  The
    transformed run rounds differently, so the symmetry holds to the mode's arithmetic tolerance."""
    from enflow_b200.data import synthetic as syn
    nf, L = 5, 3
    arrs = syn.make_batch('c2', 5, ragged=True, seed=77)
    sd = syn.make_weights(nf, 128, L, seed=3, coord_gain=0.5)
    eps = syn.make_noise(int(arrs['N'].sum()), nf, seed=5)
    model = build_model(sd, nf, L, precision=mode)
    t2 = 2 * tol

    def run(a, e):
        with torch.no_grad():
            out, ldj = model(gpu_batch(a), eps=torch.as_tensor(e))
        return {k: to_np(getattr(out, k)) for k in ('h', 'g', 'pos', 'vel')}, ldj.item(), to_np(out.ldj_mol)

    base, ldj0, ldjm0 = run(arrs, eps)
    rs = np.random.RandomState(11)
    R = _random_rotation(rs)
    rot = dict(arrs)
    rot['pos'], rot['vel'] = arrs['pos'] @ R.T, arrs['vel'] @ R.T
    o, ldj, _ = run(rot, eps)
    assert abs(ldj - ldj0) < 10 * tol * max(abs(ldj0), 1)
    assert rel_err(o['pos'], base['pos'] @ R.T) < t2 and rel_err(o['vel'], base['vel'] @ R.T) < t2
    assert rel_err(o['h'], base['h']) < t2 and rel_err(o['g'], base['g']) < t2
    tr = dict(arrs)
    tr['pos'] = arrs['pos'] + np.array([0.3, -0.2, 0.1])
    o, ldj, _ = run(tr, eps)
    assert abs(ldj - ldj0) < 10 * tol * max(abs(ldj0), 1)
    assert rel_err(o['pos'] - np.array([0.3, -0.2, 0.1]), base['pos']) < t2
    # permute atoms inside every molecule
    perm, o0 = [], 0
    for n in arrs['N']:
        perm.append(o0 + rs.permutation(int(n)))
        o0 += int(n)
    perm = np.concatenate(perm)
    pm = {k: (v[perm] if k in ('h', 'g', 'pos', 'vel', 'box') else v) for k, v in arrs.items()}
    o, ldj, ldjm = run(pm, eps[perm])
    assert abs(ldj - ldj0) < 10 * tol * max(abs(ldj0), 1)
    assert rel_err(ldjm, ldjm0) < 10 * tol
    for k in ('h', 'g', 'pos', 'vel'):
        assert rel_err(o[k], base[k][perm]) < t2, k


@modes
def test_edge_cases_single_atom_and_tiny_molecules(mode, tol, gtol):
    """Molecules with 1 atom (no edges, mean count clamps to 1: Q12) next to 2- and 3-atom ones: in the tensor-core
    modes these are tiles that are almost all padding."""
    from enflow_b200.data import synthetic as syn
    from enflow_b200.flow.loss import Alchemical_NLL
    nf, L = 5, 2
    parts = [syn.make_batch('c2', 1, n_atoms=n, seed=100 + n) for n in (1, 2, 3, 1, 7)]
    arrs = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    sd = syn.make_weights(nf, 128, L, seed=4, coord_gain=0.5)
    eps = syn.make_noise(int(arrs['N'].sum()), nf, seed=6)
    model = build_model(sd, nf, L, precision=mode)
    out, ldj = model(gpu_batch(arrs), eps=torch.as_tensor(eps))
    loss = Alchemical_NLL(kBT=syn.TRAIN_KBT, softening=0.1)(out, ldj)
    loss.backward()
    ref_loss, ref_grads, ref_state, _, _ = orc.train_step(sd, L, arrs, syn.TRAIN_DT, eps, syn.TRAIN_KBT, 0.1)
    assert abs(loss.item() - ref_loss.item()) <= tol * abs(ref_loss.item())
    for k in ('h', 'g', 'pos', 'vel'):
        check_close(to_np(getattr(out, k)), ref_state[k].numpy(), tol, f'{mode} {k}')
    worst = grad_errors(model, ref_grads)
    assert worst[0][0] < gtol, f'{mode} worst gradient errors: {worst[:5]}'


@pytest.mark.parametrize('n_mols', [1, 3])
@modes
def test_only_single_atom_molecules(n_mols, mode, tol, gtol):
    """No edge at all in the batch: every edge kernel sees an empty list."""
    from enflow_b200.data import synthetic as syn
    from enflow_b200.flow.loss import Alchemical_NLL
    nf, L = 5, 2
    parts = [syn.make_batch('c2', 1, n_atoms=1, seed=200 + i) for i in range(n_mols)]
    arrs = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    sd = syn.make_weights(nf, 128, L, seed=4, coord_gain=0.5)
    eps = syn.make_noise(int(arrs['N'].sum()), nf, seed=6)
    model = build_model(sd, nf, L, precision=mode)
    out, ldj = model(gpu_batch(arrs), eps=torch.as_tensor(eps))
    loss = Alchemical_NLL(kBT=syn.TRAIN_KBT, softening=0.1)(out, ldj)
    loss.backward()
    ref_loss, ref_grads, ref_state, _, _ = orc.train_step(sd, L, arrs, syn.TRAIN_DT, eps, syn.TRAIN_KBT, 0.1)
    assert abs(loss.item() - ref_loss.item()) <= tol * abs(ref_loss.item())
    for k in ('h', 'g', 'pos', 'vel'):
        check_close(to_np(getattr(out, k)), ref_state[k].numpy(), tol, f'{mode} {k}')
    assert torch.isfinite(model.flat_grads).all()


@modes
def test_edge_capacity_overflow_is_detected_and_retried(mode, tol, gtol):
    """A too-small edge capacity sets the device status flag; the host doubles the capacity and redoes the pass."""
    c = load_case('c1_pbc')
    model = build_model(c['sd'], c['nf'], c['L'], precision=mode)
    data = gpu_batch(c['batch'])
    B, N = int(c['batch']['N'].shape[0]), int(c['batch']['N'].sum())
    model._edge_caps[(B, N)] = 128                      # far below the ~670 edges of this batch
    with torch.no_grad():
        out, ldj = model(data, eps=torch.as_tensor(c['eps']))
    assert model._edge_caps[(B, N)] >= 1024
    for k in ('h', 'g', 'pos', 'vel'):
        check_close(to_np(getattr(out, k)), c['gold'][f'out_{k}'], tol, f'{mode} {k}')
