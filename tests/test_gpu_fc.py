"""The implicit all-pairs path of the fully connected regime (csrc/fc.cu) against K0, which carries the reference's
Data.edges semantics bit for bit (base.py:122-144; test_edges_bit_exact_vs_golden): same list, same column permutation,
the regime check accepts exactly the cases where K0 returns all pairs, and a flow that assumes the regime gives bitwise the
results of one that runs K0 at every layer."""
import numpy as np
import pytest
import torch

from enflow_b200 import _lib
from golden_util import load_case
from gpu_util import DEV, build_model, gpu_batch, to_np

pytestmark = pytest.mark.gpu
I32 = dict(dtype=torch.int32, device=DEV)


def _fc_lists(data, e_cap):
    L = _lib.lib()
    B, off, _, _ = data.meta()
    N = int(data.pos.shape[0])
    row, col, perm = (torch.full((e_cap,), -1, **I32) for _ in range(3))
    rowptr, colptr = torch.empty(N + 1, **I32), torch.empty(N + 1, **I32)
    e_dev, status, eoff = torch.zeros(2, **I32), torch.zeros(1, **I32), torch.empty(B + 2, **I32)
    p = _lib.ptr
    _lib.check(L.enflow_fc_build(p(off), B, N, e_cap, p(row), p(col), p(rowptr), p(e_dev), p(colptr), p(perm), p(eoff),
                                 p(status), _lib.stream()))
    return row, col, rowptr, colptr, perm, e_dev, status


def _fc_check(data):
    L = _lib.lib()
    B, off, _, _ = data.meta()
    status = torch.zeros(1, **I32)
    p = _lib.ptr
    rc = data.r_cut.detach().to(DEV, torch.float32).reshape(-1).contiguous()
    _lib.check(L.enflow_fc_check(p(_lib.f32c(data.pos)), p(_lib.f32c(data.box)), p(rc), p(off), B, p(status), _lib.stream()))
    return int(status.item())


def _k0_col_perm(data, csr):
    L = _lib.lib()
    B, off, _, _ = data.meta()
    N = int(data.pos.shape[0])
    row, col, rowptr, e_dev = csr
    E = int(row.numel())
    colptr, perm = torch.empty(N + 1, **I32), torch.empty(max(E, 1), **I32)
    ws = torch.empty(L.enflow_edges_workspace_ints(N), **I32)
    p = _lib.ptr
    _lib.check(L.enflow_build_col_perm(p(col), p(rowptr), p(off), B, N, E, p(e_dev), p(colptr), p(perm), p(ws), None, _lib.stream()))
    return colptr, perm[:E]


@pytest.mark.parametrize('name', ['c2_ragged', 'c2_default_init', 'c3_lj55'])
def test_fc_list_equals_k0_bit_for_bit(name):
    c = load_case(name)
    data = gpu_batch(c['batch'], dtype=torch.float32)
    e = data.build_edges(reference_order=False)
    row, col, rowptr, e_dev = e.csr
    E = int(row.numel())
    assert E == int((c['batch']['N'] * (c['batch']['N'] - 1)).sum())
    assert _fc_check(data) == 0
    frow, fcol, frowptr, fcolptr, fperm, fe, st = _fc_lists(data, E)
    assert int(st.item()) == 0 and fe.tolist() == [E, E]
    assert torch.equal(frow, row) and torch.equal(fcol, col) and torch.equal(frowptr, rowptr)
    colptr, perm = _k0_col_perm(data, e.csr)
    assert torch.equal(fcolptr, colptr) and torch.equal(fperm, perm)


def test_fc_list_with_single_atom_molecules_and_capacity_flag():
    from enflow_b200.data import synthetic as syn
    parts = [syn.make_batch('c2', 1, n_atoms=n, seed=100 + n) for n in (1, 2, 3, 1, 7)]
    arrs = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    data = gpu_batch(arrs, dtype=torch.float32)
    e = data.build_edges(reference_order=False)
    row, col, rowptr, _ = e.csr
    E = int(row.numel())
    frow, fcol, frowptr, fcolptr, fperm, fe, st = _fc_lists(data, E)
    assert _fc_check(data) == 0 and int(st.item()) == 0
    assert torch.equal(frow, row) and torch.equal(fcol, col) and torch.equal(frowptr, rowptr)
    colptr, perm = _k0_col_perm(data, e.csr)
    assert torch.equal(fcolptr, colptr) and torch.equal(fperm, perm)
    *_, fe, st = _fc_lists(data, E - 5)                   # too small a capacity: flagged like K0 does
    assert int(st.item()) & 1 and fe.tolist() == [E - 5, E]


@pytest.mark.parametrize('name', ['c1_pbc', 'c5_small', 'c5_radius'])
def test_fc_check_rejects_the_radius_graph_and_pbc_cases(name):
    c = load_case(name)
    assert _fc_check(gpu_batch(c['batch'], dtype=torch.float32)) & 8


def test_fc_check_is_exactly_as_strict_as_needed():
    """Perturbations that take a molecule out of the regime: the cut-off just below the largest pair distance, a box that
    lets a periodic image come within reach, a box small enough for the ellipsoid pre-filter to split an image."""
    from enflow_b200.data import synthetic as syn
    arrs = syn.make_batch('c2', 4, n_atoms=11, seed=9)
    assert _fc_check(gpu_batch(arrs, dtype=torch.float32)) == 0
    pos = arrs['pos'][:11].astype(np.float32).astype(np.float64)
    dmax = np.sqrt(((pos[:, None] - pos[None]) ** 2).sum(-1).max())
    for rc, flagged in ((dmax * 0.999, True), (dmax * 1.001, False)):
        a = dict(arrs)
        a['r_cut'] = arrs['r_cut'].copy()
        a['r_cut'][0] = rc
        data = gpu_batch(a, dtype=torch.float32)
        assert bool(_fc_check(data) & 8) == flagged
        E = int(data.build_edges(reference_order=False).row.numel())
        assert (E == 4 * 110) == (not flagged)          # K0 agrees: all pairs exactly when the check passes
    ext = (pos.max(0) - pos.min(0)).max()
    a = dict(arrs)
    a['box'] = arrs['box'].copy()
    a['box'][:11] = ext + 50.0                           # r_cut = 100 reaches the periodic images
    assert _fc_check(gpu_batch(a, dtype=torch.float32)) & 8


@pytest.mark.parametrize('mode', ['fp32_tc', 'fp32', 'bf16'])
def test_flow_on_the_implicit_list_is_bitwise_the_flow_on_k0(mode):
    from enflow_b200.flow.loss import Alchemical_NLL
    c = load_case('c2_ragged')
    nll = Alchemical_NLL(kBT=c['kBT'], softening=c['softening'])
    res = []
    for use_fc in (True, False):
        m = build_model(c['sd'], c['nf'], c['L'], precision=mode)
        data = gpu_batch(c['batch'])
        if not use_fc:
            m._capacity(data, __import__('enflow_b200.flow.dynamics', fromlist=['_prep'])._prep(data))
            m._fc_keys = {k: (False, 0) for k in m._fc_keys}
        out, ldj = m(data, eps=torch.as_tensor(c['eps']))
        assert m._fc_now == use_fc
        nll(out, ldj).backward()
        res.append((out.pos.detach().clone(), out.g.detach().clone(), ldj.detach().clone(), m.flat_grads.clone()))
        back = m.reverse(out, quantize=False)
        res[-1] += (back.pos.clone(), back.neg_ldj_mol.clone())
    for a, b in zip(*res):
        assert torch.equal(a, b)


def test_wrong_fc_assumption_falls_back_to_k0():
    """A layout wrongly remembered as fully connected: the device check raises status bit 8 and the pass is redone on K0."""
    c = load_case('c1_pbc')
    m = build_model(c['sd'], c['nf'], c['L'])
    data = gpu_batch(c['batch'])
    B, N = int(c['batch']['N'].shape[0]), int(c['batch']['N'].sum())
    fc = int((c['batch']['N'] * (c['batch']['N'] - 1)).sum())
    m._edge_caps[(B, N)] = fc
    m._fc_keys[(B, N)] = (True, fc)
    with torch.no_grad():
        out, ldj = m(data, eps=torch.as_tensor(c['eps']))
    assert m._fc_keys[(B, N)] == (False, 0) and not m._fc_now
    for k in ('h', 'g', 'pos', 'vel'):
        err = np.abs(to_np(getattr(out, k)) - c['gold'][f'out_{k}']).max() / np.abs(c['gold'][f'out_{k}']).max()
        assert err < 1e-5, k


@pytest.mark.parametrize('name', ['c1_pbc', 'c5_small', 'c5_radius', 'c2_ragged'])
def test_col_perm_is_the_stable_column_grouping(name):
    """enflow_build_col_perm on radius-graph lists (duplicate (row, col) pairs from periodic images included, Q11): the
    column-grouped permutation equals numpy's stable argsort of col, twice in a row (the slots are handed out by integer
    atomics and each column's list is sorted afterwards, so the result must not depend on their order)."""
    c = load_case(name)
    data = gpu_batch(c['batch'], dtype=torch.float32)
    e = data.build_edges(reference_order=False)
    col = e.csr[1].cpu().numpy().astype(np.int64)
    ref = np.argsort(col, kind='stable').astype(np.int64)
    for _ in range(2):
        colptr, perm = _k0_col_perm(data, e.csr)
        assert np.array_equal(perm.cpu().numpy().astype(np.int64), ref)
        counts = np.bincount(col, minlength=int(data.pos.shape[0]))
        assert np.array_equal(np.diff(colptr.cpu().numpy().astype(np.int64)), counts)
