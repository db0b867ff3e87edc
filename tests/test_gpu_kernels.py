"""Unit parity of the individual C-ABI kernels against plain torch fp64 restatements."""
import numpy as np
import pytest
import torch

from enflow_b200 import _lib
from gpu_util import DEV, rel_err, to_np

pytestmark = pytest.mark.gpu


def _csr(rs, N, max_deg):
    deg = rs.randint(0, max_deg + 1, size=N)
    ptr = np.zeros(N + 1, dtype=np.int32)
    ptr[1:] = np.cumsum(deg)
    return ptr, int(ptr[-1])


@pytest.mark.parametrize('silu', [0, 1])
def test_segment_sum128(silu):
    rs = np.random.RandomState(0)
    N = 1000
    ptr, E = _csr(rs, N, 40)
    x = rs.normal(size=(E, 128)).astype(np.float32)
    xt = torch.tensor(x, device=DEV)
    out = torch.empty(N, 128, dtype=torch.float32, device=DEV)
    L = _lib.lib()
    _lib.check(L.enflow_segment_sum128(_lib.ptr(xt), _lib.ptr(torch.tensor(ptr, device=DEV)), None, N, E, silu,
                                       _lib.ptr(out), _lib.stream()))
    seg = np.repeat(np.arange(N), np.diff(ptr))
    xd = torch.tensor(x, dtype=torch.float64)
    if silu:
        xd = xd * torch.sigmoid(xd)
    ref = torch.zeros(N, 128, dtype=torch.float64).index_add_(0, torch.tensor(seg), xd)
    assert rel_err(to_np(out), ref.numpy()) < 2e-6


def test_segment_sum128_perm_and_mean3():
    rs = np.random.RandomState(1)
    N = 300
    ptr, E = _csr(rs, N, 25)
    perm = rs.permutation(E).astype(np.int32)
    x = rs.normal(size=(E, 128)).astype(np.float32)
    L = _lib.lib()
    out = torch.empty(N, 128, dtype=torch.float32, device=DEV)
    xt, pt, pm = torch.tensor(x, device=DEV), torch.tensor(ptr, device=DEV), torch.tensor(perm, device=DEV)
    _lib.check(L.enflow_segment_sum128(_lib.ptr(xt), _lib.ptr(pt), _lib.ptr(pm), N, E, 0, _lib.ptr(out), _lib.stream()))
    seg = np.repeat(np.arange(N), np.diff(ptr))
    ref = torch.zeros(N, 128, dtype=torch.float64).index_add_(0, torch.tensor(seg), torch.tensor(x[perm], dtype=torch.float64))
    assert rel_err(to_np(out), ref.numpy()) < 2e-6
    v = rs.normal(size=(E, 3)).astype(np.float32)
    o3 = torch.ones(N, 3, dtype=torch.float32, device=DEV)
    _lib.check(L.enflow_segment_sum3(_lib.ptr(torch.tensor(v, device=DEV)), _lib.ptr(pt), None, N, E, 1, 0.5, 0,
                                     _lib.ptr(o3), _lib.stream()))
    tot = torch.zeros(N, 3, dtype=torch.float64).index_add_(0, torch.tensor(seg), torch.tensor(v, dtype=torch.float64))
    cnt = np.maximum(np.diff(ptr), 1)[:, None]
    assert rel_err(to_np(o3), 0.5 * tot.numpy() / cnt) < 2e-6


def test_coupling_forward_inverse_and_backward():
    rs = np.random.RandomState(2)
    sizes = np.array([1, 5, 33, 70, 2])
    N, B, nf, dt = int(sizes.sum()), len(sizes), 4, 0.05
    off = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32), device=DEV)
    mk = lambda *s: torch.tensor(rs.normal(size=s).astype(np.float32), device=DEV)
    Q, F, G, h, g, pos, vel = mk(N), mk(N, 3), mk(N, nf), mk(N, nf), mk(N, nf), mk(N, 3), mk(N, 3)
    box = torch.full((N, 3), 1.7, dtype=torch.float32, device=DEV)
    ho, go, po, vo = (torch.empty_like(t) for t in (h, g, pos, vel))
    ldj_mol = torch.zeros(B, dtype=torch.float32, device=DEV)
    L = _lib.lib()
    p = _lib.ptr
    _lib.check(L.enflow_coupling_fwd(p(Q), p(F), p(G), p(h), p(g), p(pos), p(vel), p(box), p(off), B, nf, dt, p(ho),
                                     p(go), p(po), p(vo), p(ldj_mol), _lib.stream()))
    d = lambda t: t.double().cpu()
    vel_r = torch.exp(d(Q))[:, None] * d(vel) + d(F) * dt
    g_r = d(g) + d(G) * dt
    pos_r = d(pos) + vel_r * dt
    pos_r = pos_r - (pos_r / d(box)).round() * d(box)
    h_r = d(h) + g_r * dt
    for a, b in ((vo, vel_r), (go, g_r), (po, pos_r), (ho, h_r)):
        assert rel_err(to_np(a), b.numpy()) < 2e-6
    ref_ldj = np.array([d(Q)[int(off[m]):int(off[m + 1])].sum().item() for m in range(B)])
    assert rel_err(to_np(ldj_mol), ref_ldj) < 2e-6
    # exact inverse
    h2, p2, g2, v2 = ho.clone(), po.clone(), go.clone(), vo.clone()
    _lib.check(L.enflow_coupling_inv_pre(p(g2), p(v2), p(box), N, nf, dt, p(h2), p(p2), _lib.stream()))
    neg = torch.zeros(B, dtype=torch.float32, device=DEV)
    _lib.check(L.enflow_coupling_inv_post(p(Q), p(F), p(G), p(off), B, nf, dt, p(g2), p(v2), p(neg), _lib.stream()))
    assert rel_err(to_np(h2), to_np(h)) < 1e-5 and rel_err(to_np(g2), to_np(g)) < 1e-5 and rel_err(to_np(v2), to_np(vel)) < 1e-5
    wrapped = d(pos) - (d(pos) / d(box)).round() * d(box)
    diff = to_np(p2) - wrapped.numpy()
    diff = diff - np.round(diff / 1.7) * 1.7
    assert np.abs(diff).max() < 1e-5
    assert rel_err(to_np(neg), -ref_ldj) < 2e-6
    # backward against autograd
    Qa, Fa, Ga, ha, ga, pa, va = (d(t).clone().requires_grad_(True) for t in (Q, F, G, h, g, pos, vel))
    vel_a = torch.exp(Qa)[:, None] * va + Fa * dt
    g_a = ga + Ga * dt
    pos_a = pa + vel_a * dt
    pos_a = pos_a - (pos_a / d(box)).round() * d(box)
    h_a = ha + g_a * dt
    up = [torch.tensor(rs.normal(size=t.shape)) for t in (h_a, g_a, pos_a, vel_a)]
    dl = 0.37
    ((h_a * up[0]).sum() + (g_a * up[1]).sum() + (pos_a * up[2]).sum() + (vel_a * up[3]).sum() + dl * Qa.sum()).backward()
    dh, dg, dpos, dvel = (u.float().to(DEV).contiguous() for u in up)
    dQ, dF, dG = torch.empty(N, device=DEV), torch.empty(N, 3, device=DEV), torch.empty(N, nf, device=DEV)
    dlt = torch.tensor([dl], dtype=torch.float32, device=DEV)
    _lib.check(L.enflow_coupling_bwd(p(Q), p(vel), p(dlt), N, nf, dt, p(dh), p(dg), p(dpos), p(dvel), p(dQ), p(dF), p(dG),
                                     _lib.stream()))
    for a, b in ((dQ, Qa.grad), (dF, Fa.grad), (dG, Ga.grad), (dh, ha.grad), (dg, ga.grad), (dpos, pa.grad), (dvel, va.grad)):
        assert rel_err(to_np(a), b.numpy()) < 5e-6


def test_c_abi_error_path():
    L = _lib.lib()
    dims = _lib.Dims(1, 1, 99, 1, 1, 1, 0.1, 1.0, 0, 0)
    import ctypes
    assert L.enflow_flow_workspace_bytes(ctypes.byref(dims), 0) == 0
    assert b'nf=99' in L.enflow_last_error()
