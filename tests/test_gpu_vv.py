"""VVIntegrator (dynamics.py:39-86) is dead code upstream: no oracle exists.  The fixed-forward plugin is validated by the
properties that need none: exact inverse, E(3) / permutation equivariance in the well-defined regime, determinism."""
import numpy as np
import pytest
import torch

from gpu_util import DEV, gpu_batch, rel_err, to_np

pytestmark = pytest.mark.gpu


def _vv(nf, L, precision='fp32_tc', seed=3):
    from enflow_b200.data import synthetic as syn
    from enflow_b200.flow.dynamics import VVIntegrator
    from enflow_b200.nn.argmax import ArgMax
    from enflow_b200.nn.egcl import EGCL
    sd = syn.make_weights(nf, 128, L, seed=seed, coord_gain=0.5)
    m = VVIntegrator([EGCL(nf, nf, 128) for _ in range(L)], ArgMax(nf, 128), dt=syn.TRAIN_DT)
    m.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    m = m.to(DEV)
    m.precision = precision
    return m


@pytest.mark.parametrize('config,kw', [('c2', dict(ragged=True)), ('c1', {})])
def test_vv_round_trip(config, kw):
    from enflow_b200.data import synthetic as syn
    arrs = syn.make_batch(config, 5, seed=21, **kw)
    # start inside the primary cell: the inverse evaluates network 0 on WRAPPED positions (data.pbc()), and the reference's
    # neighbour list is not invariant under box translations, so an unwrapped start has no exact inverse
    arrs['pos'] = arrs['pos'] - np.round(arrs['pos'] / arrs['box']) * arrs['box']
    nf = arrs['h'].shape[1]
    m = _vv(nf, 4)
    assert m.n_iter == 3
    eps = torch.as_tensor(syn.make_noise(int(arrs['N'].sum()), nf, seed=5))
    with torch.no_grad():
        out, ldj = m(gpu_batch(arrs, dtype=torch.float64), eps=eps)
        assert torch.isfinite(ldj) and abs(float(ldj)) > 0
        moved = rel_err(to_np(out.pos), arrs['pos'])
        back = m.reverse(out, quantize=True)
    assert moved > 1e-4                                   # the map is not the identity ...
    for k in ('vel', 'g'):                                # ... and its inverse undoes it
        assert rel_err(to_np(getattr(back, k)), arrs[k]) < 2e-5, k
    diff = to_np(back.pos) - arrs['pos']                  # positions come back up to the periodic wrap (data.pbc())
    diff = diff - np.round(diff / arrs['box']) * arrs['box']
    assert np.abs(diff).max() < 2e-5 * np.abs(arrs['pos']).max()
    assert np.array_equal(to_np(back.h), arrs['h'])


def test_vv_equivariance_and_determinism():
    from enflow_b200.data import synthetic as syn
    nf = 5
    arrs = syn.make_batch('c2', 4, ragged=True, seed=77)
    eps = syn.make_noise(int(arrs['N'].sum()), nf, seed=5)
    m = _vv(nf, 3)

    def run(a, e):
        with torch.no_grad():
            out, ldj = m(gpu_batch(a, dtype=torch.float64), eps=torch.as_tensor(e))
        return {k: to_np(getattr(out, k)) for k in ('h', 'g', 'pos', 'vel')}, float(ldj)

    base, ldj0 = run(arrs, eps)
    again, ldj1 = run(arrs, eps)
    assert ldj0 == ldj1 and all(np.array_equal(base[k], again[k]) for k in base)
    rs = np.random.RandomState(11)
    q, r = np.linalg.qr(rs.normal(size=(3, 3)))
    R = q * np.sign(np.diag(r))
    if np.linalg.det(R) < 0:
        R[:, 0] = -R[:, 0]
    rot = dict(arrs)
    rot['pos'], rot['vel'] = arrs['pos'] @ R.T, arrs['vel'] @ R.T
    o, ldj = run(rot, eps)
    assert abs(ldj - ldj0) < 1e-4 * max(abs(ldj0), 1)
    assert rel_err(o['pos'], base['pos'] @ R.T) < 2e-5 and rel_err(o['vel'], base['vel'] @ R.T) < 2e-5
    assert rel_err(o['g'], base['g']) < 2e-5
    perm, o0 = [], 0
    for n in arrs['N']:
        perm.append(o0 + rs.permutation(int(n)))
        o0 += int(n)
    perm = np.concatenate(perm)
    pm = {k: (v[perm] if k in ('h', 'g', 'pos', 'vel', 'box') else v) for k, v in arrs.items()}
    o, ldj = run(pm, eps[perm])
    assert abs(ldj - ldj0) < 1e-4 * max(abs(ldj0), 1)
    for k in ('h', 'g', 'pos', 'vel'):
        assert rel_err(o[k], base[k][perm]) < 2e-5, k


def test_vv_is_inference_only():
    from enflow_b200.data import synthetic as syn
    arrs = syn.make_batch('c2', 2, n_atoms=6, seed=1)
    m = _vv(5, 2)
    with pytest.raises(NotImplementedError):
        m(gpu_batch(arrs, dtype=torch.float32))
