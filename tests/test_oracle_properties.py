"""Size-independent properties of the oracle itself (CPU, fp64): the same properties the GPU path is held to in
tests/test_gpu_parity.py / test_gpu_fullsize.py, checked here on the checker so that a property failure on the GPU
cannot be blamed on the oracle.  Reference behaviour cited: enflow/flow/dynamics.py:10-37, enflow/nn/egcl.py:57-93,
enflow/data/base.py:122-144."""
import numpy as np
import torch

from enflow_b200.data import synthetic as syn
from oracle import enflow_oracle as orc

NF, L = 5, 3


def _setup(config='c2', B=3, seed=21, **kw):
    arrs = syn.make_batch(config, B, seed=seed, **kw)
    nf = arrs['h'].shape[1]
    sd = syn.make_weights(nf, 128, L, seed=2)
    eps = syn.make_noise(int(arrs['N'].sum()), nf, seed=3)
    return arrs, sd, eps, nf


def test_forward_then_reverse_is_identity():
    arrs, sd, eps, nf = _setup(ragged=True)
    p = orc.params_to_torch(sd)
    state, ldj, ldj_mol = orc.lf_forward(p, L, orc.to_torch(arrs), syn.TRAIN_DT, torch.as_tensor(eps))
    lat = dict(arrs)
    lat.update({k: state[k].numpy() for k in ('h', 'g', 'pos', 'vel')})
    back = orc.lf_reverse(p, L, orc.to_torch(lat), syn.TRAIN_DT, quantize=True)
    back = back[0] if isinstance(back, tuple) else back
    for k in ('pos', 'vel', 'g'):
        assert np.abs(back[k].numpy() - arrs[k]).max() < 1e-9, k
    assert np.array_equal(back['h'].numpy(), arrs['h'])          # one-hot again after ArgMax.reverse
    assert abs(float(ldj) - (float(ldj_mol.sum()) + float(ldj - ldj_mol.sum()))) < 1e-12


def test_molecules_are_independent():
    arrs, sd, eps, nf = _setup(B=4)
    p = orc.params_to_torch(sd)
    full, _, ldj_mol = orc.lf_forward(p, L, orc.to_torch(arrs), syn.TRAIN_DT, torch.as_tensor(eps))
    off = np.concatenate([[0], np.cumsum(arrs['N'])])
    m = 2
    sl = slice(off[m], off[m + 1])
    one = {k: arrs[k][sl] for k in ('h', 'g', 'pos', 'vel', 'box')}
    one['N'], one['r_cut'] = arrs['N'][m:m + 1], arrs['r_cut'][m:m + 1]
    alone, _, ldj_one = orc.lf_forward(p, L, orc.to_torch(one), syn.TRAIN_DT, torch.as_tensor(eps[sl]))
    for k in ('pos', 'vel', 'h', 'g'):
        assert np.abs(alone[k].numpy() - full[k].numpy()[sl]).max() < 1e-12, k
    assert abs(float(ldj_one[0]) - float(ldj_mol[m])) < 1e-12


def test_rotation_and_translation_equivariance_fully_connected():
    """Fully connected regime (box and r_cut far larger than the molecule): positions and velocities rotate with the
    input, scalars (h, g, per-molecule log-det) do not change; translations only shift positions."""
    arrs, sd, eps, nf = _setup(B=2)
    p = orc.params_to_torch(sd)
    rs = np.random.RandomState(5)
    q, r = np.linalg.qr(rs.normal(size=(3, 3)))
    q = q * np.sign(np.diag(r))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    shift = np.array([0.3, -0.2, 0.1])
    rot = dict(arrs)
    rot['pos'] = arrs['pos'] @ q.T + shift
    rot['vel'] = arrs['vel'] @ q.T
    a, _, la = orc.lf_forward(p, L, orc.to_torch(arrs), syn.TRAIN_DT, torch.as_tensor(eps))
    b, _, lb = orc.lf_forward(p, L, orc.to_torch(rot), syn.TRAIN_DT, torch.as_tensor(eps))
    assert np.abs(b['pos'].numpy() - (a['pos'].numpy() @ q.T + shift)).max() < 1e-9
    assert np.abs(b['vel'].numpy() - a['vel'].numpy() @ q.T).max() < 1e-9
    assert np.abs(b['h'].numpy() - a['h'].numpy()).max() < 1e-9
    assert np.abs(b['g'].numpy() - a['g'].numpy()).max() < 1e-9
    assert np.abs(lb.numpy() - la.numpy()).max() < 1e-9


def test_atom_permutation_equivariance():
    arrs, sd, eps, nf = _setup(B=1)
    p = orc.params_to_torch(sd)
    n = int(arrs['N'][0])
    perm = np.random.RandomState(9).permutation(n)
    pa = {k: (arrs[k][perm] if k in ('h', 'g', 'pos', 'vel', 'box') else arrs[k]) for k in arrs}
    a, _, la = orc.lf_forward(p, L, orc.to_torch(arrs), syn.TRAIN_DT, torch.as_tensor(eps))
    b, _, lb = orc.lf_forward(p, L, orc.to_torch(pa), syn.TRAIN_DT, torch.as_tensor(eps[perm]))
    for k in ('pos', 'vel', 'h', 'g'):
        assert np.abs(b[k].numpy() - a[k].numpy()[perm]).max() < 1e-9, k
    assert abs(float(lb[0]) - float(la[0])) < 1e-9


def test_edge_list_shapes_and_degenerate_molecules():
    # single atom: no edges; two atoms within the cutoff: both directions; beyond the cutoff: none
    box = torch.full((1, 3), 1000.0, dtype=torch.float64)
    row, col, _ = orc.build_edges(torch.zeros(1, 3, dtype=torch.float64), box, torch.tensor([1]), torch.tensor([100.0]))
    assert row.numel() == 0 and col.numel() == 0
    pos = torch.tensor([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]], dtype=torch.float64)
    box2 = torch.full((2, 3), 1000.0, dtype=torch.float64)
    row, col, _ = orc.build_edges(pos, box2, torch.tensor([2]), torch.tensor([100.0]))
    assert sorted(zip(row.tolist(), col.tolist())) == [(0, 1), (1, 0)]
    row, col, _ = orc.build_edges(pos, box2, torch.tensor([2]), torch.tensor([0.5]))
    assert row.numel() == 0
    # two molecules never share an edge
    pos4 = torch.cat([pos, pos + 0.1])
    row, col, _ = orc.build_edges(pos4, torch.full((4, 3), 1000.0, dtype=torch.float64), torch.tensor([2, 2]),
                                  torch.tensor([100.0, 100.0]))
    assert all((r < 2) == (c < 2) for r, c in zip(row.tolist(), col.tolist())) and row.numel() == 4
