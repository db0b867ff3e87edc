"""Pin the CPU oracle against outputs of the reference itself (tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

from golden_util import CASES, load_case, projection_vectors
from oracle import enflow_oracle as orc


@pytest.mark.parametrize('name', list(CASES))
def test_oracle_matches_reference(name):
    c = load_case(name)
    gold, L = c['gold'], c['L']
    p = orc.params_to_torch(c['sd'], requires_grad=True)
    b = orc.to_torch(c['batch'])
    trace = []
    state, ldj, ldj_mol = orc.lf_forward(p, L, b, c['dt'], torch.as_tensor(c['eps']), trace=trace)
    loss = orc.alchemical_nll(state, ldj, b['N'], c['kBT'], c['softening'])
    loss.backward()
    for i in range(L):
        # bit-exact edge construction, reference order (enflow/data/base.py:122-144)
        assert np.array_equal(trace[i]['row'].numpy(), gold[f'row{i}'])
        assert np.array_equal(trace[i]['col'].numpy(), gold[f'col{i}'])
        for k in 'QFG':
            np.testing.assert_allclose(trace[i][k].detach().numpy(), gold[f'{k}{i}'], rtol=1e-12, atol=1e-14)
    for k in ('h', 'g', 'pos', 'vel'):
        np.testing.assert_allclose(state[k].detach().numpy(), gold[f'out_{k}'], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(ldj.item(), gold['ldj'], rtol=1e-12)
    np.testing.assert_allclose(loss.item(), gold['loss'], rtol=1e-12)
    np.testing.assert_allclose(ldj_mol.sum().item() , (ldj - orc.argmax_forward(p, b['h'], torch.as_tensor(c['eps']))[1]).item(),
                               rtol=1e-9, atol=1e-9)
    names = list(c['sd'].keys())
    proj = projection_vectors(c['sd'])
    gs = np.array([p[k].grad.sum().item() for k in names])
    gn = np.array([p[k].grad.norm().item() for k in names])
    gp = np.array([(p[k].grad.numpy() * proj[k]).sum() for k in names])
    scale = np.maximum(gold['grad_norm'], 1e-300)
    np.testing.assert_allclose(gn / scale, gold['grad_norm'] / scale, rtol=1e-9)
    np.testing.assert_allclose(gs / scale, gold['grad_sum'] / scale, atol=1e-8)
    np.testing.assert_allclose(gp / scale, gold['grad_proj'] / scale, atol=1e-8)
    for k in names:
        if 'grad/' + k in gold:
            np.testing.assert_allclose(p[k].grad.numpy(), gold['grad/' + k], rtol=2e-6, atol=1e-6 * scale[names.index(k)])


@pytest.mark.parametrize('name', ['c1_pbc', 'c2_ragged', 'c3_lj55'])
def test_oracle_reverse_matches_reference(name):
    c = load_case(name)
    gold, L = c['gold'], c['L']
    p = orc.params_to_torch(c['sd'])
    b = orc.to_torch(c['batch'])
    lat = dict(b)
    for k in ('h', 'g', 'pos', 'vel'):
        lat[k] = torch.as_tensor(gold[f'out_{k}'])
    back, neg_ldj_mol = orc.lf_reverse(p, L, lat, c['dt'])
    for k in ('h', 'g', 'pos', 'vel'):
        np.testing.assert_allclose(back[k].numpy(), gold[f'rev_{k}'], rtol=1e-12, atol=1e-13)
    # the reference's own self-check (enflow/main.py:275-278): reverse(forward(x)) == x for pos
    # (modulo the box: inputs outside [-box/2, box/2] come back wrapped)
    diff = back['pos'].numpy() - c['batch']['pos']
    box = c['batch']['box']
    np.testing.assert_allclose(diff - np.round(diff / box) * box, 0, atol=1e-8)
    np.testing.assert_allclose(back['h'].numpy(), c['batch']['h'], atol=0)
