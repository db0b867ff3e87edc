"""Behaviour of the reference-facing module API beyond numerics: dequantiser dispatch (dynamics.py:11,35), the stand-alone
modules refusing to drop gradients silently, workspace ownership across pending backward passes."""
import numpy as np
import pytest
import torch

from gpu_util import DEV, build_model, gpu_batch, rel_err, to_np

pytestmark = pytest.mark.gpu


def _floor_flow(nf, L, sd):
    from enflow_b200.data import synthetic as syn
    from enflow_b200.flow.dynamics import LFIntegrator
    from enflow_b200.nn.egcl import EGCL
    from enflow_b200.nn.floor import Floor
    m = LFIntegrator([EGCL(nf, nf, 128) for _ in range(L)], Floor(1.0), dt=syn.TRAIN_DT)
    m.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items() if not k.startswith('dequantize.')})
    return m.to(DEV)


def test_flow_built_with_floor_runs_floor_not_argmax():
    """`data.h, ldj = self.dequantize(data.h)` (dynamics.py:11) whatever the dequantiser is (nn/floor.py:5-14)."""
    from enflow_b200.data import synthetic as syn
    nf, L = 5, 2
    arrs = syn.make_batch('c2', 3, n_atoms=9, seed=5)
    sd = syn.make_weights(nf, 128, L, seed=2, coord_gain=0.5)
    flow = _floor_flow(nf, L, sd)
    torch.manual_seed(7)
    with torch.no_grad():
        out, ldj = flow(gpu_batch(arrs, dtype=torch.float32))
    # the same uniform noise applied by hand, then the coupling stack without a dequantiser
    torch.manual_seed(7)
    d2 = gpu_batch(arrs, dtype=torch.float32)
    d2.h = d2.h + torch.rand_like(d2.h)
    with torch.no_grad():
        ref, ldj2 = flow(d2, dequantize=False)
    for k in ('h', 'g', 'pos', 'vel'):
        assert torch.equal(getattr(out, k), getattr(ref, k)), k
    assert torch.equal(ldj, ldj2)                   # Floor contributes 0 to the log-det (floor.py:12)
    h0 = torch.as_tensor(arrs['h'], dtype=torch.float32, device=DEV)
    assert bool(((out.h - ref.h) == 0).all()) and not torch.equal(d2.h, h0)
    with pytest.raises(ValueError):
        flow(gpu_batch(arrs, dtype=torch.float32), eps=torch.zeros(27, nf))
    # inverse: floor() of the recovered features (floor.py:14), not one_hot(argmax)
    back = flow.reverse(out)
    assert torch.equal(back.h, h0)
    assert rel_err(to_np(back.pos), arrs['pos']) < 1e-5


def test_floor_flow_trains():
    from enflow_b200.data import synthetic as syn
    from enflow_b200.flow.loss import Alchemical_NLL
    nf, L = 5, 2
    arrs = syn.make_batch('c2', 3, n_atoms=9, seed=5)
    flow = _floor_flow(nf, L, syn.make_weights(nf, 128, L, seed=2, coord_gain=0.5))
    out, ldj = flow(gpu_batch(arrs, dtype=torch.float32))
    Alchemical_NLL(kBT=syn.TRAIN_KBT, softening=0.1)(out, ldj).backward()
    assert torch.isfinite(flow.flat_grads).all() and float(flow.flat_grads.abs().max()) > 0


def test_stand_alone_modules_do_not_drop_gradients_silently():
    from enflow_b200.data import synthetic as syn
    nf, L = 5, 1
    arrs = syn.make_batch('c2', 2, n_atoms=6, seed=3)
    model = build_model(syn.make_weights(nf, 128, L, seed=2, coord_gain=0.5), nf, L)
    data = gpu_batch(arrs, dtype=torch.float32)
    h = data.h.clone().requires_grad_(True)
    with pytest.raises(RuntimeError):
        model.networks[0](h, data.edges)
    with pytest.raises(RuntimeError):
        model.dequantize(h)
    with torch.no_grad():                            # inference use is fine
        Q, F, G = model.networks[0](h, data.edges)
        z, log_q = model.dequantize(h)
    assert Q.shape == (12, 1) and F.shape == (12, 3) and G.shape == (12, nf) and z.shape == (12, nf)
    Q2, _, _ = model.networks[0](h.detach(), data.edges)       # detached inputs, grad mode on: allowed, no graph recorded
    assert not Q2.requires_grad and torch.equal(Q, Q2)


def test_overflow_retry_leaves_a_pending_backward_intact():
    """Two forward passes before a backward: the second one overflows its edge capacity and retries.  The retry must
    release only its own workspace; the first pass's saved activations (in the cached workspace) stay untouched."""
    from enflow_b200.data import synthetic as syn
    from enflow_b200.flow.loss import Alchemical_NLL
    nf, L = 5, 2
    sd = syn.make_weights(nf, 128, L, seed=4, coord_gain=0.5)
    a1 = syn.make_batch('c2', 6, ragged=True, seed=11)
    a2 = syn.make_batch('c2', 3, n_atoms=12, seed=12)
    a3 = syn.make_batch('c2', 4, n_atoms=10, seed=13)
    e1 = syn.make_noise(int(a1['N'].sum()), nf, seed=1)
    e2 = syn.make_noise(int(a2['N'].sum()), nf, seed=2)
    e3 = syn.make_noise(int(a3['N'].sum()), nf, seed=3)
    nll = Alchemical_NLL(kBT=syn.TRAIN_KBT, softening=0.1)
    ref = build_model(sd, nf, L)
    o, ldj = ref(gpu_batch(a1), eps=torch.as_tensor(e1))
    nll(o, ldj).backward()
    want = ref.flat_grads.clone()

    m = build_model(sd, nf, L)
    o1, ldj1 = m(gpu_batch(a1), eps=torch.as_tensor(e1))                 # pending: owns the cached workspace
    m._edge_caps[(3, 36)] = 128                                          # 3 x 12 x 11 = 396 edges: overflows, retried
    o2, ldj2 = m(gpu_batch(a2), eps=torch.as_tensor(e2))
    assert m._edge_caps[(3, 36)] >= 396
    o3, ldj3 = m(gpu_batch(a3), eps=torch.as_tensor(e3))                 # must not be handed pass 1's workspace
    nll(o1, ldj1).backward()
    assert torch.equal(m.flat_grads, want)
