"""CUDA-graph replay of the whole training step reproduces the eagerly launched step bit for bit."""
import numpy as np
import pytest
import torch

from golden_util import load_case
from gpu_util import DEV, build_model, gpu_batch

pytestmark = pytest.mark.gpu


def test_graphed_step_matches_eager_bitwise():
    from enflow_b200.flow.loss import Alchemical_NLL
    from enflow_b200.graph import GraphedTrainStep
    c = load_case('c2_ragged')
    eps = torch.as_tensor(c['eps'])
    nll = Alchemical_NLL(kBT=c['kBT'], softening=c['softening'])

    def fresh():
        m = build_model(c['sd'], c['nf'], c['L'], precision='fp32_tc')
        return m, torch.optim.Adam(m.parameters(), lr=1e-3, capturable=True)

    # eager: two optimizer steps on the same batch and noise
    m0, o0 = fresh()
    eager_losses = []
    for _ in range(5):
        o0.zero_grad(set_to_none=True)
        out, ldj = m0(gpu_batch(c['batch'], dtype=torch.float32), eps=eps)
        loss = nll(out, ldj)
        loss.backward()
        o0.step()
        eager_losses.append(loss.item())
    # graphed: the constructor runs 3 eager warm-up steps, then every call replays one captured step
    m1, o1 = fresh()
    step = GraphedTrainStep(m1, nll, o1, gpu_batch(c['batch'], dtype=torch.float32), warmup=3, eps=eps)
    l4 = step(gpu_batch(c['batch'], dtype=torch.float32)).item()      # capture itself does not execute: this is step 4
    l5 = step().item()
    assert not step.overflowed()
    assert l4 == eager_losses[3] and l5 == eager_losses[4], (eager_losses, l4, l5)
    assert torch.equal(m0.flat_params, m1.flat_params), 'parameters after 5 steps must be bit-identical'


def test_flat_adam_matches_torch_adam():
    """The fused flat-buffer Adam kernel follows torch.optim.Adam step for step and its state loads into it."""
    from enflow_b200.flow.loss import Alchemical_NLL
    from enflow_b200.optim import FlatAdam
    c = load_case('c2_ragged')
    eps = torch.as_tensor(c['eps'])
    nll = Alchemical_NLL(kBT=c['kBT'], softening=c['softening'])
    m0 = build_model(c['sd'], c['nf'], c['L'], precision='fp32_tc')
    m1 = build_model(c['sd'], c['nf'], c['L'], precision='fp32_tc')
    o0 = torch.optim.Adam(m0.parameters(), lr=1e-3)
    o1 = FlatAdam(m1, lr=1e-3)
    for _ in range(4):
        for m, o in ((m0, o0), (m1, o1)):
            o.zero_grad(set_to_none=True)
            out, ldj = m(gpu_batch(c['batch'], dtype=torch.float32), eps=eps)
            nll(out, ldj).backward()
            o.step()
    err = (m0.flat_params - m1.flat_params).abs().max().item()
    assert err < 2e-6, err              # same update up to fp32 rounding of the bias-correction arithmetic
    sd = o1.state_dict()
    assert set(sd['state'][0]) == {'step', 'exp_avg', 'exp_avg_sq'} and float(sd['state'][0]['step']) == 4.0
    o2 = torch.optim.Adam(m1.parameters(), lr=1e-3)
    o2.load_state_dict(sd)              # torch's Adam accepts the flat optimizer's state
    ref = o0.state_dict()['state'][3]['exp_avg']
    assert torch.allclose(o2.state_dict()['state'][3]['exp_avg'], ref, rtol=1e-4, atol=1e-9)
    # ... and torch's Adam can STEP with it (it reads weight_decay, amsgrad, maximize, ... from the loaded group)
    for m, o in ((m0, o0), (m1, o2)):
        o.zero_grad(set_to_none=True)
        out, ldj = m(gpu_batch(c['batch'], dtype=torch.float32), eps=eps)
        nll(out, ldj).backward()
        o.step()
    assert (m0.flat_params - m1.flat_params).abs().max().item() < 4e-6


def test_flat_adam_applies_gradients_that_do_not_alias_the_flat_buffer():
    """A .grad replaced by user code (clipping into a fresh tensor, manual accumulation) must still be applied."""
    from enflow_b200.flow.loss import Alchemical_NLL
    from enflow_b200.optim import FlatAdam
    c = load_case('c2_ragged')
    eps = torch.as_tensor(c['eps'])
    nll = Alchemical_NLL(kBT=c['kBT'], softening=c['softening'])
    ms = [build_model(c['sd'], c['nf'], c['L'], precision='fp32_tc') for _ in range(2)]
    os_ = [FlatAdam(m, lr=1e-3) for m in ms]
    for i, (m, o) in enumerate(zip(ms, os_)):
        o.zero_grad(set_to_none=True)
        out, ldj = m(gpu_batch(c['batch'], dtype=torch.float32), eps=eps)
        nll(out, ldj).backward()
        if i == 1:
            for q in m.parameters():
                q.grad = q.grad.clone() * 0.5          # fresh tensors, half the gradient
        else:
            m.flat_grads.mul_(0.5)
        o.step()
    assert torch.equal(ms[0].flat_params, ms[1].flat_params)


def test_graphed_step_follows_a_per_batch_scheduler():
    """StepLR stepped per batch (main.py:188,223, Q15) with a captured graph: the learning rate lives on the device."""
    from enflow_b200.flow.loss import Alchemical_NLL
    from enflow_b200.graph import GraphedTrainStep
    from enflow_b200.optim import FlatAdam
    c = load_case('c2_ragged')
    eps = torch.as_tensor(c['eps'])
    nll = Alchemical_NLL(kBT=c['kBT'], softening=c['softening'])
    batch = lambda: gpu_batch(c['batch'], dtype=torch.float32)
    m0 = build_model(c['sd'], c['nf'], c['L'], precision='fp32_tc')
    o0 = FlatAdam(m0, lr=1e-2)
    s0 = torch.optim.lr_scheduler.StepLR(o0, step_size=2, gamma=0.5)
    for _ in range(6):
        o0.zero_grad(set_to_none=True)
        out, ldj = m0(batch(), eps=eps)
        nll(out, ldj).backward()
        o0.step()
        s0.step()
    m1 = build_model(c['sd'], c['nf'], c['L'], precision='fp32_tc')
    o1 = FlatAdam(m1, lr=1e-2)
    s1 = torch.optim.lr_scheduler.StepLR(o1, step_size=2, gamma=0.5)
    step = GraphedTrainStep(m1, nll, o1, batch(), warmup=1, eps=eps, scheduler=s1)      # step 1 (eager warm-up)
    s1.step()
    for _ in range(5):
        step(batch())
        s1.step()
    assert o1.param_groups[0]['lr'] == o0.param_groups[0]['lr'] == 1e-2 * 0.125
    assert torch.equal(m0.flat_params, m1.flat_params)


def test_graphed_step_reports_edge_capacity_overflow_before_the_optimizer_runs():
    """A batch with more edges than the captured capacity must not be trained on a truncated neighbour list."""
    from enflow_b200.flow.loss import Alchemical_NLL
    from enflow_b200.graph import EdgeCapacityOverflow, GraphedTrainStep
    from enflow_b200.optim import FlatAdam
    c = load_case('c1_pbc')
    nll = Alchemical_NLL(kBT=c['kBT'], softening=c['softening'])
    m = build_model(c['sd'], c['nf'], c['L'], precision='fp32_tc')
    o = FlatAdam(m, lr=1e-3)
    eps = torch.as_tensor(c['eps'])
    b = gpu_batch(c['batch'], dtype=torch.float32)
    n_edges = int(b.edges.row.numel())
    m._edge_caps[(int(c['batch']['N'].shape[0]), int(c['batch']['N'].sum()))] = n_edges + 64     # just enough for this batch
    step = GraphedTrainStep(m, nll, o, b, warmup=1, eps=eps, check_overflow=True)
    step(b)                                             # fits
    before = m.flat_params.clone()
    dense = gpu_batch(c['batch'], dtype=torch.float32)
    dense.pos = dense.pos * 0.3                         # same layout, atoms pulled together: many more edges
    with pytest.raises(EdgeCapacityOverflow):
        step(dense)
    assert torch.equal(before, m.flat_params), 'the optimizer must not have run on the truncated step'
