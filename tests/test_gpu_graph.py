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
