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
