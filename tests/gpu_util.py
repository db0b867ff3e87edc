"""Helpers shared by the -m gpu parity tests (CUDA path through the C ABI vs the CPU oracle)."""
import numpy as np
import torch

from enflow_b200.data.base import batch_from_arrays
from enflow_b200.flow.dynamics import LFIntegrator
from enflow_b200.nn.argmax import ArgMax
from enflow_b200.nn.egcl import EGCL

DEV = 'cuda:0'


def build_model(sd, nf, L, H=128, dt=None, precision='fp32'):   # the FFMA path is the parity baseline; tc modes are tested explicitly
    from enflow_b200.data import synthetic as syn
    m = LFIntegrator([EGCL(nf, nf, H) for _ in range(L)], ArgMax(nf, H), dt=syn.TRAIN_DT if dt is None else dt)
    m.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    m = m.to(DEV)
    m.precision = precision
    for net in m.networks:
        net.precision = precision
    return m


def gpu_batch(arrs, dtype=torch.float64):
    return batch_from_arrays(arrs, device=DEV, dtype=dtype)


def rel_err(a, b):
    """max |a-b| / max |b| (the north star's 'relative' for a tensor of latents)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if b.size else 0.0


def to_np(t):
    return t.detach().cpu().double().numpy()
