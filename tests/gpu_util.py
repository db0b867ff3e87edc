"""Helpers shared by the -m gpu parity tests (CUDA path through the C ABI vs the CPU oracle)."""
import numpy as np
import torch

from enflow_b200.data.base import batch_from_arrays
from enflow_b200.flow.dynamics import LFIntegrator
from enflow_b200.nn.argmax import ArgMax
from enflow_b200.nn.egcl import EGCL

DEV = 'cuda:0'


# Every parity test runs in all three arithmetic modes of the edge MLP; the tolerance pair is (latents / log-likelihood,
# gradients): the north star's 1e-5 for the fp32-accurate modes ('fp32_tc' = tcgen05 with the bf16x3 operand split is the
# product default, 'fp32' = the CUDA-core cross-check) and 1e-2 under the bf16 MLP.  Gradients: 1e-4 / 5e-2 (L2, per tensor).
MODES = [('fp32_tc', 1e-5, 1e-4), ('fp32', 1e-5, 1e-4), ('bf16', 1e-2, 5e-2)]


def build_model(sd, nf, L, H=128, dt=None, precision='fp32_tc'):
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


def elem_rel_err(a, b, floor=1e-2):
    """max over elements of |a-b| / max(|b|, floor * max|b|): elementwise relative error with a floor for the
    entries that are small against the tensor's scale (reported next to the max-norm figure)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if not b.size:
        return 0.0
    den = np.maximum(np.abs(b), floor * max(np.abs(b).max(), 1e-300))
    return float((np.abs(a - b) / den).max())


def check_close(a, b, tol, what):
    """Asserts the max-norm relative error (the north star's figure for a tensor of latents) and reports the elementwise
    one (1 % floor) next to it: `pytest -rP` prints the pairs.  The elementwise figure is bounded by 100x the max-norm
    one by construction; for fp32_tc it sits at ~1e-4 on the per-layer G (entries ~1 % of the tensor's scale carry
    the same absolute error as the large ones)."""
    e1, e2 = rel_err(a, b), elem_rel_err(a, b)
    print(f'{what}: max-norm rel err {e1:.3e} (tol {tol:g}), elementwise rel err (1% floor) {e2:.3e}')
    assert e1 < tol, f'{what}: max-norm rel err {e1:.3e} (tol {tol:g}), elementwise {e2:.3e}'
    return e1, e2


def to_np(t):
    return t.detach().cpu().double().numpy()
