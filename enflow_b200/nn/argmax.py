"""ArgMax dequantiser with the reference's interface (`enflow/nn/argmax.py:6-29`)."""
import torch
from torch import nn

from .. import _lib
from ..utils.helpers import one_hot

_ORDER = ['network.0.weight', 'network.0.bias', 'network.2.weight', 'network.2.bias']


class ArgMax(nn.Module):
    PARAM_ORDER = _ORDER

    def __init__(self, node_nf, hidden_nf, act_fn=nn.SiLU()):
        super().__init__()
        self.node_nf, self.hidden_nf = node_nf, hidden_nf
        self.network = nn.Sequential(nn.Linear(node_nf, hidden_nf), act_fn, nn.Linear(hidden_nf, node_nf * 2))
        self._flat_view = None

    def _flat(self, device):
        if self._flat_view is not None and self._flat_view.device == device:
            return self._flat_view
        total, offs, cnts = _lib.param_layout(self.node_nf, 1)
        offs, cnts = offs[15:], cnts[15:]
        base = offs[0]
        flat = torch.zeros(total - base, dtype=torch.float32, device=device)
        sd = dict(self.named_parameters())
        for name, o, c in zip(_ORDER, offs, cnts):
            flat[o - base:o - base + c] = sd[name].detach().to(device, torch.float32).reshape(-1)
        return flat

    def forward(self, h, eps=None, mol_off=None):
        """Returns (z, log_q). ``eps`` defaults to ``torch.randn(h.size())`` like `argmax.py:17`.  Stand-alone
        INFERENCE entry point (no autograd; inside LFIntegrator the dequantiser is part of the fused, differentiable
        C call): inputs that require a gradient raise instead of silently losing it."""
        if torch.is_grad_enabled() and h.requires_grad:
            raise RuntimeError('enflow_b200.ArgMax.forward does not record autograd: differentiate through LFIntegrator '
                               '(or call it under torch.no_grad() / on detached inputs)')
        with torch.no_grad():
            return self._forward(h, eps, mol_off)

    def _forward(self, h, eps=None, mol_off=None):
        L = _lib.lib()
        _lib.require_cuda(h)
        dev = h.device
        N, nf = int(h.shape[0]), self.node_nf
        if eps is None:
            eps = torch.randn(h.size(), device=dev)
        hf, ef = _lib.f32c(h), _lib.f32c(eps.to(dev))
        if mol_off is None:
            mol_off = torch.tensor([0, N], dtype=torch.int32, device=dev)
        B = int(mol_off.numel()) - 1
        z = torch.empty(N, nf, dtype=torch.float32, device=dev)
        lq_atom = torch.empty(N, dtype=torch.float32, device=dev)
        lq_mol = torch.empty(B, dtype=torch.float64, device=dev)
        log_q = torch.empty(1, dtype=torch.float32, device=dev)
        ap = self._flat(dev)
        _lib.check(L.enflow_argmax_fwd(_lib.ptr(hf), _lib.ptr(ef), N, nf, _lib.ptr(ap), _lib.ptr(mol_off), B, _lib.ptr(z),
                                       _lib.ptr(lq_atom), _lib.ptr(lq_mol), _lib.ptr(log_q), _lib.stream()))
        return z, log_q[0]

    def reverse(self, z):
        return one_hot(torch.argmax(z, dim=-1), num_classes=z.shape[-1], dtype=z.dtype)
