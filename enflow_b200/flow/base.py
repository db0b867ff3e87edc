"""BaseFlow with the reference's constructor and attributes (`enflow/flow/base.py:5-18`).

Differences that make the B200 path work (the public surface is unchanged):
  * parameters are fp32 and all live in ONE flat CUDA buffer (``self.flat_params``); the individual
    ``nn.Parameter`` objects keep the reference's names/shapes and are views into it, so reference
    checkpoints load with ``load_state_dict`` (fp64 values are down-cast by ``copy_``);
  * gradients are produced into a second flat buffer with the same layout (one all-reduce for DP).
"""
import torch

from .. import _lib


class BaseFlow(torch.nn.Module):
    def __init__(self, networks, dequant_network, dt):
        super().__init__()
        self.networks = torch.nn.ModuleList(networks)
        self.dequantize = dequant_network
        self.dt = dt
        self.dt_2 = 0.5 * dt
        self.flat_params = None
        self.flat_grads = None
        self._layout = None
        self._flatten()

    # ---- flat parameter buffer -------------------------------------------------------------------
    def _ordered_params(self):
        out = []
        for i, net in enumerate(self.networks):
            sd = dict(net.named_parameters())
            out += [sd[k] for k in net.PARAM_ORDER]
        if hasattr(self.dequantize, 'PARAM_ORDER'):
            sd = dict(self.dequantize.named_parameters())
            out += [sd[k] for k in self.dequantize.PARAM_ORDER]
        return out

    def _flatten(self):
        params = self._ordered_params()
        if not params:
            return
        nf, L = self.networks[0].input_nf, len(self.networks)
        total, offs, cnts = _lib.param_layout(nf, L)
        dev = params[0].device
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, o, c in zip(params, offs, cnts):
            assert p.numel() == c, (tuple(p.shape), c)
            flat[o:o + c] = p.detach().to(torch.float32).reshape(-1)
            p.data = flat[o:o + c].view(p.shape)
        self.flat_params = flat
        self.flat_grads = torch.zeros_like(flat)
        self._layout = (offs, cnts)
        per_layer = offs[15] - offs[0] if L > 1 else None
        for i, net in enumerate(self.networks):
            end = offs[15 * (i + 1)] if 15 * (i + 1) < len(offs) else total
            net._flat_view = flat[offs[15 * i]:end]
        if hasattr(self.dequantize, 'PARAM_ORDER'):        # ArgMax: its slice of the flat buffer; others keep their own parameters
            self.dequantize._flat_view = flat[offs[15 * L]:]

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        self._flatten()           # .to()/.cuda()/.float() re-allocate parameter storage: rebuild the views
        return self

    def double(self):             # the reference casts to fp64 (base.py:12); this path computes in fp32
        return self

    def grad_views(self, flat=None):
        flat = self.flat_grads if flat is None else flat
        offs, cnts = self._layout
        return [flat[o:o + c].view(p.shape) for p, o, c in zip(self._ordered_params(), offs, cnts)]

    def forward(self, data):
        pass

    def reverse(self, data):
        pass
