"""Uniform ("floor") dequantiser with the interface of `enflow/nn/floor.py:5-14`.

Not on the hot path (Main always builds ArgMax, `enflow/main.py:153`); kept as host-side torch so that a flow
configured with it still constructs.  Integer-valued features z are spread over [z, z + scale) with uniform noise,
which leaves the log-density unchanged up to the constant log(scale) per element that the reference ignores
(it returns 0 as the log-det term).
"""
import torch


class Floor(torch.nn.Module):
    PARAM_ORDER = []          # no parameters: nothing to place in the flat buffer

    def __init__(self, dequant_scale=1):
        super().__init__()
        self.dequant_scale = float(dequant_scale)

    def forward(self, z, noise=None):
        """Returns (z + scale * u, 0) with u ~ U[0,1); ``noise`` injects u for reproducible tests."""
        u = torch.rand_like(z, dtype=z.dtype if z.is_floating_point() else torch.float32) if noise is None else noise
        return z.to(u.dtype) + self.dequant_scale * u.detach(), 0

    def reverse(self, z):
        """Quantise back: the integer part (exact inverse of forward for scale <= 1)."""
        return z.floor()
