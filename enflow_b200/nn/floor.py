"""Floor dequantiser (`enflow/nn/floor.py:5-14`); trivial, host-side torch (unused by Main)."""
import torch


class Floor(torch.nn.Module):
    def __init__(self, dequant_scale=1):
        super().__init__()
        self.dequant_scale = dequant_scale

    def forward(self, z):
        return z + self.dequant_scale * torch.rand_like(z).detach(), 0

    def reverse(self, z):
        return torch.floor(z)
