"""Alchemical_NLL with the reference's interface (`enflow/flow/loss.py:5-25`) on the K5 kernels."""
import torch

from .. import _lib


class _NLLFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, vel, h, g, ldj, off, B, max_n, kBT, softening, z_lj):
        L = _lib.lib()
        dev = pos.device
        N, nf = int(pos.shape[0]), int(h.shape[1])
        pos, vel, h, g = (_lib.f32c(t) for t in (pos, vel, h, g))
        ldj1 = _lib.f32c(ldj).reshape(1)
        mol_term = torch.empty(B * int(L.enflow_nll_slices(max_n)), dtype=torch.float64, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        p = _lib.ptr
        _lib.check(L.enflow_nll_fwd(p(pos), p(vel), p(h), p(g), p(off), B, N, nf, max_n, kBT, softening, z_lj,
                                    p(ldj1), p(mol_term), p(loss), _lib.stream()))
        ctx.saved = (pos, vel, h, g, off, B, nf, max_n, kBT, softening)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        L = _lib.lib()
        pos, vel, h, g, off, B, nf, max_n, kBT, softening = ctx.saved
        dev = pos.device
        dl = dloss.to(torch.float32).reshape(1).contiguous()
        dpos, dvel, dh, dg = (torch.empty_like(t) for t in (pos, vel, h, g))
        dldj = torch.empty(1, dtype=torch.float32, device=dev)
        p = _lib.ptr
        _lib.check(L.enflow_nll_bwd(p(pos), p(vel), p(h), p(g), p(off), B, nf, max_n, kBT, softening, p(dl), p(dpos),
                                    p(dvel), p(dh), p(dg), p(dldj), _lib.stream()))
        return dpos, dvel, dh, dg, dldj.reshape(()), None, None, None, None, None, None


class Alchemical_NLL:
    def __init__(self, kBT, partition_func=10, softening=0):
        self.kBT = kBT
        self.z_lj = partition_func
        self.softening = softening

    def __call__(self, out, ldj):
        _lib.require_cuda(out.pos)
        B, off, max_n, _ = out.meta()
        if not torch.is_tensor(ldj):
            ldj = torch.tensor(float(ldj), device=out.pos.device)
        return _NLLFn.apply(out.pos, out.vel, out.h, out.g, ldj, off, B, max_n, float(self.kBT),
                            float(self.softening), float(self.z_lj))
