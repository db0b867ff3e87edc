"""Development check for the tensor-core edge kernels: one training step per precision mode on a few batch
shapes, gradients and latents of the tcgen05 modes compared with the FFMA mode on the same GPU, plus run-to-run
bitwise determinism.  Meant to be run under `timeout` (a hung kernel must not hold the box)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))

from enflow_b200.data import synthetic as syn  # noqa: E402
from enflow_b200.flow.loss import Alchemical_NLL  # noqa: E402
from gpu_util import build_model, gpu_batch  # noqa: E402


def step(model, arrs, eps, nll):
    model.zero_grad(set_to_none=True)
    out, ldj = model(gpu_batch(arrs), eps=torch.as_tensor(eps))
    loss = nll(out, ldj)
    loss.backward()
    torch.cuda.synchronize()
    return loss.detach().clone(), out.pos.detach().clone(), out.g.detach().clone(), model.flat_grads.clone()


def main():
    shapes = [('c2', 2, dict(n_atoms=5)), ('c2', 4, dict(n_atoms=12)), ('c2', 64, dict(ragged=True)), ('c1', 64, {}),
              ('c2', 1024, {})]
    if len(sys.argv) > 1:
        shapes = shapes[:int(sys.argv[1])]
    nll = Alchemical_NLL(kBT=syn.TRAIN_KBT, softening=syn.TRAIN_SOFTENING)
    for cfg, B, kw in shapes:
        arrs = syn.make_batch(cfg, B, **kw)
        nf = arrs['h'].shape[1]
        L = 2 if B < 1024 else 5
        sd = syn.make_weights(nf, 128, L, seed=0, coord_gain=0.5)
        eps = syn.make_noise(int(arrs['N'].sum()), nf)
        ref = None
        for prec in ('fp32', 'fp32_tc', 'bf16'):
            model = build_model(sd, nf, L, precision=prec)
            t0 = time.time()
            r1 = step(model, arrs, eps, nll)
            r2 = step(model, arrs, eps, nll)
            dt = time.time() - t0
            det = all(torch.equal(a, b) for a, b in zip(r1, r2))
            if ref is None:
                ref = r1
            gerr = float((r1[3] - ref[3]).norm() / ref[3].norm())
            perr = float((r1[1] - ref[1]).abs().max() / ref[1].abs().max())
            gg = float((r1[2] - ref[2]).abs().max() / ref[2].abs().max())
            lerr = float(abs(r1[0] - ref[0]) / abs(ref[0]))
            print(f'{cfg} B={B} {kw} {prec:8s} loss {float(r1[0]):.6f} dloss {lerr:.2e} dpos {perr:.2e} dg {gg:.2e} '
                  f'dgrad {gerr:.2e} deterministic {det} finite {bool(torch.isfinite(r1[3]).all())} ({dt:.2f}s)', flush=True)


if __name__ == '__main__':
    main()
