"""Race hunt without compute-sanitizer: the same seeded training step many times, every result compared bit for
bit with the first (loss, latents, per-molecule log-det, all gradients).  Races in the software-pipelined kernels
(shared-memory buffer reuse, TMEM accumulator reuse) would show up as run-to-run differences.

    python tools/stress_determinism.py [iters]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from enflow_b200.data import synthetic as syn  # noqa: E402
from enflow_b200.flow.loss import Alchemical_NLL  # noqa: E402
from gpu_util import build_model, gpu_batch  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
bad = 0
for config, B, nf, kw, prec in [('c2', 1024, 5, {'ragged': True}, 'fp32_tc'), ('c2', 333, 5, {}, 'bf16'),
                                ('c1', 64, 4, {}, 'fp32_tc'), ('c5', 4, 5, {}, 'fp32_tc'), ('c3', 97, 1, {}, 'fp32_tc')]:
    arrs = syn.make_batch(config, B, **kw)
    eps = torch.as_tensor(syn.make_noise(int(arrs['N'].sum()), nf))
    model = build_model(syn.make_weights(nf, 128, 5, seed=0), nf, 5, precision=prec)
    nll = Alchemical_NLL(kBT=syn.TRAIN_KBT, softening=syn.TRAIN_SOFTENING)
    first = None
    for it in range(iters):
        model.zero_grad(set_to_none=True)
        out, ldj = model(gpu_batch(arrs), eps=eps)
        loss = nll(out, ldj)
        loss.backward()
        cur = (loss.detach().clone(), out.pos.detach().clone(), out.g.detach().clone(), out.ldj_mol.detach().clone(),
               model.flat_grads.clone())
        if first is None:
            first = cur
        elif not all(torch.equal(a, b) for a, b in zip(first, cur)):
            bad += 1
            print(f'{config} B={B} {prec}: iteration {it} differs from iteration 0')
    print(f'{config} B={B} {prec}: {iters} iterations, loss {first[0].item():.6f}, finite grads {bool(torch.isfinite(first[4]).all())}')
print('DIFFERENCES:', bad)
sys.exit(1 if bad else 0)
