"""HBM roofline of the aggregation (K2) and coupling (K3) kernels in isolation, through the C ABI.

    python tools/kernel_roofline.py > profiles/r1_kernel_roofline.json

Each kernel is timed with CUDA events on its launch stream, 20 iterations after 3 warm-ups, with the L2 flushed
between iterations (a 512 MB buffer is overwritten).  achieved = algorithmic bytes / mean duration; peak = the
measured copy bandwidth in MEASURED_PEAKS.json (burst figure: the kernel is timed alone).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from enflow_b200 import _lib  # noqa: E402

dev = torch.device('cuda', 0)
L = _lib.lib()
p = _lib.ptr
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(np.min(ts))


def seg_sum(B, n):
    N, deg = B * n, n - 1
    E = N * deg
    x = torch.randn(E, 128, device=dev)
    ptr = (torch.arange(N + 1, device=dev, dtype=torch.int64) * deg).to(torch.int32)
    out = torch.empty(N, 128, device=dev)
    fn = lambda: _lib.check(L.enflow_segment_sum128(p(x), p(ptr), None, N, E, 0, p(out), _lib.stream()))
    mean, best = timeit(fn)
    nbytes = E * 512 + (N + 1) * 4 + N * 512
    return {'kernel': 'k_segment_sum128', 'shape': f'B={B} n={n} E={E}', 'bytes': nbytes, 'ms_mean': mean, 'ms_min': best,
            'achieved_gbs': nbytes / mean / 1e6, 'peak_gbs': peak, 'frac': nbytes / mean / 1e6 / peak}


def coupling(B, n, nf):
    N = B * n
    mk = lambda *s: torch.randn(*s, device=dev)
    Q, F, G, h, g, pos, vel = mk(N), mk(N, 3), mk(N, nf), mk(N, nf), mk(N, nf), mk(N, 3), mk(N, 3)
    box = torch.full((N, 3), 5.0, device=dev)
    off = (torch.arange(B + 1, device=dev, dtype=torch.int64) * n).to(torch.int32)
    ho, go, po, vo = (torch.empty_like(t) for t in (h, g, pos, vel))
    ldj = torch.zeros(B, device=dev)
    fn = lambda: _lib.check(L.enflow_coupling_fwd(p(Q), p(F), p(G), p(h), p(g), p(pos), p(vel), p(box), p(off), B, nf, 0.01,
                                                   p(ho), p(go), p(po), p(vo), p(ldj), _lib.stream()))
    mean, best = timeit(fn)
    nbytes = N * (19 + 5 * nf) * 4 + B * 4
    return {'kernel': 'k_coupling_fwd', 'shape': f'B={B} n={n} nf={nf} N={N}', 'bytes': nbytes, 'ms_mean': mean, 'ms_min': best,
            'achieved_gbs': nbytes / mean / 1e6, 'peak_gbs': peak, 'frac': nbytes / mean / 1e6 / peak}


res = [seg_sum(1024, 29), seg_sum(1024, 55), seg_sum(4096, 29),
       coupling(1024, 29, 5), coupling(16384, 22, 4), coupling(125000, 22, 4), coupling(1 << 20, 22, 4)]
print(json.dumps({'peak_source': 'MEASURED_PEAKS.json hbm_gbs (copy, burst)', 'l2': 'flushed between iterations', 'results': res}, indent=1))
