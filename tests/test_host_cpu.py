"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the flat
parameter layout matches the reference's state_dict, and the product path refuses to run without CUDA."""
import os
import re

import numpy as np
import pytest
import torch

import __graft_entry__ as entry
from enflow_b200 import _lib
from enflow_b200.data import synthetic as syn
from enflow_b200.data.base import DataLoader, Data, batch_from_arrays
from enflow_b200.flow.dynamics import LFIntegrator
from enflow_b200.nn.argmax import ArgMax
from enflow_b200.nn.egcl import EGCL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module', autouse=True)
def built():
    entry.build()


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'enflow_b200.h')).read()
    declared = set(re.findall(r'\b(enflow_[a-z0-9_]+)\s*\(', header))
    declared.discard('enflow_dims_t')
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    handle = _lib.lib()
    for name in declared:
        assert hasattr(handle, name)
    assert handle.enflow_hidden() == 128


def test_param_layout_matches_reference_state_dict():
    for nf, L in ((5, 5), (4, 2), (1, 3)):
        total, offs, cnts = _lib.param_layout(nf, L)
        shapes = syn.flow_param_shapes(nf, 128, L)
        assert [int(np.prod(s)) for _, s in shapes] == cnts
        assert all(o % 32 == 0 for o in offs) and all(a + c <= b for a, c, b in zip(offs, cnts, offs[1:] + [total]))
    assert sum(_lib.param_layout(5, 5)[2]) == 268968          # total parameter count at nf=5, L=5


def test_flat_buffer_and_state_dict_round_trip():
    nf, L = 5, 2
    sd = syn.make_weights(nf, 128, L, seed=2)
    m = LFIntegrator([EGCL(nf, nf, 128) for _ in range(L)], ArgMax(nf, 128), dt=0.01)
    assert list(m.state_dict().keys()) == [k for k, _ in syn.flow_param_shapes(nf, 128, L)]
    m.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})      # fp64 checkpoint -> fp32 views
    total, offs, cnts = _lib.param_layout(nf, L)
    for (k, v), o, c in zip(sd.items(), offs, cnts):
        assert np.array_equal(m.flat_params[o:o + c].numpy(), v.astype(np.float32).reshape(-1)), k
    p0 = next(m.parameters())
    p0.data.add_(1.0)                                                    # optimizer-style in-place update
    assert torch.equal(m.flat_params[offs[0]:offs[0] + cnts[0]].view(p0.shape), p0.data)


def test_collater_layout_matches_reference():
    arrs = syn.make_batch('c2', 3, ragged=True)
    mols, o = [], 0
    for m, n in enumerate(arrs['N']):
        sl = slice(o, o + int(n))
        mols.append(Data(z=None, h=torch.tensor(arrs['h'][sl]), g=torch.tensor(arrs['g'][sl]),
                         pos=torch.tensor(arrs['pos'][sl]), vel=torch.tensor(arrs['vel'][sl]), N=int(n),
                         r_cut=float(arrs['r_cut'][m]), box=torch.tensor(arrs['box'][sl])))
        o += int(n)
    b = DataLoader.collater(mols)
    assert b.r_cut.dtype == torch.float32 and b.N.dtype == torch.int64          # enflow/data/base.py:170-171
    assert torch.equal(b.pos, torch.tensor(arrs['pos']))
    B, off, max_n, _ = b.meta()
    assert B == 3 and off.tolist() == [0] + list(np.cumsum(arrs['N']))
    assert [int(m.pos.shape[0]) for m in b] == list(arrs['N'])


def test_no_cpu_fallback():
    nf = 5
    m = LFIntegrator([EGCL(nf, nf, 128)], ArgMax(nf, 128), dt=0.01)
    data = batch_from_arrays(syn.make_batch('c2', 1, n_atoms=4))
    with pytest.raises(RuntimeError, match='CUDA'):
        m(data)


def test_bench_reference_arm_runs_the_vendored_reference():
    """`bench.py --impl reference` / the cpu_baseline leg: the unmodified reference from baseline/_ref when build() vendored
    it (kind 'reference'), else the port under oracle/ (kind 'port'); both give a finite molecules/s on a tiny sample."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(root, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    mols_s, cores, sec, kind = bench.cpu_reference_throughput('c2', 5, 2, 1, 0, {'n_atoms': 6})
    vendored = os.path.exists(os.path.join(root, 'baseline', '_ref', 'enflow', 'flow', 'dynamics.py'))
    assert kind == ('reference' if vendored else 'port')
    assert mols_s > 0 and cores >= 1 and sec > 0
