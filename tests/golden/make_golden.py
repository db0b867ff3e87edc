"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
Writes tests/golden/<case>.npz.  Inputs and weights are NOT stored: they are
regenerated from seeds by ``enflow_b200.data.synthetic`` (numpy RandomState);
a sha256 of them is stored so generator drift is detected.

How the reference is driven (nothing in it is edited):
  * ``rdkit`` is absent here and is needed only for one constant
    (`enflow/utils/constants.py:2`), so a two-file stub is written to a temp dir.
  * batches are built from per-molecule ``Data`` objects through the reference's own
    ``DataLoader.collater`` (`enflow/data/base.py:162-174`).
  * the model is ``LFIntegrator([EGCL]*L distinct, ArgMax, dt)`` exactly as
    `enflow/main.py:150-153` builds it, weights loaded with ``load_state_dict``.
  * the ArgMax noise (`enflow/nn/argmax.py:17`, ``torch.randn`` float32) is injected by
    temporarily replacing ``torch.randn`` so that oracle/GPU can share the same eps.
"""
import hashlib
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from enflow_b200.data import synthetic as syn  # noqa: E402

CASES = {
    # name: (config, num_mols, kwargs for make_batch, nf, L, weight seed, coord gain, softening)
    'c1_pbc':    ('c1', 4, {}, 4, 5, 0, 0.5, syn.TRAIN_SOFTENING),
    'c2_ragged': ('c2', 6, {'ragged': True}, 5, 5, 0, 0.5, syn.TRAIN_SOFTENING),
    'c2_default_init': ('c2', 3, {'n_atoms': 29}, 5, 2, 1, 1e-3, 0.0),
    'c3_lj55':   ('c3', 2, {}, 1, 5, 0, 0.5, syn.TRAIN_SOFTENING),
    'c5_radius': ('c5', 1, {'n_atoms': 500, 'seed': 1239}, 5, 2, 0, 0.5, syn.TRAIN_SOFTENING),
    'c5_small':  ('c5', 2, {'n_atoms': 96}, 5, 3, 0, 0.5, syn.TRAIN_SOFTENING),
}
H = 128


def input_digest(batch, sd, eps):
    m = hashlib.sha256()
    for k in sorted(batch):
        m.update(np.ascontiguousarray(batch[k]).tobytes())
    for k in sorted(sd):
        m.update(np.ascontiguousarray(sd[k]).tobytes())
    m.update(np.ascontiguousarray(eps).tobytes())
    return m.hexdigest()


def projection_vectors(sd, seed=7):
    rs = np.random.RandomState(seed)
    return {k: rs.normal(size=v.shape) for k, v in sd.items()}


def import_reference():
    stub = tempfile.mkdtemp(prefix='rdkit_stub_')
    os.makedirs(os.path.join(stub, 'rdkit'))
    with open(os.path.join(stub, 'rdkit', '__init__.py'), 'w') as f:
        f.write('from . import Chem\n')
    with open(os.path.join(stub, 'rdkit', 'Chem.py'), 'w') as f:
        f.write('class _PT:\n    def GetAtomicWeight(self, s):\n        return 39.948\n'
                'def GetPeriodicTable():\n    return _PT()\n')
    sys.path.insert(0, stub)
    sys.path.insert(0, '/root/reference')
    import enflow.data.base as rbase
    import enflow.nn.egcl as regcl
    import enflow.nn.argmax as rargmax
    import enflow.flow.dynamics as rdyn
    import enflow.flow.loss as rloss
    return rbase, regcl, rargmax, rdyn, rloss


def ref_batch(rbase, batch):
    mols, o = [], 0
    for m, n in enumerate(batch['N']):
        n = int(n)
        sl = slice(o, o + n)
        mols.append(rbase.Data(z=['X'] * n, h=torch.tensor(batch['h'][sl]), g=torch.tensor(batch['g'][sl]),
                               pos=torch.tensor(batch['pos'][sl]), vel=torch.tensor(batch['vel'][sl]),
                               N=n, r_cut=float(batch['r_cut'][m]), box=torch.tensor(batch['box'][sl]),
                               label=[0] * n))
        o += n
    # collater is an ordinary method that does not touch self (base.py:162-174)
    return rbase.DataLoader.collater(None, mols)


def run_case(name, mods):
    rbase, regcl, rargmax, rdyn, rloss = mods
    config, B, kw, nf, L, wseed, gain, soft = CASES[name]
    batch = syn.make_batch(config, B, **kw)
    sd = syn.make_weights(nf, H, L, seed=wseed, coord_gain=gain)
    n_atoms = int(batch['N'].sum())
    eps = syn.make_noise(n_atoms, nf)
    dt, kBT = syn.TRAIN_DT, syn.TRAIN_KBT

    model = rdyn.LFIntegrator([regcl.EGCL(nf, nf, H) for _ in range(L)], rargmax.ArgMax(nf, H), dt=dt)
    model.load_state_dict({k: torch.tensor(v, dtype=torch.float64) for k, v in sd.items()})

    out = {'digest': np.array(input_digest(batch, sd, eps)), 'n_atoms': np.array(n_atoms)}
    layer_io = []

    def hook(mod, inp, outp):
        h_in, edges = inp
        layer_io.append((edges.row.clone(), edges.col.clone(), [o.detach().clone() for o in outp]))
    handles = [net.register_forward_hook(hook) for net in model.networks]

    data = ref_batch(rbase, batch)
    assert data.r_cut.dtype == torch.float32
    real_randn = torch.randn
    torch.randn = lambda *a, **k: torch.tensor(eps)
    try:
        res, ldj = model(data)
    finally:
        torch.randn = real_randn
    loss = rloss.Alchemical_NLL(kBT=kBT, softening=soft)(res, ldj)
    loss.backward()
    for hd in handles:
        hd.remove()

    for i, (row, col, (Q, Fo, G)) in enumerate(layer_io):
        out[f'row{i}'] = row.numpy().astype(np.int32)
        out[f'col{i}'] = col.numpy().astype(np.int32)
        out[f'Q{i}'] = Q.numpy()
        out[f'F{i}'] = Fo.numpy()
        out[f'G{i}'] = G.numpy()
    for k in ('h', 'g', 'pos', 'vel'):
        out[f'out_{k}'] = getattr(res, k).detach().numpy()
    out['ldj'] = ldj.detach().numpy()
    out['loss'] = loss.detach().numpy()

    proj = projection_vectors(sd)
    names = list(sd.keys())
    grads = dict(model.named_parameters())
    out['grad_sum'] = np.array([grads[k].grad.sum().item() for k in names])
    out['grad_norm'] = np.array([grads[k].grad.norm().item() for k in names])
    out['grad_proj'] = np.array([(grads[k].grad.numpy() * proj[k]).sum() for k in names])
    if L <= 2 and n_atoms < 200:
        for k in names:
            out['grad/' + k] = grads[k].grad.numpy().astype(np.float32)

    # inverse pass from the flow output (dynamics.py:25-37); ArgMax.reverse quantises h
    with torch.no_grad():
        lat = rbase.Data(z=res.z, h=res.h.detach().clone(), g=res.g.detach().clone(),
                         pos=res.pos.detach().clone(), vel=res.vel.detach().clone(),
                         N=res.N, r_cut=res.r_cut, box=res.box, label=res.label)
        back = model.reverse(lat)
    for k in ('h', 'g', 'pos', 'vel'):
        out[f'rev_{k}'] = getattr(back, k).numpy()

    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print(f'{name}: atoms={n_atoms} E0={len(layer_io[0][0])} ldj={float(ldj):.6f} loss={float(loss):.6f} '
          f'-> {os.path.getsize(path) / 1024:.0f} KiB')


if __name__ == '__main__':
    torch.set_num_threads(os.cpu_count())
    mods = import_reference()
    for name in (sys.argv[1:] or CASES):
        run_case(name, mods)
