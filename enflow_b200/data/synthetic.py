"""Seeded synthetic conformers and weights for the BASELINE.json configs.

Everything here is numpy (legacy ``RandomState``, whose streams are frozen by
numpy's compatibility policy), so the same seed gives bit-identical inputs in
this container, on the GPU box, and inside ``tests/golden/make_golden.py``.

Every float produced is exactly representable in fp32 and returned as fp64:
the reference/oracle (fp64, `enflow/flow/base.py:12`) and the B200 path (fp32
state) therefore start from *identical* numbers, which is what makes the
bit-exact edge-construction tests meaningful.

Batch layout follows the reference collater (`enflow/data/base.py:162-174`):
molecules concatenated along axis 0, ``N`` int64 per molecule, ``r_cut``
float32 per molecule (`base.py:171`), ``box`` tiled per atom (`base.py:235`).
"""
import math
import numpy as np

from ..utils.conversion import kelvin_to_lj, time_to_lj, dist_to_lj

SIGMA_ANG = 3.4

# constants of example/train.yaml:13-16,24-26 in LJ units
TRAIN_DT = time_to_lj(1.0, 'pico')
TRAIN_KBT = kelvin_to_lj(300.0)
TRAIN_SOFTENING = 0.1
TRAIN_RCUT = dist_to_lj(3.0, 'ang')


def _f32(x):
    return np.asarray(x, dtype=np.float32).astype(np.float64)


def _self_avoiding(rs, n, extent, min_dist):
    pts = np.zeros((n, 3))
    k = 0
    while k < n:
        p = rs.uniform(0.0, extent, size=3)
        if k == 0 or np.min(np.sum((pts[:k] - p) ** 2, axis=1)) >= min_dist ** 2:
            pts[k] = p
            k += 1
    return pts


def _one_hot(types, nf):
    h = np.zeros((len(types), nf))
    h[np.arange(len(types)), types] = 1.0
    return h


def _cluster55(rs, n, spacing, jitter):
    g = np.arange(4) - 1.5
    grid = np.stack(np.meshgrid(g, g, g, indexing='ij'), -1).reshape(-1, 3)
    order = np.argsort(np.sum(grid ** 2, axis=1), kind='stable')
    side = 4
    while len(grid) < n:          # larger clusters than 64: grow the cube
        side += 1
        g = np.arange(side) - (side - 1) / 2
        grid = np.stack(np.meshgrid(g, g, g, indexing='ij'), -1).reshape(-1, 3)
        order = np.argsort(np.sum(grid ** 2, axis=1), kind='stable')
    pts = grid[order[:n]] * spacing
    return pts + rs.uniform(-jitter, jitter, size=pts.shape)


def make_batch(config, num_mols, seed=None, n_atoms=None, ragged=False):
    """Return one collated batch (dict of numpy arrays) for a named config.

    config: 'c1' train.yaml shape (n=22, nf=4, radius graph in the PBC-quirk regime)
            'c2' QM9-sized (n<=29, nf=5, fully connected: box=1000, r_cut=100)
            'c3' LJ-55 (n=55, nf=1, fully connected)
            'c4' generate latents (C1 shape, h,g ~ N(0, 1/sqrt(kBT)))
            'c5' protein fragment (n=500, nf=5, r_cut=5 A, box=100 A)
            'c5fc' as c5 but r_cut/box large: equivariance-testable regime
    """
    cid = {'c1': 1, 'c2': 2, 'c3': 3, 'c4': 4, 'c5': 5, 'c5fc': 5}[config]
    rs = np.random.RandomState(1234 + cid if seed is None else seed)
    kbt = TRAIN_KBT
    mols = []
    box0 = None
    for m in range(num_mols):
        if config in ('c1', 'c4'):
            n, nf = n_atoms or 22, 4
            types = np.array(([0] * 12 + [1] * 6 + [2] * 2 + [3] * 2) * ((n + 21) // 22))[:n]
            rs.shuffle(types)
            pos_ang = _self_avoiding(rs, n, 8.0, 1.0)
            if box0 is None:   # base.py:212-213: box of the FIRST molecule, reused
                box0 = np.round(pos_ang.max(0) - pos_ang.min(0)) / SIGMA_ANG
            pos = (pos_ang - pos_ang.mean(0, keepdims=True)) / SIGMA_ANG
            box, r_cut = box0, TRAIN_RCUT
        elif config == 'c2':
            nf = 5
            n = n_atoms or 29
            if ragged:
                n = int(np.clip(np.round(rs.normal(18.0, 5.0)), 3, 29))
            types = rs.randint(0, nf, size=n)
            pos_ang = _self_avoiding(rs, n, 9.0, 1.0)
            pos = (pos_ang - pos_ang.mean(0, keepdims=True)) / SIGMA_ANG
            box, r_cut = np.full(3, 1000.0), 100.0
        elif config == 'c3':
            n, nf = n_atoms or 55, 1
            types = np.zeros(n, dtype=np.int64)
            pos = _cluster55(rs, n, 1.0, 0.05)
            pos = pos - pos.mean(0, keepdims=True)
            box, r_cut = np.full(3, 1000.0), 100.0
        else:
            n, nf = n_atoms or 500, 5
            types = rs.randint(0, nf, size=n)
            side = (n / 0.1) ** (1.0 / 3.0)
            pos_ang = rs.uniform(0.0, side, size=(n, 3))
            pos = (pos_ang - pos_ang.mean(0, keepdims=True)) / SIGMA_ANG
            if config == 'c5':
                box, r_cut = np.full(3, 100.0 / SIGMA_ANG), 5.0 / SIGMA_ANG
            else:
                box, r_cut = np.full(3, 1000.0), 100.0
        vel = rs.normal(0.0, math.sqrt(kbt), size=(n, 3))
        if config == 'c4':
            h = rs.normal(0.0, 1.0 / math.sqrt(kbt), size=(n, nf))
            g = rs.normal(0.0, 1.0 / math.sqrt(kbt), size=(n, nf))
        else:
            h = _one_hot(types, nf)
            g = rs.normal(0.0, 1.0, size=(n, nf))
        mols.append((h, g, pos, vel, np.tile(box, (n, 1)), n, r_cut))
    return {
        'h': _f32(np.concatenate([m[0] for m in mols])),
        'g': _f32(np.concatenate([m[1] for m in mols])),
        'pos': _f32(np.concatenate([m[2] for m in mols])),
        'vel': _f32(np.concatenate([m[3] for m in mols])),
        'box': _f32(np.concatenate([m[4] for m in mols])),
        'N': np.array([m[5] for m in mols], dtype=np.int64),
        'r_cut': np.array([m[6] for m in mols], dtype=np.float32),
    }


def make_noise(num_atoms, nf, seed=99):
    """ArgMax noise eps. The reference draws it in float32 (`enflow/nn/argmax.py:17`)."""
    rs = np.random.RandomState(seed)
    return rs.normal(0.0, 1.0, size=(num_atoms, nf)).astype(np.float32)


def egcl_param_shapes(nf, H):
    """state_dict names/shapes of one EGCL (`enflow/nn/egcl.py:21-55`), in flat-buffer order."""
    return [
        ('edge_nn.0.weight', (H, 2 * nf + 1)), ('edge_nn.0.bias', (H,)),
        ('edge_nn.2.weight', (H, H)), ('edge_nn.2.bias', (H,)),
        ('node_nn.0.weight', (H, H + nf)), ('node_nn.0.bias', (H,)),
        ('node_nn.2.weight', (nf, H)), ('node_nn.2.bias', (nf,)),
        ('coord_nn.0.weight', (H, H)), ('coord_nn.0.bias', (H,)),
        ('coord_nn.2.weight', (1, H)),
        ('vel_scaling_nn.0.weight', (H, nf)), ('vel_scaling_nn.0.bias', (H,)),
        ('vel_scaling_nn.2.weight', (1, H)), ('vel_scaling_nn.2.bias', (1,)),
    ]


def argmax_param_shapes(nf, H):
    """state_dict names/shapes of ArgMax (`enflow/nn/argmax.py:9-12`)."""
    return [
        ('network.0.weight', (H, nf)), ('network.0.bias', (H,)),
        ('network.2.weight', (2 * nf, H)), ('network.2.bias', (2 * nf,)),
    ]


def flow_param_shapes(nf, H, L):
    """Full LFIntegrator state_dict order (`enflow/flow/base.py:8-9`): networks.{i}.*, dequantize.*"""
    out = []
    for i in range(L):
        out += [(f'networks.{i}.{k}', s) for k, s in egcl_param_shapes(nf, H)]
    out += [(f'dequantize.{k}', s) for k, s in argmax_param_shapes(nf, H)]
    return out


def make_weights(nf, H, L, seed=0, coord_gain=0.5):
    """Random weights with the distributions of torch's default ``nn.Linear`` init.

    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weights and biases; ``coord_nn.2.weight``
    is xavier-uniform with ``coord_gain`` (reference default 1e-3, `egcl.py:32-33`;
    SURVEY 8d asks for a second 'trained-like' set at 0.5 so the force branch is visible).
    Values are fp32-representable, returned as fp64 arrays keyed by state_dict name.
    """
    rs = np.random.RandomState(seed)
    sd = {}
    for name, shape in flow_param_shapes(nf, H, L):
        if name.endswith('coord_nn.2.weight'):
            bound = coord_gain * math.sqrt(6.0 / (shape[0] + shape[1]))
        else:
            fan_in = shape[1] if len(shape) == 2 else None
            if fan_in is None:   # bias: fan_in of the matching weight
                fan_in = sd[name[:-4] + 'weight'].shape[1]
            bound = 1.0 / math.sqrt(fan_in)
        sd[name] = _f32(rs.uniform(-bound, bound, size=shape))
    return sd


class SYNTHETICDataset:
    """Dataset plugin following the reference's convention (`enflow/main.py:67-68`: module ``enflow.data.<type>``
    exporting ``<TYPE>Dataset``): seeded synthetic conformers of one of the BASELINE.json shapes, held in memory as
    per-molecule ``Data`` objects like ``InMemoryBaseDataset`` (`enflow/data/base.py:250-283`)."""

    def __init__(self, config='c2', num_mols=1024, seed=None, n_atoms=None, ragged=False, **_ignored):
        import torch
        from .base import Data
        kw = {'seed': seed, 'ragged': ragged}
        if n_atoms:
            kw['n_atoms'] = int(n_atoms)
        arrs = make_batch(config, int(num_mols), **kw)
        self.data_list, o = [], 0
        for m, n in enumerate(arrs['N']):
            sl = slice(o, o + int(n))
            self.data_list.append(Data(z=None, h=torch.tensor(arrs['h'][sl], dtype=torch.float32),
                                       g=torch.tensor(arrs['g'][sl], dtype=torch.float32),
                                       pos=torch.tensor(arrs['pos'][sl], dtype=torch.float32),
                                       vel=torch.tensor(arrs['vel'][sl], dtype=torch.float32), N=int(n),
                                       r_cut=float(arrs['r_cut'][m]),
                                       box=torch.tensor(arrs['box'][sl], dtype=torch.float32), label=None))
            o += int(n)

    def __len__(self):
        return len(self.data_list)

    def __getitem__(self, idx):
        return self.data_list[idx]

    @property
    def node_nf(self):
        return self.data_list[0].h.shape[1]

    @property
    def num_atoms_per_mol(self):
        return self.data_list[0].N
