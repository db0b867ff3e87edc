"""Shared loader for tests/golden/*.npz (inputs/weights are regenerated from seeds)."""
import hashlib
import importlib.util
import os

import numpy as np

from enflow_b200.data import synthetic as syn

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location('make_golden', os.path.join(HERE, 'golden', 'make_golden.py'))
_mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mg)
CASES = _mg.CASES
H = _mg.H
projection_vectors = _mg.projection_vectors


def load_case(name):
    config, B, kw, nf, L, wseed, gain, soft = CASES[name]
    batch = syn.make_batch(config, B, **kw)
    sd = syn.make_weights(nf, H, L, seed=wseed, coord_gain=gain)
    eps = syn.make_noise(int(batch['N'].sum()), nf)
    gold = dict(np.load(os.path.join(HERE, 'golden', name + '.npz')))
    assert _mg.input_digest(batch, sd, eps) == str(gold['digest']), 'synthetic generator drifted from the golden inputs'
    return {'batch': batch, 'sd': sd, 'eps': eps, 'gold': gold, 'nf': nf, 'L': L, 'softening': soft,
            'dt': syn.TRAIN_DT, 'kBT': syn.TRAIN_KBT}
