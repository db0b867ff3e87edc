"""Prior sampler for generate through the C ABI against the oracle, and the `lj` dataset plugin."""
import math

import numpy as np
import pytest
import torch

from oracle import enflow_oracle as orc

pytestmark = pytest.mark.gpu


def _setup(n=300, box=8.0, seed=5):
    from enflow_b200.data.lj import arrange_points_on_grid
    rs = np.random.RandomState(seed)
    pos = arrange_points_on_grid(n, np.full(3, box), 0.5) + rs.uniform(-0.2, 0.2, size=(n, 3))
    return pos, np.full(3, box), rs


def _call(n):
    from enflow_b200 import _lib
    L = _lib.lib()
    dev = torch.device('cuda', 0)
    ws = torch.empty(int(L.enflow_lj_prior_workspace_doubles(n)), dtype=torch.float64, device=dev)
    return _lib, L, dev, ws


def test_forces_and_energy_match_oracle():
    pos, box, _ = _setup()
    n = len(pos)
    _lib, L, dev, ws = _call(n)
    p = torch.tensor(pos, device=dev)
    f = torch.empty_like(p)
    e = torch.zeros(2, dtype=torch.float64, device=dev)
    cbox = (_lib.C.c_double * 3)(*box)
    _lib.check(L.enflow_lj_prior_forces(_lib.ptr(p), n, cbox, 0.1, 3.0, _lib.ptr(ws), _lib.ptr(f), _lib.ptr(e), _lib.stream()))
    u_ref, f_ref = orc.lj_prior_energy_forces(pos, box, 0.1, 3.0)
    assert abs(e[0].item() - u_ref) <= 1e-11 * abs(u_ref)
    assert np.abs(f.cpu().numpy() - f_ref).max() <= 1e-10 * np.abs(f_ref).max()


def test_frictionless_trajectory_matches_oracle():
    pos, box, rs = _setup(n=150, box=6.5)
    n = len(pos)
    vel = rs.normal(0, 1.0, size=pos.shape)
    _lib, L, dev, ws = _call(n)
    p, v = torch.tensor(pos, device=dev), torch.tensor(vel, device=dev)
    cbox = (_lib.C.c_double * 3)(*box)
    _lib.check(L.enflow_lj_prior_run(_lib.ptr(p), _lib.ptr(v), n, cbox, 0.1, 3.0, 0.002, 1.0, 1.0, 25, 7, 0, _lib.ptr(ws), None,
                                     _lib.stream()))
    x_ref, v_ref = pos, vel
    for _ in range(25):
        x_ref, v_ref = orc.langevin_middle_step(x_ref, v_ref, box, 0.1, 3.0, 0.002, 1.0, 1.0, np.zeros_like(pos))
    assert np.abs(p.cpu().numpy() - x_ref).max() < 1e-10
    assert np.abs(v.cpu().numpy() - v_ref).max() < 1e-9


def test_thermostat_noise_statistics():
    """a = 0: every step replaces the velocities by sqrt(kBT) N(0,1): mean, variance and independence of the stream."""
    pos, box, _ = _setup(n=2000, box=16.0)
    n = len(pos)
    _lib, L, dev, ws = _call(n)
    p, v = torch.tensor(pos, device=dev), torch.zeros(n, 3, dtype=torch.float64, device=dev)
    cbox = (_lib.C.c_double * 3)(*box)
    run = lambda seed, step0: _lib.check(L.enflow_lj_prior_run(_lib.ptr(p), _lib.ptr(v), n, cbox, 0.1, 3.0, 1e-9, 0.0, 0.7, 1, seed,
                                                               step0, _lib.ptr(ws), None, _lib.stream()))
    run(11, 0)
    v0 = v.clone()
    assert abs(v0.mean().item()) < 0.04 and abs(v0.var().item() / 0.7 - 1.0) < 0.06
    kurt = ((v0 / math.sqrt(0.7)) ** 4).mean().item()
    assert abs(kurt - 3.0) < 0.35
    run(11, 1)
    c = torch.corrcoef(torch.stack([v0.flatten(), v.flatten()]))[0, 1].item()
    assert abs(c) < 0.05                                          # consecutive steps are independent
    v1 = v.clone()
    v.zero_(); run(11, 1)
    assert torch.equal(v, v1)                                     # same (seed, step) reproduces the noise bit for bit
    v.zero_(); run(12, 1)
    assert not torch.equal(v, v1)


def test_dataset_samples_the_prior_and_feeds_reverse():
    from enflow_b200.data.lj import LJDataset
    from enflow_b200.utils.conversion import lj_to_kelvin, lj_to_dist
    kBT = 1.2
    box_ang = [lj_to_dist(7.0)] * 3
    kw = dict(n_atoms=256, box=box_ang, temp=lj_to_kelvin(kBT), friction=5.0, dt=0.004, n_iter=3000, interval=100,
              discard=1000, softening=0.1, node_nf=4, seed=3)
    ds = LJDataset(**kw)
    assert len(ds) == 21 and ds.node_nf == 4 and ds.num_atoms_per_mol == 256
    temps = np.array([t for _, _, t in ds.log])
    assert abs(temps.mean() / kBT - 1.0) < 0.05, temps.mean()
    # equilibrated fluid, not a lattice that has barely moved: the potential energy has left the minimised value and
    # sits on a plateau (second half of the frames: no drift beyond the fluctuations), and the atoms have diffused
    pot = np.array([u for _, u, _ in ds.log]) / 256
    half = len(pot) // 2
    a, b = pot[half:half + half // 2], pot[half + half // 2:]
    assert abs(a.mean() - b.mean()) < 4 * max(pot[half:].std(), 1e-3), (a.mean(), b.mean(), pot[half:].std())
    msd = float(((ds[20].pos - ds[0].pos).double() ** 2).sum(-1).mean())
    assert msd > 0.05, msd
    d = ds[5]
    assert d.pos.shape == (256, 3) and d.vel.shape == (256, 3) and d.h.shape == (256, 4)
    assert float(d.pos.abs().max()) <= 7.0 / 2 + 7.0 / 2        # centred, inside one cell width of the origin
    assert abs(float(d.h.double().var()) * kBT - 1.0) < 0.25
    ds2 = LJDataset(**kw)
    assert torch.equal(ds2[5].pos, d.pos) and torch.equal(ds2[5].vel, d.vel)
    # a frame goes through the inverse pass like any latent batch
    from gpu_util import build_model
    from enflow_b200.data.synthetic import make_weights
    from enflow_b200.data.base import DataLoader
    model = build_model(make_weights(4, 128, 2, seed=1), 4, 2)
    batch = next(iter(DataLoader(ds, batch_size=1))).to(0)
    out = model.reverse(batch)
    assert torch.isfinite(out.pos).all() and torch.isfinite(out.h).all()
