"""Oracle of the generate-mode prior sampler: internal consistency on CPU.  The oracle itself is
parity-unpinned (the reference delegates to OpenMM, absent here); these tests pin it to the published definitions:
forces = -grad U, minimum image, plain cutoff, and the LangevinMiddle update with a = 1 being time reversible."""
import numpy as np

from oracle import enflow_oracle as orc
from enflow_b200.data.lj import arrange_points_on_grid


def _system(n=40, box=6.0, seed=3):
    rs = np.random.RandomState(seed)
    pos = arrange_points_on_grid(n, np.full(3, box), 0.6) + rs.uniform(-0.15, 0.15, size=(n, 3))
    return pos, np.full(3, box)


def test_forces_are_minus_gradient():
    pos, box = _system()
    u0, f = orc.lj_prior_energy_forces(pos, box, 0.1, 2.9)
    h = 1e-6
    for (i, c) in [(0, 0), (7, 1), (23, 2), (39, 0)]:
        p = pos.copy(); p[i, c] += h
        m = pos.copy(); m[i, c] -= h
        up, _ = orc.lj_prior_energy_forces(p, box, 0.1, 2.9)
        um, _ = orc.lj_prior_energy_forces(m, box, 0.1, 2.9)
        assert abs(-(up - um) / (2 * h) - f[i, c]) < 1e-5 * max(1.0, abs(f[i, c]))
    assert np.abs(f.sum(0)).max() < 1e-9          # Newton's third law


def test_minimum_image_and_cutoff():
    box = np.full(3, 5.0)
    pos = np.array([[0.2, 0.0, 0.0], [4.9, 0.0, 0.0]])          # 0.3 apart through the boundary
    u, f = orc.lj_prior_energy_forces(pos, box, 0.1, 3.0)
    q = 1.0 / (0.1 + 0.3)
    assert abs(u - 4 * (q ** 12 - q ** 6)) < 1e-12 * abs(u)
    assert f[0, 0] > 0 and f[1, 0] < 0                          # repelled away from each other across the wall
    pos[1, 0] = 3.3                                             # image distance 1.9, direct 3.1: image counts
    u, _ = orc.lj_prior_energy_forces(pos, box, 0.1, 3.0)
    q = 1.0 / (0.1 + 1.9)
    assert abs(u - 4 * (q ** 12 - q ** 6)) < 1e-12
    u, _ = orc.lj_prior_energy_forces(pos, box, 0.1, 1.5)       # beyond the cutoff: nothing
    assert u == 0.0


def test_middle_scheme_is_reversible_without_friction():
    pos, box = _system(n=20, box=5.0)
    vel = np.random.RandomState(1).normal(0, 1, size=pos.shape)
    z = np.zeros_like(pos)
    x, v = pos, vel
    for _ in range(10):
        x, v = orc.langevin_middle_step(x, v, box, 0.1, 2.4, 0.002, 1.0, 1.0, z)
    v = -v
    # the scheme kicks before drifting: its time reverse drifts before kicking, i.e. a forward run of the reversed
    # velocities returns to the start up to one kick (O(dt) in velocity, O(dt^2) in position)
    for _ in range(10):
        x, v = orc.langevin_middle_step(x, v, box, 0.1, 2.4, 0.002, 1.0, 1.0, z)
    assert np.abs(x - pos).max() < 5e-4


def test_grid_matches_reference_layout():
    pts = arrange_points_on_grid(10, np.array([4.0, 4.0, 4.0]), 0.5)
    assert pts.shape == (10, 3)
    assert pts.min() >= 0.5 - 1e-12 and pts.max() <= 3.5 + 1e-12
    assert len({tuple(np.round(p, 9)) for p in pts}) == 10
