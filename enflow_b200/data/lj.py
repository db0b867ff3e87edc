"""`lj` dataset plugin: the prior sampler of generate mode.

The reference builds this dataset by running OpenMM (`enflow/data/lj.py:32-89` system and soft Lennard-Jones force,
`enflow/data/simulated.py:83-132` LangevinMiddleIntegrator, minimisation, Maxwell velocities, a reporter that turns
every `interval`-th frame from `discard` on into a `Data` with random `h`, `g` ~ N(0, 1/sqrt(kBT)),
`simulated.py:57-77`).  Here the same run happens on the GPU through the C ABI (`enflow_lj_prior_*`,
`csrc/lj_prior.cu`) in the reduced units of the likelihood, so `generate.yaml` needs neither OpenMM nor a CPU
simulation.  Same YAML keys as the reference (`example/generate.yaml`): n_atoms, box, temp, friction, dt, n_iter,
interval, discard, softening, cutoff (in sigma), gap, node_nf, log, traj; `seed` is new (the run is reproducible).

What is and is not the same as upstream: potential, cutoff, minimum image, integrator scheme, friction and the
frame schedule follow the reference's configuration; the energy minimiser is a capped steepest descent instead
of OpenMM's L-BFGS; the random streams differ (Philox here), so frames agree in distribution, not bit for bit.
The run advances PHYSICAL time like OpenMM does upstream: the step is dt / tau with tau = sigma sqrt(M/eps) = 4.405 ps
(`time_to_lj_physical`; the reference's own `time_to_lj`, which it applies to the flow's dt, carries an amu-for-kg slip
and is 31.6 times smaller: with it 2000 steps of 4 fs moved the atoms ~0.2 sigma off the minimised lattice instead of
equilibrating the fluid).  Particles have unit mass, so velocities come out with <v^2> = kBT per component, the kinetic
term `Alchemical_NLL` assigns (`enflow/flow/loss.py:16-22`); positions sample the periodic, cut-off soft-LJ fluid of
`enflow/data/lj.py:65-76` ((s + r) softening, cutoff 3 sigma), which is the reference's prior but not literally the
non-periodic r^2 + s expression inside the likelihood.  Parity unpinned: OpenMM is not available here and the
reference holds no vectors for this path; tests check forces/trajectories against the oracle's restatement, the
thermostat and the potential-energy plateau statistically.
"""
import math

import numpy as np
import torch

from .. import _lib
from ..utils.conversion import dist_to_lj, kelvin_to_lj, time_to_lj_physical, lj_to_dist
from ..utils.helpers import apply_pbc
from .base import Data


def arrange_points_on_grid(n, box, gap):
    """The first n sites of a regular lattice spanning [gap, box - gap] on every axis: the start configuration of
    the reference's run (`enflow/data/lj.py:9-30`).  Sites per axis follow the reference (z: ceil(n^(1/3)), then y,
    then x) and the sites are enumerated z fastest, then x, then y, which is the order its flattened meshgrid has."""
    nz = int(np.ceil(n ** (1.0 / 3.0)))
    ny = int(np.ceil(np.sqrt(n / nz)))
    nx = int(np.ceil(n / (ny * nz)))
    site = np.arange(n)
    index = {2: site % nz, 0: (site // nz) % nx, 1: site // (nz * nx)}
    out = np.empty((n, 3), dtype=np.float64)
    for axis, count in ((0, nx), (1, ny), (2, nz)):
        step = (box[axis] - 2.0 * gap) / (count - 1) if count > 1 else 0.0
        out[:, axis] = gap + step * index[axis]
    return out


class LJDataset:
    def __init__(self, n_atoms, box, temp, friction, dt, n_iter, interval, softening, node_nf, discard=-1, cutoff=3.0,
                 gap=1.0, dist_unit='ang', time_unit='pico', r_cut=None, seed=0, log=None, traj=None,
                 minimize_iters=200, device=None, **_ignored):
        if not torch.cuda.is_available():
            raise RuntimeError('LJDataset runs its Langevin sampler on the GPU (enflow_lj_prior_*); no CUDA device is available')
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        L = _lib.lib()
        n = int(n_atoms)
        box_lj = np.array([dist_to_lj(float(b), dist_unit) for b in box], dtype=np.float64)
        gap_lj = dist_to_lj(float(gap), dist_unit)
        kBT = kelvin_to_lj(float(temp))
        dt_lj = time_to_lj_physical(float(dt), time_unit)      # OpenMM's physical step in units of tau (see module docstring)
        # friction is per picosecond (`simulated.py:109`: friction/(scale*ps)), dt in `time_unit`: a = exp(-gamma dt)
        dt_ps = float(dt) * (1.0 if time_unit == 'pico' else 1e-3)
        a = math.exp(-float(friction) * dt_ps)
        interval, n_iter = int(interval), int(n_iter)
        discard = interval if int(discard) == -1 else int(discard)          # simulated.py:87
        self.kBT, self.box, self.softening, self.cutoff = kBT, box_lj, float(softening), float(cutoff)
        self.r_cut = float(cutoff) if r_cut is None else dist_to_lj(float(r_cut), dist_unit)
        self.data_list, self.log = [], []

        pos = torch.tensor(arrange_points_on_grid(n, box_lj, gap_lj), dtype=torch.float64, device=dev).contiguous()
        vel = torch.zeros_like(pos)
        ws = torch.empty(int(L.enflow_lj_prior_workspace_doubles(n)), dtype=torch.float64, device=dev)
        energy = torch.zeros(2, dtype=torch.float64, device=dev)
        cbox = (_lib.C.c_double * 3)(*box_lj)
        p, st = _lib.ptr, _lib.stream()
        with torch.cuda.device(dev):
            _lib.check(L.enflow_lj_prior_minimize(p(pos), n, cbox, self.softening, self.cutoff, int(minimize_iters), 1e-3,
                                                  0.05, p(ws), st))
            _lib.check(L.enflow_lj_prior_velocities(p(vel), n, kBT, int(seed), st))       # setVelocitiesToTemperature
            gen = torch.Generator(device='cpu').manual_seed(int(seed))
            step = 0
            while step < n_iter:
                todo = min(interval - step % interval, n_iter - step)
                _lib.check(L.enflow_lj_prior_run(p(pos), p(vel), n, cbox, self.softening, self.cutoff, dt_lj, a, kBT, todo,
                                                 int(seed), step, p(ws), p(energy), st))
                step += todo
                if step % interval or step < discard:
                    continue
                e = energy.cpu()
                temp_now = float(e[1]) * 2.0 / (3.0 * n)
                self.log.append((step, float(e[0]), temp_now))
                sd = 1.0 / math.sqrt(kBT)
                h = torch.normal(0.0, sd, size=(n, int(node_nf)), generator=gen, dtype=torch.float64)
                g = torch.normal(0.0, sd, size=(n, int(node_nf)), generator=gen, dtype=torch.float64)
                box_t = torch.tensor(box_lj, dtype=torch.float64)
                wrapped = apply_pbc(pos.cpu(), box_t)                                     # simulated.py:45
                wrapped = wrapped - wrapped.mean(dim=0, keepdim=True)                     # transforms.Center (main.py:74)
                self.data_list.append(Data(z=['Ar'] * n, h=h.float(), g=g.float(), pos=wrapped.float(),
                                           vel=vel.cpu().float(), N=n, r_cut=self.r_cut,
                                           box=box_t.repeat(n, 1).float(),
                                           label=f'Simulated dataset: LJ Frame: {step}'))
        if log:
            with open(log, 'w') as f:
                f.write('#"Step","Potential Energy (eps)","Temperature (eps/kB)"\n')
                for s, u, t in self.log:
                    f.write(f'{s},{u},{t}\n')
        if traj:
            with open(traj, 'w') as f:
                for d in self.data_list:
                    f.write(f'{d.N}\n{d.label}\n')
                    for x in lj_to_dist(d.pos.double(), dist_unit).tolist():
                        f.write(f'Ar {x[0]:.6f} {x[1]:.6f} {x[2]:.6f}\n')

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
