"""CPU oracle for the enflow hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
package ``enflow_b200`` never does: it calls the CUDA library and fails loudly
when that library is missing.

What this is: a functional restatement, in float64 on the CPU, of the
reference algorithm for the path
    ArgMax dequantiser -> L x (neighbour list -> EGCL -> leap-frog coupling) -> ldj
    -> Alchemical_NLL, its gradients, and the inverse pass.
Each function cites the reference file:line it follows (paths relative to the
reference tree).  It is written against torch's CPU ATen ops rather than numpy
on purpose: the reference's arithmetic *is* ATen float64 (`enflow/flow/base.py:12`),
so the same primitives (``round`` half-to-even, ``scatter_add_`` in index order,
``nonzero`` row-major, 0-dim fp32 promotion) reproduce it bit for bit, and
autograd supplies the gradient oracle.  No reference source is copied.

Parity pin: the reference ships no tests (`.gitignore:1` hides them), so the
oracle is pinned against outputs of the reference itself, imported from
/root/reference in the build container by ``tests/golden/make_golden.py`` and
committed under ``tests/golden/*.npz`` (see ``tests/test_oracle_golden.py``).
"""
import math
import torch

F64 = torch.float64


# --------------------------------------------------------------------------- helpers
def wrap(x, period):
    """`enflow/utils/helpers.py:7-8` apply_pbc: x - round(x/period)*period (half-to-even)."""
    return x - (x / period).round() * period


def log_gaussian(z):
    """`helpers.py:4-5`: log(2*pi) is added once per call, not per element (quirk Q3)."""
    return -0.5 * ((z ** 2).sum() + math.log(2 * math.pi))


def segment_sum(data, seg, num):
    """`helpers.py:54-60` unsorted_segment_sum."""
    out = data.new_zeros((num, data.size(1)))
    out.scatter_add_(0, seg.unsqueeze(-1).expand(-1, data.size(1)), data)
    return out


def segment_mean(data, seg, num):
    """`helpers.py:63-70` unsorted_segment_mean; count clamped to >= 1 (quirk Q12)."""
    idx = seg.unsqueeze(-1).expand(-1, data.size(1))
    tot = data.new_zeros((num, data.size(1))).scatter_add_(0, idx, data)
    cnt = data.new_zeros((num, data.size(1))).scatter_add_(0, idx, torch.ones_like(data))
    return tot / cnt.clamp(min=1)


def silu(x):
    return x * torch.sigmoid(x)


def linear(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


# --------------------------------------------------------------------------- neighbour list
def periodic_images_within(pos, box, r_cut):
    """`helpers.py:15-29`: 27 images, order c outer / b / a inner over [-L, +L, 0] (Q10),
    kept if sum((p/(box+r_cut))^2) <= 1 (Q7).  r_cut is a 0-dim fp32 tensor (Q9)."""
    shifts = [torch.stack([a, b, c]) for c in (-box[2], box[2], box[2] * 0)
              for b in (-box[1], box[1], box[1] * 0) for a in (-box[0], box[0], box[0] * 0)]
    images = torch.cat([pos + s for s in shifts])
    keep = ((images / (box + r_cut)) ** 2).sum(dim=1) <= 1
    ids = torch.arange(pos.shape[0]).repeat(27)
    return images[keep], ids[keep]


def build_edges(pos, box, N, r_cut):
    """`enflow/data/base.py:122-144` Data.edges.

    pos [sumN,3] f64, box [sumN,3] f64 (row 0 of each molecule is used, base.py:130),
    N [B] int64, r_cut [B] float32.  Returns (row, col, edge_box).  Both columns of
    the (image point, atom) hit list are sent through id_mapping (quirk Q6); self
    pairs are dropped on the remapped labels, duplicates are kept (Q11).
    """
    rows, cols, boxes = [], [], []
    start = 0
    for m in range(N.numel()):
        n = int(N[m])
        p = pos[start:start + n]
        b = box[start]
        rc = r_cut[m]
        images, idmap = periodic_images_within(p, b, rc)
        r_sq = rc * rc                                    # fp32 product (Q9)
        dist_sq = (images.unsqueeze(1) - p).pow(2).sum(dim=2)
        hits = (dist_sq < r_sq).nonzero()
        lab = idmap[hits] + start
        lab = lab[lab[:, 0] != lab[:, 1]]
        rows.append(lab[:, 0])
        cols.append(lab[:, 1])
        boxes.append(b.repeat(lab.shape[0], 1))
        start += n
    return torch.cat(rows), torch.cat(cols), torch.cat(boxes)


def coord_diff(pos, row, col, edge_box):
    """`base.py:15-19`: minimum image with period box/2 (quirk Q8)."""
    return wrap(pos[row] - pos[col], edge_box * 0.5)


# --------------------------------------------------------------------------- EGCL
def egcl_forward(p, pre, h, pos, row, col, edge_box, coords_weight=1.0, trace=None):
    """`enflow/nn/egcl.py:77-93` with the default options (attention/norm_diff/tanh off).

    p: dict of parameters, pre: key prefix ('networks.3.').  Returns (Q, F, G).
    """
    d = coord_diff(pos, row, col, edge_box)
    radial = (d ** 2).sum(1, keepdim=True)                              # egcl.py:80
    e_in = torch.cat([h[row], h[col], radial], dim=1)                   # egcl.py:58
    x1 = silu(linear(e_in, p[pre + 'edge_nn.0.weight'], p[pre + 'edge_nn.0.bias']))
    edge_attr = silu(linear(x1, p[pre + 'edge_nn.2.weight'], p[pre + 'edge_nn.2.bias']))
    x3 = silu(linear(edge_attr, p[pre + 'coord_nn.0.weight'], p[pre + 'coord_nn.0.bias']))
    s = linear(x3, p[pre + 'coord_nn.2.weight'])                        # egcl.py:31,38 no bias
    trans = torch.clamp(d * s, min=-100, max=100)                       # egcl.py:72-73
    force = segment_mean(trans, row, h.size(0)) * coords_weight         # egcl.py:74-75
    agg = segment_sum(edge_attr, row, h.size(0))                        # egcl.py:66
    x4 = silu(linear(torch.cat([h, agg], dim=1), p[pre + 'node_nn.0.weight'], p[pre + 'node_nn.0.bias']))
    G = linear(x4, p[pre + 'node_nn.2.weight'], p[pre + 'node_nn.2.bias'])
    x6 = silu(linear(h, p[pre + 'vel_scaling_nn.0.weight'], p[pre + 'vel_scaling_nn.0.bias']))
    Q = linear(x6, p[pre + 'vel_scaling_nn.2.weight'], p[pre + 'vel_scaling_nn.2.bias'])
    if trace is not None:
        trace.append({'d': d, 'edge_attr': edge_attr, 'trans': trans, 'agg': agg, 's': s})
    return Q, force, G


# --------------------------------------------------------------------------- dequantiser
def argmax_forward(p, h, eps, pre='dequantize.'):
    """`enflow/nn/argmax.py:14-26`; eps is the float32 noise the reference draws at :17."""
    net = linear(silu(linear(h, p[pre + 'network.0.weight'], p[pre + 'network.0.bias'])),
                 p[pre + 'network.2.weight'], p[pre + 'network.2.bias'])
    log_scale, translate = torch.chunk(net, 2, dim=-1)
    u = translate + eps * log_scale.exp()
    log_q = log_gaussian(u) - log_scale.sum()
    T = (h * u).sum(-1, keepdim=True)
    z = h * u + (1 - h) * (T - torch.nn.functional.softplus(T - u))
    log_q = log_q - ((1 - h) * torch.nn.functional.logsigmoid(T - u)).sum()
    return z, log_q


def argmax_reverse(z):
    """`argmax.py:28-29`: one_hot(argmax)."""
    idx = torch.argmax(z, dim=-1)
    out = torch.zeros(z.shape, dtype=F64)
    return out.scatter_(1, idx.unsqueeze(1), 1)


# --------------------------------------------------------------------------- flow
def lf_forward(p, L, batch, dt, eps, trace=None, dequantize=True):
    """`enflow/flow/dynamics.py:10-23` LFIntegrator.forward.

    batch: dict h,g,pos,vel,box (f64), N (int64), r_cut (f32).  Returns (state dict, ldj, ldj_mol)
    where ldj_mol[B] is the per-molecule sum of Q (the scalar ldj = log_q + ldj_mol.sum()).
    """
    h, g, pos, vel = batch['h'], batch['g'], batch['pos'], batch['vel']
    box, N, r_cut = batch['box'], batch['N'], batch['r_cut']
    if dequantize:
        h, ldj = argmax_forward(p, h, eps)                              # dynamics.py:11 (Q4: + sign)
    else:
        ldj = torch.zeros((), dtype=F64)
    mol_of_atom = torch.repeat_interleave(torch.arange(N.numel()), N)
    ldj_mol = torch.zeros(N.numel(), dtype=F64)
    for i in range(L):
        row, col, ebox = build_edges(pos.detach(), box, N, r_cut)       # dynamics.py:13 (Q5)
        layer_trace = [] if trace is not None else None
        Q, Fo, G = egcl_forward(p, f'networks.{i}.', h, pos, row, col, ebox, trace=layer_trace)
        vel = torch.exp(Q) * vel + Fo * dt                              # dynamics.py:14
        g = g + G * dt                                                  # :15
        pos = wrap(pos + vel * dt, box)                                 # :17-18, base.py:119-120
        h = h + g * dt                                                  # :19
        ldj = ldj + Q.sum()                                             # :21 (Q2: not x3)
        ldj_mol = ldj_mol + segment_sum(Q, mol_of_atom, N.numel())[:, 0]
        if trace is not None:
            t = layer_trace[0]
            t.update({'row': row, 'col': col, 'Q': Q, 'F': Fo, 'G': G,
                      'h': h, 'g': g, 'pos': pos, 'vel': vel})
            trace.append(t)
    return {'h': h, 'g': g, 'pos': pos, 'vel': vel}, ldj, ldj_mol


def lf_reverse(p, L, batch, dt, quantize=True):
    """`dynamics.py:25-37` LFIntegrator.reverse; also returns per-molecule -sum(Q) (C4 extension)."""
    h, g, pos, vel = batch['h'], batch['g'], batch['pos'], batch['vel']
    box, N, r_cut = batch['box'], batch['N'], batch['r_cut']
    mol_of_atom = torch.repeat_interleave(torch.arange(N.numel()), N)
    neg_ldj_mol = torch.zeros(N.numel(), dtype=F64)
    for i in reversed(range(L)):
        h = h - g * dt
        pos = wrap(pos - vel * dt, box)
        row, col, ebox = build_edges(pos, box, N, r_cut)
        Q, Fo, G = egcl_forward(p, f'networks.{i}.', h, pos, row, col, ebox)
        g = g - G * dt
        vel = (vel - Fo * dt) / torch.exp(Q)
        neg_ldj_mol = neg_ldj_mol - segment_sum(Q, mol_of_atom, N.numel())[:, 0]
    if quantize:
        h = argmax_reverse(h)
    return {'h': h, 'g': g, 'pos': pos, 'vel': vel}, neg_ldj_mol


# --------------------------------------------------------------------------- likelihood
def lj_potential(pos, N, softening):
    """`enflow/flow/loss.py:11-19`: per molecule, upper triangle, r^2 == 0 dropped (Q13)."""
    H = 0
    start = 0
    for m in range(N.numel()):
        n = int(N[m])
        p = pos[start:start + n]
        dist_sq = torch.triu((p.unsqueeze(1) - p).pow(2).sum(dim=2))
        r_sq = dist_sq[dist_sq != 0] + softening
        r_6 = r_sq.pow(3)
        r_12 = r_6.pow(2)
        H = H + 4 * (1 / r_12 - 1 / r_6).sum()
        start += n
    return H


def alchemical_nll(state, ldj, N, kBT, softening=0.0, z_lj=10.0):
    """`loss.py:21-25` Alchemical_NLL.__call__."""
    H = lj_potential(state['pos'], N, softening) + 0.5 * (state['vel'] ** 2).sum()
    logZ = -int(N.sum()) * (math.log(z_lj) - 1.5 * math.log(2 * math.pi / kBT))
    log_px = -H / kBT + logZ + ldj + log_gaussian(state['h']) + log_gaussian(state['g'])
    return -log_px / N.numel()


# --------------------------------------------------------------------------- convenience
def to_torch(batch):
    out = {}
    for k, v in batch.items():
        t = torch.as_tensor(v)
        out[k] = t.to(F64) if (t.is_floating_point() and k != 'r_cut') else t
    return out


def params_to_torch(sd, requires_grad=False):
    return {k: torch.as_tensor(v, dtype=F64).clone().requires_grad_(requires_grad) for k, v in sd.items()}


def train_step(sd, L, batch, dt, eps, kBT, softening):
    """forward + NLL + backward (the `enflow/main.py:219-221` sequence). Returns loss, grads, state, ldj."""
    p = params_to_torch(sd, requires_grad=True)
    b = to_torch(batch)
    state, ldj, ldj_mol = lf_forward(p, L, b, dt, torch.as_tensor(eps))
    loss = alchemical_nll(state, ldj, b['N'], kBT, softening)
    loss.backward()
    grads = {k: v.grad for k, v in p.items()}
    return loss.detach(), grads, {k: v.detach() for k, v in state.items()}, ldj.detach(), ldj_mol.detach()


# ---- prior sampler for generate -----------------------------------------------------------------
# PARITY UNPINNED: the reference runs this through OpenMM (CustomNonbondedForce + LangevinMiddleIntegrator), which
# is not installed here and has no golden vectors in the reference.  These restate the PUBLISHED definitions the
# reference configures (enflow/data/lj.py:65 energy expression, :74-75 CutoffPeriodic at cutoff*sigma, no switching;
# OpenMM LangevinMiddleIntegrator: v += dt f/m; x += dt/2 v; v = a v + sqrt(kT(1-a^2)/m) R; x += dt/2 v,
# a = exp(-friction dt)) in reduced units (sigma = eps = m = 1) and serve as the checker of lj_prior.cu.
def lj_prior_energy_forces(pos, box, softening, cutoff):
    """numpy fp64: total energy and forces [N,3] of the periodic soft-LJ fluid (minimum image, plain cutoff)."""
    import numpy as np
    pos = np.asarray(pos, dtype=np.float64)
    box = np.asarray(box, dtype=np.float64)
    d = pos[:, None, :] - pos[None, :, :]
    d -= box * np.rint(d / box)
    r = np.sqrt((d ** 2).sum(-1))
    mask = (r < cutoff) & ~np.eye(len(pos), dtype=bool)
    q = np.where(mask, 1.0 / (softening + np.where(mask, r, 1.0)), 0.0)
    u_pair = 4.0 * (q ** 12 - q ** 6)
    g = np.where(mask, 4.0 * (12.0 * q ** 13 - 6.0 * q ** 7) / np.where(mask, r, 1.0), 0.0)       # -dU/dr / r
    return 0.5 * u_pair.sum(), (g[:, :, None] * d).sum(1)


def langevin_middle_step(pos, vel, box, softening, cutoff, dt, a, kBT, noise):
    """One LangevinMiddle step with the given standard-normal noise [N,3] (numpy fp64)."""
    import numpy as np
    _, f = lj_prior_energy_forces(pos, box, softening, cutoff)
    v = vel + dt * f
    x = pos + 0.5 * dt * v
    v = a * v + np.sqrt(kBT * (1.0 - a * a)) * noise
    x = x + 0.5 * dt * v
    return x, v
