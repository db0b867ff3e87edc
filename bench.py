"""Benchmark of the enflow hot path on B200: flow molecules/s for one training step
(LFIntegrator.forward + Alchemical_NLL + backward + gradient all-reduce + Adam), BASELINE.json metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c2]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` goes through
the public API with pinned HOST buffers (H2D of the batch and D2H of the loss inside the timed region).
`--impl reference` times the reference's own CPU path (baseline/_ref, vendored by build(); else the port under oracle/)
on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from enflow_b200.data import synthetic as syn  # noqa: E402

METRIC = 'flow molecules/sec (log-lik fwd+bwd)'
UNIT = 'molecules/s'
H = 128
L_LAYERS = 5
CONFIGS = {
    # name: (synthetic config, per-GPU batch, make_batch kwargs, nf, description)
    'c1': ('c1', 64, {}, 4, 'example/train.yaml shape: 22-atom conformers, radius graph, batch 64'),
    'c2': ('c2', 1024, {'n_atoms': 29}, 5, 'QM9-sized synthetic molecules, 29 atoms fully connected, batch 1024 per GPU'),
    'c3': ('c3', 1024, {}, 1, 'LJ-55 clusters, 55 particles fully connected, batch 1024 per GPU'),
    'c4': ('c4', 16384, {}, 4, 'example/generate.yaml inverse pass: 22-atom latents, 16384 conformers per launch per GPU, '
                               'state + per-molecule log-det out'),
    'c5': ('c5', 32, {'n_atoms': 500}, 5, '500-atom fragments, radius-cutoff graph, batch 32 per GPU'),
}


def workload_config(config, batch, world, nf, desc, n_atoms, E):
    """`config` of the JSON line: describes the WORKLOAD only, identical for the b200 arm and the reference arm (the
    arm-specific facts - arithmetic mode, launch mode, the CPU sample size - live in their own keys)."""
    generate = config == 'c4'
    return {'workload': desc, 'per_gpu_batch': batch, 'global_batch': batch * world, 'atoms_per_gpu': n_atoms,
            'edges_per_layer_per_gpu': E, 'layers': L_LAYERS, 'hidden': H, 'nf': nf,
            'step': ('LFIntegrator.reverse (inverse pass with per-molecule log-det, no collective)' if generate else
                     'LFIntegrator.forward + Alchemical_NLL + backward (all parameter grads) + gradient all-reduce (N > 1) + Adam'),
            'parallelism': f'dp{world}',
            'l2': 'no explicit flush: every layer of every step streams far more than the 126 MB L2 through HBM '
                  '(C2: E*H*4 = 426 MB of edge gradients per layer)'}


def static_edge_count(config, arrs):
    """edges per layer: all ordered pairs for the fully connected configs, else measured once with the oracle-free K0 path"""
    if config in ('c2', 'c3'):
        return int((arrs['N'] * (arrs['N'] - 1)).sum())
    return None


NCU_SUMMARY = os.path.join('profiles', 'r2_edge_kernels_summary.csv')


def ncu_summary(kernel):
    """dram traffic per launch and tensor-pipe activity of `kernel` from the committed `ncu --set full` summary of this
    bench command (profiles/r2_edge_kernels_summary.csv, written by tools/ncu_summary.py); None if not captured."""
    path = os.path.join(ROOT, NCU_SUMMARY)
    if not os.path.exists(path):
        return None
    import csv
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    for r in rows[1:]:
        if r and r[0].startswith(kernel):
            def col(prefix):
                for i, h in enumerate(hdr):
                    if h.startswith(prefix):
                        unit = h[h.index('[') + 1:h.index(']')] if '[' in h else ''
                        v = float(r[i])
                        return v * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}.get(unit, 1.0)
                return None
            rd, wr = col('dram__bytes_read.sum'), col('dram__bytes_write.sum')
            return {'traffic': (rd + wr) if rd is not None and wr is not None else None,
                    'pipe_tensor_active_pct': col('sm__pipe_tensor_cycles_active'), 'ncu_us': col('gpu__time_duration.sum')}
    return None


def edge_flops(E, nf, train):
    """per edge 2H(2nf+1) + 4H^2 + 2H forward; backward = 2x (dgrad + wgrad)."""
    fwd = E * (2 * H * (2 * nf + 1) + 4 * H * H + 2 * H)
    return 3 * fwd if train else fwd


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return {'hbm_gbs': p['hbm_gbs'], 'bf16_burst': p['bf16_tflops'],
                'bf16_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_burst': 1590.0, 'bf16_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler:
    """nvidia-smi clock/throttle samples during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r for r in self.rows if len(r) >= 7 and r[0].isdigit()]
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[3 + i] == 'Active' for r in rows)]
        return {'sm_mhz': float(np.median([float(r[0]) for r in rows])), 'sm_max_mhz': float(rows[0][1]),
                'power_w_max': max(float(r[2]) for r in rows if r[2].replace('.', '').isdigit()),
                'samples': len(rows), 'reasons': reasons}


REF_DIR = os.path.join(ROOT, 'baseline', '_ref')      # the unmodified reference tree + a 2-file rdkit stub (build())


def _import_reference():
    """enflow.* of the UNMODIFIED reference, vendored by __graft_entry__.build() to baseline/_ref (git-ignored, travels to the
    GPU box).  None when it was never vendored (the port under oracle/ then stands in)."""
    if not os.path.exists(os.path.join(REF_DIR, 'enflow', 'flow', 'dynamics.py')):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import enflow.data.base as rbase
    import enflow.flow.dynamics as rdyn
    import enflow.flow.loss as rloss
    import enflow.nn.argmax as rargmax
    import enflow.nn.egcl as regcl
    return rbase, regcl, rargmax, rdyn, rloss


def cpu_reference_throughput(config, nf, sample_mols, steps, warmup, kwargs):
    """Time the reference's own CPU path (fp64 PyTorch, all host threads) on a bounded sample of the workload:
    enflow/main.py:150-153 model, batch through DataLoader.collater (data/base.py:162-174), LFIntegrator.forward +
    Alchemical_NLL + backward (main.py:219-221), or LFIntegrator.reverse for the generate config (main.py:269).
    Falls back to the CPU port of the same algorithm (oracle/, kind 'port') when baseline/_ref is absent.
    Returns (molecules/s, threads, seconds per step, kind)."""
    torch.set_num_threads(os.cpu_count())
    arrs = syn.make_batch(config, sample_mols, **kwargs)
    sd = syn.make_weights(nf, H, L_LAYERS, seed=0, coord_gain=0.5)
    eps = syn.make_noise(int(arrs['N'].sum()), nf)
    mods = _import_reference()
    if mods is None:
        from oracle import enflow_oracle as orc

        def one():
            if config == 'c4':
                with torch.no_grad():
                    orc.lf_reverse(orc.params_to_torch(sd), L_LAYERS, orc.to_torch(arrs), syn.TRAIN_DT)
            else:
                orc.train_step(sd, L_LAYERS, arrs, syn.TRAIN_DT, eps, syn.TRAIN_KBT, syn.TRAIN_SOFTENING)
        kind = 'port'
    else:
        rbase, regcl, rargmax, rdyn, rloss = mods
        model = rdyn.LFIntegrator([regcl.EGCL(nf, nf, H) for _ in range(L_LAYERS)], rargmax.ArgMax(nf, H), dt=syn.TRAIN_DT)
        model.load_state_dict({k: torch.tensor(v, dtype=torch.float64) for k, v in sd.items()})
        nll = rloss.Alchemical_NLL(kBT=syn.TRAIN_KBT, softening=syn.TRAIN_SOFTENING)

        def batch():
            mols, o = [], 0
            for m, n in enumerate(arrs['N']):
                n = int(n)
                sl = slice(o, o + n)
                mols.append(rbase.Data(z=['X'] * n, h=torch.tensor(arrs['h'][sl]), g=torch.tensor(arrs['g'][sl]),
                                       pos=torch.tensor(arrs['pos'][sl]), vel=torch.tensor(arrs['vel'][sl]), N=n,
                                       r_cut=float(arrs['r_cut'][m]), box=torch.tensor(arrs['box'][sl]), label=[0] * n))
                o += n
            return rbase.DataLoader.collater(None, mols)      # an ordinary method that does not touch self

        def one():
            data = batch()
            if config == 'c4':
                with torch.no_grad():
                    model.reverse(data)
            else:
                model.zero_grad(set_to_none=True)
                out, ldj = model(data)
                nll(out, ldj).backward()
        kind = 'reference'
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return sample_mols / sec, torch.get_num_threads(), sec, kind


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores: the unmodified tree vendored
    to baseline/_ref by build() (kind 'reference'); the port under oracle/ (pinned against it by tests/golden) only if that
    tree is absent (kind 'port').  Same config / steps / warmup as the b200 arm; each step is a bounded sample."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    config, batch, kwargs, nf, desc = CONFIGS[args.config]
    batch = args.batch or batch
    world = int(os.environ.get('WORLD_SIZE', 1))
    sample = {'c1': 64, 'c2': 48, 'c3': 12, 'c4': 256, 'c5': 1}[args.config]
    arrs = syn.make_batch(config, 2, **kwargs)            # layout only: every bench config has a fixed atom count per molecule
    n_mol = int(arrs['N'][0])
    n_atoms = n_mol * batch
    E = n_mol * (n_mol - 1) * batch if config in ('c2', 'c3') else None      # radius-graph configs: data dependent
    mols_s, cores, sec, kind = cpu_reference_throughput(config, nf, sample, args.steps, args.warmup, kwargs)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': mols_s, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(config, batch, world, nf, desc, n_atoms, E),
        'cpu_baseline': {'value': mols_s, 'unit': UNIT, 'cores': cores, 'kind': kind,
                         'sample': f'{sample} molecules of the same config per step (per-molecule throughput), fp64 torch CPU, '
                                   f'all host threads; {args.warmup} warm-up + {args.steps} timed steps'},
        'e2e': {'value': mols_s, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c2', choices=list(CONFIGS))
    ap.add_argument('--batch', type=int, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--precision', default=None, choices=['fp32', 'fp32_tc', 'bf16'],
                    help='edge-MLP arithmetic; default: fp32_tc (fp32-accurate, tensor cores), bf16 for c5 as BASELINE.json asks')
    ap.add_argument('--no-graph', action='store_true', help='launch every step eagerly instead of replaying a CUDA graph')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.precision is None:
        args.precision = 'bf16' if args.config == 'c5' else 'fp32_tc'
    if args.impl == 'reference':
        return run_reference(args)

    import torch.distributed as dist
    from enflow_b200 import _lib
    from enflow_b200.data.base import Data
    from enflow_b200.flow.dynamics import LFIntegrator
    from enflow_b200.flow.loss import Alchemical_NLL
    from enflow_b200.nn.argmax import ArgMax
    from enflow_b200.nn.egcl import EGCL

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    config, batch, kwargs, nf, desc = CONFIGS[args.config]
    batch = args.batch or batch

    # model: L distinct EGCLs + ArgMax (enflow/main.py:150-153), seeded random weights
    sd = syn.make_weights(nf, H, L_LAYERS, seed=0, coord_gain=0.5)
    model = LFIntegrator([EGCL(nf, nf, H) for _ in range(L_LAYERS)], ArgMax(nf, H), dt=syn.TRAIN_DT)
    model.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    model = model.to(dev)
    model.precision = args.precision
    if world > 1:
        dist.broadcast(model.flat_params, 0)
        model._dp_group = dist.group.WORLD
    nll = Alchemical_NLL(kBT=syn.TRAIN_KBT, softening=syn.TRAIN_SOFTENING)
    from enflow_b200.optim import FlatAdam
    opt = FlatAdam(model, lr=1e-3)          # one fused kernel over the flat parameter / gradient buffers

    # per-rank synthetic batch (weak scaling: fixed per-GPU batch), pinned host copy + resident device copy
    arrs = syn.make_batch(config, batch, seed=1234 + 10 * rank + {'c1': 1, 'c2': 2, 'c3': 3, 'c4': 4, 'c5': 5}[config], **kwargs)
    f32 = lambda k: torch.tensor(arrs[k], dtype=torch.float32)
    host = Data(h=f32('h'), g=f32('g'), pos=f32('pos'), vel=f32('vel'), N=torch.tensor(arrs['N']),
                r_cut=torch.tensor(arrs['r_cut']), box=f32('box')).pin_memory()
    resident = host.to(dev)
    resident.meta()
    n_atoms = int(arrs['N'].sum())
    E_static = static_edge_count(config, arrs)          # all ordered pairs (fully connected configs) or None
    E = E_static
    h2d = sum(t.numel() * t.element_size() for t in (host.h, host.g, host.pos, host.vel, host.box, host.N, host.r_cut))

    def view(d):      # fresh wrapper: forward() rebinds the fields of the Data it is given
        v = Data(h=d.h, g=d.g, pos=d.pos, vel=d.vel, N=d.N, r_cut=d.r_cut, box=d.box, device=d.device)
        v._meta = d._meta
        return v

    generate = config == 'c4'

    def step(d):
        if generate:                                          # dynamics.py:25-37 + per-molecule -sum(Q)
            out = model.reverse(d)
            return out.neg_ldj_mol.sum()
        opt.zero_grad(set_to_none=True)
        out, ldj = model(d)                                   # eps drawn with torch.randn like argmax.py:17
        loss = nll(out, ldj)
        loss.backward()                                       # includes the flat-gradient all-reduce when world > 1
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dbg = (lambda m: print(f'[bench dbg rank {rank}] {m}', file=sys.stderr, flush=True)) if os.environ.get('ENFLOW_GRAPH_DEBUG') \
        else (lambda m: None)
    model.check_status = True
    for _ in range(args.warmup):
        step(view(resident))
    dbg('eager warm-up done')
    model.check_status = False      # capacity is now known; no host sync inside the timed device loop
    barrier()
    if E is None:          # radius-graph configs: the first layer's neighbour list, measured
        E = int(view(resident).build_edges(reference_order=False).row.numel())

    # The whole step (forward C call, likelihood, backward C call, all-reduce, Adam) has no host synchronisation, so it
    # is captured once in a CUDA graph and replayed (enflow_b200.graph); --no-graph launches every kernel eagerly.
    L = _lib.lib()
    gstep, graph_note, launches_per_step = None, 'eager launches', None
    # (with more than one rank the NCCL all-reduce is not captured: two graphs with one eager all-reduce between them)
    if not args.no_graph:
        try:
            from enflow_b200.graph import GraphedReverse, GraphedTrainStep
            L.enflow_launch_count(1)
            if generate:
                gstep = GraphedReverse(model, resident, warmup=1)
            else:
                gstep = GraphedTrainStep(model, nll, opt, resident, warmup=1)
            launches_per_step = int(L.enflow_launch_count(1)) // 2      # 1 eager warm-up + 1 captured step
            graph_note = 'CUDA graph replay of the whole step'
            if world > 1 and not generate:
                graph_note += (' (two graphs around one eager NCCL all-reduce of the flat gradient buffer)' if gstep.dp_eager
                               else ' (ONE graph: the NCCL all-reduce of the flat gradient buffer is captured with the kernels)')
        except Exception as exc:      # keep measuring: eager path
            gstep, graph_note = None, f'eager launches (graph capture failed: {type(exc).__name__}: {exc})'
            torch.cuda.synchronize()

    def run_resident():
        if gstep is None:
            return step(view(resident))
        return gstep().neg_ldj_mol.sum() if generate else gstep()

    def run_host():          # H2D of the batch from pinned host memory every step
        if gstep is None:
            return step(host.to(dev))
        return gstep(host).neg_ldj_mol.sum() if generate else gstep(host)

    # ---- end-to-end through the public API with host buffers: H2D of the batch + D2H of the loss every step
    dbg(f'graph built: {graph_note}')
    model.check_status = gstep is None
    for _ in range(args.warmup):       # warm the host-buffer path too (allocator growth for the per-step device batch)
        run_host().item()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss_host = run_host().item()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    dbg('e2e loop done')
    model.check_status = False

    # ---- device-timed region: inputs resident in HBM, K steps between CUDA events
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)        # let nvidia-smi spin up so samples fall inside the (short) timed region
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.enflow_launch_count(1)
    dbg('before timed barrier')
    barrier()
    dbg('timed loop')
    ev0.record()
    for _ in range(args.steps):
        run_resident()
    ev1.record()
    barrier()
    dbg('timed loop done')
    ms_total = ev0.elapsed_time(ev1)
    launches = launches_per_step * args.steps if gstep is not None else int(L.enflow_launch_count(1))
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel durations: the same steps launched eagerly with CUDA events around every kernel family
    L.enflow_timing_enable(1)
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev2.record()
    for _ in range(args.steps):
        step(view(resident))
    ev3.record()
    barrier()
    dbg('eager kernel-timing loop done')
    ms_eager = ev2.elapsed_time(ev3)
    fam = _lib.timing_read()
    L.enflow_timing_enable(0)
    dbg('timing read')

    t = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    dbg('max over ranks done')

    if rank == 0:
        pk = peaks()
        ms_step = ms_total / args.steps
        mols = batch * world
        # one timing family per kernel: average launch duration (CUDA events on the launch stream, eager launches) and
        # share of the eagerly launched step
        avg = {k: (v[0] / v[1] if v[1] else 0.0) for k, v in fam.items()}
        share = {k: v[0] / ms_eager for k, v in fam.items()}
        dom = max(fam, key=lambda k: fam[k][0])
        tc_mode = args.precision != 'fp32'
        flops = {'edge_fwd': edge_flops(E, nf, False), 'edge_bwd': 2 * edge_flops(E, nf, False)}
        roof = None
        if dom in flops and avg[dom] > 0:
            ach = flops[dom] / (avg[dom] * 1e-3) / 1e12
            kname = 'k_' + dom + ('_tc' if tc_mode else '')
            ncu = ncu_summary(kname + ('<1>' if args.precision == 'fp32_tc' else '<0>') if tc_mode else kname) \
                if (args.config == 'c2' and batch == 1024) else None
            mma_per_gemm = {'fp32': 0, 'fp32_tc': 3, 'bf16': 1}[args.precision]
            gemms = 6 if dom == 'edge_bwd' else 2
            roof = {'kernel': kname, 'bound': 'tensor', 'achieved': ach, 'peak': pk['bf16_sustained'],
                    'unit': 'TFLOP/s', 'frac': ach / pk['bf16_sustained'],
                    'traffic': ncu['traffic'] if ncu else None,
                    'traffic_source': NCU_SUMMARY + ' (dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full '
                                      'of this command)' if ncu and ncu['traffic'] else None,
                    'pipe_tensor_active_pct': ncu['pipe_tensor_active_pct'] if ncu else None,
                    'avg_launch_ms': avg[dom], 'launches_per_step': fam[dom][1] / args.steps,
                    'peak_source': pk['source'] + ' bf16 dense sustained (kernel timed inside a long step)',
                    'algorithmic_flops_per_launch': flops[dom],
                    'note': ('achieved = algorithmic FLOPs of the two HxH layers'
                             + (' (x2: dgrad + wgrad)' if dom == 'edge_bwd' else '') + ' / CUDA-event time of this kernel alone. '
                             f'The tensor pipe executes {mma_per_gemm} bf16 MMA(s) per GEMM term (the bf16x3 operand split keeps '
                             f'fp32 parity) and {gemms} GEMMs per tile'
                             + (' (2 recompute + 2 dgrad + 2 wgrad)' if dom == 'edge_bwd' else '')
                             + '; pipe_tensor_active_pct is what ncu measured for the executed MMAs. '
                             + ('The kernel is bound by the L1/shared-memory data pipe (MMA operand reads + epilogue traffic)'
                                if dom == 'edge_bwd' else 'The kernel is bound by the SFU (three sigmoid passes per element)')
                             + ', DESIGN.md section 4') if tc_mode
                            else 'fp32 FFMA cross-check path: dense layers on the CUDA cores'}
        elif avg.get('edges'):
            # radius-graph / small-batch configs are dominated by K0 (neighbour list): integer + fp64 work whose
            # algorithmic traffic is the positions in and the edge list out
            k0_bytes = n_atoms * 24 + E * 8 + (n_atoms + 1) * 4
            a = k0_bytes / (avg['edges'] * 1e-3) / 1e9
            roof = {'kernel': 'K0 neighbour list (k_edges_survivors + k_edges_hits x2 + scans)', 'bound': 'hbm',
                    'achieved': a, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': a / pk['hbm_gbs'], 'traffic': None,
                    'peak_source': pk['source'] + ' HBM copy bandwidth', 'algorithmic_bytes_per_launch': k0_bytes,
                    'note': f'dominant family of this config is {dom}; K0 is latency-bound fp64 image tests over 27 n points per '
                            'molecule, far from the HBM roofline by construction (DESIGN.md section 4)'}
        # the other tensor-core edge kernel (same arithmetic as `roofline`, its own timing family)
        tensor_other = {}
        if tc_mode and roof is not None and roof.get('bound') == 'tensor':
            for other in ('edge_fwd', 'edge_bwd'):
                if other != dom and avg.get(other):
                    a = flops[other] / (avg[other] * 1e-3) / 1e12
                    kn = 'k_' + other + '_tc'
                    n2 = ncu_summary(kn + ('<1>' if args.precision == 'fp32_tc' else '<0>')) \
                        if (args.config == 'c2' and batch == 1024) else None
                    tensor_other[other] = {'kernel': kn, 'bound': 'tensor', 'achieved': a, 'peak': pk['bf16_sustained'],
                                           'unit': 'TFLOP/s', 'frac': a / pk['bf16_sustained'], 'avg_launch_ms': avg[other],
                                           'algorithmic_flops_per_launch': flops[other],
                                           'traffic': n2['traffic'] if n2 else None,
                                           'pipe_tensor_active_pct': n2['pipe_tensor_active_pct'] if n2 else None}
        # the HBM-bound kernels the north star names (SURVEY 8d byte counts), each from its OWN timing family
        hbm = {}
        seg_bytes = E * H * 4 + (n_atoms + 1) * 4 + n_atoms * H * 4
        cpl_bytes = n_atoms * (19 + 5 * nf) * 4 + batch * 4
        for fam_name, kernel, nbytes in (('seg_cols', 'k_segment_sum128<0,1> (dS: column-grouped sum of dz1)', seg_bytes),
                                         ('coupling_fwd', 'k_coupling_fwd', cpl_bytes)):
            if avg.get(fam_name):
                a = nbytes / (avg[fam_name] * 1e-3) / 1e9
                hbm[fam_name] = {'kernel': kernel, 'bound': 'hbm', 'achieved': a, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                                 'frac': a / pk['hbm_gbs'], 'bytes': nbytes, 'avg_launch_ms': avg[fam_name]}
        cfg = workload_config(config, batch, world, nf, desc, n_atoms, E_static)
        line = {
            'metric': METRIC, 'value': mols * args.steps / (ms_total * 1e-3), 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': cfg,
            'edge_mlp': args.precision, 'edges_per_layer_measured': E,
            'clocks': clocks,
            'e2e': {'value': mols * args.steps / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': 4, 'last_loss': loss_host},
            'gpu_launches': launches,
            'launch_mode': graph_note,
            'eager_ms_per_step': ms_eager / args.steps,
            'roofline': roof,
            'roofline_tensor_kernels': tensor_other,
            'roofline_hbm_kernels': hbm,
            'kernel_ms_per_step': {k: v[0] / args.steps for k, v in fam.items()},
            'kernel_share_of_step': share,
            'kernel_timing_note': 'kernel_ms_per_step / kernel_share_of_step come from the eagerly launched pass '
                                  '(eager_ms_per_step, CUDA events around every kernel family, everything on one stream); ms_per_step '
                                  'is the graph replay, in which the weight-gradient reductions and the row run sums of the '
                                  'backward pass run on a side stream beside the main chain, so the family times sum to more '
                                  'than ms_per_step',
        }
        if world == 1 and not args.no_cpu_baseline:
            sample = {'c1': 64, 'c2': 48, 'c3': 12, 'c4': 256, 'c5': 1}[args.config]
            mols_s, cores, sec, kind = cpu_reference_throughput(config, nf, sample, 2, 1, kwargs)
            line['cpu_baseline'] = {'value': mols_s, 'unit': UNIT, 'cores': cores, 'kind': kind,
                                    'sample': f'{sample} molecules of the same config per step, fp64 torch CPU, '
                                              + ('the unmodified reference (baseline/_ref)' if kind == 'reference' else
                                                 'port of the reference algorithm (oracle/)') + ', 1 warm-up + 2 timed steps'}
        print(json.dumps(line), flush=True)
    if world > 1:
        # tear down: the graphs that captured the gradient all-reduce go first (destroying an NCCL communicator while a
        # captured graph still references it does not return); if the teardown stalls anyway, leave without it
        import gc
        gstep = None
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        done = threading.Event()
        threading.Thread(target=lambda: (dist.destroy_process_group(), done.set()), daemon=True).start()
        if not done.wait(20.0):
            dbg('destroy_process_group stalled: exiting without it')
            sys.stdout.flush()
            os._exit(0)
        dbg('process group destroyed')


if __name__ == '__main__':
    main()
