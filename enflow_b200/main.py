"""YAML-driven train / generate driver with the reference's interface (`enflow/main.py:34-288`).

``Main(world_size, world_rank, local_rank, num_cpus_per_task)(yaml_path)``.  What changes:
  * the model runs on the B200 path (``enflow_b200.flow`` / ``enflow_b200.nn``); there is no CPU mode;
  * data parallelism is one all-reduce of the flat gradient buffer (``enflow_b200.parallel``) instead of DDP;
    ranks come from the explicit arguments (SLURM, `__main__.py:6`) or from torchrun's env;
  * ``batch_size`` is read from ``training`` or ``dataset`` (the shipped `example/train.yaml:7` puts it under
    ``dataset`` while `main.py:126` reads ``training``: KeyError upstream);
  * the checkpoint is written without the ``.module`` indirection that crashes serial runs (`main.py:238`);
    the dict schema and key names are the reference's (`main.py:236-250`), so checkpoints interchange.
Dataset plugins follow `main.py:67-68`: module ``enflow_b200.data.<type>`` exporting ``<TYPE>Dataset``.
"""
import importlib
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist
import yaml
from torch.utils.data.distributed import DistributedSampler

from .data.base import DataLoader
from .flow.loss import Alchemical_NLL
from .nn.argmax import ArgMax
from .nn.egcl import EGCL
from .optim import FlatAdam
from .parallel import env_ranks, init_data_parallel
from .utils.conversion import SIGMA_M, kelvin_to_lj, lj_to_kelvin, time_to_lj


def eprint(*args, **kwargs):
    print(*args, file=sys.stderr, **kwargs)


def write_xyz(out, file):
    """`main.py:27-32`: every atom labelled Ar, positions in Angstrom."""
    with open(file, 'w') as f:
        f.write('%d\n%s\n' % (int(out.pos.shape[0]), ' '))
        for x in out.pos.detach().cpu().double() * SIGMA_M * 1e10:
            f.write('%s %.18g %.18g %.18g\n' % ('Ar', x[0].item(), x[1].item(), x[2].item()))


class Main:
    def __init__(self, world_size=None, world_rank=None, local_rank=None, num_cpus_per_task=None):
        self.world_size, self.world_rank, local = env_ranks(world_size, world_rank, local_rank)
        self.ddp = self.world_size > 1
        if not torch.cuda.is_available():
            raise RuntimeError('enflow_b200.Main needs a CUDA device (the B200 path has no CPU fallback)')
        torch.cuda.set_device(local)
        self.local_rank = torch.device('cuda', local)
        self.num_cpus_per_task = int(num_cpus_per_task or 0)
        if self.ddp:
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            os.environ.setdefault('MASTER_PORT', '29500')
            if not dist.is_initialized():
                dist.init_process_group('nccl', init_method='env://', rank=self.world_rank,
                                        world_size=self.world_size, device_id=self.local_rank)
            if self.world_rank == 0:
                eprint(f'Running data parallel on {self.world_size} ranks', flush=True)
        else:
            print('Running serially', flush=True)

    def _setup_dataset(self, label, args):
        dtype = args[label]['type']
        cls = getattr(importlib.import_module(f'enflow_b200.data.{dtype}'), f'{dtype.upper()}Dataset')
        kwargs = {k: v for k, v in args[label].items() if k not in ('batch_size', 'type')}
        if 'units' in args:                                         # main.py:71-72
            kwargs.setdefault('dist_unit', args['units'].get('dist', 'ang'))
            kwargs.setdefault('time_unit', args['units'].get('time', 'pico'))
        return cls(**kwargs)

    def setup(self, input):
        self.start_epoch = 0
        checkpoint = None
        with open(input) as f:
            args = yaml.load(f, Loader=yaml.FullLoader)
        self.mode = {'train': 'train', 'generate': 'gen', 'dataset': 'data'}.get(args['mode'])
        if self.mode is None:
            raise ValueError(f"unknown mode {args['mode']!r}")
        dyn = args.get('dynamics', {})
        self.checkpoint_path = dyn.get('checkpoint_path', '')
        if self.checkpoint_path and os.path.exists(self.checkpoint_path):
            if self.world_rank == 0:
                print('Loading from saved state', flush=True)
            checkpoint = torch.load(self.checkpoint_path, weights_only=False, map_location='cpu')
            node_nf, self.hidden_nf, self.n_iter = checkpoint['node_nf'], checkpoint['hidden_nf'], checkpoint['n_iter']
            dt, self.integrator = checkpoint['dt'], checkpoint['integrator']
            lj_kBT, softening = checkpoint['lj_kBT'], checkpoint['softening']
        elif self.mode != 'data':
            self.hidden_nf = int(dyn['network']['hidden_nf'])
            self.n_iter = int(dyn['n_iter'])
            dt = time_to_lj(float(dyn['dt']), unit=args['units']['time'])
            self.integrator = dyn['integrator'].lower()
            lj_kBT = kelvin_to_lj(float(args['training']['loss']['temp']))
            softening = float(args['training']['loss']['softening'])
        if self.mode == 'gen':
            batch_size = int(args['dataset'].get('batch_size', 1))
            if checkpoint and args['dataset']['type'] == 'lj':      # main.py:118-123: the prior is the trained model's
                ds = args['dataset']
                ds['node_nf'], ds['softening'], ds['temp'] = node_nf, softening, lj_to_kelvin(lj_kBT)
                ds['box'] = [float(b) for b in ds['box']]
                ds['n_atoms'] = int(ds['n_atoms'])
        else:
            tr = args.get('training', {})
            batch_size = int(tr['batch_size'] if 'batch_size' in tr else args['dataset']['batch_size'])
        self.dataset = self._setup_dataset('dataset', args)
        if self.mode == 'data':
            return
        if self.ddp:
            self.sampler = DistributedSampler(self.dataset, num_replicas=self.world_size, rank=self.world_rank, shuffle=True)
            self.train_loader = DataLoader(self.dataset, batch_size=batch_size, num_workers=self.num_cpus_per_task,
                                           pin_memory=False, shuffle=False, sampler=self.sampler, drop_last=False)
        else:
            self.train_loader = DataLoader(self.dataset, batch_size=batch_size, shuffle=self.mode == 'train')
        if not checkpoint:
            node_nf = self.dataset.node_nf
        networks = [EGCL(node_nf, node_nf, self.hidden_nf) for _ in range(self.n_iter)]       # main.py:150-151 (Q14)
        cls = getattr(importlib.import_module('enflow_b200.flow.dynamics'), f'{self.integrator.upper()}Integrator')
        self.model = cls(networks, ArgMax(node_nf, self.hidden_nf), dt=dt).to(self.local_rank)
        if 'precision' in dyn:
            self.model.precision = dyn['precision']
        if checkpoint:
            self.model.load_state_dict(checkpoint['model_state_dict'])
            self.start_epoch = checkpoint['epoch'] + 1
        if self.ddp:
            init_data_parallel(self.model)
        if self.mode == 'gen':
            return
        tr = args['training']
        self.log_interval = int(tr['log_interval'])
        # not in the reference: replay the whole step as one CUDA graph when the batches share one layout
        self.cuda_graph = bool(tr.get('cuda_graph', False))
        self.num_epochs = int(tr['num_epochs'])
        # Adam as in main.py:177, as one fused kernel over the flat buffers; its state_dict has torch.optim.Adam's layout
        self.optimizer = FlatAdam(self.model, lr=float(tr['lr']))
        self.scheduler = None
        if tr.get('scheduler'):
            step, gamma = float(tr['scheduler_step']), float(tr['gamma'])
            if step != 0 and gamma != 0:
                self.scheduler = torch.optim.lr_scheduler.StepLR(self.optimizer, step_size=step, gamma=gamma)
        if self.world_rank == 0:
            eprint(f'Loss function parameters: softening={softening}, kBT={lj_kBT}', flush=True)
        self.nll = Alchemical_NLL(kBT=lj_kBT, softening=softening)
        if checkpoint:
            self.optimizer.load_state_dict(checkpoint['optimizer_state_dict'])
            if self.scheduler and 'scheduler_state_dict' in checkpoint:
                self.scheduler.load_state_dict(checkpoint['scheduler_state_dict'])

    def _graphed_step(self, data):
        """One optimizer step through `GraphedTrainStep`; None when this batch does not have the captured layout
        (e.g. the last, shorter batch of an epoch), in which case the caller launches the step eagerly."""
        from .graph import EdgeCapacityOverflow, GraphedTrainStep
        n_cpu = data.meta()[3]
        g = getattr(self, '_gstep', None)
        if g is None:
            # the capture's warm-up IS this batch's step (one eager step), the capture itself executes nothing
            self._gstep = GraphedTrainStep(self.model, self.nll, self.optimizer, data, warmup=1, scheduler=self.scheduler,
                                           check_overflow=True)
            return self._gstep.warmup_loss
        if not torch.equal(n_cpu, g._n_cpu):
            return None
        try:
            return g(data)
        except EdgeCapacityOverflow:
            # more edges than the captured capacity (radius graphs): the optimizer has not run.  Launch this step eagerly
            # (its own retry doubles model._edge_caps) and capture again at the next batch with the larger capacity.
            self._gstep = None
            return None

    def train(self):
        if self.world_rank == 0:
            print('Epoch \tTraining Loss \t   Time (s)', flush=True)
        for epoch in range(self.start_epoch, self.start_epoch + self.num_epochs):
            losses = []
            if self.ddp:
                self.sampler.set_epoch(epoch)
            self.model.train()
            torch.cuda.synchronize()
            start_time = time.time()
            for i, data in enumerate(self.train_loader):
                data = data.to(self.local_rank)
                loss = self._graphed_step(data) if self.cuda_graph else None
                if loss is None:
                    self.optimizer.zero_grad()
                    out, ldj = self.model(data)
                    loss = self.nll(out, ldj)
                    loss.backward()
                    self.optimizer.step()
                if self.scheduler:
                    self.scheduler.step()                        # per batch (main.py:223, Q15)
                losses.append(loss.detach().clone())
            epoch_loss = torch.stack(losses).mean()
            if self.ddp:
                dist.all_reduce(epoch_loss, op=dist.ReduceOp.SUM)
                epoch_loss /= self.world_size
            if self.world_rank == 0:
                to_save = {'epoch': epoch, 'model_state_dict': self.model.state_dict(),
                           'optimizer_state_dict': self.optimizer.state_dict(), 'node_nf': self.dataset.node_nf,
                           'hidden_nf': self.hidden_nf, 'softening': self.nll.softening, 'lj_kBT': self.nll.kBT,
                           'integrator': self.integrator, 'n_iter': self.n_iter, 'dt': self.model.dt}
                if self.scheduler:
                    to_save['scheduler_state_dict'] = self.scheduler.state_dict()
                if self.checkpoint_path:
                    torch.save(to_save, self.checkpoint_path)
                torch.cuda.synchronize()
                if epoch % self.log_interval == 0:
                    print('%.5i \t    %.2f \t    %.2f \t    %.2e' % (epoch, epoch_loss.item(), time.time() - start_time,
                                                                    self.optimizer.param_groups[0]['lr']), flush=True)
            if self.ddp:
                dist.barrier()
        return epoch_loss.item()

    def generate(self, out_prefix=''):
        """`main.py:263-278`: inverse pass on the first batch, writers, and the forward round-trip self-check."""
        data = next(iter(self.train_loader)).to(self.local_rank)
        start = data.clone()
        out = self.model.reverse(data)
        np.savetxt(out_prefix + 'h.out', out.h.detach().cpu().numpy(), delimiter=' ')
        write_xyz(out, out_prefix + 'test_out.xyz')
        # Round-trip self-check.  Upstream compares a tensor with itself (reverse and forward mutate and return
        # the same Data object, main.py:269-278); here the flow is actually inverted: un-quantised inverse,
        # then forward without re-dequantising, must give back the latents.
        with torch.no_grad():
            chk = self.model.reverse(start.clone(), quantize=False)
            back, _ = self.model(chk, dequantize=False)
        diff = back.pos - start.pos.to(back.pos.dtype)
        box = start.box.to(diff.dtype)
        ok = bool(((diff - (diff / box).round() * box).abs() < 1e-4).all())
        ok = ok and torch.allclose(back.h, start.h.to(back.h.dtype), atol=1e-4)
        print(ok)
        return out, ok

    def __call__(self, input):
        self.setup(input)
        res = None
        if self.mode == 'train':
            res = self.train()
        elif self.mode == 'gen':
            res = self.generate()
        if self.ddp:
            self._gstep = None                  # captured graphs reference the NCCL communicator: release them first
            import gc
            gc.collect()
            torch.cuda.synchronize()
            dist.destroy_process_group()
        return res
