"""CUDA-graph capture of one whole training step.

A step of the flow is ~250 kernel launches from one forward C call, one backward C call, the likelihood and the
optimizer.  None of them synchronises with the host (edge counts stay on the device), so the step can be captured
once and replayed: launch overhead disappears, which matters for small batches (`example/train.yaml` shape: 1.9 ms
of kernels per step against ~3.8 ms of launch/Python time when issued eagerly).

    step = GraphedTrainStep(model, nll, optimizer, example_batch)      # batch already on the device
    loss = step(batch)                                                 # same B / atoms per molecule as the example

Constraints: every batch must have the example's layout (same number of molecules and atoms per molecule; the
dims of the C calls are baked into the graph); the optimizer must support capture (``torch.optim.Adam(...,
capturable=True)`` or ``FlatAdam``); the edge capacity is fixed at capture time: with ``check_overflow=True`` the step
is two graphs (gradients; optimizer) and the device flag is read between them, so an overflowing batch raises
``EdgeCapacityOverflow`` BEFORE the optimizer runs (``Main`` recaptures with twice the capacity and redoes the step);
otherwise ``step.overflowed()`` reports it on demand.  With ``FlatAdam`` the learning rate lives on the device, so a
scheduler stepped between replays (the reference's per-batch StepLR) is followed.
"""
import torch

from .data.base import Data


class EdgeCapacityOverflow(RuntimeError):
    """The replayed step met more edges than the capacity captured in the graph (gradients are from a truncated list;
    the optimizer has not run).  Recapture with a larger capacity (``model._edge_caps``) and redo the step."""


class GraphedTrainStep:
    def __init__(self, model, nll, optimizer, example, warmup=3, eps=None, scheduler=None, check_overflow=False):
        if not example.pos.is_cuda:
            raise RuntimeError('GraphedTrainStep needs the example batch on the CUDA device')
        self.model, self.nll, self.optimizer = model, nll, optimizer
        # A scheduler changes param_groups[0]['lr'] between replays: only an optimizer that keeps the learning rate on
        # the device (FlatAdam.sync_lr) can follow it inside a captured graph
        if scheduler is not None and not hasattr(optimizer, 'sync_lr'):
            raise ValueError('GraphedTrainStep with a scheduler needs enflow_b200.optim.FlatAdam (device-resident lr)')
        self.check_overflow = bool(check_overflow)
        # optional static ArgMax-noise buffer (refill it before a replay); default: torch.randn inside the graph
        self.eps = None if eps is None else eps.detach().to(example.pos.device, torch.float32).contiguous().clone()
        f32 = lambda t: t.detach().to(torch.float32).contiguous().clone()
        self.static = Data(z=example.z, h=f32(example.h), g=f32(example.g), pos=f32(example.pos), vel=f32(example.vel),
                           N=example.N.clone(), r_cut=example.r_cut.detach().to(example.pos.device, torch.float32).clone(),
                           box=f32(example.box), label=example.label, device=example.device)
        B, off, max_n, n_cpu = example.meta()
        self.static._meta = (B, off.clone(), max_n, n_cpu.clone())
        self._n_cpu = n_cpu.clone()
        self.status = None
        self._check = model.check_status
        # eager warm-up on a side stream (allocator, lazy kernel attributes, edge capacity), as torch requires
        model.check_status = True
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.warmup_loss = self._step_body().detach().clone()       # real optimizer steps on the example batch
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        model.check_status = False          # no host read-back inside the graph
        try:      # the warm-up ran on a side stream: the AccumulateGrad stream-mismatch warning is expected and harmless
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:
            pass
        # Data parallel: the gradient all-reduce (issued from the flow's backward node, NCCL AVG over the flat buffer) is
        # captured with everything else: ONE graph per step, as on a single GPU.  Should the capture of the collective
        # fail (an NCCL build without graph support), fall back to two graphs (everything up to the gradients; the
        # optimizer) with one eager all-reduce between their replays.
        self.dp = getattr(model, '_dp_group', None) is not None
        self.split = self.check_overflow                 # gradients and optimizer as two graphs
        self.dp_eager = False
        if self.dp and not self.split:
            import torch.distributed as dist
            ok, g1, loss1 = True, torch.cuda.CUDAGraph(), None
            import os, sys
            dbg = (lambda m: print(f'[graph dbg rank {dist.get_rank()}] {m}', file=sys.stderr, flush=True)) \
                if os.environ.get('ENFLOW_GRAPH_DEBUG') else (lambda m: None)
            dbg('capturing single graph')
            # the captured collective gets a communicator of its own: eager collectives on the training group (logging
            # all-reduce, barriers, an eagerly launched step of another layout) then never interleave with graph replays
            # on one NCCL communicator
            train_group = model._dp_group
            if getattr(model, '_dp_graph_group', None) is None:
                model._dp_graph_group = dist.new_group(ranks=list(range(dist.get_world_size(train_group))))
                warm = torch.zeros(1, device=example.pos.device)
                dist.all_reduce(warm, group=model._dp_graph_group)      # create the communicator before capturing on it
                torch.cuda.synchronize()
            model._dp_group = model._dp_graph_group
            try:
                # thread-local capture mode: the NCCL watchdog thread queries events while this thread captures
                with torch.cuda.graph(g1, capture_error_mode='thread_local'):
                    loss1 = self._step_body()
            except Exception as exc:                     # noqa: BLE001 - keep training: two-graph form below
                ok = False
                dbg(f'capture failed: {type(exc).__name__}: {exc}')
                torch.cuda.synchronize()
            model._dp_group = train_group
            dbg(f'capture done ok={ok}')
            # every rank must replay the same form: one that failed to capture the collective takes all of them along
            agree = torch.tensor([1.0 if ok else 0.0], device=example.pos.device)
            dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=model._dp_group)
            dbg(f'agree={float(agree.item())}')
            if float(agree.item()) >= 1.0:
                self.graph, self.loss = g1, loss1
                self.status = model.last_status
                model.check_status = self._check
                return
            del g1, loss1
            self.split = self.dp_eager = True
        elif self.dp:
            self.dp_eager = True
        self.graph = torch.cuda.CUDAGraph()
        if self.split:
            model._dp_defer = True
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss = self._grad_body()
            with torch.cuda.graph(self.graph_opt, pool=self.graph.pool()):
                self.optimizer.step()
            model._dp_defer = False
        else:
            with torch.cuda.graph(self.graph):
                self.loss = self._step_body()
        self.status = model.last_status
        model.check_status = self._check

    def _view(self):
        d = self.static
        v = Data(z=d.z, h=d.h, g=d.g, pos=d.pos, vel=d.vel, N=d.N, r_cut=d.r_cut, box=d.box, label=d.label, device=d.device)
        v._meta = d._meta
        return v

    def _grad_body(self):
        self.optimizer.zero_grad(set_to_none=True)
        out, ldj = self.model(self._view(), eps=self.eps)
        loss = self.nll(out, ldj)
        loss.backward()
        return loss

    def _step_body(self):
        loss = self._grad_body()
        self.optimizer.step()
        return loss

    def load(self, batch):
        """Copy a batch (host-pinned or device) into the graph's static input buffers (async on the current stream)."""
        n_cpu = batch.N.detach().to('cpu', torch.int64).reshape(-1) if not batch.N.is_cuda else None
        if n_cpu is not None and not torch.equal(n_cpu, self._n_cpu):
            raise ValueError('GraphedTrainStep: batch layout (atoms per molecule) differs from the captured example')
        s = self.static
        for name in ('h', 'g', 'pos', 'vel', 'box'):
            getattr(s, name).copy_(getattr(batch, name), non_blocking=True)
        s.r_cut.copy_(batch.r_cut.reshape(-1), non_blocking=True)

    def __call__(self, batch=None):
        if batch is not None:
            self.load(batch)
        if hasattr(self.optimizer, 'sync_lr'):
            self.optimizer.sync_lr()                    # outside the graph: the captured kernel reads the device scalar
        self.graph.replay()
        if self.split:
            if self.check_overflow and self.overflowed():
                raise EdgeCapacityOverflow('edge capacity captured in the CUDA graph exceeded by this batch')
            if self.dp_eager:
                from .parallel import allreduce_mean_
                allreduce_mean_(self.model.flat_grads, self.model._dp_group)
            self.graph_opt.replay()
        return self.loss

    def overflowed(self):
        """True if the captured edge capacity was exceeded in the last replay, or the fully connected regime the capture
        assumed no longer holds (status bits 1 and 8; synchronises)."""
        return bool(int(self.status.item()) & 9)


class GraphedReverse:
    """CUDA-graph replay of ``LFIntegrator.reverse`` (the generate.yaml inverse pass) for batches of one layout.

        inv = GraphedReverse(model, example_batch)       # batch on the device
        out = inv(batch)                                 # out.h/g/pos/vel, out.neg_ldj_mol: static output buffers

    Same constraints as GraphedTrainStep: fixed layout, edge capacity fixed at capture (``overflowed()`` on demand)."""

    def __init__(self, model, example, quantize=True, warmup=2):
        if not example.pos.is_cuda:
            raise RuntimeError('GraphedReverse needs the example batch on the CUDA device')
        self.model, self.quantize = model, quantize
        f32 = lambda t: t.detach().to(torch.float32).contiguous().clone()
        self.static = Data(z=example.z, h=f32(example.h), g=f32(example.g), pos=f32(example.pos), vel=f32(example.vel),
                           N=example.N.clone(), r_cut=example.r_cut.detach().to(example.pos.device, torch.float32).clone(),
                           box=f32(example.box), label=example.label, device=example.device)
        B, off, max_n, n_cpu = example.meta()
        self.static._meta = (B, off.clone(), max_n, n_cpu.clone())
        self._n_cpu = n_cpu.clone()
        check = model.check_status
        model.check_status = True
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(warmup):
                model.reverse(self._view(), quantize=quantize)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        model.check_status = False
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = model.reverse(self._view(), quantize=quantize)
        self.status = model.last_status
        model.check_status = check

    def _view(self):
        d = self.static
        v = Data(z=d.z, h=d.h, g=d.g, pos=d.pos, vel=d.vel, N=d.N, r_cut=d.r_cut, box=d.box, label=d.label, device=d.device)
        v._meta = d._meta
        return v

    def __call__(self, batch=None):
        if batch is not None:
            s = self.static
            for name in ('h', 'g', 'pos', 'vel', 'box'):
                getattr(s, name).copy_(getattr(batch, name), non_blocking=True)
            s.r_cut.copy_(batch.r_cut.reshape(-1), non_blocking=True)
        self.graph.replay()
        return self.out

    def overflowed(self):
        return bool(int(self.status.item()) & 9)
