"""Fused Adam over the flow's flat parameter / gradient buffers.

``FlatAdam(model, lr=...)`` is a ``torch.optim.Optimizer`` (schedulers such as the reference's per-batch StepLR,
`enflow/main.py:188,223`, work unchanged) whose ``step()`` is ONE kernel over ``model.flat_params`` instead of a
multi-tensor sweep over 79 tensors.  ``state_dict()`` / ``load_state_dict()`` use torch.optim.Adam's layout
(per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``), so optimizer states interchange with the checkpoints the
reference writes (`main.py:236-250`).
"""
import torch

from . import _lib


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.model = model
        params = model._ordered_params()
        # torch.optim.Adam's full set of group defaults: a state_dict written here loads into torch.optim.Adam (the
        # reference, main.py:177,203) and steps there (Adam.step reads every one of these keys)
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False)
        super().__init__(params, defaults)
        flat = model.flat_params
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=flat.device)
        # the kernel reads the learning rate from this device scalar, so a captured CUDA graph follows a scheduler:
        # sync_lr() (outside the graph) refreshes it whenever param_groups[0]['lr'] changed
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=flat.device)
        self._lr_seen = float(lr)

    def sync_lr(self):
        lr = float(self.param_groups[0]['lr'])
        if lr != self._lr_seen:
            self.lr_dev.fill_(lr)
            self._lr_seen = lr

    def _gather_foreign_grads(self):
        """Gradients normally ARE views of model.flat_grads (the fused backward wrote them there).  A .grad that is
        its own tensor (accumulated by autograd into a clone, clipped into a fresh tensor, set by user code) is copied
        into its slice so the kernel never applies stale values."""
        m = self.model
        offs, cnts = m._layout
        base, esz = m.flat_grads.data_ptr(), m.flat_grads.element_size()
        for p, o, c in zip(self.param_groups[0]['params'], offs, cnts):
            if p.grad is None:
                raise RuntimeError('FlatAdam.step() needs gradients from the fused backward pass on every parameter')
            if p.grad.data_ptr() != base + o * esz or p.grad.dtype != m.flat_grads.dtype:
                m.flat_grads[o:o + c].copy_(p.grad.reshape(-1))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        m = self.model
        if m.flat_params.device != self.exp_avg.device:
            raise RuntimeError('FlatAdam: the model moved to another device after the optimizer was built')
        grp = self.param_groups[0]
        self._gather_foreign_grads()
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        _lib.require_cuda(m.flat_params)
        _lib.check(_lib.lib().enflow_adam_step(_lib.ptr(m.flat_params), _lib.ptr(m.flat_grads), _lib.ptr(self.exp_avg),
                                               _lib.ptr(self.exp_avg_sq), m.flat_params.numel(), _lib.ptr(self.step_dev),
                                               float(grp['lr']), _lib.ptr(self.lr_dev), float(grp['betas'][0]),
                                               float(grp['betas'][1]), float(grp['eps']), _lib.stream()))
        return loss

    # ---- torch.optim.Adam-compatible (de)serialisation ------------------------------------------------------
    def state_dict(self):
        offs, cnts = self.model._layout
        params = self.param_groups[0]['params']
        step = self.step_dev.to(torch.float32).cpu().reshape(())
        state = {i: {'step': step.clone(), 'exp_avg': self.exp_avg[o:o + c].view(p.shape).clone(),
                     'exp_avg_sq': self.exp_avg_sq[o:o + c].view(p.shape).clone()}
                 for i, (p, o, c) in enumerate(zip(params, offs, cnts))}
        group = {k: v for k, v in self.param_groups[0].items() if k != 'params'}
        group['params'] = list(range(len(params)))
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd):
        offs, cnts = self.model._layout
        for i, (o, c) in enumerate(zip(offs, cnts)):
            st = sd['state'].get(i)
            if st is None:
                continue
            self.exp_avg[o:o + c] = st['exp_avg'].to(self.exp_avg.device, torch.float32).reshape(-1)
            self.exp_avg_sq[o:o + c] = st['exp_avg_sq'].to(self.exp_avg.device, torch.float32).reshape(-1)
            self.step_dev.fill_(int(float(st['step'])))
        for k, v in sd['param_groups'][0].items():
            if k != 'params':
                self.param_groups[0][k] = v
        self.sync_lr()
