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
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        flat = model.flat_params
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=flat.device)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        m = self.model
        if m.flat_params.device != self.exp_avg.device:
            raise RuntimeError('FlatAdam: the model moved to another device after the optimizer was built')
        grp = self.param_groups[0]
        # every parameter gradient is a view of model.flat_grads (the backward pass wrote them there)
        if any(p.grad is None for p in grp['params']):
            raise RuntimeError('FlatAdam.step() needs gradients from the fused backward pass on every parameter')
        _lib.require_cuda(m.flat_params)
        _lib.check(_lib.lib().enflow_adam_step(_lib.ptr(m.flat_params), _lib.ptr(m.flat_grads), _lib.ptr(self.exp_avg),
                                               _lib.ptr(self.exp_avg_sq), m.flat_params.numel(), _lib.ptr(self.step_dev),
                                               float(grp['lr']), float(grp['betas'][0]), float(grp['betas'][1]),
                                               float(grp['eps']), _lib.stream()))
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
