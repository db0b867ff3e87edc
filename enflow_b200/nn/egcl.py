"""EGCL with the reference's constructor, parameter names and forward signature
(`enflow/nn/egcl.py:5-93`), evaluated by the CUDA kernels of csrc/.

``forward(h, edges) -> (Q [N,1], F [N,3], G [N,nf])`` is the stand-alone INFERENCE entry point: its outputs do not carry
autograd (training runs through ``LFIntegrator``, whose fused C calls hold the hand-written backward of every layer).
Called with autograd enabled on inputs that require a gradient it raises instead of silently dropping that gradient.
"""
import torch
from torch import nn

from .. import _lib

_ORDER = ['edge_nn.0.weight', 'edge_nn.0.bias', 'edge_nn.2.weight', 'edge_nn.2.bias',
          'node_nn.0.weight', 'node_nn.0.bias', 'node_nn.2.weight', 'node_nn.2.bias',
          'coord_nn.0.weight', 'coord_nn.0.bias', 'coord_nn.2.weight',
          'vel_scaling_nn.0.weight', 'vel_scaling_nn.0.bias', 'vel_scaling_nn.2.weight', 'vel_scaling_nn.2.bias']


class EGCL(nn.Module):
    PARAM_ORDER = _ORDER

    def __init__(self, input_nf, output_nf, hidden_nf, act_fn=nn.SiLU(), coords_weight=1.0, attention=False,
                 clamp=False, norm_diff=False, tanh=False):
        super().__init__()
        if attention or norm_diff or tanh:
            # `enflow/main.py:151` never enables them; tanh is half-implemented upstream (egcl.py:40-42)
            raise NotImplementedError('enflow_b200.EGCL implements the options Main uses (attention/norm_diff/tanh off)')
        if not isinstance(act_fn, nn.SiLU):
            raise NotImplementedError('enflow_b200.EGCL kernels are specialised for SiLU')
        if input_nf != output_nf:
            raise ValueError('the flow needs output_nf == input_nf (enflow/main.py:151)')
        self.input_nf, self.hidden_nf = input_nf, hidden_nf
        self.coords_weight = coords_weight
        self.attention, self.norm_diff, self.tanh, self.clamp = attention, norm_diff, tanh, clamp
        self.edge_nn = nn.Sequential(nn.Linear(2 * input_nf + 1, hidden_nf), act_fn, nn.Linear(hidden_nf, hidden_nf), act_fn)
        self.node_nn = nn.Sequential(nn.Linear(hidden_nf + input_nf, hidden_nf), act_fn, nn.Linear(hidden_nf, output_nf))
        last = nn.Linear(hidden_nf, 1, bias=False)
        torch.nn.init.xavier_uniform_(last.weight, gain=0.001)            # egcl.py:32-33
        self.coord_nn = nn.Sequential(nn.Linear(hidden_nf, hidden_nf), act_fn, last)
        self.vel_scaling_nn = nn.Sequential(nn.Linear(input_nf, hidden_nf), act_fn, nn.Linear(hidden_nf, 1))
        self._flat_view = None     # set by BaseFlow when the layer lives inside the flat parameter buffer
        self.precision = 'fp32'    # see LFIntegrator.precision

    def _layer_flat(self, device):
        if self._flat_view is not None and self._flat_view.device == device:
            return self._flat_view
        total, offs, cnts = _lib.param_layout(self.input_nf, 1)
        flat = torch.zeros(total, dtype=torch.float32, device=device)
        sd = dict(self.named_parameters())
        for name, o, c in zip(_ORDER, offs, cnts):
            flat[o:o + c] = sd[name].detach().to(device, torch.float32).reshape(-1)
        return flat

    def forward(self, h, edges):
        self._grad_mode = torch.is_grad_enabled()
        with torch.no_grad():
            return self._forward(h, edges)

    def _forward(self, h, edges):
        L = _lib.lib()
        if self.hidden_nf != L.enflow_hidden():
            raise ValueError(f'kernels are built for hidden_nf={L.enflow_hidden()}')
        _lib.require_cuda(h, edges.coord)
        if self._grad_mode and (h.requires_grad or edges.coord.requires_grad):
            raise RuntimeError('enflow_b200.EGCL.forward does not record autograd: differentiate through LFIntegrator '
                               '(or call it under torch.no_grad() / on detached inputs)')
        if edges.csr is None:
            raise ValueError('EGCL.forward needs Edges built by Data.edges on the GPU (row-grouped CSR attached)')
        dev = h.device
        nf, H = self.input_nf, self.hidden_nf
        row, col, rowptr, e_dev = edges.csr
        E = int(row.numel())
        N = int(h.shape[0])
        hf = _lib.f32c(h)
        pos = _lib.f32c(edges.coord)
        # per-atom box: every edge of an atom carries its molecule's box (base.py:140)
        box = torch.zeros(N, 3, dtype=torch.float32, device=dev)
        if E:
            box[edges.row] = edges.box.to(torch.float32)
        lp = self._layer_flat(dev)
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        packed = new(L.enflow_pack_floats(nf))
        P, S, Q = new(N, H), new(N, H), new(N)
        z2, z3, s, trans = new(max(E, 1), H), new(max(E, 1), H), new(max(E, 1)), new(max(E, 1), 3)
        agg, z4, F, G, wr = new(N, H), new(N, H), new(N, 3), new(N, nf), new(H)
        st = _lib.stream()
        p = _lib.ptr
        _lib.check(L.enflow_pack_layer(p(lp), nf, p(packed), st))
        _lib.check(L.enflow_node_pre_fwd(p(hf), N, nf, p(lp), p(P), p(S), p(Q), st))
        mode = _lib.MODES[self.precision]
        if mode == 0:
            _lib.check(L.enflow_edge_fwd(p(row), p(col), p(e_dev), E, p(pos), p(box), p(P), p(S), p(lp), p(packed), nf,
                                         p(wr), p(z2), p(z3), p(s), p(trans), st))
            _lib.check(L.enflow_segment_sum128(p(z2), p(rowptr), None, N, E, 1, p(agg), st))
        else:
            wimg = torch.empty(L.enflow_tc_pack_bytes(), dtype=torch.uint8, device=dev)
            mis = torch.empty(N + 2, dtype=torch.int32, device=dev)
            scratch = torch.empty(L.enflow_run_scratch_ints(N), dtype=torch.int32, device=dev)
            runs = new(L.enflow_run_rows(E, N), H)
            _lib.check(L.enflow_tc_pack_layer(p(lp), nf, p(wimg), st))
            _lib.check(L.enflow_run_index(p(rowptr), N, p(mis), p(scratch), st))
            _lib.check(L.enflow_edge_fwd_tc(mode, p(row), p(col), p(e_dev), E, p(pos), p(box), p(P), p(S), p(lp),
                                            p(wimg), nf, p(rowptr), p(mis), p(runs), p(s), p(trans), st))
            _lib.check(L.enflow_run_sum128(p(runs), p(rowptr), p(mis), N, E, p(agg), st))
        _lib.check(L.enflow_segment_sum3(p(trans), p(rowptr), None, N, E, 1, float(self.coords_weight), 0, p(F), st))
        _lib.check(L.enflow_node_post_fwd(p(hf), p(agg), N, nf, p(lp), p(packed), p(z4), p(G), st))
        self.last = {'agg': agg, 'trans': trans, 's': s, 'z2': z2}
        return Q.unsqueeze(1), F, G
