"""Debug only: clock64 phase timeline of k_edge_bwd_tc (profiles/r1c_phase_times.txt).

    make -C enflow_b200/csrc clean && make -C enflow_b200/csrc PHASE=1 && python tools/phase_times.py
    make -C enflow_b200/csrc clean && make -C enflow_b200/csrc          # back to the product library
"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from enflow_b200 import _lib
from enflow_b200.data import synthetic as syn
from enflow_b200.flow.loss import Alchemical_NLL
from gpu_util import build_model, gpu_batch
arrs = syn.make_batch('c2', 1024)
eps = torch.as_tensor(syn.make_noise(int(arrs['N'].sum()), 5))
model = build_model(syn.make_weights(5, 128, 5, seed=0), 5, 5, precision='fp32_tc')
nll = Alchemical_NLL(kBT=syn.TRAIN_KBT, softening=syn.TRAIN_SOFTENING)
for _ in range(3):
    model.zero_grad(set_to_none=True)
    out, ldj = model(gpu_batch(arrs), eps=eps)
    nll(out, ldj).backward()
torch.cuda.synchronize()
L = ctypes.CDLL(os.path.join(ROOT, 'enflow_b200', 'libenflow_b200.so'))
buf = (ctypes.c_longlong * 64)()
print('rc', L.enflow_debug_phase_times(buf))
names = ['top', 'waitG1', 'E1', 'issueG2', 'waitG2', 'E2', 'issueDG2', 'gather', 'waitDG2', 'E3', 'P1', 'issueDG1', 'loadz1', 'waitDG1', 'P0next', 'E4']
for t in range(4):
    s = [buf[t * 16 + k] for k in range(16)]
    nxt = buf[(t + 1) * 16] if t < 3 else None
    print('tile', t, ' '.join(f'{names[k]}={s[k] - s[k - 1]}' for k in range(1, 16)), 'total', (nxt - s[0]) if nxt else s[15] - s[0])
