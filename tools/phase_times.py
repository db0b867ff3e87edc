"""Debug only: clock64 timeline of k_edge_bwd_tc (one epilogue thread and the MMA-issuing lane of one CTA, four periods
= eight tiles in the middle of the launch).

    make -C enflow_b200/csrc clean && make -C enflow_b200/csrc PHASE=1 -j8 && python tools/phase_times.py [precision]
    make -C enflow_b200/csrc clean && make -C enflow_b200/csrc -j8          # back to the product library
"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from enflow_b200.data import synthetic as syn
from enflow_b200.flow.loss import Alchemical_NLL
from gpu_util import build_model, gpu_batch
prec = sys.argv[1] if len(sys.argv) > 1 else 'fp32_tc'
arrs = syn.make_batch('c2', 1024)
eps = torch.as_tensor(syn.make_noise(int(arrs['N'].sum()), 5))
model = build_model(syn.make_weights(5, 128, 5, seed=0), 5, 5, precision=prec)
nll = Alchemical_NLL(kBT=syn.TRAIN_KBT, softening=syn.TRAIN_SOFTENING)
for _ in range(3):
    model.zero_grad(set_to_none=True)
    out, ldj = model(gpu_batch(arrs), eps=eps)
    nll(out, ldj).backward()
torch.cuda.synchronize()
L = ctypes.CDLL(os.path.join(ROOT, 'enflow_b200', 'libenflow_b200.so'))
buf = (ctypes.c_longlong * 256)()
print('rc', L.enflow_debug_phase_times(buf), 'precision', prec)
tn = ['E2(a)', 'E4(b-)+X(b)', 'E3(a)', 'E1(b)', 'E4(a)+X(a+)', 'E2(b)', 'E1(a+)', 'E3(b)']
gn = ['G3(a)', 'G1(b)', 'G4(a)', 'G2(b)', 'G1(a+)', 'G3(b)', 'G2(a+)', 'G4(b)']
t0 = buf[0]
for k in range(4):
    e = [buf[k * 32 + i] - t0 for i in range(25)]
    m = [buf[128 + k * 32 + i] - t0 for i in range(16)]
    print(f'period {k}: starts at {e[0]}, ends at {e[24]} (length {e[24] - e[0]})')
    for i in range(8):
        print(f'   {tn[i]:13s} start {e[3*i]:7d}  wait {e[3*i+1]-e[3*i]:5d}  work {e[3*i+2]-e[3*i+1]:5d}   |  {gn[i]:7s} operands ready {m[2*i]:7d}  issue took {m[2*i+1]-m[2*i]:5d}')
