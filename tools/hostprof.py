import sys, time, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enflow_b200.data import synthetic as syn
from enflow_b200.data.base import Data
from enflow_b200.flow.dynamics import LFIntegrator
from enflow_b200.flow.loss import Alchemical_NLL
from enflow_b200.nn.argmax import ArgMax
from enflow_b200.nn.egcl import EGCL
dev=torch.device('cuda',0)
nf,H,L=5,128,5
sd=syn.make_weights(nf,H,L)
model=LFIntegrator([EGCL(nf,nf,H) for _ in range(L)],ArgMax(nf,H),dt=syn.TRAIN_DT)
model.load_state_dict({k:torch.tensor(v) for k,v in sd.items()}); model=model.to(dev)
nll=Alchemical_NLL(kBT=syn.TRAIN_KBT,softening=0.1); opt=torch.optim.Adam(model.parameters(),lr=1e-3)
cfg=sys.argv[1] if len(sys.argv)>1 else 'c2'
arrs=syn.make_batch('c2',1024,n_atoms=29) if cfg=='c2' else syn.make_batch('c5',32,n_atoms=500)
f32=lambda k: torch.tensor(arrs[k],dtype=torch.float32)
host=Data(h=f32('h'),g=f32('g'),pos=f32('pos'),vel=f32('vel'),N=torch.tensor(arrs['N']),r_cut=torch.tensor(arrs['r_cut']),box=f32('box')).pin_memory()
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(6):
    t0=T(); d=host.to(dev); t1=T()
    opt.zero_grad(set_to_none=True); t2=T()
    model.check_status = (it<3)
    out,ldj=model(d); t3h=time.perf_counter(); t3=T()
    loss=nll(out,ldj); t4=T()
    loss.backward(); t5h=time.perf_counter(); t5=T()
    opt.step(); t6h=time.perf_counter(); t6=T()
    l=loss.item(); t7=T()
    print(f"it{it}: h2d {1e3*(t1-t0):.2f} zero {1e3*(t2-t1):.2f} fwd host {1e3*(t3h-t2):.2f} total {1e3*(t3-t2):.2f} | nll {1e3*(t4-t3):.2f} | bwd host {1e3*(t5h-t4):.2f} total {1e3*(t5-t4):.2f} | opt host {1e3*(t6h-t5):.2f} total {1e3*(t6-t5):.2f} | item {1e3*(t7-t6):.2f}")
# no-sync loop timing
for name,chk in (('check',True),('nocheck',False)):
    model.check_status=chk
    t0=T()
    for it in range(10):
        d=host.to(dev); opt.zero_grad(set_to_none=True); out,ldj=model(d); loss=nll(out,ldj); loss.backward(); opt.step(); l=loss.item()
    t1=T(); print(name, 'e2e ms/step', 1e3*(t1-t0)/10)
import cProfile,pstats
model.check_status=False
pr=cProfile.Profile(); pr.enable()
for it in range(5):
    d=host.to(dev); opt.zero_grad(set_to_none=True); out,ldj=model(d); loss=nll(out,ldj); loss.backward(); opt.step(); l=loss.item()
pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
