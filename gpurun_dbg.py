import os, sys, ctypes as C
os.environ["REGT_TC_DEBUG"]="1"
sys.path.insert(0,"/root/repo/regt-gcn_b200"); sys.path.insert(0,"/root/repo")
import numpy as np, torch
from regt_b200 import workloads as W, _lib, engine
from models import TemporalGCN
prec=sys.argv[1] if len(sys.argv)>1 else "bf16"
which=sys.argv[2] if len(sys.argv)>2 else "fwd"
w=W.make_workload(2); m=TemporalGCN(8,w.T,w.O,hidden=w.H,precision=prec); W.init_params_synthetic(m,1); m=m.cuda()
x,y=w.inputs(); x=x.cuda(); y=y.cuda()
g=tuple(a.cuda() for a in w.graph_args())
for i in range(2): m.fused_step(x,y,*g) if prec=="bf16" else None
plan=m._plan(x,*g)
params=m._param_dict()
st=engine.build_state(m._mode,m._prec(),plan,x,m._hidden,m.output_dim,params,y,None,True,fuse_head=True)
engine.run_forward(st,True); torch.cuda.synchronize()
def dump(ncol, names):
    v=st.workspace.view(torch.int64).cpu().numpy()
    idx=np.where(v==0x5245475444424721)[0]
    i0=idx[0]-9
    t=v[i0:i0+24*10].reshape(24,10)
    for row in t[:12]:
        if row[9]!=0x5245475444424721: break
        print(" ".join(f"{names[i]}={row[i+1]-row[i]:6d}" for i in range(ncol-1)), "| step=%6d start=%d"%(row[ncol-1]-row[0], row[0]-t[0,0]))
    e=v[i0+240:i0+243]
    print("entry->first step", t[0,0]-e[0], " first step->epilogue done", e[1]-t[0,0], " wait others", e[2]-e[1])
if which=="fwd":
    dump(7,["P","st","wM1","E1","wM2","E2"])
else:
    grads={k:torch.zeros_like(p) for k,p in params.items()}
    st.workspace.view(torch.int64)[:]=st.workspace.view(torch.int64)  # no-op
    engine.run_backward(st,grads,st.d_out,None,False,True); torch.cuda.synchronize()
    dump(8,["h","E0","xs","wM1+E1a","E1","wM2","E2","dp"])
