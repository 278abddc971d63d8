import os, sys, ctypes as C
os.environ["REGT_TC_DEBUG"]="1"
sys.path.insert(0,"/root/repo/regt-gcn_b200"); sys.path.insert(0,"/root/repo")
import torch
from regt_b200 import workloads as W, _lib, engine
from models import TemporalGCN
w=W.make_workload(2); m=TemporalGCN(8,w.T,w.O,hidden=w.H,precision=sys.argv[1] if len(sys.argv)>1 else "bf16"); W.init_params_synthetic(m,1); m=m.cuda()
x,y=w.inputs(); x=x.cuda()
g=tuple(a.cuda() for a in w.graph_args())
with torch.no_grad():
    for i in range(3): m(x,*g)
    plan=m._plan(x,*g)
    st=engine.build_state(m._mode,m._prec(),plan,x,m._hidden,m.output_dim,m._param_dict(),None,None,True)
    engine.run_forward(st,True); torch.cuda.synchronize()
    # find tc_wpart offset: read first 24*8 int64 from the workspace region: easier: search for plausible clock values
    ws=st.workspace.view(torch.int64)
    # debugging: the layout is internal; scan for a run of increasing large values
    v=ws.cpu().numpy()
    import numpy as np
    idx=np.where((v[:-8]>1e9)&(v[1:-7]>v[:-8])&(v[1:-7]-v[:-8]<1e7)&(v[2:-6]>v[1:-7])&(v[2:-6]-v[1:-7]<1e7)&(v[3:-5]>v[2:-6]) & (v[3:-5]-v[2:-6]<1e7))[0]
    print(len(idx), idx[:5])
    i0=idx[0]
    t=v[i0:i0+24*8].reshape(24,8)[:, :7]
    base=t[0,0]
    for row in t[:12]:
        d=row-base
        print("P=%6d st=%5d wM1=%6d E1=%6d wM2=%6d E2=%6d | step=%6d" % (row[1]-row[0], row[2]-row[1], row[3]-row[2], row[4]-row[3], row[5]-row[4], row[6]-row[5], row[6]-row[0]), "start", d[0])
