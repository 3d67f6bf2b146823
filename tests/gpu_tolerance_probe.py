import sys, os
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, orc
from helpers import make_pair
for name,app,kw in [("upwelling",orc.APP_UPWELLING,{}),("seamount",orc.APP_SEAMOUNT,{}),("benchmark",orc.APP_BENCHMARK,dict(Lm=128,Mm=64,N=30))]:
    o,t=make_pair(app,strict=False,spinup=0,**kw)
    done=0
    for nst in (10,50,200):
        for s in range(nst-done):
            o.step(1); t.set("sustr",o.field("sustr")); t.main3d(1)
        done=nst
        out=[]
        for n in ["zeta1","u1","v1","t1_0"]+(["t1_1"] if int(o.opt("NT"))>1 else []):
            a=o.field(n); b=t.get(n)
            out.append((n, float(np.max(np.abs(a-b))/max(np.max(np.abs(a)),1e-300))))
        print(name,nst,[(n,f"{v:.2e}") for n,v in out], flush=True)
