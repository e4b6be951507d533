import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from util import *
from pelec_b200.reactor import ReactorB200
from oracle.pyoracle import Oracle
for mech in ["LiDryer","drm19"]:
    g=golden(mech); orc=Oracle(mech)
    r=ReactorB200.create("ReactorRK64",mech,ode__atol=1e-10); r.init(1,1)
    for variant in ["golden","nosrc","only_esrc","only_ysrc"]:
        s=golden_state(g)
        if variant=="nosrc": s.rYsrc[:]=0; s.rEsrc[:]=0
        if variant=="only_esrc": s.rYsrc[:]=0
        if variant=="only_ysrc": s.rEsrc[:]=0
        b=s.copy(); orc.react_rk64(b,1e-6,atol=1e-10)
        r.react_host(s.ncells,1,1,s.rY,s.rYsrc,s.T,s.rE,s.rEsrc,s.FC,s.mask,1e-6,0.0)
        rho=b.rY.sum(axis=0,keepdims=True)
        e=np.abs(s.rY-b.rY)/np.maximum(np.abs(b.rY),1e-12*rho)
        k,c=np.unravel_index(np.nanargmax(e),e.shape)
        print(mech,variant,"max rel",e.max(),"species",orc.species[k],"cell",c,"val",b.rY[k,c],"rho",rho[0,c],"abs",abs(s.rY[k,c]-b.rY[k,c]),"FC eq",np.mean(s.FC==b.FC),"dFCmax",np.abs(s.FC-b.FC).max(),"relT",(np.abs(s.T-b.T)/b.T).max())
        print("   per-species max rel:", " ".join("%s:%.1e"%(orc.species[i],e[i].max()) for i in range(orc.N)))
    r.close()
