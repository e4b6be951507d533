"""TEST INFRASTRUCTURE: numpy prototype of the ARKODE ERKStep control logic restated in pelec_b200/csrc/prx_arkode.cuh + prx_api.cu
(react_arkode), with the RHS of the pinned C oracle - used on the CPU to debug the port against the reference fixture
(tests/golden/arkode_<mech>.npz): python tools/arkode_prototype.py LiDryer default  ->  per call nfe (ref nfe), steps, error-test
failures, deviation in units of the tolerance.  It reproduces the reference RHS-evaluation counts of all twelve fixture calls."""
import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle.pyoracle import Oracle, Reference, reacteval_states, BoxState

def table(method):
    """the Butcher table of ode.rk_method from pelec_b200/csrc/prx_ark_tables.inc (the product's generated tables)"""
    import re
    txt=open(os.path.join(os.path.dirname(os.path.abspath(__file__)),'..','pelec_b200','csrc','prx_ark_tables.inc')).read()
    blk=re.search(r"case %d:.*?return true;"%method, txt, re.S).group(0)
    s,q,p=[int(v) for v in re.search(r"T->stages = (\d+); T->q = (\d+); T->p = (\d+);",blk).groups()]
    A=np.zeros((s,s)); b=np.zeros(s); d=np.zeros(s); c=np.zeros(s)
    for m in re.finditer(r"T->(A\[(\d+)\]\[(\d+)\]|([bdc])\[(\d+)\]) = ([^;]*);", blk):
        v=eval(m.group(6))
        if m.group(2) is not None: A[int(m.group(2)),int(m.group(3))]=v
        else: {"b":b,"d":d,"c":c}[m.group(4)][int(m.group(5))]=v
    return s,A,b,d,c,q,p

class Ctl:
    def __init__(s, kind):
        s.k1,s.k2,s.k3=(0.58,-0.21,0.1) if kind==0 else ((0.8,-0.31,0.0) if kind==1 else (1.0,0,0))
        s.bias=1.5; s.ep=s.epp=s.hp=s.hpp=1.0; s.first=0
    def est(s,h,p,dsm):
        # SUNMAX(A,B) = A > B ? A : B: a NaN error norm (a stage left the physical range) gives TINY, as in the reference
        smax=lambda a,b: a if a>b else b
        o=p+1; e1=smax(s.bias*dsm,1e-10); e2=smax(s.ep,1e-10); e3=smax(s.epp,1e-10)
        if s.first<1: return h*e1**(-s.k1/o)
        if s.first<2: return h*e1**(-s.k1/o)*e2**(-s.k2/o)
        return h*e1**(-s.k1/o)*e2**(-s.k2/o)*e3**(-s.k3/o)
    def upd(s,h,dsm):
        s.epp=s.ep; s.ep=s.bias*dsm; s.hpp=s.hp; s.hp=h
        if s.first<2: s.first+=1

def arkode_box(orc, st, dt, ctl, rtol, atol, method, typ=None, log=None):
    N=orc.N; nc=st.ncells; neq=N+1
    s,A,b,d,c,q,p=table(method)
    rho=st.rY.sum(axis=0); Y=(st.rY/rho).T.copy()
    T0=orc.ey2t(st.rE/rho, Y, st.T)
    yn=np.concatenate([st.rY.T, T0[:,None]],axis=1).copy()   # [nc, neq]
    atolv=np.full(neq, atol) if typ is None else typ*atol
    nfe=[0]
    def f(t,y):
        nfe[0]+=1
        return orc.fkernelspec(t, np.ascontiguousarray(y), st.rE, st.rEsrc, np.ascontiguousarray(st.rYsrc.T))
    wr=lambda v,w: np.sqrt(np.sum((v*w)**2)/v.size)
    ewt=1.0/(rtol*np.abs(yn)+atolv[None,:])
    ur=2.220446049250313e-16
    t0=0.0; tout=dt; tn=tcur=t0
    tdist=abs(tout-tcur); tround=ur*max(abs(tcur),abs(tout))
    F=[None]*s
    F[0]=f(tn,yn)
    hlb=100*tround
    hub_inv=np.max(np.abs(F[0])/(0.1*np.abs(yn)+1.0/ewt)); hub=0.1*tdist
    if hub*hub_inv>1: hub=1/hub_inv
    hg=np.sqrt(hlb*hub)
    if hub<hlb: h=hg
    else:
        for c1 in range(1,5):
            ycur=hg*F[0]+yn; f1=f(tcur+hg,ycur); ydd=wr((1/hg)*f1+(-1/hg)*F[0], ewt)
            hnew=np.sqrt(2/ydd) if ydd*hub*hub>2 else np.sqrt(hg*hub)
            if c1==4: break
            hrat=hnew/hg
            if 0.5<hrat<2: break
            if c1>1 and hrat>2: hnew=hg; break
            hg=hnew
        h=min(max(0.5*hnew,hlb),hub)
    etamax=10000.; eta=1.; hprime=h; init=True; fcur=True; nst=0; nloc=0; netf=0
    stiff=np.allclose(b[:s],A[s-1,:s],atol=1e-14)
    while True:
        if not init: ewt=1.0/(rtol*np.abs(yn)+atolv[None,:])
        if nloc>=500: break
        if ur*wr(yn,ewt)>1: break
        if hprime!=h: h=h*eta
        nef=0; ok=False
        while True:
            if not fcur:
                F[0]=F[s-1].copy() if stiff else f(tn,yn); fcur=True
            for i in range(1,s):
                yc=None
                for j in range(i):
                    if A[i,j]!=0:
                        yc=(h*A[i,j])*F[j] if yc is None else yc+(h*A[i,j])*F[j]
                yc=yc+yn
                F[i]=f(tn+c[i]*h,yc)
            ynew=None; ye=None
            for j in range(s):
                ynew=(h*b[j])*F[j] if ynew is None else ynew+(h*b[j])*F[j]
                ye=(h*(b[j]-d[j]))*F[j] if ye is None else ye+(h*(b[j]-d[j]))*F[j]
            ynew=ynew+yn
            dsm=wr(ye,ewt)
            hacc=ctl.est(h,p-1,dsm); hacc*=0.99
            hacc=min(abs(hacc),abs(etamax*h)); hacc=max(hacc,0.1*h)
            if dsm<=1 and (hacc>h*0.999999 and hacc<h*1.5*1.000001): hacc=h
            eta=hacc/h; eta=min(eta,etamax)
            if log is not None: log.append((tn,h,dsm,eta))
            if dsm<=1: ok=True; break
            nef+=1; netf+=1
            if nef==7: break
            etamax=1.0
            if nef>=2: eta=min(eta,0.3)
            eta=min(eta,etamax)
            h*=eta; hprime=h
        if not ok: break
        tcur=tn+h; yold=yn; fold=F[0]; hi=h
        yn=ynew; fcur=False; ctl.upd(h,dsm); nst+=1; tn=tcur; hprime=h*eta; etamax=25.; init=False; nloc+=1
        if (tcur-tout)*h>=0:
            tau=(tout-tcur)/h
            fn=F[s-1].copy() if stiff else f(tn,yn)
            t2=tau*tau; t3=tau*t2
            if q-1>=5:
                # quintic (arkode_interp.c case 5): fa = f at the CUBIC's -1/3 point (the nested quartic evaluations leave it there),
                # fb = f at the quartic's -2/3 point; two more evaluations are made and discarded by the reference (counted)
                ta=-1.0/3.0; a2=ta*ta; a3=ta*a2
                ya=(3*a2+2*a3)*yold+(1-3*a2-2*a3)*yn+hi*(a2+a3)*fold+hi*(ta+2*a2+a3)*fn
                fa=f(tn-hi/3.0,ya); nfe[0]+=2
                tb=-2.0/3.0; b2=tb*tb; b3=tb*b2; b4=tb*b3
                yb=(-6*b2-16*b3-9*b4)*yold+(1+6*b2+16*b3+9*b4)*yn+hi*0.25*(-5*b2-14*b3-9*b4)*fold+hi*(tb+2*b2+b3)*fn+hi*27*0.25*(-b4-2*b3-b2)*fa
                fb=f(tn-hi*2.0/3.0,yb); t4=tau*t3; t5=tau*t4
                a0=54*t5+135*t4+110*t3+30*t2
                yout=a0*yold+(1-a0)*yn+hi/4*(27*t5+63*t4+49*t3+13*t2)*fold+hi/4*(27*t5+72*t4+67*t3+26*t2+4*tau)*fn+hi/4*(81*t5+189*t4+135*t3+27*t2)*fa+hi/4*(81*t5+216*t4+189*t3+54*t2)*fb
            elif q-1>=4:
                ta=-1.0/3.0; a2=ta*ta; a3=ta*a2
                ya=(3*a2+2*a3)*yold+(1-3*a2-2*a3)*yn+hi*(a2+a3)*fold+hi*(ta+2*a2+a3)*fn
                fa=f(tn-hi/3.0,ya); t4=tau*t3
                yout=(-6*t2-16*t3-9*t4)*yold+(1+6*t2+16*t3+9*t4)*yn+hi*0.25*(-5*t2-14*t3-9*t4)*fold+hi*(tau+2*t2+t3)*fn+hi*27*0.25*(-t4-2*t3-t2)*fa
            elif q-1>=3: yout=(3*t2+2*t3)*yold+(1-3*t2-2*t3)*yn+hi*(t2+t3)*fold+hi*(tau+2*t2+t3)*fn
            elif q-1==2: yout=t2*yold+(1-t2)*yn+hi*(t2+tau)*fn
            else: yout=-tau*yold+(1+tau)*yn
            yn=yout; break
    return yn, nfe[0], nst, netf

def tables_fixture():
    z=np.load('/root/repo/tests/golden/arkode_tables_LiDryer.npz')
    so=os.dup(1); os.dup2(os.open(os.devnull,os.O_WRONLY),1)
    orc=Oracle("LiDryer"); N=orc.N
    for m in z["methods"]:
        ctl=Ctl(0)
        for i in range(3):
            pre="in_%d_"%i; po="m%d_%d_"%(m,i)
            st=BoxState(z[pre+"rY0"].copy(),z[pre+"T0"].copy(),z[pre+"rE0"].copy(),z[pre+"rYsrc"].copy(),z[pre+"rEsrc"].copy())
            y,nfe,nst,netf=arkode_box(orc,st,float(z["dt"]),ctl,1e-6,1e-10,int(m))
            lim=1e-10+1e-6*np.abs(z[po+"rY"])
            dy=np.max(np.abs(y[:,:N].T-z[po+"rY"])/lim)
            sys.stderr.write("rk_method %d call %d: nfe %d (ref %d) nst %d netf %d dy/tol %.3e\n"%(m,i,nfe,int(z["m%d_nfe"%m][i]),nst,netf,dy))

if __name__=="__main__":
    if sys.argv[1]=="tables":
        tables_fixture(); sys.exit(0)
    mech=sys.argv[1]; tag=sys.argv[2]
    z=np.load('/root/repo/tests/golden/arkode_%s.npz'%mech)
    os.dup2(os.open(os.devnull,os.O_WRONLY),1)
    orc=Oracle(mech)
    kw={"default":dict(rtol=1e-6,atol=1e-10,method=40,kind=0,typ=False),"bs_pi_typ":dict(rtol=1e-8,atol=1e-10,method=30,kind=1,typ=True)}[tag]
    ctl=Ctl(kw["kind"]); N=orc.N
    typ=np.concatenate([np.full(N,1e-4),[2000.0]]) if kw["typ"] else None
    for i in range(3):
        pre="%s_%d_"%(tag,i)
        st=BoxState(z[pre+"rY0"].copy(),z[pre+"T0"].copy(),z[pre+"rE0"].copy(),z[pre+"rYsrc"].copy(),z[pre+"rEsrc"].copy())
        log=[]
        y,nfe,nst,netf=arkode_box(orc,st,float(z["dt"]),ctl,kw["rtol"],kw["atol"],kw["method"],typ,log)
        lim=kw["atol"]+kw["rtol"]*np.abs(z[pre+"rY"])
        dy=np.max(np.abs(y[:,:N].T-z[pre+"rY"])/lim)
        sys.stderr.write("%s %s call %d: nfe %d (ref %d) nst %d netf %d dy/tol %.3e dT %.2e h0 %.3e\n"%(mech,tag,i,nfe,int(z[tag+"_nfe"][i]),nst,netf,dy,np.max(np.abs(y[:,N]-z[pre+"T"])/z[pre+"T"]),log[0][1]))
