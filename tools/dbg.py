"""TEST TOOLING (not product): one small ReactorCvode call through the C ABI with progress prints, checked against the
oracle - used by the tools/gpu_session*.sh scripts to localise crashes and to spot-check parity of A/B builds
(PRX_LIB_PATH, PRX_BDF_VARIANT).  usage: python tools/dbg.py [mechanism]"""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
def P(*a):
    print(*a, flush=True)
P("start", os.environ.get("PRX_BDF_VARIANT"))
import torch
P("torch", torch.cuda.is_available())
from oracle.pyoracle import Oracle, reacteval_states
from pelec_b200.reactor import ReactorB200
mech = sys.argv[1] if len(sys.argv) > 1 else "LiDryer"
orc = Oracle(mech)
st = reacteval_states(orc, (4, 16, 2))
a, b = st.copy(), st.copy()
r = ReactorB200.create("ReactorCvode", mech, ode__rtol=1e-10, ode__atol=1e-10, cvode__max_order=4)
P("created")
r.init(1, 1)
P("init done")
r.react_host(4, 16, 2, a.rY, a.rYsrc, a.T, a.rE, a.rEsrc, a.FC, a.mask, 1e-6, 0.0)
P("react done", r.stats())
orc.react_cvode(b, 1e-6, rtol=1e-10, atol=1e-10, max_order=4)
P("dY", np.abs(a.rY - b.rY).max(), "dT", (np.abs(a.T - b.T) / b.T).max(), "FC", a.FC.sum(), b.FC.sum())
