"""debug: slab pipeline vs single launch for RK64 (session r02aj)"""
import os, sys
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from oracle.pyoracle import Oracle, reacteval_states
from pelec_b200.reactor import ReactorB200
mech = "LiDryer"; orc = Oracle(mech)
nx, ny, nz = 64, 64, 67
st0 = reacteval_states(orc, (nx, ny, nz))
rng = np.random.default_rng(11)
st0.mask[rng.random(st0.ncells) < 0.05] = -1
st0.rYsrc[:] = 1e-3 * rng.standard_normal(st0.rYsrc.shape) * st0.rY
st0.rYsrc -= st0.rYsrc.sum(axis=0, keepdims=True) / orc.N
st0.rEsrc[:] = 1e2 * rng.standard_normal(st0.ncells)
for key in sys.argv[1:] or ["ReactorRK64"]:
    r = ReactorB200.create(key, mech, **({"ode__atol": 1e-10} if key == "ReactorRK64" else {"ode__rtol": 1e-8, "ode__atol": 1e-10, "cvode__max_order": 4, "cvode__solve_type": "denseAJ_direct"}))
    r.init(1, 1)
    outs = []
    for slabs, streams in ((1, 1), (1, 1), (2, 1), (8, 1), (8, 3), (5, 1)):
        os.environ["PRX_HOST_SLABS"] = str(slabs); os.environ["PRX_HOST_REACT_STREAMS"] = str(streams)
        st = st0.copy()
        r.react_host(nx, ny, nz, st.rY, st.rYsrc, st.T, st.rE, st.rEsrc, st.FC, st.mask, 1e-6, 0.0)
        outs.append(st)
        ref = outs[0]
        d = (st.rY != ref.rY) & ~(np.isnan(st.rY) & np.isnan(ref.rY))
        cells = np.where(d.any(axis=0))[0]
        rel = np.abs(st.rY - ref.rY) / np.maximum(np.abs(ref.rY), 1e-300)
        print(key, "slabs", slabs, "streams", streams, "nan", int(np.isnan(st.rY).sum()), "cells differing", cells.size,
              "max rel", float(np.nanmax(rel)) if cells.size else 0.0, "FC differing", int((st.FC != ref.FC).sum()), "T differing", int((st.T != ref.T).sum()),
              "k-planes of the first few", (cells[:6] // (nx * ny)).tolist(), "species of first", np.where(d[:, cells[0]])[0].tolist() if cells.size else [])
    r.close()
