#!/usr/bin/env python
"""Generate tests/golden/arkode_tables_LiDryer.npz from the UNMODIFIED reference's ReactorArkode (oracle/_ref/libpeleref_LiDryer.so;
build container only): every other explicit Butcher table ode.rk_method selects (ReactorArkode.cpp:25-90 -
31, 41, 42, 43: order 3 / 4, cubic Hermite dense output; 50 - 54: order 5, quartic dense output with its extra right-hand side;
60, 80: order 6 / 8, quintic dense output with four more),
ERKStep defaults (PID controller, rtol 1e-6, atol 1e-10), three successive react calls of ONE reactor object on the boxes of
make_golden5.py (the controller's history carries over).  Inputs are stored once."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden5 import DT, boxes  # noqa: E402

METHODS = (31, 41, 42, 43, 50, 51, 52, 53, 54, 60, 80)


def main():
    so = os.dup(1)
    os.dup2(os.open(os.devnull, os.O_WRONLY), 1)
    from oracle.pyoracle import Reference
    mech = "LiDryer"
    ref = Reference(mech)
    out, msgs = {}, []
    ins = boxes(ref, mech)
    for i, s0 in enumerate(ins):
        pre = "in_%d_" % i
        out.update({pre + "rY0": s0.rY, pre + "T0": s0.T, pre + "rE0": s0.rE, pre + "rYsrc": s0.rYsrc, pre + "rEsrc": s0.rEsrc})
    for m in METHODS:
        sts = [s.copy() for s in ins]
        nfe = ref.react_arkode(sts, DT, rtol=1e-6, atol=1e-10, rk_method=m, rk_controller=0)
        for i, s1 in enumerate(sts):
            pre = "m%d_%d_" % (m, i)
            out.update({pre + "rY": s1.rY, pre + "T": s1.T, pre + "rE": s1.rE, pre + "FC": s1.FC})
        out["m%d_nfe" % m] = np.array(nfe)
        msgs.append("%s rk_method %d: nfe per call %s" % (mech, m, nfe))
    np.savez_compressed(os.path.join(HERE, "arkode_tables_%s.npz" % mech), dt=DT, methods=np.array(METHODS), **out)
    sys.stdout.flush()
    os.dup2(so, 1)
    print("\n".join(msgs))


if __name__ == "__main__":
    main()
