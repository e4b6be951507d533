#!/usr/bin/env python
"""Generate tests/golden/arkode_<mech>.npz from the UNMODIFIED reference's ReactorArkode (oracle/_ref/libpeleref_<mech>.so; run in
the build container only): three successive react calls of ONE reactor object on three different boxes (each a coupled system
of 48 / 64 / 32 cells; ReactEval states, flame-table states, ReactEval states again) - the step-size controller carries its
history across the calls - for the default configuration (ERKStep, Zonneveld 5-3-4, PID, rtol 1e-6, atol 1e-10) and for
Bogacki-Shampine 4-2-3 with the PI controller and typical values, and for ode.use_erkstep = 0 (ARKStep, explicit table)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
DT = 1e-6


def boxes(ref, mech):
    from oracle.pyoracle import BoxState, reacteval_states
    a = reacteval_states(ref, (1, 48, 1))
    z = np.load(os.path.join(HERE, "pmf_states_%s.npz" % mech))
    idx = np.linspace(0, z["T"].size - 1, 64).astype(int)
    rng = np.random.default_rng(9)
    b = BoxState(z["rY"][:, idx].copy(), z["T"][idx].copy(), z["rE"][idx].copy(),
                 1e-3 * rng.uniform(-1, 1, (z["rY"].shape[0], 64)) * z["rY"][:, idx] / DT, 1e4 * rng.uniform(-1, 1, 64))
    c = reacteval_states(ref, (1, 32, 1))
    c.T += 30.0  # a poor temperature guess: box_flatten re-solves T
    return [a, b, c]


def main():
    so = os.dup(1)
    os.dup2(os.open(os.devnull, os.O_WRONLY), 1)
    from oracle.pyoracle import Reference
    msgs = []
    for mech in ("LiDryer", "drm19"):
        ref = Reference(mech)
        out = {}
        for tag, kw in (("default", dict(rtol=1e-6, atol=1e-10, rk_method=40, rk_controller=0)),
                        ("bs_pi_typ", dict(rtol=1e-8, atol=1e-10, rk_method=30, rk_controller=1,
                                           typ_vals=np.concatenate([np.full(ref.N, 1e-4), [2000.0]]))),
                        ("arkstep", dict(rtol=1e-6, atol=1e-10, rk_method=40, rk_controller=0, use_erkstep=0))):
            sts = boxes(ref, mech)
            ins = [s.copy() for s in sts]
            nfe = ref.react_arkode(sts, DT, **kw)
            for i, (s0, s1) in enumerate(ins and zip(ins, sts)):
                pre = "%s_%d_" % (tag, i)
                out.update({pre + "rY0": s0.rY, pre + "T0": s0.T, pre + "rE0": s0.rE, pre + "rYsrc": s0.rYsrc, pre + "rEsrc": s0.rEsrc,
                            pre + "rY": s1.rY, pre + "T": s1.T, pre + "rE": s1.rE, pre + "FC": s1.FC})
            out[tag + "_nfe"] = np.array(nfe)
            msgs.append("%s %s: nfe per call %s" % (mech, tag, nfe))
        np.savez_compressed(os.path.join(HERE, "arkode_%s.npz" % mech), dt=DT, **out)
    sys.stdout.flush()
    os.dup2(so, 1)
    print("\n".join(msgs))


if __name__ == "__main__":
    main()
