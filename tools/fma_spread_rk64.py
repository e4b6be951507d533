"""TEST TOOLING: how far does the reference's RK64 ALGORITHM move when only floating-point contraction changes?
Builds oracle/chem_oracle.c a second time with `-mfma -ffp-contract=fast` into a scratch copy of the oracle package and
compares both builds on the golden states, for both energy forms (reactor_type 1 = e, 2 = h), in the units of the RK64
parity bar (tests/util.py assert_rk64_parity).  Measured in this container (gcc 13, x86-64):
    LiDryer  type 1: species 6.5e-12  T 2.6e-13      type 2: species 1.4e-12  T 4.5e-14
    drm19    type 1: species 1.2e-10  T 5.9e-12      type 2: species 1.7e-08  T 9.9e-10
(identical step counts in every case).  The GPU always contracts to FMA, so the enthalpy-form test uses a 1e-7 bar.
Usage: python tools/fma_spread_rk64.py"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN = r'''
import sys, numpy as np
root, tag, out = sys.argv[1:4]
sys.path.insert(0, root); sys.path.insert(0, %r)
from util import golden, golden_state
import oracle.pyoracle as po
for mech in ("LiDryer", "drm19"):
    g = golden(mech); o = po.Oracle(mech)
    for rt in (1, 2):
        s = golden_state(g); s.rE = s.rE + (1013250.0 if rt == 2 else 0.0)
        o.react_rk64(s, float(g["dt"]), atol=1e-10, reactor_type=rt)
        np.save("%%s/%%s_%%s_%%d.npy" %% (out, tag, mech, rt), np.vstack([s.rY, s.T[None, :], s.FC[None, :]]))
''' % os.path.join(ROOT, "tests")

if __name__ == "__main__":
    import numpy as np
    tmp = tempfile.mkdtemp()
    pkg = os.path.join(tmp, "oracle")
    os.makedirs(os.path.join(pkg, "_build"))
    for f in ("__init__.py", "pyoracle.py"):
        shutil.copy(os.path.join(ROOT, "oracle", f), pkg)
    for m in ("LiDryer", "drm19"):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-mfma", "-ffp-contract=fast", "-std=gnu11", "-I" + os.path.join(ROOT, "oracle"),
                               '-DPRX_MECH_HEADER="gen/mech_%s.h"' % m, "-shared", "-o", os.path.join(pkg, "_build", "liboracle_%s.so" % m),
                               os.path.join(ROOT, "oracle", "chem_oracle.c"), "-lm"])
    script = os.path.join(tmp, "run.py")
    open(script, "w").write(RUN)
    subprocess.check_call([sys.executable, script, ROOT, "strict", tmp])
    subprocess.check_call([sys.executable, script, tmp, "fma", tmp])
    for mech in ("LiDryer", "drm19"):
        for rt in (1, 2):
            a = np.load("%s/strict_%s_%d.npy" % (tmp, mech, rt))
            f = np.load("%s/fma_%s_%d.npy" % (tmp, mech, rt))
            rho = a[:-2].sum(axis=0, keepdims=True)
            e = np.abs(f[:-2] - a[:-2]) / np.maximum(np.abs(a[:-2]), 1e-2 * rho)
            print("%-8s reactor_type %d  species %.3e  T %.3e  FC equal %.2f" % (mech, rt, e.max(), (np.abs(f[-2] - a[-2]) / a[-2]).max(),
                                                                             np.mean(f[-1] == a[-1])))
