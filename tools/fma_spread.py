"""TEST TOOLING: how far does the reference ALGORITHM move when only floating-point contraction changes?
Builds oracle/chem_oracle.c a second time with `-mfma -ffp-contract=fast` (into /tmp/ofma) and compares both builds with
the golden vectors of the unmodified reference, in units of (atol + rtol*|y|).  Usage: python tools/fma_spread.py <mech>"""
import subprocess
import sys, os, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from util import *
import oracle.pyoracle as po
mech = sys.argv[1]
os.makedirs('/tmp/ofma', exist_ok=True)
subprocess.check_call(['gcc', '-O2', '-fPIC', '-mfma', '-ffp-contract=fast', '-std=gnu11',
                       '-DPRX_MECH_HEADER="gen/mech_%s.h"' % mech, '-shared', '-o', '/tmp/ofma/liboracle_%s.so' % mech,
                       'chem_oracle.c', '-lm'], cwd='/root/repo/oracle')
g = golden(mech)
class OF(po.Oracle):
    def __init__(self, mech, path):
        self._path = path
        import oracle.pyoracle as m
        old = m.HERE
        # replicate init with a different path
        self.lib = L = C.CDLL(path); self.mech=mech; self.N=L.orc_num_species(); self.neq=self.N+1
        self.CvodeParams = po._cvode_params_type(self.neq)
        dp, ip = po.dp, po.ip
        L.orc_react_cvode.argtypes = [C.POINTER(self.CvodeParams), C.c_int, dp, dp, dp, dp, dp, dp, ip, C.c_double, C.c_double, C.POINTER(po.CvodeStats)]
strict = po.Oracle(mech)
fma = OF(mech, '/tmp/ofma/liboracle_%s.so' % mech)
for i, kw in enumerate(CVODE_SETTINGS):
    a, b = golden_state(g), golden_state(g)
    strict.react_cvode(a, float(g["dt"]), **kw); fma.react_cvode(b, float(g["dt"]), **kw)
    ref = g["cv%d_rY"%i]
    lim = (kw["atol"] + kw["rtol"]*np.abs(ref))
    print(mech, i, kw["solve_type"], "strict-vs-golden %.3g" % (np.abs(a.rY-ref)/lim).max(), "fma-vs-golden %.3g" % (np.abs(b.rY-ref)/lim).max(),
          "dT/T fma %.2e" % (np.abs(b.T-g["cv%d_T"%i])/b.T).max(), "FC", a.FC.sum(), b.FC.sum(), g["cv%d_FC"%i].sum())
