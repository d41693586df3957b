import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gp_mpc_b200
from gp_mpc_b200 import _lib as L
from oracle import gp_oracle as orc
from tests._util import relinf
for N, nout in ((1000, 1), (4096, 1), (8192, 1), (16384, 1), (4096, 6)):
    p = orc.synthetic_problem(N, 10, nout, config_id=5, H=50)
    for ov in (0, 1):
        eng = gp_mpc_b200.Engine(N, 10, nout, device=0)
        eng.set_option('overlap', ov)
        eng.set_data(p['X'], p['Y']); eng.set_hyper(p['hyper'])
        ms = eng.profile(L.PROF_FACTORIZE, reps=3)
        import time
        t0 = time.perf_counter(); eng.factorize(); tf = time.perf_counter() - t0
        err = float('nan')
        if N <= 4096:
            ch = eng.get(L.GET_CHOL, 0); li = eng.get(L.GET_LINV, 0)
            Lo = np.linalg.cholesky(orc.covSEard(p['X'], p['X'], p['hyper'][0, :10], 1.0) + 1e-4 * np.eye(N))
            err = max(relinf(ch, Lo), np.abs(li @ Lo - np.eye(N)).max())
        print('N=%d outputs=%d overlap=%d  potrf+trtri(1 output) %.2f ms (%.2f TF/s)  factorize() all outputs %.1f ms  err %.2e' % (
            N, nout, ov, ms, 2.0 * N ** 3 / 3 / ms / 1e9, tf * 1e3, err), flush=True)
        eng.close()
