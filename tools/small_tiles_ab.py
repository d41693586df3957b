import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gp_mpc_b200
from gp_mpc_b200 import _lib as L
from oracle import gp_oracle as orc
from tests._util import relinf
for N in (1000, 4096, 8192, 16384):
    p = orc.synthetic_problem(N, 10, 1, config_id=5, H=50)
    for stv in (148, 296, 444, 592):
        eng = gp_mpc_b200.Engine(N, 10, 1, device=0)
        eng.set_option('small_tiles', stv)
        eng.set_data(p['X'], p['Y']); eng.set_hyper(p['hyper'])
        ms = min(eng.profile(L.PROF_FACTORIZE, reps=3) for _ in range(2))
        eng.factorize()
        err = float('nan')
        if N <= 4096:
            Lo = np.linalg.cholesky(orc.covSEard(p['X'], p['X'], p['hyper'][0, :10], 1.0) + 1e-4 * np.eye(N))
            err = max(relinf(eng.get(L.GET_CHOL, 0), Lo), np.abs(eng.get(L.GET_LINV, 0) @ Lo - np.eye(N)).max())
        print('N=%d small_tiles=%d potrf+trtri %.3f ms (%.2f TF/s) err %.2e' % (N, stv, ms, 2.0 * N ** 3 / 3 / ms / 1e9, err), flush=True)
        eng.close()
