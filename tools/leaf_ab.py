import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gp_mpc_b200
from gp_mpc_b200 import _lib as L
from oracle import gp_oracle as orc
from tests._util import relinf, load_fixture
for name in ('tank', 'car'):
    m = load_fixture(name)
    for lv in (0, 1):
        eng = gp_mpc_b200.Engine(m['X'].shape[0], m['X'].shape[1], m['Y'].shape[1], device=0)
        eng.set_option('leaf_variant', lv)
        eng.set_data(m['X'], m['Y']); eng.set_hyper(m['hyper']); info = eng.factorize()
        N = m['X'].shape[0]
        ch = eng.get(L.GET_CHOL, 0); li = eng.get(L.GET_LINV, 0)
        print(name, 'leaf', lv, 'info', info, 'chol %.2e' % relinf(ch, m['chol'][0]), 'Linv*L-I %.2e' % np.abs(li @ ch - np.eye(N)).max(), flush=True)
        eng.close()
for N in (1000, 4096, 16384):
    p = orc.synthetic_problem(N, 10, 1, config_id=5, H=50)
    for lv in (0, 1):
        eng = gp_mpc_b200.Engine(N, 10, 1, device=0)
        eng.set_option('leaf_variant', lv)
        eng.set_data(p['X'], p['Y']); eng.set_hyper(p['hyper'])
        ms = eng.profile(L.PROF_FACTORIZE, reps=3)
        eng.factorize()
        ch = eng.get(L.GET_CHOL, 0)
        if N <= 4096:
            Lo = np.linalg.cholesky(orc.covSEard(p['X'], p['X'], p['hyper'][0, :10], 1.0) + 1e-4 * np.eye(N))
            err = relinf(ch, Lo)
        else:
            err = float('nan')
        print('N=%d leaf=%d factorize %.2f ms (%.2f TF/s)  chol err %.2e' % (N, lv, ms, 2.0 * N ** 3 / 3 / ms / 1e9, err), flush=True)
        eng.close()
