import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gp_mpc_b200
from gp_mpc_b200 import _lib as L
from oracle import gp_oracle as orc
from tests._util import relinf
# correctness of the TMA-fed variants through the whole pipeline
p = orc.synthetic_problem(1000, 8, 2, config_id=2, H=30)
post = orc.postfit(p['X'], p['Y'], p['hyper'], lapack_general_solve=False)
mo, vo = orc.gp_mean_var(p['X'], p['hyper'], post['alpha'], post['chol'], p['Z'])
for gv, tv in ((2, 1), (1, 2), (3, 1), (1, 3), (3, 3)):
    eng = gp_mpc_b200.Engine(1000, 8, 2, device=0)
    eng.set_option('gemm_variant', gv); eng.set_option('tri_variant', tv)
    eng.set_data(p['X'], p['Y']); eng.set_hyper(p['hyper']); eng.factorize()
    mean, var, _, _ = eng.predict(p['Z'], p['Sigma'], L.METHOD_TA)
    print('gemm_variant=%d tri_variant=%d chol %.2e mean %.2e var %.2e' % (gv, tv, relinf(eng.get(L.GET_CHOL, 0), post['chol'][0]), relinf(mean, mo), relinf(var, vo)), flush=True)
    eng.close()
for N in (4096, 16384):
    pp = orc.synthetic_problem(N, 10, 1, config_id=5, H=50)
    eng = gp_mpc_b200.Engine(N, 10, 1, device=0)
    eng.set_data(pp['X'], pp['Y']); eng.set_hyper(pp['hyper'])
    n1 = (N // 128 // 2) * 128; n2 = N - n1
    for gv in (1, 2, 3):
        eng.set_option('gemm_variant', gv)
        ms = eng.profile(L.PROF_SYRK, reps=5); msf = eng.profile(L.PROF_FACTORIZE, reps=2)
        print('N=%d gemm_variant=%d (%s) syrk %.3f ms %.2f TF/s  factorize %.2f ms' % (N, gv, {1: 'cp.async', 2: 'TMA rows', 3: 'TMA tensor-map'}[gv], ms, n2 * (n2 + 128.0) * n1 / ms / 1e9, msf), flush=True)
    eng.set_option('gemm_variant', 1); eng.factorize()
    for tv in (1, 2, 3):
        eng.set_option('tri_variant', tv)
        for H in (50, 30):
            ms = eng.profile(L.PROF_TRIGEMM, n=H, reps=10)
            print('N=%d tri_variant=%d (%s) H=%d %.4f ms %.2f TF/s' % (N, tv, {1: 'cp.async', 2: 'TMA rows', 3: 'TMA tensor-map'}[tv], H, ms, H * float(N) * N / ms / 1e9), flush=True)
    eng.close()
