"""A/B of the factorisation switches on the GPU box (development aid; prints only):
GEMM feed (gemm_variant 1 cp.async / 2 per-row bulk copies / 3 TMA tensor maps), leaf kernel (leaf_variant 0/1/2/3 = v1/v2/v3/v3 with two pivots per step),
small-tile threshold, side-stream overlap -- correctness through the whole pipeline, then potrf+trtri / SYRK timings.
    python tools/factor_ab.py [N ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gp_mpc_b200
from gp_mpc_b200 import _lib as L
from oracle import gp_oracle as orc
from tests._util import relinf

p = orc.synthetic_problem(1000, 8, 2, config_id=2, H=30)
post = orc.postfit(p['X'], p['Y'], p['hyper'], lapack_general_solve=False)
for gv in (1, 2, 3):
    for lv in (0, 1, 2, 3):
        eng = gp_mpc_b200.Engine(1000, 8, 2, device=0)
        eng.set_option('gemm_variant', gv); eng.set_option('leaf_variant', lv)
        eng.set_data(p['X'], p['Y']); eng.set_hyper(p['hyper']); eng.factorize()
        print('gemm_variant=%d leaf_variant=%d chol %.2e alpha %.2e' % (gv, lv, relinf(eng.get(L.GET_CHOL, 0), post['chol'][0]),
                                                                      relinf(eng.get(L.GET_ALPHA, 0), post['alpha'][0])), flush=True)
        eng.close()
for N in [int(a) for a in sys.argv[1:]] or [1024, 4096, 8192, 16384]:
    pp = orc.synthetic_problem(N, 10, 1, config_id=5, H=50)
    eng = gp_mpc_b200.Engine(N, 10, 1, device=0)
    eng.set_data(pp['X'], pp['Y']); eng.set_hyper(pp['hyper'])
    n1 = (N // 128 // 2) * 128; n2 = N - n1
    for name, vals in (('gemm_variant', (1, 3)), ('leaf_variant', (1, 2, 3)), ('overlap', (0, 1)), ('lookahead', (0, 1)), ('lookahead_min', (512, 2048, 1024))):
        for v in vals:
            eng.set_option(name, v)
            ms = eng.profile(L.PROF_SYRK, reps=3); msf = eng.profile(L.PROF_FACTORIZE, reps=3)
            print('N=%d %s=%d  syrk %.3f ms %.2f TF/s   potrf+trtri %.3f ms %.2f TF/s' % (
                N, name, v, ms, n2 * (n2 + 1.0) * n1 / ms / 1e9, msf, 2.0 * N ** 3 / 3 / msf / 1e9), flush=True)
    eng.close()
