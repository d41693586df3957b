"""A/B of the persistent stream-K predict product on the GPU box: time per launch and per-CTA
load balance for 1 and 8 outputs at the C5 size, over persistent-grid sizes.
    python tools/predict_ab.py [N] [H]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gp_mpc_b200
from gp_mpc_b200 import _lib as L
from bench import make_workload

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
H = int(sys.argv[2]) if len(sys.argv) > 2 else 50
for nout in (1, 8):
    w = make_workload(N, 10, nout, 5, H)
    eng = gp_mpc_b200.Engine(N, 10, nout, device=0)
    eng.set_data(w['X'], w['Y']); eng.set_hyper(w['hyper']); eng.factorize()
    eng.predict(w['Z'], w['Sigma'], L.METHOD_TA)
    print('N=%d H=%d outputs=%d  ks kernel alone %.1f us' % (N, H, nout, eng.profile(L.PROF_KS, n=H, reps=20) * 1e3), flush=True)
    print('N=%d H=%d outputs=%d  product only %.1f us   product + finalize + assemble %.1f us' % (
        N, H, nout, eng.profile(L.PROF_TRIGEMM, n=H, reps=20) * 1e3, eng.profile(L.PROF_PREDICT_TAIL, n=H, reps=20) * 1e3), flush=True)
    print('   tail phases (us after the last other CTA ended):', {k: round(v, 1) for k, v in eng.profile_tail(H).items()}, flush=True)
    for ctas in (0,):
        eng.set_option('predict_ctas', ctas)
        ms = eng.profile(L.PROF_TRIGEMM, n=H, reps=20)
        bal = eng.profile_balance(H)
        print('N=%d H=%d outputs=%d ctas=%s  %.4f ms  %.2f TF/s  balance(us) min %.1f max %.1f mean %.1f span %.1f' % (
            N, H, nout, ctas or 'auto', ms, nout * H * float(N) * N / ms / 1e9, bal['min_us'], bal['max_us'], bal['mean_us'], bal['span_us']), flush=True)
    eng.close()
