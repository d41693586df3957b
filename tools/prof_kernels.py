"""Launch each hot kernel a couple of times at benchmark size -- the target of the
`ncu --set full` captures kept under profiles/ (never a source of bench numbers)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gp_mpc_b200
from gp_mpc_b200 import _lib as L
from bench import make_workload

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
which = sys.argv[2] if len(sys.argv) > 2 else 'all'
w = make_workload(N, 10, 1, 5, 50)
eng = gp_mpc_b200.Engine(N, 10, 1, device=0)
eng.set_data(w['X'], w['Y']); eng.set_hyper(w['hyper'])
if which in ('all', 'kbuild'):
    print('kbuild_full ms', eng.profile(L.PROF_KBUILD_FULL, reps=1))
if which in ('all', 'syrk'):
    print('syrk ms', eng.profile(L.PROF_SYRK, reps=1))
if which == 'factor':
    eng.factorize()
if which in ('all', 'trigemm'):
    eng.factorize()
    eng.predict(w['Z'], w['Sigma'], L.METHOD_TA)
    print('trigemm ms', eng.profile(L.PROF_TRIGEMM, n=50, reps=1))
eng.close()
