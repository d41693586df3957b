import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gp_mpc_b200
from gp_mpc_b200 import _lib as L
from bench import make_workload
for N, nout in ((16384, 1), (16384, 8), (8192, 1), (4096, 1), (4096, 6), (1024, 6)):
    w = make_workload(N, 10, nout, 5, 50)
    eng = gp_mpc_b200.Engine(N, 10, nout, device=0)
    eng.set_data(w['X'], w['Y']); eng.set_hyper(w['hyper']); eng.factorize()
    for ks in (0, 256, 512, 1024, 2048, 5632):
        if ks and (ks < N // 32 or ks >= N): continue
        eng.set_option('ksplit', ks)
        ms = min(eng.profile(L.PROF_TRIGEMM, n=50, reps=10) for _ in range(2))
        print('N=%d outputs=%d ksplit=%s  %.4f ms  %.2f TF/s' % (N, nout, ks or 'auto', ms, nout * 50 * float(N) * N / ms / 1e9), flush=True)
    eng.close()
