import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gp_mpc_b200
from gp_mpc_b200 import _lib as L
from bench import make_workload
for N in (16384, 8192):
    w = make_workload(N, 10, 1, 5, 50)
    eng = gp_mpc_b200.Engine(N, 10, 1, device=0)
    eng.set_data(w['X'], w['Y']); eng.set_hyper(w['hyper'])
    for occ in (4, 5, 6):
        eng.set_option('kbuild_occ', occ)
        ms = min(eng.profile(L.PROF_KBUILD_FULL, reps=10) for _ in range(3))
        msl = min(eng.profile(L.PROF_KBUILD_LOWER, reps=10) for _ in range(3))
        print('N=%d occ=%d full %.4f ms %.1f GB/s (%.3f of 6570)   lower %.4f ms %.1f GB/s' % (N, occ, ms, 8.0 * N * N / ms / 1e6, 8.0 * N * N / ms / 1e6 / 6570, msl, 4.0 * N * N / msl / 1e6), flush=True)
    eng.close()
