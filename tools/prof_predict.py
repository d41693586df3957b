"""Short driver for ncu captures of the predict step at the C5 size: python tools/prof_predict.py [nout] [N] [H]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_mpc_b200
from gp_mpc_b200 import _lib as L
from bench import make_workload

nout = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
H = int(sys.argv[3]) if len(sys.argv) > 3 else 50
w = make_workload(N, 10, nout, 5, H)
eng = gp_mpc_b200.Engine(N, 10, nout, device=0)
eng.set_data(w['X'], w['Y']); eng.set_hyper(w['hyper']); eng.factorize()
for _ in range(4):
    eng.predict(w['Z'], w['Sigma'], L.METHOD_TA)
eng.close()
