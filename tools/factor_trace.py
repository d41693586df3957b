"""Short driver for a launch list of one factorisation: python tools/factor_trace.py [N] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_mpc_b200
from bench import make_workload

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = make_workload(N, 8, 1, 3, 8)
eng = gp_mpc_b200.Engine(N, 8, 1, device=0)
eng.set_data(w['X'], w['Y']); eng.set_hyper(w['hyper'])
for _ in range(reps):
    eng.factorize()
eng.close()
if len(sys.argv) > 3:          # phase stamps of one leaf
    eng = gp_mpc_b200.Engine(N, 8, 1, device=0)
    eng.set_data(w['X'], w['Y']); eng.set_hyper(w['hyper'])
    st = eng.profile_leaf()
    names = ['start', 'loaded', 'panel0'] + ['step%d' % k for k in range(1, 8)] + ['L stored', 'inv16', 'inv32', 'inv64', 'Linv stored']
    prev = 0.0
    for n_, v in zip(names, st):
        print('%-12s %9.0f cycles  (+%.0f)' % (n_, v, v - prev)); prev = v
    eng.close()
