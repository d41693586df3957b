"""C4: NLML + analytic gradient evaluations at N=8192 (and the SLSQP fit wall time at a smaller N)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gp_mpc_b200
from gp_mpc_b200 import _lib as L
from bench import make_workload
out = {}
for N in (4096, 8192):
    w = make_workload(N, 8, 1, 4, 30)
    eng = gp_mpc_b200.Engine(N, 8, 1, device=0)
    eng.set_data(w['X'], w['Y'])
    th = w['hyper'][0].copy()
    eng.nlml(0, th, grad=True)
    t0 = time.perf_counter()
    for k in range(5):
        f, g = eng.nlml(0, th * (1 + 0.01 * k), grad=True)
    dt = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    for k in range(5):
        f = eng.nlml(0, th * (1 + 0.01 * k), grad=False)
    dt0 = (time.perf_counter() - t0) / 5
    out['N=%d' % N] = {'nlml_grad_ms': dt * 1e3, 'nlml_only_ms': dt0 * 1e3, 'flops_TF_per_s(N^3)': N ** 3 / dt / 1e12}
    eng.close()
# full fit (reference init, SLSQP, analytic gradient) at N=2048, one output
w = make_workload(2048, 8, 1, 4, 30)
t0 = time.perf_counter()
gp = gp_mpc_b200.GP(w['X'], w['Y'], normalize=False, optimizer_opts={'maxiter': 60})
out['fit_N2048_s'] = time.perf_counter() - t0
gp.close()
print(json.dumps(out, indent=1))
