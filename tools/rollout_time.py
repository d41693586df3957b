"""Open-loop roll-out, device-resident (gpmpc_rollout) vs one host call per step, on the tank fixture (N=60) and C2-sized data."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gp_mpc_b200
from tests._util import load_fixture, load_golden
from bench import make_workload

def run(gp, x0, useq, tag):
    for dev in (True, False):
        gp.rollout(x0, useq[:3], methods=['TA'], device_rollout=dev)
        t0 = time.perf_counter()
        for _ in range(5):
            rm, rv = gp.rollout(x0, useq, methods=['TA'], device_rollout=dev)
        dt = (time.perf_counter() - t0) / 5
        print('%s  Nt=%d  %s: %.2f ms per roll-out = %.1f us per step' % (tag, len(useq), 'device-resident' if dev else 'host loop      ', dt * 1e3, dt / len(useq) * 1e6), flush=True)

m = load_fixture('tank'); d = load_golden('derived', 'tank')
gp = gp_mpc_b200.GP(m['X'], m['Y'], normalize=True, meta=m['meta'], xlb=m['xlb'], xub=m['xub'], ulb=m['ulb'], uub=m['uub'],
                    hyper=dict(hyper=m['hyper']), device=0)
useq = np.tile(d['u0'], (200, 1))
run(gp, d['x0'], useq, 'tank N=60 Ny=4 Nu=2')
gp.close()
rng = np.random.default_rng(3)
N, Ny, Nu = 1000, 6, 2
X = rng.uniform(-1, 1, (N, Ny + Nu)); Y = 0.9 * X[:, :Ny] + 0.1 * np.tanh(X[:, :Ny] + X[:, Ny:].sum(1, keepdims=True)) + 1e-2 * rng.standard_normal((N, Ny))
hyper = np.column_stack([np.full((Ny, Ny + Nu), 2.0), np.full(Ny, 1.0), np.full(Ny, 0.05)])
gp = gp_mpc_b200.GP(X, Y, normalize=False, xlb=[-1] * Ny, xub=[1] * Ny, ulb=[-1] * Nu, uub=[1] * Nu, hyper=dict(hyper=hyper), device=0)
run(gp, X[0, :Ny], np.tile(X[:1, Ny:], (200, 1)), 'synthetic N=1000 Ny=6 Nu=2')
gp.close()
