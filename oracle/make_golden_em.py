"""Offline generator of tests/golden/em_mp_{tank,car}.npz: the reference's exact-moment-matching
formula (gp_functions.py:344-418) evaluated in 40-digit arithmetic from the stored (X, Y, hyper) of
the reference's saved models -- see gp_oracle.gp_exact_moment_mp.  Takes minutes (pure-Python
mpmath); the GPU tests only load the result.

    python oracle/make_golden_em.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gp_oracle as orc          # noqa: E402
from tests._util import load_fixture         # noqa: E402


def main():
    for name, npts in (('tank', 3), ('car', 1)):
        m = load_fixture(name)
        X, Y, hyper = m['X'], m['Y'], m['hyper']
        Nx = X.shape[1]
        rng = np.random.default_rng(4 if name == 'tank' else 14)
        # test inputs near the data (the GP's standardised space), input covariances of the size the
        # MPC propagates (mpc_class.py:265-275 / gp_class.py:764,780)
        Z = X[rng.choice(X.shape[0], npts, replace=False)] + 0.05 * rng.standard_normal((npts, Nx)) * X.std(0)
        A = rng.standard_normal((Nx, Nx))
        S0 = 1e-3 * np.diag(X.var(0)) + 1e-4 * (A * X.std(0)) @ (A * X.std(0)).T
        Sg = np.stack([S0 * (1 + 0.3 * h) for h in range(npts)])
        mean = np.zeros((npts, Y.shape[1])); cov = np.zeros((npts, Y.shape[1], Y.shape[1]))
        for h in range(npts):
            mean[h], cov[h] = orc.gp_exact_moment_mp(X, Y, hyper, Z[h], Sg[h], dps=40)
            print(name, h, 'var', np.diag(cov[h]), flush=True)
        np.savez(os.path.join(ROOT, 'tests', 'golden', 'em_mp_%s.npz' % name), Z=Z, Sigma=Sg, mean=mean, cov=cov)


if __name__ == '__main__':
    main()
