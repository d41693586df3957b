"""Prior mean functions of the reference (``get_mean_function``, gp_functions.py:25-69) as plain
numpy -- they are O(N Nx) host work (the reference evaluates them through a CasADi Function once
per fit, optimize.py:492-493).

    'zero'        m(x) = 0
    'const'       m(x) = a
    'linear'      m(x) = a^T x + b
    'polynomial'  m(x) = a^T x^2 + b^T x + c

The parameters are the TAIL of the hyper row (``hyp_s[-1]``, ``hyp_s[-Nx-1:-1]``,
``hyp_s[-2*Nx-1:-Nx-1]``), after ``[ell_1..ell_Nx, sf, sn]``.  Where they enter:
alpha = K^-1 (y - m(X))  (optimize.py:492-494) and the NLL of the CasADi training twin
(optimize.py:41,72,97).  The reference's prediction never adds m(z) back; GP exposes
that as a flag.
"""
from __future__ import annotations

import numpy as np


def count_mean_params(meanFunc, Nx):
    """optimize.py:402-412 (same names, same NameError)."""
    if meanFunc == 'zero':
        return 0
    if meanFunc == 'const':
        return 1
    if meanFunc == 'linear':
        return Nx + 1
    if meanFunc == 'polynomial':
        return 2 * Nx + 1
    raise NameError('No mean function called: ' + meanFunc)


def _split(hyper_row, Nx, func):
    h = np.asarray(hyper_row, dtype=np.float64).reshape(-1)
    if func == 'const':
        return None, None, h[-1]
    if func == 'linear':
        return None, h[-Nx - 1:-1], h[-1]
    if func == 'polynomial':
        return h[-2 * Nx - 1:-Nx - 1], h[-Nx - 1:-1], h[-1]
    raise NameError('No mean function called: ' + func)


def mean_function(hyper_row, X, func='zero'):
    """m(X) for X:(n,Nx) -> (n,)   (gp_functions.py:44-67)."""
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    if func == 'zero':
        return np.zeros(X.shape[0])
    a2, a1, c = _split(hyper_row, X.shape[1], func)
    m = np.full(X.shape[0], float(c))
    if a1 is not None:
        m = m + X @ a1
    if a2 is not None:
        m = m + (X * X) @ a2
    return m


def mean_jacobian(hyper_row, Z, func='zero'):
    """d m / d z for Z:(H,Nx) -> (H,Nx)."""
    Z = np.atleast_2d(np.asarray(Z, dtype=np.float64))
    J = np.zeros_like(Z)
    if func in ('zero', 'const'):
        return J
    a2, a1, _ = _split(hyper_row, Z.shape[1], func)
    J += a1[None, :]
    if a2 is not None:
        J += 2.0 * Z * a2[None, :]
    return J


def mean_design(X, func='zero'):
    """Phi with m(X) = Phi @ params (params in hyper-row order): the mean is linear in its parameters,
    so d NLL / d params = -Phi^T alpha."""
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    n = X.shape[0]
    if func == 'zero':
        return np.zeros((n, 0))
    if func == 'const':
        return np.ones((n, 1))
    if func == 'linear':
        return np.hstack([X, np.ones((n, 1))])
    if func == 'polynomial':
        return np.hstack([X * X, X, np.ones((n, 1))])
    raise NameError('No mean function called: ' + func)


def mean_bounds(y, Nx, func):
    """Bounds of the mean parameters, optimize.py:452-459 (numeric path; the IPOPT twin :227-232 sets
    the same offset bounds).  For a negative data mean the reference's interval
    [meanF/10 - 1e-8, meanF*10 + 1e-8] is inverted; it is returned sorted so SLSQP accepts it."""
    h_m = count_mean_params(func, Nx)
    lb = -np.inf * np.ones(h_m); ub = np.inf * np.ones(h_m)
    if h_m == 0:
        return np.zeros((0, 2))
    meanF = float(np.mean(y))
    if func == 'const':
        lb[-1], ub[-1] = -1e2, 1e2
    else:
        lo, hi = meanF / 10 - 1e-8, meanF * 10 + 1e-8
        lb[-1], ub[-1] = min(lo, hi), max(lo, hi)
        lb[:-1] = -1e-2
        ub[:-1] = 1e-2
    return np.column_stack([lb, ub])
