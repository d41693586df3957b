"""Hyper-parameter fit driver on top of the CUDA engine.

Mirrors ``train_gp_numpy`` (reference optimize.py:359-503): per output, SLSQP
from the reference's initial point and bounds, then the post-fit block (K ->
chol -> alpha) -- except that every NLML evaluation is one call into
libgpmpc (`gpmpc_nlml`: K build + tensor-core Cholesky + logdet, on the GPU)
and, by default, SLSQP receives the ANALYTIC gradient from the same call instead
of the reference's forward differences (optimize.py:466-467).  ``jac='fd'``
reproduces the reference's finite-difference trajectory (each FD probe is again
a GPU NLML evaluation).
"""
from __future__ import annotations

import time

import numpy as np


from .mean_functions import count_mean_params, mean_bounds, mean_design


def bounds_and_init(X, y, fixed_bounds=False):
    """Bounds / start of optimize.py:433-451 for the zero-mean model.
    ``lb[:Nx] = 1-2`` (= -1) is the reference's typo for 1e-2; it is
    replicated unless ``fixed_bounds``."""
    N, Nx = X.shape
    num_hyp = Nx + 2
    lb = -np.inf * np.ones(num_hyp)
    ub = np.inf * np.ones(num_hyp)
    lb[:Nx] = 1e-2 if fixed_bounds else 1 - 2
    ub[:Nx] = 2e2
    lb[Nx] = 1e-8
    ub[Nx] = 1e2
    lb[Nx + 1] = 10 ** -10
    ub[Nx + 1] = 10 ** -2
    init = np.zeros(num_hyp)
    init[:Nx] = np.std(X, 0)
    init[Nx] = np.std(y)
    init[Nx + 1] = 1e-5
    return np.hstack((lb.reshape(num_hyp, 1), ub.reshape(num_hyp, 1))), init


def train_gp_b200(engine, X, Y, meanFunc='zero', hyper_init=None, multistart=1,
                  optimizer_opts=None, verbose=True):
    """Fit the outputs owned by `engine`; returns hyper rows for those outputs
    (shape (out_count, Nx+2)) -- the caller gathers them across ranks."""
    from scipy.optimize import minimize

    # 'const' / 'linear' / 'polynomial' add h_m mean parameters to every hyper row
    # (optimize.py:402-417).  In the reference's NUMERIC path (the default, and the one mirrored
    # here) the objective ignores them (calc_NLL_numpy reads hyper[:Nx+2] only, optimize.py:337-340;
    # "only support a zero-mean function", :377-379), their bounds are set after `bounds` was built
    # (:435-460) and therefore never reach SLSQP, so they stay at their initial value 0 and the
    # fitted prior mean is identically zero.  That behaviour is the default here: zero columns.
    # optimizer_opts={'fit_mean': True} instead fits them jointly with the kernel parameters on the
    # objective of the reference's CasADi/IPOPT twin (calc_NLL, optimize.py:22-97: NLL of the
    # residual y - m(X)), with the mean bounds of optimize.py:452-459; the extra gradient block is
    # d NLL / d params = -Phi^T alpha (m = Phi params is linear in its parameters).
    h_m = count_mean_params(meanFunc, X.shape[1])
    N, Nx = X.shape
    options = {'disp': False, 'maxiter': 10000}
    jac_mode = 'analytic'
    fixed_bounds = False
    fit_mean = False
    parallel_fits = True
    if optimizer_opts is not None:
        optimizer_opts = dict(optimizer_opts)
        jac_mode = optimizer_opts.pop('jac', jac_mode)
        fixed_bounds = bool(optimizer_opts.pop('fixed_bounds', False))
        fit_mean = bool(optimizer_opts.pop('fit_mean', False)) and h_m > 0
        parallel_fits = bool(optimizer_opts.pop('parallel_fits', True))
        options.update(optimizer_opts)
    if jac_mode not in ('analytic', 'fd'):
        raise ValueError("optimizer_opts['jac'] must be 'analytic' or 'fd'")

    if verbose:
        print('\n________________________________________')
        print('# Optimizing hyperparameters (N=%d)' % N)
        print('----------------------------------------')
    rows = np.zeros((engine.out_count, Nx + 2 + h_m))

    def fit_one(eng, a):
        """SLSQP for output a on `eng` (any engine that owns a); returns (hyper row, seconds)."""
        bounds, init = bounds_and_init(X, Y[:, a], fixed_bounds)
        if hyper_init is not None:
            init = np.asarray(hyper_init, dtype=np.float64)[a, :Nx + 2].copy()
        if fit_mean:
            from ._lib import GET_ALPHA_NLML
            Phi = mean_design(X, meanFunc)
            bounds = np.vstack([bounds, mean_bounds(Y[:, a], Nx, meanFunc)])
            init = np.concatenate([init[:Nx + 2], np.zeros(h_m) if hyper_init is None
                                   else np.asarray(hyper_init, dtype=np.float64)[a, Nx + 2:Nx + 2 + h_m]])
            init = np.clip(init, bounds[:, 0], bounds[:, 1])

        def fun(theta):
            if fit_mean:
                eng.set_y(a, Y[:, a] - Phi @ theta[Nx + 2:])
                if jac_mode != 'analytic':
                    return eng.nlml(a, theta[:Nx + 2], grad=False)
                nll, g = eng.nlml(a, theta[:Nx + 2], grad=True)
                return nll, np.concatenate([g, -Phi.T @ eng.get(GET_ALPHA_NLML, a)])
            if jac_mode == 'analytic':
                return eng.nlml(a, theta, grad=True)
            return eng.nlml(a, theta, grad=False)

        # multistart re-runs from the SAME init (optimize.py:462-469, q8): identical results,
        # so one run decides
        t0 = time.time()
        res = minimize(fun, init, method='SLSQP', jac=(jac_mode == 'analytic'), options=options,
                       bounds=bounds, tol=1e-12)
        if fit_mean:
            eng.set_y(a, Y[:, a])
        return res.x, time.time() - t0

    outs = list(engine.local_outputs)
    workers = None
    if parallel_fits and len(outs) > 1 and hasattr(engine, 'device'):
        # The per-output fits are independent (optimize.py:433 loops over them) and each NLML evaluation
        # is latency-bound below N ~ 8192 (its potrf recursion leaves most SMs idle): run them
        # concurrently, one scratch engine (own CUDA streams) and one host thread per output.  ctypes
        # releases the GIL inside the library calls.  Falls back to the sequential loop when the
        # scratch engines do not fit in device memory.
        try:
            workers = []
            for a in outs:
                w = type(engine)(engine.N, Nx, engine.Ny, a, 1, engine.device)
                w.set_data(X, Y)
                workers.append(w)
        except Exception:
            for w in workers or []:
                w.close()
            workers = None
    if workers:
        from concurrent.futures import ThreadPoolExecutor
        try:
            with ThreadPoolExecutor(max_workers=len(outs)) as ex:
                futs = [ex.submit(fit_one, w, a) for w, a in zip(workers, outs)]
                results = [f.result() for f in futs]
        finally:
            for w in workers:
                w.close()
    else:
        results = [fit_one(engine, a) for a in outs]
    for k, (a, (x, secs)) in enumerate(zip(outs, results)):
        if verbose:
            print("* State %d:  %f s" % (a, secs))
        rows[k, :len(x)] = x
    if verbose:
        print('----------------------------------------')
    return rows
