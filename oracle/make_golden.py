"""Generate the committed golden fixtures under ``tests/golden/``.

Run in the BUILD container (where /root/reference exists):

    python oracle/make_golden.py

What is written (and where each number comes from):

  tests/golden/fixture_{tank,car}.npz
      The reference's two saved models (examples/models/gp_*_example.json,
      written by gp_class.py:693-734) re-encoded as compact binary: X, Y,
      hyper, alpha, chol (lower triangle packed row-major, np.tril_indices
      order), invK (same packing; it is symmetric to rounding in the files),
      meta/normalize.  These are the reference's OWN stored outputs.

  tests/golden/ref_verbatim_{tank,car}.npz
      Outputs of the reference's numpy-only functions executed VERBATIM from
      /root/reference (casadi/pyDOE/matplotlib stubbed, oracle/ref_loader.py):
      optimize.calc_cov_matrix, optimize.calc_NLL_numpy, GP.covSEard, GP.covar
      on the fixture data and on seeded test points.

  tests/golden/derived_{tank,car}.npz
      Known answers from the oracle restatement (oracle/gp_oracle.py) of the
      CasADi-only graph builders on the fixture data (ME / TA / EM predictions
      at the examples' operating points).  NOT CasADi output -- CasADi is not
      installed; see the oracle header for the pinning status.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import gp_oracle as orc          # noqa: E402
from oracle import ref_loader                # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')

# operating points of the example scripts
POINTS = {
    'tank': dict(x0=[8., 10., 8., 19.], u0=[45., 45.]),     # tank_example.py:83-84
    'car': dict(x0=[13.89, 0., 0.], u0=[0., 0.]),           # car_example.py:202-204
}


def pack_lower(A):
    A = np.asarray(A)
    n = A.shape[-1]
    i, j = np.tril_indices(n)
    return A[..., i, j]


def load_fixture_json(name):
    path = os.path.join(ref_loader.REFERENCE_ROOT, 'examples', 'models', 'gp_%s_example.json' % name)
    with open(path) as f:
        d = json.load(f)
    m = dict(X=np.array(d['X']), Y=np.array(d['Y']),
             hyper=np.array(d['hyper']['hyper']), alpha=np.array(d['hyper']['alpha']),
             chol=np.array(d['hyper']['chol']), invK=np.array(d['hyper']['invK']),
             length_scale=np.array(d['hyper']['length_scale']),
             signal_var=np.array(d['hyper']['signal_var']),
             noise_var=np.array(d['hyper']['noise_var']),
             mean=np.array(d['hyper']['mean']),
             normalize=bool(d['normalize']), mean_func=d['mean_func'])
    if d.get('meta'):
        m['meta'] = {k: np.array(v) for k, v in d['meta'].items()}
        for k in ('xlb', 'xub', 'ulb', 'uub'):
            m[k] = np.array(d[k])
    return m


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load_reference()
    for name in ('tank', 'car'):
        m = load_fixture_json(name)
        N, Nx = m['X'].shape
        Ny = m['Y'].shape[1]
        # upper triangle of the stored chol must be exact zeros
        assert np.all(np.triu(m['chol'], 1) == 0.0)
        fx = dict(X=m['X'], Y=m['Y'], hyper=m['hyper'], alpha=m['alpha'],
                  chol_packed=pack_lower(m['chol']), invK_packed=pack_lower(m['invK']),
                  invK_asym=np.array([np.abs(m['invK'][a] - m['invK'][a].T).max() for a in range(Ny)]),
                  length_scale=m['length_scale'], signal_var=m['signal_var'],
                  noise_var=m['noise_var'], mean=m['mean'],
                  normalize=np.array(m['normalize']))
        if name == 'tank':
            # EM subtracts invK from beta beta^T (gp_functions.py:410-411) and is sensitive to
            # the 1e-10-level asymmetry of the stored invK: keep the tank's full square verbatim
            fx['invK_full'] = m['invK']
        if 'meta' in m:
            for k, v in m['meta'].items():
                fx['meta_' + k] = v
            for k in ('xlb', 'xub', 'ulb', 'uub'):
                fx[k] = m[k]
        np.savez_compressed(os.path.join(OUT, 'fixture_%s.npz' % name), **fx)

        # ---- verbatim reference outputs --------------------------------------
        rng = np.random.default_rng(2024)
        idx = rng.integers(0, N, 12)
        Zt = m['X'][idx] + 0.3 * m['X'].std(0) * rng.standard_normal((12, Nx))
        ver = dict(Zt=Zt)
        nll = np.zeros(Ny)
        Ksum = np.zeros(Ny); Krow = np.zeros((Ny, N)); Kdiag_probe = np.zeros((Ny, 8))
        ks_cls = np.zeros((Ny, N, 12))
        for a in range(Ny):
            ell = m['hyper'][a, :Nx]; sf2 = m['hyper'][a, Nx] ** 2
            nll[a] = float(ref.optimize.calc_NLL_numpy(m['hyper'][a].copy(), m['X'].copy(), m['Y'][:, a].copy()))
            K = ref.optimize.calc_cov_matrix(m['X'].copy(), ell, sf2)
            Ksum[a] = K.sum(); Krow[a] = K[N // 3]
            Kdiag_probe[a] = K[np.arange(8) * (N // 8), (np.arange(8) * 7) % N]
            ks_cls[a] = ref.GP.covSEard(None, m['X'].copy(), Zt.copy(), ell, sf2)
        g = ref_loader.reference_gp_shell(m)
        cov_full = g.covar(Zt.copy())                     # (D,n,n), first Ny slabs filled
        ver.update(nll=nll, K_sum=Ksum, K_row=Krow, K_probe=Kdiag_probe, ks=ks_cls,
                   covar=cov_full)
        np.savez_compressed(os.path.join(OUT, 'ref_verbatim_%s.npz' % name), **ver)

        # ---- derived known answers (oracle restatement) ----------------------
        p = POINTS[name]
        model = dict(m)
        Nu = Nx - Ny
        # input covariance recipe of predict_compare, gp_class.py:758-764,780
        Sigma = np.zeros((Nx, Nx))
        Sigma[:Ny, :Ny] = np.diag(m['hyper'][:, Nx + 1] ** 2)
        Sigma[Ny:, Ny:] = 1e-6 * np.eye(Nu)
        mean_me, cov_me = orc.predict(model, p['x0'], p['u0'], Sigma, 'ME')
        mean_ta, cov_ta = orc.predict(model, p['x0'], p['u0'], Sigma, 'TA')
        A, B = orc.discrete_linearize(model, p['x0'], p['u0'])
        der = dict(x0=np.array(p['x0']), u0=np.array(p['u0']), Sigma=Sigma,
                   mean_me=mean_me, cov_me=cov_me, mean_ta=mean_ta, cov_ta=cov_ta,
                   A=A, B=B)
        if name == 'tank':   # EM is numerically meaningless on the car fixture
            mean_em, cov_em = orc.predict(model, p['x0'], p['u0'], Sigma, 'EM')
            der.update(mean_em=mean_em, cov_em=cov_em)
        # a batch of test points in RAW units around the data, ME + TA
        rng = np.random.default_rng(99)
        if m['normalize']:
            Zraw = m['meta']['meanZ'] + m['meta']['stdZ'] * (0.6 * rng.standard_normal((16, Nx)))
            Zs = (Zraw - m['meta']['meanZ']) / m['meta']['stdZ']
        else:
            sel = rng.integers(0, N, 16)
            Zraw = m['X'][sel] + 0.3 * m['X'].std(0) * rng.standard_normal((16, Nx))
            Zs = Zraw
        mean_b, var_b = orc.gp_mean_var(m['X'], m['hyper'], m['alpha'], m['chol'], Zs)
        J_b = orc.gp_mean_jac(m['X'], m['hyper'], m['alpha'], Zs)
        cov_b = orc.ta_cov(var_b, J_b, Sigma)
        der.update(Zraw=Zraw, Zs=Zs, mean_b=mean_b, var_b=var_b, J_b=J_b, cov_b=cov_b)
        np.savez_compressed(os.path.join(OUT, 'derived_%s.npz' % name), **der)
        print(name, 'N=%d Nx=%d Ny=%d' % (N, Nx, Ny), 'NLL', nll)
        print('  ME mean', mean_me.ravel(), 'var', np.diag(cov_me))
        print('  TA diag', np.diag(cov_ta), 'cov01', cov_ta[0, 1])
        if name == 'tank':
            print('  EM mean(std)', ((mean_em.ravel() - m['meta']['meanY']) / m['meta']['stdY']),
                  'EM diag', np.diag(cov_em))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
