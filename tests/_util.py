"""Shared helpers for the test-suite (fixture loading, error norms)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def unpack_lower(packed, n, symmetric=False):
    packed = np.asarray(packed)
    out = np.zeros(packed.shape[:-1] + (n, n))
    i, j = np.tril_indices(n)          # row-major lower order (index work: bit-exact)
    out[..., i, j] = packed
    if symmetric:
        out[..., j, i] = packed
    return out


def load_fixture(name):
    """The reference's saved model gp_<name>_example.json as a model dict."""
    z = np.load(os.path.join(GOLDEN, 'fixture_%s.npz' % name))
    N = z['X'].shape[0]
    m = dict(X=z['X'], Y=z['Y'], hyper=z['hyper'], alpha=z['alpha'],
             chol=unpack_lower(z['chol_packed'], N),
             invK=unpack_lower(z['invK_packed'], N, symmetric=True),
             length_scale=z['length_scale'], signal_var=z['signal_var'],
             noise_var=z['noise_var'], mean=z['mean'],
             normalize=bool(z['normalize']), mean_func='zero')
    if 'invK_full' in z.files:
        m['invK'] = z['invK_full']
    if m['normalize']:
        m['meta'] = {k[5:]: z[k] for k in z.files if k.startswith('meta_')}
        for k in ('xlb', 'xub', 'ulb', 'uub'):
            m[k] = z[k]
    return m


def load_golden(kind, name):
    return np.load(os.path.join(GOLDEN, '%s_%s.npz' % (kind, name)))


def relinf(a, b):
    """batch-infinity-norm relative error  max|a-b| / max|b|  (SURVEY 8d gate)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)
