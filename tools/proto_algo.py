"""numpy prototype of the GPU algorithm's numerics (recursive potrf+trtri with
explicit-inverse panel solves, Linv-based alpha / predict with refinement).
Development aid only: lets the numerical design be checked without a GPU."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gp_oracle as orc
from tests._util import load_fixture, relinf
import scipy.linalg as sl

def leaf(A):
    L = np.linalg.cholesky(A)
    return L, sl.solve_triangular(L, np.eye(len(A)), lower=True)

def potrf_inv(A, LEAF):
    n = len(A)
    if n <= LEAF:
        return leaf(A)
    n1 = (n // 2 + LEAF - 1) // LEAF * LEAF
    L11, Li11 = potrf_inv(A[:n1, :n1], LEAF)
    L21 = A[n1:, :n1] @ Li11.T
    S = A[n1:, n1:] - L21 @ L21.T
    L22, Li22 = potrf_inv(S, LEAF)
    P = L21 @ Li11
    Li21 = -Li22 @ P
    L = np.zeros_like(A); Li = np.zeros_like(A)
    L[:n1, :n1] = L11; L[n1:, :n1] = L21; L[n1:, n1:] = L22
    Li[:n1, :n1] = Li11; Li[n1:, :n1] = Li21; Li[n1:, n1:] = Li22
    return L, Li

def solve_alpha(L, Li, y, nref):
    a = Li.T @ (Li @ y)
    for _ in range(nref):
        r = y - L @ (L.T @ a)
        a = a + Li.T @ (Li @ r)
    return a

def predict_var(L, Li, ks, sf2, nref):
    v = Li @ ks
    for _ in range(nref):
        r = ks - L @ v
        v = v + Li @ r
    return sf2 - (v * v).sum(0)

for name, LEAF in [('tank', 16), ('car', 32), ('car', 128)]:
    m = load_fixture(name)
    N, Nx = m['X'].shape
    rng = np.random.default_rng(0)
    Z = m['X'][rng.integers(0, N, 30)] + 0.3 * rng.standard_normal((30, Nx)) * m['X'].std(0)
    mean_o, var_o = orc.gp_mean_var(m['X'], m['hyper'], m['alpha'], m['chol'], Z)
    for a in range(m['hyper'].shape[0]):
        ell = m['hyper'][a, :Nx]; sf2 = m['hyper'][a, Nx]**2; sn2 = m['hyper'][a, Nx+1]**2
        K = orc.covSEard(m['X'], m['X'], ell, sf2) + sn2 * np.eye(N)
        L, Li = potrf_inv(K, LEAF)
        ks = orc.covSEard(m['X'], Z, ell, sf2)
        out = [f'{name} leaf{LEAF} a{a} chol {relinf(L, m["chol"][a]):.1e}']
        for nref in (0, 1, 2):
            al = solve_alpha(L, Li, m['Y'][:, a], nref)
            out.append(f'alpha[{nref}] {relinf(al, m["alpha"][a]):.1e} mean {relinf(ks.T@al, mean_o[:, a]):.1e}')
        for nref in (0, 1):
            out.append(f'var[{nref}] {relinf(predict_var(L, Li, ks, sf2, nref), var_o[:, a]):.1e}')
        print(' '.join(out))

p = orc.synthetic_problem(1536, 8, 2, config_id=3, H=30)
post = orc.postfit(p['X'], p['Y'], p['hyper'], lapack_general_solve=False)
mean_o, var_o = orc.gp_mean_var(p['X'], p['hyper'], post['alpha'], post['chol'], p['Z'])
for a in range(2):
    ell = p['hyper'][a, :8]; sf2 = p['hyper'][a, 8]**2; sn2 = p['hyper'][a, 9]**2
    K = orc.covSEard(p['X'], p['X'], ell, sf2) + sn2 * np.eye(1536)
    L, Li = potrf_inv(K, 128)
    ks = orc.covSEard(p['X'], p['Z'], ell, sf2)
    al0 = solve_alpha(L, Li, p['Y'][:, a], 0); al1 = solve_alpha(L, Li, p['Y'][:, a], 1)
    print('synth a%d chol %.1e alpha0 %.1e alpha1 %.1e mean0 %.1e mean1 %.1e var0 %.1e var1 %.1e  var/sf2 min %.1e' % (
        a, relinf(L, post['chol'][a]), relinf(al0, post['alpha'][a]), relinf(al1, post['alpha'][a]),
        relinf(ks.T@al0, mean_o[:, a]), relinf(ks.T@al1, mean_o[:, a]),
        relinf(predict_var(L, Li, ks, sf2, 0), var_o[:, a]), relinf(predict_var(L, Li, ks, sf2, 1), var_o[:, a]), (var_o[:, a]/sf2).min()))
