"""First-contact GPU diagnostics: step-by-step errors of every kernel against the oracle,
then kernel timings.  Writes gpurun_out/diag.log (stdout) -- development aid."""
import json, os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gp_mpc_b200
from gp_mpc_b200 import _lib as L
from oracle import gp_oracle as orc
from tests._util import relinf

def check(N, Nx, Ny, H, cfg):
    print('=== N=%d Nx=%d Ny=%d H=%d' % (N, Nx, Ny, H), flush=True)
    p = orc.synthetic_problem(N, Nx, Ny, config_id=cfg, H=H)
    eng = gp_mpc_b200.Engine(N, Nx, Ny, device=0)
    eng.set_data(p['X'], p['Y']); eng.set_hyper(p['hyper'])
    K = eng.build_K(0)
    Ko = orc.covSEard(p['X'], p['X'], p['hyper'][0, :Nx], p['hyper'][0, Nx]**2) + p['hyper'][0, Nx+1]**2*np.eye(N)
    print('  K build err %.2e sym %s' % (relinf(K, Ko), np.array_equal(K, K.T)), flush=True)
    info = eng.factorize()
    print('  info', info, flush=True)
    post = orc.postfit(p['X'], p['Y'], p['hyper'], lapack_general_solve=False)
    for a in range(Ny):
        chol = eng.get(L.GET_CHOL, a); linv = eng.get(L.GET_LINV, a); al = eng.get(L.GET_ALPHA, a)
        print('  a=%d chol %.2e  Linv*L-I %.2e  alpha %.2e logdet %.3e vs %.3e' % (
            a, relinf(chol, post['chol'][a]), np.abs(linv @ post['chol'][a] - np.eye(N)).max(), relinf(al, post['alpha'][a]),
            eng.get(L.GET_LOGDET, a)[0], 2*np.log(np.diag(post['chol'][a])).sum()), flush=True)
        if not np.isfinite(chol).all():
            bad = np.argwhere(~np.isfinite(chol)); print('   non-finite chol entries, first at', bad[0])
        else:
            e = np.abs(chol - post['chol'][a]); i, j = np.unravel_index(e.argmax(), e.shape)
            print('   worst chol entry at', (i, j), 'tile', (i // 128, j // 128))
    mo, vo = orc.gp_mean_var(p['X'], p['hyper'], post['alpha'], post['chol'], p['Z'])
    Jo = orc.gp_mean_jac(p['X'], p['hyper'], post['alpha'], p['Z'])
    co = orc.ta_cov(vo, Jo, p['Sigma'])
    mean, var, cov, jac = eng.predict(p['Z'], p['Sigma'], L.METHOD_TA)
    print('  predict: mean %.2e var %.2e jac %.2e cov %.2e' % (relinf(mean, mo), relinf(var, vo), relinf(jac, Jo), relinf(cov, co)), flush=True)
    eng.set_option('refine', 1)
    mean, var, cov, jac = eng.predict(p['Z'], p['Sigma'], L.METHOD_TA)
    print('  predict(refine): mean %.2e var %.2e' % (relinf(mean, mo), relinf(var, vo)), flush=True)
    th = p['hyper'][0].copy(); th[:Nx] *= 0.7
    nll, g = eng.nlml(0, th, grad=True)
    print('  nlml %.10e vs %.10e  grad err %.2e' % (nll, orc.calc_NLL(th, p['X'], p['Y'][:, 0]),
          relinf(g, orc.calc_NLL_grad_analytic(th, p['X'], p['Y'][:, 0]))), flush=True)
    eng.close()

def timings(N, Nx=10, Ny=1, H=50):
    print('=== timings N=%d' % N, flush=True)
    p = orc.synthetic_problem(N, Nx, Ny, config_id=5, H=H)
    eng = gp_mpc_b200.Engine(N, Nx, Ny, device=0)
    eng.set_data(p['X'], p['Y']); eng.set_hyper(p['hyper'])
    out = {'N': N}
    ms = eng.profile(L.PROF_KBUILD_FULL, reps=5); out['kbuild_full_ms'] = ms; out['kbuild_full_GBs'] = 8.0*N*N/ms/1e6
    ms = eng.profile(L.PROF_KBUILD_LOWER, reps=5); out['kbuild_lower_ms'] = ms; out['kbuild_lower_GBs'] = 4.0*N*N/ms/1e6
    n1 = (N//128//2)*128; n2 = N - n1
    ms = eng.profile(L.PROF_SYRK, reps=5); out['syrk_ms'] = ms; out['syrk_TF'] = (n2*(n2+128.0))*n1/ms/1e9
    t0 = time.perf_counter(); ms = eng.profile(L.PROF_FACTORIZE, reps=2); out['factorize_ms'] = ms
    out['factorize_TF(2N^3/3)'] = 2.0*N**3/3/ms/1e9
    t0 = time.perf_counter(); eng.factorize(); out['factorize_call_s'] = time.perf_counter()-t0
    ms = eng.profile(L.PROF_TRIGEMM, n=H, reps=10); out['trigemm_ms'] = ms; out['trigemm_TF'] = H*float(N)*N/ms/1e9
    out['trigemm_GBs'] = 4.0*N*N/ms/1e6
    Z = p['Z']; t0 = time.perf_counter()
    for _ in range(10): eng.predict(Z, p['Sigma'], L.METHOD_TA)
    out['predict_host_ms'] = (time.perf_counter()-t0)/10*1e3
    print(json.dumps(out), flush=True)
    eng.close()
    return out

if __name__ == '__main__':
    import torch
    print(torch.cuda.get_device_name(0), flush=True)
    for args in [(300, 5, 2, 30, 1), (1000, 8, 3, 30, 2), (130, 10, 2, 70, 3)]:
        try: check(*args)
        except Exception: traceback.print_exc()
    res = []
    for N in (1024, 4096, 8192, 16384):
        try: res.append(timings(N))
        except Exception: traceback.print_exc()
    a = torch.randn(8192, 8192, dtype=torch.float64, device='cuda'); b = torch.randn_like(a)
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        print('cuBLAS DGEMM 8192^3: %.2f ms  %.1f TF/s' % (e0.elapsed_time(e1), 2*8192.0**3/e0.elapsed_time(e1)/1e9), flush=True)
    x = torch.empty(1 << 28, dtype=torch.float64, device='cuda'); y = torch.empty_like(x)
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize()
        print('copy 2 GiB: %.1f GB/s' % (2*x.numel()*8/e0.elapsed_time(e1)/1e6), flush=True)
    json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'diag_timings.json'), 'w'), indent=1)
