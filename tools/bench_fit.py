"""Hyper-parameter fit wall time (SLSQP + analytic gradients, a7): the independent per-output fits run sequentially
('parallel_fits': False) vs concurrently (default) on one GPU.   python tools/bench_fit.py [N ...]"""
import os, sys, time, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gp_mpc_b200
from bench import make_workload

for N in [int(a) for a in sys.argv[1:]] or [1000, 2048, 4096]:
    w = make_workload(N, 8, 6, 2, 8)
    for par in (False, True):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            gp = gp_mpc_b200.GP(w['X'], w['Y'], normalize=False, optimizer_opts={'maxiter': 30, 'parallel_fits': par})
        dt = time.perf_counter() - t0
        hy = gp.get_hyper_parameters()['length_scale']
        print('N=%d Ny=6 maxiter=30  parallel_fits=%s  fit wall %.2f s   (checksum %.6f)' % (N, par, dt, float(np.sum(hy))), flush=True)
        gp.close()
