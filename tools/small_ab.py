"""Small-N (C2: N=1000, Nx=8, Ny=6, H=30) predict step: device time per step (CUDA events on the engine's
stream), product kernel alone, host C-ABI call, and sequential H=1 calls -- over persistent-grid sizes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gp_mpc_b200
from gp_mpc_b200 import _lib as L
from bench import make_workload

for (N, Nx, Ny, H, cfg) in [(1000, 8, 6, 30, 2), (4096, 8, 6, 30, 3)]:
    w = make_workload(N, Nx, Ny, cfg, H)
    eng = gp_mpc_b200.Engine(N, Nx, Ny, device=0)
    eng.set_data(w['X'], w['Y']); eng.set_hyper(w['hyper']); eng.factorize()
    st = torch.cuda.ExternalStream(eng.stream())
    dZ = torch.from_numpy(w['Z']).cuda(); dS = torch.from_numpy(w['Sigma']).cuda()
    dm = torch.empty(H, Ny, dtype=torch.float64, device='cuda'); dv = torch.empty_like(dm)
    dc = torch.empty(H, Ny, Ny, dtype=torch.float64, device='cuda'); dj = torch.empty(H, Ny, Nx, dtype=torch.float64, device='cuda')
    fl = torch.empty(64 << 20, dtype=torch.float32, device='cuda')
    torch.cuda.synchronize()
    for ctas in (0, 24, 48, 96, 148, 296):
        eng.set_option('predict_ctas', ctas)
        for _ in range(5):
            eng.predict_device(L.METHOD_TA, H, dZ.data_ptr(), dS.data_ptr(), 0, dm.data_ptr(), dv.data_ptr(), dc.data_ptr(), dj.data_ptr())
        eng.synchronize()
        tot = 0.0; reps = 30
        with torch.cuda.stream(st):
            for _ in range(reps):
                fl.zero_()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(st)
                eng.predict_device(L.METHOD_TA, H, dZ.data_ptr(), dS.data_ptr(), 0, dm.data_ptr(), dv.data_ptr(), dc.data_ptr(), dj.data_ptr())
                e1.record(st); eng.synchronize()
                tot += e0.elapsed_time(e1)
        ms_prod = eng.profile(L.PROF_TRIGEMM, n=H, reps=20)
        t0 = time.perf_counter()
        for _ in range(200):
            eng.predict(w['Z'], w['Sigma'], L.METHOD_TA)
        host = (time.perf_counter() - t0) / 200
        print('N=%d ctas=%s  step (cold L2) %.1f us  product alone (warm) %.1f us  host call %.1f us -> %.0f pred/s e2e' % (
            N, ctas or 'auto', tot / reps * 1e3, ms_prod * 1e3, host * 1e6, H / host), flush=True)
    eng.set_option('predict_ctas', 0)
    z1 = w['Z'][:1].copy()
    t0 = time.perf_counter()
    for _ in range(500):
        eng.predict(z1, None, L.METHOD_ME, want_jac=False)
    print('N=%d sequential H=1 ME calls: %.1f us each' % (N, (time.perf_counter() - t0) / 500 * 1e6), flush=True)
    eng.close()
