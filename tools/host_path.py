"""Where a small-N host call spends its time (C2: N=1000, Nx=8, Ny=6, H=30): Python wrapper vs raw ctypes call vs
asynchronous enqueue (device pointers) vs stream synchronisation vs the device step itself."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gp_mpc_b200
from gp_mpc_b200 import _lib as L
from bench import make_workload

N, Nx, Ny, H = (int(a) for a in (sys.argv[1:5] if len(sys.argv) > 4 else (1000, 8, 6, 30)))
w = make_workload(N, Nx, Ny, 2, H)
eng = gp_mpc_b200.Engine(N, Nx, Ny, device=0)
eng.set_data(w['X'], w['Y']); eng.set_hyper(w['hyper']); eng.factorize()
Z, S = w['Z'], w['Sigma']
R = 2000
def timeit(f, reps=R):
    for _ in range(50): f()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter() - t0) / reps * 1e6
t_py = timeit(lambda: eng.predict(Z, S, L.METHOD_TA))
mean = np.empty((H, Ny)); var = np.empty((H, Ny)); cov = np.empty((H, Ny, Ny)); jac = np.empty((H, Ny, Nx))
args = (eng.h, int(L.METHOD_TA), H, L._ptr(Z), L._ptr(S), 0, L._ptr(mean), L._ptr(var), L._ptr(cov), L._ptr(jac))
fn = eng.lib.gpmpc_predict
t_raw = timeit(lambda: fn(*args))
dZ = torch.from_numpy(Z).cuda(); dS = torch.from_numpy(S).cuda()
dm = torch.empty(H, Ny, dtype=torch.float64, device='cuda'); dv = torch.empty_like(dm)
dc = torch.empty(H, Ny, Ny, dtype=torch.float64, device='cuda'); dj = torch.empty(H, Ny, Nx, dtype=torch.float64, device='cuda')
pd = lambda: eng.predict_device(L.METHOD_TA, H, dZ.data_ptr(), dS.data_ptr(), 0, dm.data_ptr(), dv.data_ptr(), dc.data_ptr(), dj.data_ptr())
def enq_sync():
    pd(); eng.synchronize()
t_dev_call = timeit(enq_sync)
# enqueue cost alone: 20 calls back to back, then one sync
def enq20():
    for _ in range(20): pd()
    eng.synchronize()
t_enq20 = timeit(enq20, 200) / 20
st = torch.cuda.ExternalStream(eng.stream())
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(st):
    e0.record(st)
    for _ in range(200): pd()
    e1.record(st)
eng.synchronize()
t_dev = e0.elapsed_time(e1) / 200 * 1e3
t_sync_only = timeit(lambda: eng.synchronize())
print('N=%d Nx=%d Ny=%d H=%d' % (N, Nx, Ny, H))
print('python Engine.predict (host arrays)      %.1f us' % t_py)
print('raw ctypes gpmpc_predict (host arrays)   %.1f us' % t_raw)
print('predict_device + synchronize             %.1f us' % t_dev_call)
print('back-to-back steps, per step (pipelined) %.1f us wall, %.1f us on the device (events)' % (t_enq20, t_dev))
print('synchronize on an idle stream            %.1f us' % t_sync_only)
eng.close()
