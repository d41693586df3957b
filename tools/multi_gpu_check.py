"""torchrun --nproc-per-node W tools/multi_gpu_check.py : parity of the multi-GPU path
(outputs sharded, peer-memory fused gather AND NCCL gather, 'points' fallback) vs the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import gp_mpc_b200
from gp_mpc_b200 import _lib as L
from oracle import gp_oracle as orc
from tests._util import load_fixture, load_golden, relinf

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
ok = True
def report(tag, **errs):
    global ok
    bad = {k: v for k, v in errs.items() if not (v < 1e-6)}
    if bad: ok = False
    if rank == 0: print(tag, {k: '%.2e' % v for k, v in errs.items()}, 'FAIL' if bad else 'ok', flush=True)

for peer in ('0', '1'):
    os.environ['GPMPC_NO_PEER'] = '1' if peer == '0' else '0'
    # fixture through the GP class
    m = load_fixture('tank'); d = load_golden('derived', 'tank')
    gp = gp_mpc_b200.GP(m['X'], m['Y'], normalize=True, meta=m['meta'], xlb=m['xlb'], xub=m['xub'], ulb=m['ulb'], uub=m['uub'],
                        hyper=dict(hyper=m['hyper']), device=lr)
    mean, cov = gp.predict(d['x0'], d['u0'], d['Sigma'])
    chol = gp.get_chol()
    report('tank peer=%s mode/owned=%s' % (peer, list(gp.engine.local_outputs)), mean=relinf(mean, d['mean_ta']), cov=relinf(cov, d['cov_ta']), chol=relinf(chol, m['chol']) * 1e3)
    gp.close()
    # synthetic, many steps back to back (exercises the double-buffered flags), H > 64 chunks too
    for (N, Nx, Ny, H) in [(1000, 10, 8, 50), (700, 10, 8, 130), (500, 5, 3, 9)]:
        p = orc.synthetic_problem(N, Nx, Ny, config_id=N, H=H)
        post = orc.postfit(p['X'], p['Y'], p['hyper'], lapack_general_solve=False)
        mo, vo = orc.gp_mean_var(p['X'], p['hyper'], post['alpha'], post['chol'], p['Z'])
        Jo = orc.gp_mean_jac(p['X'], p['hyper'], post['alpha'], p['Z']); co = orc.ta_cov(vo, Jo, p['Sigma'])
        gp = gp_mpc_b200.GP(p['X'], p['Y'], normalize=False, hyper=dict(hyper=p['hyper']), device=lr)
        worst = 0.0
        for it in range(25):
            mean, cov = gp.predict_batch(p['Z'][:, :Ny], p['Z'][:, Ny:], p['Sigma'])
            worst = max(worst, relinf(mean, mo), relinf(cov, co))
        report('synthetic N=%d Ny=%d H=%d peer=%s' % (N, Ny, H, peer), worst=worst)
        gp.close()
flag = torch.tensor([0 if ok else 1], device='cuda'); dist.all_reduce(flag)
if rank == 0: print('MULTI_GPU_CHECK', 'PASS' if flag.item() == 0 else 'FAIL', flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 0 else 1)
