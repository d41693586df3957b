"""CPU stand-in for gp_mpc_b200.Engine used ONLY by the world_size-2 gloo tests of the
host-side multi-GPU logic (partitioning, id broadcast, hyper gather, result assembly).
The numerics come from the oracle; the product never imports this."""
import numpy as np
import torch.distributed as dist

from oracle import gp_oracle as orc


class OracleEngine:
    def __init__(self, N, Nx, Ny, out_begin=0, out_count=None, device=0):
        self.N, self.Nx, self.Ny = N, Nx, Ny
        self.out_begin = out_begin
        self.out_count = Ny - out_begin if out_count is None else out_count
        self.rank, self.world = 0, 1
        self.uid = None
        self.closed = False

    @property
    def local_outputs(self):
        return range(self.out_begin, self.out_begin + self.out_count)

    @staticmethod
    def comm_unique_id():
        return b'u' * 128

    def comm_init(self, uid, rank, world):
        assert uid == b'u' * 128 and len(uid) == 128        # the id broadcast by rank 0 arrived intact
        self.rank, self.world, self.uid = rank, world, uid

    def set_data(self, X, Y):
        self.X, self.Y = np.array(X), np.array(Y)

    def set_hyper(self, hyper):
        self.hyper = np.array(hyper)

    def set_y(self, a, y):
        self.Y[:, a] = np.asarray(y)

    def set_option(self, name, value):
        pass

    def factorize(self, jitter=1e-8):
        rows = list(self.local_outputs)
        self.post = orc.postfit(self.X, self.Y[:, rows], self.hyper[rows], lapack_general_solve=False)
        return np.zeros(self.out_count, dtype=np.int32)

    def nlml(self, a, theta, grad=True):
        f = orc.calc_NLL(theta, self.X, self.Y[:, a], lapack_general_solve=False)
        return (f, orc.calc_NLL_grad_analytic(theta, self.X, self.Y[:, a])) if grad else f

    def get(self, what, a):
        k = a - self.out_begin
        return {0: self.post['chol'][k], 1: self.post['alpha'][k], 2: self.post['invK'][k]}[what]

    def predict(self, Z, Sigma=None, method=1, want_cov=True, want_jac=True):
        rows = list(self.local_outputs)
        Z = np.asarray(Z).reshape(-1, self.Nx)
        m, v = orc.gp_mean_var(self.X, self.hyper[rows], self.post['alpha'], self.post['chol'], Z)
        J = orc.gp_mean_jac(self.X, self.hyper[rows], self.post['alpha'], Z)
        if self.world > 1:                                   # what ncclAllGather does in libgpmpc
            parts = [None] * self.world
            dist.all_gather_object(parts, (self.out_begin, m, v, J))
            parts.sort(key=lambda t: t[0])
            m = np.concatenate([p[1] for p in parts], 1); v = np.concatenate([p[2] for p in parts], 1)
            J = np.concatenate([p[3] for p in parts], 1)
        cov = None
        if want_cov:
            cov = orc.ta_cov(v, J, Sigma) if (method == 1 and Sigma is not None) else orc.me_cov(v)
        return m, v, cov, (J if want_jac else None)

    # rank-1 append stand-in: refits with the oracle; `fail_on` = (rank, N) simulates ONE rank losing
    # positive definiteness (what gpmpc_append reports as GPMPC_ERR_NOTPD -> Engine.append False)
    fail_on = None

    def append(self, x_new, y_new):
        if OracleEngine.fail_on is not None and (self.rank, self.N) == tuple(OracleEngine.fail_on):
            return False
        self.X = np.vstack([self.X, np.asarray(x_new).reshape(1, -1)])
        self.Y = np.vstack([self.Y, np.asarray(y_new).reshape(1, -1)])
        self.N += 1
        self.factorize()
        return True

    def close(self):
        self.closed = True


class OracleEngineWithRollout(OracleEngine):
    """Adds a numpy restatement of gpmpc_rollout (include/gpmpc.h): the feedback arithmetic of
    rollout_feedback_kernel in the same operation order, the predict step from the oracle."""

    def rollout(self, z0, U, Sigma0, method=1, scale=None):
        Ny, Nx = self.Ny, self.Nx
        Nu = Nx - Ny
        U = np.asarray(U, dtype=np.float64).reshape(-1, Nu) if Nu > 0 else np.zeros((np.shape(U)[0], 0))
        Nt = U.shape[0]
        z = np.asarray(z0, dtype=np.float64).copy(); Sg = np.asarray(Sigma0, dtype=np.float64).copy()
        means = np.empty((Nt, Ny)); var = np.empty((Nt, Ny)); cov = None
        for t in range(Nt):
            m, v, cov, _ = self.predict(z.reshape(1, -1), Sg, method, True, False)
            means[t] = m[0]; var[t] = np.diag(cov[0])
            if t + 1 < Nt:
                zx = m[0]
                if scale is not None:
                    zx = ((zx * scale[0] + scale[1]) - scale[2]) / scale[3]
                z = np.concatenate([zx, U[t + 1]])
                Sg[:Ny, :Ny] = cov[0]
        return means, var, cov[0]

