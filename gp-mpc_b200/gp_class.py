"""``GP`` -- the host-side mirror of the reference's ``gp_mpc.gp_class.GP``
(gp_class.py:20-861) whose arithmetic runs in libgpmpc (hand-written sm_100a CUDA).

Same constructor, method names, argument meaning, return shapes, printed banners
and error types as the reference; what changes is where the numbers are made:

  reference                                          here
  -------------------------------------------------  ----------------------------------
  train_gp_numpy + SLSQP/FD  (optimize.py:359-503)   optimize.train_gp_b200 -> gpmpc_nlml
  post-fit chol/alpha/invK   (optimize.py:479-494)   gpmpc_factorize / gpmpc_get
  build_gp / build_TA_cov    (gp_functions.py:72-173) gpmpc_predict (ME / TA, batched)
  GP.covSEard / GP.covar     (gp_class.py:314-381)   gpmpc_build_K / gpmpc_predict

Host work kept in numpy is exactly what the reference keeps on the host:
standardisation of a handful of numbers (gp_class.py:253-262), JSON I/O.
There is no CPU fallback for the dense algebra.
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import _lib
from .comm import Comm
from .mean_functions import mean_function, mean_jacobian
from .optimize import train_gp_b200
from .partition import choose_mode, output_block, point_block

_GPU_METHODS = {'ME': _lib.METHOD_ME, 'TA': _lib.METHOD_TA, 'EM': _lib.METHOD_EM}
_KNOWN_METHODS = ('ME', 'TA', 'EM', 'old_ME', 'old_TA')      # gp_class.py:197-205


def _is_symbolic(v):
    """CasADi SX/MX arguments (mpc_class.py:390-412 calls predict with MX symbols)."""
    return type(v).__module__.split('.')[0] == 'casadi' and type(v).__name__ in ('MX', 'SX')


class GP:
    def __init__(self, X, Y, mean_func="zero", gp_method="TA",
                 optimizer_opts=None, hyper=None, normalize=True, multistart=1,
                 xlb=None, xub=None, ulb=None, uub=None, meta=None,
                 optimize_nummeric=True, device=None, comm=None, engine_factory=None,
                 prior_mean_in_predict=False):
        """ Initialize and optimize GP model  (reference gp_class.py:21-75)

        Extra keyword arguments (not in the reference): ``device`` (CUDA ordinal,
        default $LOCAL_RANK or 0), ``comm`` (a ``Comm``; default: the initialised
        torch.distributed world, else single process), ``engine_factory`` (tests),
        ``prior_mean_in_predict``: the reference fits alpha on y - m(X) (optimize.py:492-494) but
        never adds the prior mean m(z) back when predicting (build_gp is always called without
        meanFunc, gp_class.py:69-71).  False (default) replicates that; True adds m(z)
        to the predicted mean and d m/d z to the Jacobian / Taylor covariance.
        """
        X = np.array(X, dtype=np.float64).copy()
        Y = np.array(Y, dtype=np.float64).copy()
        if X.ndim != 2 or Y.ndim != 2 or X.shape[0] != Y.shape[0]:
            raise ValueError('X must be (N, Nx) and Y (N, Ny) with the same N')
        self.__X = X
        self.__Y = Y
        self.__Ny = Y.shape[1]
        self.__Nx = X.shape[1]
        self.__N = X.shape[0]
        self.__Nu = self.__Nx - self.__Ny            # gp_class.py:36 (q5)

        self.__gp_method = gp_method
        self.__mean_func = mean_func
        self.__normalize = normalize
        self.__comm = comm if comm is not None else Comm()
        self.__device = int(device if device is not None else os.environ.get('LOCAL_RANK', 0))
        self.__engine_factory = engine_factory or _lib.Engine
        self.__engine = None
        self.__invK = None
        self.__prior_mean_in_predict = bool(prior_mean_in_predict)
        self.__xlb = self.__xub = self.__ulb = self.__uub = None

        if meta is not None:                         # gp_class.py:42-50
            self.__meanY = np.array(meta['meanY'])
            self.__stdY = np.array(meta['stdY'])
            self.__meanZ = np.array(meta['meanZ'])
            self.__stdZ = np.array(meta['stdZ'])
            self.__meanX = np.array(meta['meanX'])
            self.__stdX = np.array(meta['stdX'])
            self.__meanU = np.array(meta['meanU'])
            self.__stdU = np.array(meta['stdU'])
        if xlb is not None:                          # kept so load_model -> save_model round-trips
            self.__xlb, self.__xub = np.array(xlb), np.array(xub)
            self.__ulb, self.__uub = np.array(ulb), np.array(uub)

        """ Optimize hyperparameters """
        if hyper is None:
            self.optimize(X=X, Y=Y, opts=optimizer_opts, mean_func=mean_func,
                          xlb=xlb, xub=xub, ulb=ulb, uub=uub,
                          multistart=multistart, normalize=normalize,
                          optimize_nummeric=optimize_nummeric)
        else:
            # gp_class.py:58-66: a saved model carries (hyper, invK, alpha, chol).  The stored
            # X,Y are already standardised (:697-698) and are not re-standardised.  The factors
            # are recomputed on the GPU from (X, hyper): L^-1, which the predict kernels need,
            # is not part of the saved model.
            self.__hyper = np.array(hyper['hyper'], dtype=np.float64)
            self.__set_hyper_views()
            self.__build_engine()
            self.__factorize()

        self.set_method(gp_method)

    # ------------------------------------------------------------------ engine plumbing
    def __set_hyper_views(self):
        Nx = self.__Nx                               # gp_class.py:139-142 (q1)
        self.__hyper_length_scales = self.__hyper[:, :Nx]
        self.__hyper_signal_variance = self.__hyper[:, Nx] ** 2
        self.__hyper_noise_variance = self.__hyper[:, Nx + 1] ** 2
        self.__hyper_mean = self.__hyper[:, (Nx + 1):]

    def __build_engine(self):
        if self.__engine is not None:
            self.__engine.close()
        c = self.__comm
        self.__mode = choose_mode(self.__Ny, c.world)
        if self.__mode == 'outputs':
            b, n = output_block(self.__Ny, c.rank, c.world)
        else:
            b, n = 0, self.__Ny
        self.__engine = self.__engine_factory(self.__N, self.__Nx, self.__Ny, b, n, self.__device)
        self.__engine.set_data(self.__X, self.__Y)
        if c.world > 1 and self.__mode == 'outputs':
            uid = self.__engine_factory.comm_unique_id() if c.rank == 0 else None
            uid = c.broadcast_object(uid, src=0)
            self.__engine.comm_init(uid, c.rank, c.world)
            # fused epilogue + all-gather over NVLink peer memory (CUDA IPC); NCCL stays the fallback
            if hasattr(self.__engine, 'peer_export') and os.environ.get('GPMPC_NO_PEER', '0') != '1':
                self.__attach_peers(c)

    def __attach_peers(self, c):
        """CUDA-IPC exchange blocks for the fused epilogue/all-gather.  The decision is collective:
        if any rank cannot export or map a peer block (no P2P, IPC disabled in the container),
        every rank falls back to the NCCL gather."""
        eng = self.__engine
        try:
            mine = eng.peer_export(int(os.environ.get('GPMPC_PEER_HCAP', 256)))
        except Exception:
            mine = None
        handles = c.allgather_object(mine)
        ok = all(hd is not None for hd in handles)
        if ok:
            try:
                eng.peer_attach(handles)
            except Exception:
                ok = False
        if not all(c.allgather_object(ok)):
            eng.set_option('peer', 0)

    def __has_prior_mean(self):
        return self.__mean_func != 'zero' and self.__hyper.shape[1] > self.__Nx + 2

    def __factorize(self):
        if self.__has_prior_mean():
            # alpha = K^-1 (y - m(X))  (optimize.py:492-494): the engine factorises on the residual
            for a in self.__engine.local_outputs:
                self.__engine.set_y(a, self.__Y[:, a] - mean_function(self.__hyper[a], self.__X, self.__mean_func))
        self.__engine.set_hyper(self.__hyper)
        info = self.__engine.factorize(1e-8)
        for k, a in enumerate(self.__engine.local_outputs):
            if info[k] == 1:                         # optimize.py:486
                print("K matrix is not positive definit, adding jitter!")
        self.__invK = None
        # ranks factorise independently (a jitter retry on one rank can take seconds at large N):
        # line them up again before the first predict step's peer exchange starts its timeout clock
        if self.__comm.world > 1:
            self.__comm.barrier()

    @property
    def engine(self):
        return self.__engine

    # ------------------------------------------------------------------ training
    def optimize(self, X=None, Y=None, opts=None, mean_func='zero',
                 xlb=None, xub=None, ulb=None, uub=None,
                 multistart=1, normalize=True, warm_start=False,
                 optimize_nummeric=True):
        """reference gp_class.py:78-142"""
        self.__mean_func = mean_func
        self.__normalize = normalize

        if normalize and X is not None:              # :85-99  (population std, ddof=0)
            self.__xlb = np.array(xlb)
            self.__xub = np.array(xub)
            self.__ulb = np.array(ulb)
            self.__uub = np.array(uub)
            self.__meanY = np.mean(Y, 0)
            self.__stdY = np.std(Y, 0)
            self.__meanZ = np.mean(X, 0)
            self.__stdZ = np.std(X, 0)
            self.__meanX = np.mean(X[:, :self.__Ny], 0)
            self.__stdX = np.std(X[:, :self.__Ny], 0)
            self.__meanU = np.mean(X[:, self.__Ny:], 0)
            self.__stdU = np.std(X[:, self.__Ny:], 0)

        if X is not None:                            # :101-117
            X = np.array(X).copy()
            self.__X = self.standardize(X, self.__meanZ, self.__stdZ) if normalize else X.copy()
        if Y is not None:
            Y = np.array(Y).copy()
            self.__Y = self.standardize(Y, self.__meanY, self.__stdY) if (normalize and X is not None) else Y.copy()
        self.__N = self.__X.shape[0]

        hyp_init = self.__hyper if warm_start else None
        self.__build_engine()
        # optimize_nummeric=False selects the CasADi/IPOPT twin in the reference
        # (optimize.py:100-294, same objective with AD gradients).  Here both settings use the
        # GPU NLML with its analytic gradient under SLSQP.
        rows = train_gp_b200(self.__engine, self.__X, self.__Y, meanFunc=self.__mean_func,
                             optimizer_opts=opts, multistart=multistart, hyper_init=hyp_init)
        blocks = self.__comm.allgather_object((self.__engine.out_begin, rows))
        hyper = np.zeros((self.__Ny, blocks[0][1].shape[1]))      # Nx+2 (+ mean parameters, all zero)
        for b, r in blocks:
            hyper[b:b + len(r)] = r
        self.__hyper = hyper
        self.__lam_x = 0
        self.__set_hyper_views()
        self.__factorize()

    def validate(self, X_test, Y_test):
        """ Validate GP model with test data  (reference gp_class.py:145-190; one batched
        GPU predict instead of the per-row Python loop) """
        Y_test = np.array(Y_test, dtype=np.float64).copy()
        X_test = np.array(X_test, dtype=np.float64).copy()
        if self.__normalize:
            Y_test = self.standardize(Y_test, self.__meanY, self.__stdY)
            X_test = self.standardize(X_test, self.__meanZ, self.__stdZ)

        N, Ny = Y_test.shape
        mean, var = self.__predict_std(X_test, None, 'ME', want_cov=False, want_jac=False)[:2]
        var = var + self.noise_variance()[None, :]                      # :161
        loss = np.sum((Y_test - mean) ** 2, 0) / N
        NLP = np.sum(0.5 * np.log(2 * np.pi * var) + ((Y_test - mean) ** 2) / (2 * var), 0)
        SMSE = loss / np.std(Y_test, 0)                                 # :166 (q15)
        MNLP = NLP / N

        print('\n________________________________________')
        print('# Validation of GP model ')
        print('----------------------------------------')
        print('* Num training samples: ' + str(self.__N))
        print('* Num test samples: ' + str(N))
        print('----------------------------------------')
        print('* Mean squared error: ')
        for i in range(Ny):
            print('\t- State %d: %f' % (i + 1, loss[i]))
        print('----------------------------------------')
        print('* Standardized mean squared error:')
        for i in range(Ny):
            print('\t* State %d: %f' % (i + 1, SMSE[i]))
        print('----------------------------------------')
        print('* Mean Negative log Probability:')
        for i in range(Ny):
            print('\t* State %d: %f' % (i + 1, MNLP[i]))
        print('----------------------------------------\n')

        self.__SMSE = np.max(SMSE)
        return np.array(SMSE).flatten(), np.array(MNLP).flatten()

    # ------------------------------------------------------------------ prediction
    def set_method(self, gp_method='TA'):
        """ Select wich GP function to use  (reference gp_class.py:193-242)

            'ME': Mean Equivalence (normal GP), 'TA': 1st order Taylor Approximation and
            'EM': exact moment matching run on the GPU.  The deprecated 'old_ME'/'old_TA' are
            valid names in the reference but out of scope here.
        """
        if gp_method not in _KNOWN_METHODS:
            raise NameError('No GP method called: ' + gp_method)        # gp_class.py:237
        if gp_method not in _GPU_METHODS:
            raise NotImplementedError("gp_method %r is not implemented by the B200 engine "
                                      "(available: 'ME', 'TA', 'EM')" % gp_method)
        if gp_method == 'EM' and self.__sharded_outputs():
            # exact moment matching couples every pair of outputs (gp_functions.py:394-412): it
            # needs all Ny factors on one GPU, which the by-output sharding does not provide
            raise NotImplementedError("gp_method 'EM' needs all outputs on one GPU; this GP is sharded by "
                                      "output over %d ranks (use 'TA'/'ME', or build the GP with a "
                                      "single-process Comm)" % self.__comm.world)
        self.__gp_method = gp_method

    def __sharded_outputs(self):
        return self.__comm.world > 1 and getattr(self, '_GP__mode', 'outputs') == 'outputs'

    def __predict_std(self, Z, Sigma, method, want_cov=True, want_jac=True):
        """Batched predict in the GP's standardised space.  Z:(H,Nx)."""
        Z = np.ascontiguousarray(Z, dtype=np.float64).reshape(-1, self.__Nx)
        c = self.__comm
        add_pm = self.__prior_mean_in_predict and self.__has_prior_mean() and method != 'EM'
        need_jac = want_jac or add_pm
        if c.world > 1 and self.__mode == 'points':
            b, n = point_block(Z.shape[0], c.rank, c.world)
            Sg = Sigma[b:b + n] if (Sigma is not None and np.ndim(Sigma) == 3) else Sigma
            part = self.__engine.predict(Z[b:b + n], Sg, _GPU_METHODS[method], want_cov, need_jac) if n else None
            parts = [p for p in c.allgather_object(part) if p is not None]
            out = tuple(None if parts[0][k] is None else np.concatenate([p[k] for p in parts], 0)
                        for k in range(4))
        else:
            out = self.__engine.predict(Z, Sigma, _GPU_METHODS[method], want_cov, need_jac)
        if add_pm:
            out = self.__add_prior_mean(Z, Sigma, method, out, want_cov, want_jac)
        return out

    def __add_prior_mean(self, Z, Sigma, method, out, want_cov, want_jac):
        """flag-gated fix of SURVEY q2: mean += m(z), J += dm/dz, and for 'TA' the J Sigma J^T term is
        rebuilt with the full Jacobian (O(H Ny Nx^2) host work)."""
        mean, var, cov, jac = out
        M = np.column_stack([mean_function(self.__hyper[a], Z, self.__mean_func) for a in range(self.__Ny)])
        Jm = np.stack([mean_jacobian(self.__hyper[a], Z, self.__mean_func) for a in range(self.__Ny)], 1)
        mean = mean + M
        if jac is not None:
            jfull = jac + Jm
            if cov is not None and method == 'TA' and Sigma is not None:
                S = np.broadcast_to(np.asarray(Sigma, dtype=np.float64), (Z.shape[0], self.__Nx, self.__Nx)) \
                    if np.ndim(Sigma) == 2 else np.asarray(Sigma, dtype=np.float64)
                cov = cov + np.einsum('had,hde,hbe->hab', jfull, S, jfull) - np.einsum('had,hde,hbe->hab', jac, S, jac)
            jac = jfull
        return mean, var, cov, (jac if want_jac else None)

    def predict_batch(self, x, u, cov=None, method=None):
        """Horizon batch: x:(H,Ny) u:(H,Nu) cov:(Nx,Nx)|(H,Nx,Nx)|None ->
        mean:(H,Ny) (de-standardised), cov:(H,Ny,Ny) (standardised units, q4).
        This is the entry a batched MPC shooting loop (mpc_class.py:361-423) calls: one GPU
        pass for all nodes instead of Nt symbolic graph copies."""
        method = method or self.__gp_method
        x = np.asarray(x, dtype=np.float64).reshape(-1, self.__Ny)
        u = np.asarray(u, dtype=np.float64).reshape(x.shape[0], self.__Nu)
        if self.__normalize:
            x = self.standardize(x, self.__meanX, self.__stdX)
            u = self.standardize(u, self.__meanU, self.__stdU)
        Z = np.hstack([x, u])
        if cov is None and method in ('TA', 'EM'):
            cov = np.zeros((self.__Nx, self.__Nx))      # no input uncertainty: TA reduces to diag(var)
        mean, var, c, _ = self.__predict_std(Z, cov if method in ('TA', 'EM') else None, method, True, False)
        if self.__normalize:
            mean = self.inverse_mean(mean, self.__meanY, self.__stdY)
        return mean, c

    def predict_batch_grad(self, x, u, cov=None, method=None):
        """predict_batch plus the first derivatives CasADi's AD extracts from the symbolic GP when nlpsol
        differentiates the MPC's NLP (mpc_class.py:390-412, :496-513): a dict with
            mean (H,Ny)           de-standardised, as predict_batch
            cov  (H,Ny,Ny)        standardised units (q4)
            dmean_dz (H,Ny,Nx)    d mean / d [x,u] in the CALLER's units (chain rule through the scalers)
            dcov_dz  (H,Ny,Ny,Nx) d cov / d [x,u]  (cov itself is not rescaled, so only 1/stdZ enters)
            dcov_dSigma_factor (H,Ny,Nx)  J with d cov[a][b] / d Sigma[d][e] = J[a][d] J[b][e] ('TA')
        Methods 'ME' and 'TA'.  This is what a casadi.Callback's Jacobian function returns; the same
        numbers are available to `casadi.external` through gp_b200 / jac_gp_b200 (include/gpmpc_casadi.h)."""
        method = method or self.__gp_method
        if method not in ('ME', 'TA'):
            raise NotImplementedError("derivatives are available for gp_method 'ME' and 'TA'")
        if self.__comm.world > 1 and self.__mode == 'outputs':
            raise NotImplementedError('predict_batch_grad needs all outputs on one GPU (build the GP with a single-process Comm)')
        x = np.asarray(x, dtype=np.float64).reshape(-1, self.__Ny)
        u = np.asarray(u, dtype=np.float64).reshape(x.shape[0], self.__Nu)
        if self.__normalize:
            x = self.standardize(x, self.__meanX, self.__stdX)
            u = self.standardize(u, self.__meanU, self.__stdU)
        Z = np.hstack([x, u])
        if cov is None and method == 'TA':
            cov = np.zeros((self.__Nx, self.__Nx))
        g = self.__engine.predict_grad(Z, cov if method == 'TA' else None, _GPU_METHODS[method])
        mean, jac, dcov = g['mean'], g['jac'], g['dcov_dz']
        out = dict(cov=g['cov'], dcov_dSigma_factor=jac.copy())
        if self.__normalize:
            mean = self.inverse_mean(mean, self.__meanY, self.__stdY)
            jac = jac * self.__stdY[None, :, None] / self.__stdZ[None, None, :]
            dcov = dcov / self.__stdZ[None, None, None, :]
        out.update(mean=mean, dmean_dz=jac, dcov_dz=dcov)
        return out

    def predict(self, x, u, cov):
        """ Predict future state  (reference gp_class.py:245-263)

        # Arguments:
            x: State vector (Nx x 1)
            u: Input vector (Nu x 1)
            cov: Covariance matrix of input z=[x, u] (Nx+nu x Nx+Nu)
        # Returns mean (Ny,1) [de-standardised] and cov (Ny,Ny) [NOT rescaled, q4]
        """
        if _is_symbolic(x) or _is_symbolic(u) or _is_symbolic(cov):
            raise NotImplementedError(
                'symbolic (CasADi MX/SX) predict: bind the engine with casadi.external("gp_b200", libgpmpc.so) or wrap '
                'GP.predict_batch / predict_batch_grad in a casadi.Callback as described in INTEGRATION.md section 3 '
                '(SURVEY 8f row 1); CasADi is not installed in this build, so that last Python step is not exercised here')
        x = np.asarray(x, dtype=np.float64).reshape(-1)
        u = np.asarray(u, dtype=np.float64).reshape(-1)
        mean, c = self.predict_batch(x.reshape(1, -1), u.reshape(1, -1),
                                     None if cov is None else np.asarray(cov, dtype=np.float64))
        return mean.reshape(self.__Ny, 1), c[0]

    def rollout(self, x0, u, methods=None, device_rollout=True):
        """ The numeric multi-step prediction of ``predict_compare`` (reference
        gp_class.py:746-804, open-loop branch) without the plotting / plant simulation:
        for every method, propagate (mean, covariance) through ``predict`` over the input
        sequence u:(Nt,Nu), starting from x0 with covariance diag(sn2) (+1e-6 on the inputs).
        Returns mean, var of shape (len(methods), Nt+1, Ny); var is rescaled by stdY^2 when
        normalize (:795-796)."""
        Nx, Ny = self.__Nx, self.__Ny
        u = np.asarray(u, dtype=np.float64)
        u = u.reshape(-1, self.__Nu) if self.__Nu > 0 else np.zeros((u.shape[0] if u.ndim else 0, 0))   # Nu = 0: Nt = len(u)
        Nt = u.shape[0]
        initVar = self.__hyper[:, Nx + 1] ** 2
        if methods is None:                             # gp_class.py:747 default; 'EM' only where it can run
            methods = ['TA', 'ME'] if self.__sharded_outputs() else ['EM', 'TA', 'ME']
        methods = list(methods)
        mean = np.zeros((len(methods), Nt + 1, Ny))
        var = np.zeros((len(methods), Nt + 1, Ny))
        covar = np.eye(Nx) * 1e-6                       # shared across methods, as in the reference
        keep = self.__gp_method
        # 'ME' / 'TA' on a single handle: all Nt steps run on the device (gpmpc_rollout), same arithmetic as the loop below
        on_device = (device_rollout and hasattr(self.__engine, 'rollout') and self.__comm.world == 1
                     and not (self.__prior_mean_in_predict and self.__has_prior_mean()))
        for i, meth in enumerate(methods):
            self.set_method(meth)
            mean_t = np.asarray(x0, dtype=np.float64).reshape(-1)
            covar[:Ny, :Ny] = np.diag(initVar)
            mean[i, 0, :] = mean_t
            if on_device and meth in ('ME', 'TA') and Nt > 0:
                un = u
                z_x = mean_t
                scale = None
                if self.__normalize:
                    z_x = self.standardize(mean_t, self.__meanX, self.__stdX)
                    un = self.standardize(u, self.__meanU, self.__stdU)
                    scale = np.stack([self.__stdY, self.__meanY, self.__meanX, self.__stdX])
                z0 = np.concatenate([np.asarray(z_x, dtype=np.float64).reshape(-1), un[0].reshape(-1)])
                m_std, v_std, c_last = self.__engine.rollout(z0, un, covar, _GPU_METHODS[meth], scale)
                mean[i, 1:, :] = self.inverse_mean(m_std, self.__meanY, self.__stdY) if self.__normalize else m_std
                var[i, 1:, :] = self.inverse_variance(v_std) if self.__normalize else v_std
                covar[:Ny, :Ny] = c_last
                continue
            for t in range(1, Nt + 1):
                mean_t, covar_x = self.predict(mean_t, u[t - 1, :], covar)
                mean_t = np.array(mean_t).reshape(Ny)
                mean[i, t, :] = mean_t
                var[i, t, :] = np.diag(covar_x)
                if self.__normalize:
                    var[i, t, :] = self.inverse_variance(var[i, t, :])
                covar[:Ny, :Ny] = covar_x
        self.set_method(keep)
        return mean, var

    def get_size(self):
        """ (N, Ny, Nu)  (reference gp_class.py:266-274) """
        return self.__N, self.__Ny, self.__Nu

    def get_hyper_parameters(self):
        """ reference gp_class.py:277-290 """
        return dict(length_scale=self.__hyper_length_scales, signal_var=self.__hyper_signal_variance,
                    noise_var=self.__hyper_noise_variance, mean=self.__hyper_mean)

    def print_hyper_parameters(self):
        """ Print out all hyperparameters  (reference gp_class.py:293-312) """
        print('\n________________________________________')
        print('# Hyper-parameters')
        print('----------------------------------------')
        print('* Num samples:', self.__N)
        print('* Ny:', self.__Ny)
        print('* Nu:', self.__Nu)
        print('* Normalization:', self.__normalize)
        for state in range(self.__Ny):
            print('----------------------------------------')
            print('* Lengthscale: ', state)
            for i in range(self.__Ny + self.__Nu):
                print(('-- l{a}: {l}').format(a=i, l=self.__hyper_length_scales[state, i]))
            print('* Signal variance: ', state)
            print('-- sf2:', self.__hyper_signal_variance[state])
            print('* Noise variance: ', state)
            print('-- sn2:', self.__hyper_noise_variance[state])
        print('----------------------------------------')

    def covSEard(self, X, Z, ell, sf2):
        """ GP Squared Exponential Kernel k(X,Z)  (reference gp_class.py:314-350).
        Same argument handling and ValueError; evaluated by the engine's K-build kernel
        on a scratch handle (X and Z stacked, off-diagonal block returned). """
        X = np.asarray(X, dtype=np.float64); Z = np.asarray(Z, dtype=np.float64)
        X = X.reshape(1, -1) if X.ndim == 1 else X
        Z = Z.reshape(1, -1) if Z.ndim == 1 else Z
        n1, D = X.shape
        n2, D2 = Z.shape
        if D != D2:
            raise ValueError('Input dimensions are not the same! D_x=' + str(D) + ', D_z=' + str(D2))
        hyp = np.concatenate([np.asarray(ell, dtype=np.float64).reshape(-1), [np.sqrt(sf2), 0.0]])[None, :]
        eng = self.__engine_factory(n1 + n2, D, 1, 0, 1, self.__device)
        try:
            eng.set_data(np.vstack([X, Z]), np.zeros((n1 + n2, 1)))
            eng.set_hyper(hyp)
            K = eng.build_K(0)
        finally:
            eng.close()
        return K[:n1, n1:].copy()

    def covar(self, X_new):
        """ Compute covariance of input data  (reference gp_class.py:353-381)

        # Arguments:
            X_new: Input matrix or vector of size (n x D), in the GP's (standardised) space.
        # Returns:
            covar: (D x (n x n)) -- the reference allocates D = input-dimension slabs and fills
                   the first Ny with  kss - v^T v  (q12); (D x n) for a 1-D input.
        """
        X_new = np.asarray(X_new, dtype=np.float64)
        one_d = X_new.ndim == 1
        Z = X_new.reshape(1, -1) if one_d else X_new
        n, D = Z.shape
        eng = self.__engine
        mine = [(eng.out_begin, eng.posterior_cov(Z))]
        if self.__comm.world > 1 and self.__mode == 'outputs':
            mine = self.__comm.allgather_object(mine[0])
        covar = np.zeros((D, n) if one_d else (D, n, n))
        for b, blk in mine:
            for k in range(blk.shape[0]):
                covar[b + k] = blk[k].reshape(n) if one_d else blk[k]
        return covar

    def update_data_all(self, X_new, Y_new):
        """ Update training data with all new observations  (reference gp_class.py:474-550):
        append, keep the hyper-parameters, rebuild chol / alpha on the GPU. """
        X_new = np.array(X_new, dtype=np.float64).copy()
        Y_new = np.array(Y_new, dtype=np.float64).copy()
        if self.__normalize:
            Y_new = self.standardize(Y_new, self.__meanY, self.__stdY)
            X_new = self.standardize(X_new, self.__meanZ, self.__stdZ)
        print('\n________________________________________')
        print('# Updating training data with ' + str(X_new.shape[0]) + ' new samples')
        print('----------------------------------------')
        self.__X = np.vstack([self.__X, X_new])
        self.__Y = np.vstack([self.__Y, Y_new])
        self.__N = self.__X.shape[0]
        self.__build_engine()
        self.__factorize()
        self.set_method(self.__gp_method)

    def append_data(self, X_new, Y_new):
        """ Add observations one at a time with the O(N^2) rank-1 update of L, L^-1 and alpha
        (what the reference's ``update_data``, gp_class.py:384-471, set out to do); hyper-
        parameters are kept.  Falls back to a full refactorisation (``update_data_all``) when the
        padded capacity is exhausted or an update loses positive definiteness. """
        X_new = np.array(X_new, dtype=np.float64).reshape(-1, self.__Nx)
        Y_new = np.array(Y_new, dtype=np.float64).reshape(-1, self.__Ny)
        Xs, Ys = X_new, Y_new
        if self.__normalize:
            Ys = self.standardize(Y_new, self.__meanY, self.__stdY)
            Xs = self.standardize(X_new, self.__meanZ, self.__stdZ)
        for k in range(Xs.shape[0]):
            ok = self.__engine.append(Xs[k], Ys[k])
            if self.__comm.world > 1:
                # the fallback below runs collectives (engine rebuild): the decision must be collective
                # too -- a rank that alone lost positive definiteness would otherwise hang the others
                ok = all(self.__comm.allgather_object(bool(ok)))
            if not ok:
                self.__X = np.vstack([self.__X, Xs[k:]])
                self.__Y = np.vstack([self.__Y, Ys[k:]])
                self.__N = self.__X.shape[0]
                self.__build_engine()
                self.__factorize()
                return
            self.__X = np.vstack([self.__X, Xs[k:k + 1]])
            self.__Y = np.vstack([self.__Y, Ys[k:k + 1]])
            self.__N = self.__X.shape[0]
        self.__invK = None

    def replace_data_all(self, X_new, Y_new):
        """ Replace training data with new observations  (reference gp_class.py:553-626) """
        X_new = np.array(X_new, dtype=np.float64).copy()
        Y_new = np.array(Y_new, dtype=np.float64).copy()
        if self.__normalize:
            Y_new = self.standardize(Y_new, self.__meanY, self.__stdY)
            X_new = self.standardize(X_new, self.__meanZ, self.__stdZ)
        print('\n________________________________________')
        print('# Replacing training data with ' + str(X_new.shape[0]) + ' new samples')
        print('----------------------------------------')
        self.__X = X_new
        self.__Y = Y_new
        self.__N = self.__X.shape[0]
        self.__build_engine()
        self.__factorize()
        self.set_method(self.__gp_method)

    def update_data(self, X_new, Y_new, N_new=None):
        """reference gp_class.py:384-471 is self-declared broken ("NOT working as intended",
        :397).  Not replicated."""
        raise NotImplementedError('GP.update_data is broken in the reference (gp_class.py:397); '
                                  'use update_data_all / replace_data_all')

    def standardize(self, Y, mean, std):
        return (Y - mean) / std                      # gp_class.py:629-630

    def normalize(self, u, lb, ub):
        return (u - lb) / (ub - lb)                  # gp_class.py:632-633

    def inverse_mean(self, x, mean, std):
        """ Inverse standardization of the mean  (gp_class.py:635-638) """
        return (x * std) + mean

    def inverse_variance(self, variance):
        """ Inverse standardization of the variance  (gp_class.py:640-644) """
        return variance * self.__stdY ** 2

    def discrete_linearize(self, x0, u0, cov0):
        """ Linearize the GP around the operating point  x[k+1] = Ax[k] + Bu[k]
        (reference gp_class.py:647-661): Jacobian of the predicted mean in standardised
        space, inputs standardised when normalize, outputs not rescaled.  The reference
        differentiates the ACTIVE method's mean; for 'ME'/'TA' that is the posterior-mean Jacobian
        returned here, for 'EM' (mean depends on the input covariance) the reference's A, B differ --
        this engine always linearises the 'ME' mean. """
        x0 = np.asarray(x0, dtype=np.float64).reshape(-1)
        u0 = np.asarray(u0, dtype=np.float64).reshape(-1)
        if self.__normalize:
            x0 = self.standardize(x0, self.__meanX, self.__stdX)
            u0 = self.standardize(u0, self.__meanU, self.__stdU)
        J = self.__predict_std(np.concatenate([x0, u0])[None, :], None, 'ME', False, True)[3][0]
        return J[:, :self.__Ny].copy(), J[:, self.__Ny:].copy()

    def jacobian(self, x0, u0, cov0):
        """ Jacobian of posterior mean J = dmu/dx  (reference gp_class.py:664-672; no
        standardisation there either) """
        z = np.concatenate([np.asarray(x0, dtype=np.float64).reshape(-1),
                            np.asarray(u0, dtype=np.float64).reshape(-1)])
        J = self.__predict_std(z[None, :], None, 'ME', False, True)[3][0]
        return J[:, :self.__Ny].copy()

    def noise_variance(self):
        """ Get the noise variance  (gp_class.py:675-678) """
        return self.__hyper_noise_variance

    def sparse(self, M):
        """ Sparse Gaussian Process -- an empty stub in the reference too (gp_class.py:682-689) """

    # ------------------------------------------------------------------ factors / model I/O
    def __gather_factor(self, what):
        eng = self.__engine
        mine = [(a, eng.get(what, a)) for a in eng.local_outputs]
        if self.__comm.world > 1 and self.__mode == 'outputs':
            mine = [p for blk in self.__comm.allgather_object(mine) for p in blk]
        return np.stack([m for _, m in sorted(mine, key=lambda t: t[0])], 0)

    def get_chol(self):
        return self.__gather_factor(_lib.GET_CHOL)

    def get_alpha(self):
        return self.__gather_factor(_lib.GET_ALPHA)

    def get_invK(self):
        if self.__invK is None:
            self.__invK = self.__gather_factor(_lib.GET_INVK)
        return self.__invK

    def _GP__to_dict(self):
        """ Store model data in a dictionary  (reference gp_class.py:693-726, same schema) """
        gp_dict = {}
        gp_dict['X'] = self.__X.tolist()
        gp_dict['Y'] = self.__Y.tolist()
        gp_dict['hyper'] = dict(
            hyper=self.__hyper.tolist(),
            invK=self.get_invK().tolist(),
            alpha=self.get_alpha().tolist(),
            chol=self.get_chol().tolist(),
            length_scale=self.__hyper_length_scales.tolist(),
            signal_var=self.__hyper_signal_variance.tolist(),
            noise_var=self.__hyper_noise_variance.tolist(),
            mean=self.__hyper_mean.tolist())
        gp_dict['mean_func'] = self.__mean_func
        gp_dict['normalize'] = self.__normalize
        if self.__normalize:
            gp_dict['xlb'] = np.asarray(self.__xlb).tolist()
            gp_dict['xub'] = np.asarray(self.__xub).tolist()
            gp_dict['ulb'] = np.asarray(self.__ulb).tolist()
            gp_dict['uub'] = np.asarray(self.__uub).tolist()
            gp_dict['meta'] = dict(
                meanY=self.__meanY.tolist(), stdY=self.__stdY.tolist(),
                meanZ=self.__meanZ.tolist(), stdZ=self.__stdZ.tolist(),
                meanX=self.__meanX.tolist(), stdX=self.__stdX.tolist(),
                meanU=self.__meanU.tolist(), stdU=self.__stdU.tolist())
        return gp_dict

    def save_model(self, filename):
        """ Save model to a json file  (reference gp_class.py:729-734) """
        output_dict = self._GP__to_dict()
        with open(filename + ".json", "w") as outfile:
            json.dump(output_dict, outfile)

    def save_model_npz(self, filename):
        """ Binary side-car of save_model for large N (a 16384^2 factor is ~5 GB as JSON text):
        same fields, one compressed .npz; chol / invK are NOT stored (they are recomputed on the
        GPU at load time, as load_model does anyway). """
        d = dict(X=self.__X, Y=self.__Y, hyper=self.__hyper, mean_func=np.array(self.__mean_func),
                 normalize=np.array(bool(self.__normalize)))
        if self.__normalize:
            d.update(xlb=np.asarray(self.__xlb, dtype=np.float64), xub=np.asarray(self.__xub, dtype=np.float64),
                     ulb=np.asarray(self.__ulb, dtype=np.float64), uub=np.asarray(self.__uub, dtype=np.float64),
                     meanY=self.__meanY, stdY=self.__stdY, meanZ=self.__meanZ, stdZ=self.__stdZ,
                     meanX=self.__meanX, stdX=self.__stdX, meanU=self.__meanU, stdU=self.__stdU)
        np.savez_compressed(filename + '.npz', **d)

    @classmethod
    def load_model_npz(cls, filename, **kwargs):
        z = np.load(filename + '.npz')
        kw = dict(X=z['X'], Y=z['Y'], hyper=dict(hyper=z['hyper']), mean_func=str(z['mean_func']),
                  normalize=bool(z['normalize']))
        if kw['normalize']:
            kw.update(xlb=z['xlb'], xub=z['xub'], ulb=z['ulb'], uub=z['uub'],
                      meta={k: z[k] for k in ('meanY', 'stdY', 'meanZ', 'stdZ', 'meanX', 'stdX', 'meanU', 'stdU')})
        kw.update(kwargs)
        return cls(**kw)

    @classmethod
    def load_model(cls, filename, **kwargs):
        """ Create a new model from file  (reference gp_class.py:737-743) """
        with open(filename + ".json") as json_data:
            input_dict = json.load(json_data)
        input_dict.update(kwargs)
        return cls(**input_dict)

    def close(self):
        if self.__engine is not None:
            self.__engine.close()
            self.__engine = None
