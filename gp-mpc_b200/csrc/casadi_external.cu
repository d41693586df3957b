// CasADi `external`-shaped entry points around gpmpc_predict_grad.
//
// mpc_class.py:361-423 calls gp.predict(mean_t, u_t, covar_t) once per shooting node with MX
// symbols and nlpsol (:496-513) differentiates the resulting graph.  With
//     F = casadi.external('gp_b200', 'libgpmpc.so')
// the GP becomes ONE opaque call for all Nt nodes per NLP iterate, evaluated on the GPU, with the
// analytic Jacobian function `jac_gp_b200` CasADi looks up by name.  The signatures follow
// CasADi's C API for external functions (casadi_int = long long, casadi_real = double; dense
// matrices are column-major, sparsity patterns are compressed-column: {nrow, ncol, colind[ncol+1],
// row[nnz]}, with the 3-entry form {nrow, ncol, 1} meaning dense).  Nothing here needs CasADi to
// compile or to be tested: tests drive these entry points through ctypes.
//
//   inputs : i0 = Z     (Nx  x Nt)     test inputs of all shooting nodes, standardised space
//            i1 = Sigma (Nx  x Nx*Nt)  input covariance of every node (Nt blocks side by side)
//   outputs: o0 = mean  (Ny  x Nt)
//            o1 = cov   (Ny  x Ny*Nt)  'TA' / 'ME' covariance blocks
//   jac_gp_b200: inputs (i0, i1, o0, o1), outputs block-diagonal sparse
//            jac_o0_i0 (Ny*Nt x Nx*Nt), jac_o0_i1 (empty), jac_o1_i0 (Ny*Ny*Nt x Nx*Nt),
//            jac_o1_i1 (Ny*Ny*Nt x Nx*Nx*Nt;  d cov[a][b] / d Sigma[d][e] = J_a[d] J_b[e] for 'TA')
#include "../../include/gpmpc.h"

#include <mutex>
#include <vector>

typedef long long casadi_int;
typedef double casadi_real;

namespace {
struct Bound {
    gpmpc_handle_t h = nullptr;
    int method = GPMPC_METHOD_TA, Nt = 0, Nx = 0, Ny = 0;
    std::vector<casadi_int> sp_in[2], sp_out[2], sp_jac[4];
    std::vector<double> sig, mean, var, cov, jac, dvar, dcov;
    int refs = 0;
};
Bound g_b;
std::mutex g_mtx;

std::vector<casadi_int> dense_sp(casadi_int r, casadi_int c) { return {r, c, 1}; }

// block-diagonal CCS pattern: Nt dense blocks of R rows x Cc columns
std::vector<casadi_int> blockdiag_sp(casadi_int R, casadi_int Cc, casadi_int Nt)
{
    std::vector<casadi_int> sp;
    sp.reserve(2 + Cc * Nt + 1 + R * Cc * Nt);
    sp.push_back(R * Nt); sp.push_back(Cc * Nt);
    for (casadi_int c = 0; c <= Cc * Nt; ++c) sp.push_back(c * R);
    for (casadi_int t = 0; t < Nt; ++t)
        for (casadi_int c = 0; c < Cc; ++c)
            for (casadi_int r = 0; r < R; ++r) sp.push_back(t * R + r);
    return sp;
}
std::vector<casadi_int> empty_sp(casadi_int r, casadi_int c)
{
    std::vector<casadi_int> sp(2 + c + 1, 0);
    sp[0] = r; sp[1] = c;
    return sp;
}

// evaluate mean/cov (+ derivatives when grad) for the bound handle; Sigma blocks are transposed to
// the engine's row-major convention (a symmetric Sigma is unchanged)
int eval(const casadi_real* Z, const casadi_real* Sigma, bool grad)
{
    Bound& b = g_b;
    if (!b.h || !Z) return 1;
    const int Nt = b.Nt, Nx = b.Nx;
    const bool ta = b.method == GPMPC_METHOD_TA;
    if (ta) {
        if (!Sigma) return 1;
        for (int t = 0; t < Nt; ++t)
            for (int d = 0; d < Nx; ++d)
                for (int e = 0; e < Nx; ++e)
                    b.sig[((size_t)t * Nx + d) * Nx + e] = Sigma[((size_t)t * Nx + e) * Nx + d];
    }
    if (!grad)
        return gpmpc_predict(b.h, b.method, Nt, Z, ta ? b.sig.data() : nullptr, 1, b.mean.data(), b.var.data(),
                             b.cov.data(), b.jac.data()) == GPMPC_OK ? 0 : 1;
    return gpmpc_predict_grad(b.h, b.method, Nt, Z, ta ? b.sig.data() : nullptr, 1, b.mean.data(), b.var.data(),
                              b.cov.data(), b.jac.data(), b.dvar.data(), b.dcov.data(), nullptr) == GPMPC_OK ? 0 : 1;
}
}  // namespace

// Bind the (process-global) external to a factorised engine handle: method GPMPC_METHOD_ME / _TA,
// Nt shooting nodes per call.  Call again to re-bind (e.g. after a refit or another horizon).
extern "C" int gp_b200_bind(gpmpc_handle_t h, int method, int Nt)
{
    std::lock_guard<std::mutex> lock(g_mtx);
    int N = 0, Nx = 0, Ny = 0;
    if (!h || Nt < 1 || (method != GPMPC_METHOD_ME && method != GPMPC_METHOD_TA)) return GPMPC_ERR_ARG;
    if (gpmpc_get_size(h, &N, &Nx, &Ny) != GPMPC_OK) return GPMPC_ERR_ARG;
    Bound& b = g_b;
    b.h = h; b.method = method; b.Nt = Nt; b.Nx = Nx; b.Ny = Ny;
    b.sp_in[0] = dense_sp(Nx, Nt); b.sp_in[1] = dense_sp(Nx, (casadi_int)Nx * Nt);
    b.sp_out[0] = dense_sp(Ny, Nt); b.sp_out[1] = dense_sp(Ny, (casadi_int)Ny * Nt);
    b.sp_jac[0] = blockdiag_sp(Ny, Nx, Nt);
    b.sp_jac[1] = empty_sp((casadi_int)Ny * Nt, (casadi_int)Nx * Nx * Nt);
    b.sp_jac[2] = blockdiag_sp((casadi_int)Ny * Ny, Nx, Nt);
    b.sp_jac[3] = (method == GPMPC_METHOD_TA) ? blockdiag_sp((casadi_int)Ny * Ny, (casadi_int)Nx * Nx, Nt)
                                              : empty_sp((casadi_int)Ny * Ny * Nt, (casadi_int)Nx * Nx * Nt);
    b.sig.assign((size_t)Nt * Nx * Nx, 0.0);
    b.mean.assign((size_t)Nt * Ny, 0.0); b.var.assign((size_t)Nt * Ny, 0.0);
    b.cov.assign((size_t)Nt * Ny * Ny, 0.0); b.jac.assign((size_t)Nt * Ny * Nx, 0.0);
    b.dvar.assign((size_t)Nt * Ny * Nx, 0.0); b.dcov.assign((size_t)Nt * Ny * Ny * Nx, 0.0);
    return GPMPC_OK;
}

extern "C" void gp_b200_unbind(void)
{
    std::lock_guard<std::mutex> lock(g_mtx);
    g_b = Bound();
}

// ---- the function itself
extern "C" casadi_int gp_b200_n_in(void) { return 2; }
extern "C" casadi_int gp_b200_n_out(void) { return 2; }
extern "C" const char* gp_b200_name_in(casadi_int i) { return i == 0 ? "z" : (i == 1 ? "sigma" : nullptr); }
extern "C" const char* gp_b200_name_out(casadi_int i) { return i == 0 ? "mean" : (i == 1 ? "cov" : nullptr); }
extern "C" const casadi_int* gp_b200_sparsity_in(casadi_int i) { return (i >= 0 && i < 2 && g_b.h) ? g_b.sp_in[i].data() : nullptr; }
extern "C" const casadi_int* gp_b200_sparsity_out(casadi_int i) { return (i >= 0 && i < 2 && g_b.h) ? g_b.sp_out[i].data() : nullptr; }
extern "C" int gp_b200_work(casadi_int* sz_arg, casadi_int* sz_res, casadi_int* sz_iw, casadi_int* sz_w)
{
    if (sz_arg) *sz_arg = 2;
    if (sz_res) *sz_res = 2;
    if (sz_iw) *sz_iw = 0;
    if (sz_w) *sz_w = 0;
    return 0;
}
extern "C" void gp_b200_incref(void) { std::lock_guard<std::mutex> lock(g_mtx); ++g_b.refs; }
extern "C" void gp_b200_decref(void) { std::lock_guard<std::mutex> lock(g_mtx); --g_b.refs; }

extern "C" int gp_b200(const casadi_real** arg, casadi_real** res, casadi_int* iw, casadi_real* w, int mem)
{
    (void)iw; (void)w; (void)mem;
    std::lock_guard<std::mutex> lock(g_mtx);
    if (!arg || !res) return 1;
    if (eval(arg[0], arg[1], false)) return 1;
    const Bound& b = g_b;
    // (Nt,Ny) row-major == Ny x Nt column-major; each Ny x Ny block is written column-major
    // (element (a,b) at a + Ny*b) -- the blocks are symmetric only up to rounding
    if (res[0]) std::copy(b.mean.begin(), b.mean.end(), res[0]);
    if (res[1])
        for (int t = 0; t < b.Nt; ++t)
            for (int bb = 0; bb < b.Ny; ++bb)
                for (int a = 0; a < b.Ny; ++a)
                    res[1][((size_t)t * b.Ny + bb) * b.Ny + a] = b.cov[((size_t)t * b.Ny + a) * b.Ny + bb];
    return 0;
}

// ---- its Jacobian: inputs (z, sigma, mean, cov), outputs the four blocks in CCS nonzero order
extern "C" casadi_int jac_gp_b200_n_in(void) { return 4; }
extern "C" casadi_int jac_gp_b200_n_out(void) { return 4; }
extern "C" const char* jac_gp_b200_name_in(casadi_int i)
{
    static const char* n[] = {"z", "sigma", "out_mean", "out_cov"};
    return (i >= 0 && i < 4) ? n[i] : nullptr;
}
extern "C" const char* jac_gp_b200_name_out(casadi_int i)
{
    static const char* n[] = {"jac_mean_z", "jac_mean_sigma", "jac_cov_z", "jac_cov_sigma"};
    return (i >= 0 && i < 4) ? n[i] : nullptr;
}
extern "C" const casadi_int* jac_gp_b200_sparsity_in(casadi_int i)
{
    if (!g_b.h || i < 0 || i > 3) return nullptr;
    return i < 2 ? g_b.sp_in[i].data() : g_b.sp_out[i - 2].data();
}
extern "C" const casadi_int* jac_gp_b200_sparsity_out(casadi_int i) { return (i >= 0 && i < 4 && g_b.h) ? g_b.sp_jac[i].data() : nullptr; }
extern "C" int jac_gp_b200_work(casadi_int* sz_arg, casadi_int* sz_res, casadi_int* sz_iw, casadi_int* sz_w)
{
    if (sz_arg) *sz_arg = 4;
    if (sz_res) *sz_res = 4;
    if (sz_iw) *sz_iw = 0;
    if (sz_w) *sz_w = 0;
    return 0;
}

extern "C" int jac_gp_b200(const casadi_real** arg, casadi_real** res, casadi_int* iw, casadi_real* w, int mem)
{
    (void)iw; (void)w; (void)mem;
    std::lock_guard<std::mutex> lock(g_mtx);
    if (!arg || !res) return 1;
    if (eval(arg[0], arg[1], true)) return 1;
    const Bound& b = g_b;
    const int Nt = b.Nt, Nx = b.Nx, Ny = b.Ny;
    if (res[0])          // block t, column d, row a:  d mean_a / d z_d
        for (int t = 0; t < Nt; ++t)
            for (int d = 0; d < Nx; ++d)
                for (int a = 0; a < Ny; ++a) res[0][((size_t)t * Nx + d) * Ny + a] = b.jac[((size_t)t * Ny + a) * Nx + d];
    if (res[2])          // block t, column e, row a + Ny*b (column-major vec of the Ny x Ny block)
        for (int t = 0; t < Nt; ++t)
            for (int e = 0; e < Nx; ++e)
                for (int bb = 0; bb < Ny; ++bb)
                    for (int a = 0; a < Ny; ++a)
                        res[2][(((size_t)t * Nx + e) * Ny + bb) * Ny + a] = b.dcov[(((size_t)t * Ny + a) * Ny + bb) * Nx + e];
    if (res[3] && b.method == GPMPC_METHOD_TA)   // column d + Nx*e (vec of Sigma), row a + Ny*b:  J_a[d] J_b[e]
        for (int t = 0; t < Nt; ++t)
            for (int e = 0; e < Nx; ++e)
                for (int d = 0; d < Nx; ++d)
                    for (int bb = 0; bb < Ny; ++bb)
                        for (int a = 0; a < Ny; ++a)
                            res[3][((((size_t)t * Nx + e) * Nx + d) * Ny + bb) * Ny + a] =
                                b.jac[((size_t)t * Ny + a) * Nx + d] * b.jac[((size_t)t * Ny + bb) * Nx + e];
    return 0;
}
