// Per-NX entry points: each ipoc_nx.cu translation unit (compiled with -DIPOC_NX=<n>) defines the
// specialisation for its NX; ipoc_api.cu dispatches on the runtime nx.  nu is dispatched inside.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace ipoc {

// ---- optional side jobs of the fused entry points (ipoc_newton_attempt_f64, ipoc_costates_f64) ----------
// Every pointer may be NULL.  `handled` reports which of them the scan kernels took care of themselves
// (hierarchical / one-warp-per-sequence plans); the C ABI wrapper runs the stand-alone kernels for the rest.
enum { IPOC_X_HU = 1, IPOC_X_TRIAL = 2, IPOC_X_CONS = 4, IPOC_X_ACCEPT = 8, IPOC_X_NORM = 16 };
struct AttemptExtras {
    const double* reg_scale;   // reg_eff = reg * reg_scale            (rp * ||cu||, ref :117)
    double* hu;                // max|ru|                               (ref :158)
    const double *x, *u;       // trial point tx = x + dx, tu = u + du  (ref :156-157)
    double *tx, *tu;
    const double* cons;        // given constraint array (batch, N, nc) -> traj_feasible (ref :45-47)
    int nc;
    int32_t* traj_feasible;
    const double *cost, *new_cost;   // accept / regularisation update (ref :159-173); rp == NULL = off
    const int32_t *traj_feas_in, *active;
    double *rp, *r_inc;
    int32_t* success;
    double* gain;
    int handled;
};

// Instantiated (nx, nu) pairs: nu <= min(nx, 4).  X(nx, nu)
#define IPOC_FOR_PAIRS(X)                                                                              \
    X(1, 1) X(2, 1) X(2, 2) X(3, 1) X(3, 2) X(3, 3) X(4, 1) X(4, 2) X(4, 3) X(4, 4) X(5, 1) X(5, 2) X(5, 3) X(5, 4) \
    X(6, 1) X(6, 2) X(6, 3) X(6, 4) X(7, 1) X(7, 2) X(7, 3) X(7, 4) X(8, 1) X(8, 2) X(8, 3) X(8, 4)
#define IPOC_FOR_NX(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8)

// nu-independent entry points, one specialisation per NX (defined in the NU = 1 unit of that NX)
template <int NX> size_t nx_ws_bytes(int kind, int N, int batch, bool sharded);
template <int NX>
int nx_affine_scan(int reverse, int transpose, int N, int batch, const double* F, const double* c, const double* seed,
                   double* out, void* ws, size_t ws_bytes, cudaStream_t st, const double* sq_src, int sq_width,
                   double* cu_norm, int* handled, const int32_t* fresh);
template <int NX>
int nx_affine_reduce(int reverse, int transpose, int N, const double* F, const double* c, double* carry_out, void* ws,
                     size_t ws_bytes, cudaStream_t st);
template <int NX>
int nx_affine_apply(int reverse, int transpose, int N, int rank, int nranks, const double* F, const double* c,
                    const double* carries, const double* seed, double* out, void* ws, size_t ws_bytes, cudaStream_t st);

// per-(NX, NU) entry points, one specialisation per translation unit
template <int NX, int NU>
int nxu_newton_step(int N, int batch, const double* fx, const double* fu, const double* ru, const double* Q,
                    const double* R, const double* M, const double* reg, double* dx, double* du, double* Kx, double* d,
                    double* pred, int32_t* feasible, void* ws, size_t ws_bytes, cudaStream_t st, AttemptExtras* xtra);
template <int NX, int NU>
int nxu_lqt_bwd(int N, int batch, const double* A, const double* B, const double* c, const double* X, const double* U,
                const double* M, const double* q, const double* p, const double* ST, const double* vT, double* Kx,
                double* d, double* S, double* v, double* pred, int32_t* feasible, void* ws, size_t ws_bytes,
                cudaStream_t st);
template <int NX, int NU>
int nxu_lqt_fwd(int N, int batch, const double* A, const double* B, const double* c, const double* Kx, const double* d,
                const double* x0, double* u, double* x, void* ws, size_t ws_bytes, cudaStream_t st);
template <int NX, int NU>
int nxu_newton_bwd_reduce(int N, const double* fx, const double* fu, const double* ru, const double* Q, const double* R,
                          const double* M, const double* reg, double* carry_out, void* ws, size_t ws_bytes,
                          cudaStream_t st);
template <int NX, int NU>
int nxu_newton_bwd_apply(int N, int rank, int nranks, const double* fx, const double* fu, const double* ru,
                         const double* Q, const double* R, const double* M, const double* reg, const double* carries,
                         const double* ST, double* Kx, double* d, double* pred, int32_t* feasible, double* fwd_carry_out,
                         void* ws, size_t ws_bytes, cudaStream_t st);
template <int NX, int NU>
int nxu_newton_fwd_apply(int N, int rank, int nranks, const double* fx, const double* fu, const double* Kx,
                         const double* d, const double* fwd_carries, double* dx, double* du, void* ws, size_t ws_bytes,
                         cudaStream_t st);

}  // namespace ipoc
