// One translation unit per (state dimension, control dimension):
//   nvcc -DIPOC_NX=<nx> -DIPOC_NU=<nu> -c ipoc_nx.cu -o ipoc_nx<nx>_nu<nu>.o
// (parallel build; the NX = 8 instantiations take minutes each).  The nu-independent entry points of an NX
// (affine scans, workspace sizes) live in its NU = 1 unit.  Instantiated pairs: nu <= min(nx, 4)
// (ipoc_dispatch.h: IPOC_FOR_PAIRS).
#if !defined(IPOC_NX) || !defined(IPOC_NU)
#error "compile with -DIPOC_NX=<state dimension> -DIPOC_NU=<control dimension>"
#endif
#include "ipoc_impl.cuh"

namespace ipoc {

constexpr int NXc = IPOC_NX;
constexpr int NUc = IPOC_NU;

template <>
int nxu_newton_step<NXc, NUc>(int N, int batch, const double* fx, const double* fu, const double* ru, const double* Q,
                              const double* R, const double* M, const double* reg, double* dx, double* du, double* Kx,
                              double* d, double* pred, int32_t* feasible, void* ws, size_t ws_bytes, cudaStream_t st,
                              AttemptExtras* xtra) {
    return newton_step_impl<NXc, NUc>(N, batch, fx, fu, ru, Q, R, M, reg, dx, du, Kx, d, pred, feasible, ws, ws_bytes, st,
                                      xtra);
}
template <>
int nxu_lqt_bwd<NXc, NUc>(int N, int batch, const double* A, const double* B, const double* c, const double* Xm,
                          const double* U, const double* M, const double* q, const double* p, const double* ST,
                          const double* vT, double* Kx, double* d, double* S, double* v, double* pred,
                          int32_t* feasible, void* ws, size_t ws_bytes, cudaStream_t st) {
    return lqt_bwd_impl<NXc, NUc>(N, batch, A, B, c, Xm, U, M, q, p, ST, vT, Kx, d, S, v, pred, feasible, ws, ws_bytes, st);
}
template <>
int nxu_lqt_fwd<NXc, NUc>(int N, int batch, const double* A, const double* B, const double* c, const double* Kx,
                          const double* d, const double* x0, double* u, double* x, void* ws, size_t ws_bytes,
                          cudaStream_t st) {
    return lqt_fwd_impl<NXc, NUc>(N, batch, A, B, c, Kx, d, x0, u, x, ws, ws_bytes, st);
}
template <>
int nxu_newton_bwd_reduce<NXc, NUc>(int N, const double* fx, const double* fu, const double* ru, const double* Q,
                                    const double* R, const double* M, const double* reg, double* carry_out, void* ws,
                                    size_t ws_bytes, cudaStream_t st) {
    return newton_bwd_reduce_impl<NXc, NUc>(N, fx, fu, ru, Q, R, M, reg, carry_out, ws, ws_bytes, st);
}
template <>
int nxu_newton_bwd_apply<NXc, NUc>(int N, int rank, int nranks, const double* fx, const double* fu, const double* ru,
                                   const double* Q, const double* R, const double* M, const double* reg,
                                   const double* carries, const double* ST, double* Kx, double* d, double* pred,
                                   int32_t* feasible, double* fwd_carry_out, void* ws, size_t ws_bytes,
                                   cudaStream_t st) {
    return newton_bwd_apply_impl<NXc, NUc>(N, rank, nranks, fx, fu, ru, Q, R, M, reg, carries, ST, Kx, d, pred, feasible,
                                           fwd_carry_out, ws, ws_bytes, st);
}
template <>
int nxu_newton_fwd_apply<NXc, NUc>(int N, int rank, int nranks, const double* fx, const double* fu, const double* Kx,
                                   const double* d, const double* fwd_carries, double* dx, double* du, void* ws,
                                   size_t ws_bytes, cudaStream_t st) {
    return newton_fwd_apply_impl<NXc, NUc>(N, rank, nranks, fx, fu, Kx, d, fwd_carries, dx, du, ws, ws_bytes, st);
}

#if IPOC_NU == 1
template <> size_t nx_ws_bytes<NXc>(int kind, int N, int batch, bool sharded) {
    return ws_bytes_impl<NXc>(kind, N, batch, sharded);
}
template <>
int nx_affine_scan<NXc>(int reverse, int transpose, int N, int batch, const double* F, const double* c,
                        const double* seed, double* out, void* ws, size_t ws_bytes, cudaStream_t st,
                        const double* sq_src, int sq_width, double* cu_norm, int* handled, const int32_t* fresh) {
    return affine_scan_impl<NXc>(reverse, transpose, N, batch, F, c, seed, out, ws, ws_bytes, st, sq_src, sq_width, cu_norm,
                                 handled, fresh);
}
template <>
int nx_affine_reduce<NXc>(int reverse, int transpose, int N, const double* F, const double* c, double* carry_out,
                          void* ws, size_t ws_bytes, cudaStream_t st) {
    return affine_reduce_impl<NXc>(reverse, transpose, N, F, c, carry_out, ws, ws_bytes, st);
}
template <>
int nx_affine_apply<NXc>(int reverse, int transpose, int N, int rank, int nranks, const double* F, const double* c,
                         const double* carries, const double* seed, double* out, void* ws, size_t ws_bytes,
                         cudaStream_t st) {
    return affine_apply_impl<NXc>(reverse, transpose, N, rank, nranks, F, c, carries, seed, out, ws, ws_bytes, st);
}
#endif

}  // namespace ipoc
