// One translation unit per state dimension: nvcc -DIPOC_NX=<n> -c ipoc_nx.cu -o ipoc_nx<n>.o
// (parallel build; the NX = 8 instantiation alone takes about a minute).
#ifndef IPOC_NX
#error "compile with -DIPOC_NX=<state dimension>"
#endif
#include "ipoc_impl.cuh"

namespace ipoc {

// nu values instantiated for this NX (every extra pair costs compile time: the leaf kernels are
// instantiated per (NX, NU) and per loader)
#if IPOC_NX == 1
#define IPOC_FOR_NU(X) X(1)
#elif IPOC_NX == 6
#define IPOC_FOR_NU(X) X(1) X(2) X(3)
#else
#define IPOC_FOR_NU(X) X(1) X(2)
#endif
constexpr int NXc = IPOC_NX;

template <> int nx_supported<NXc>(int nu) {
#define X(b) if (nu == b) return 1;
    IPOC_FOR_NU(X)
#undef X
    return 0;
}
template <> size_t nx_ws_bytes<NXc>(int kind, int N, int batch, bool sharded) {
    return ws_bytes_impl<NXc>(kind, N, batch, sharded);
}
template <>
int nx_newton_step<NXc>(int nu, int N, int batch, const double* fx, const double* fu, const double* ru,
                        const double* Q, const double* R, const double* M, const double* reg, double* dx, double* du,
                        double* Kx, double* d, double* pred, int32_t* feasible, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
#define X(b) if (nu == b) return newton_step_impl<NXc, b>(N, batch, fx, fu, ru, Q, R, M, reg, dx, du, Kx, d, pred, feasible, ws, ws_bytes, st);
    IPOC_FOR_NU(X)
#undef X
    return IPOC_EUNSUPPORTED_DIM;
}
template <>
int nx_lqt_bwd<NXc>(int nu, int N, int batch, const double* A, const double* B, const double* c, const double* Xm,
                    const double* U, const double* M, const double* q, const double* p, const double* ST,
                    const double* vT, double* Kx, double* d, double* S, double* v, double* pred, int32_t* feasible,
                    void* ws, size_t ws_bytes, cudaStream_t st) {
#define X(b) if (nu == b) return lqt_bwd_impl<NXc, b>(N, batch, A, B, c, Xm, U, M, q, p, ST, vT, Kx, d, S, v, pred, feasible, ws, ws_bytes, st);
    IPOC_FOR_NU(X)
#undef X
    return IPOC_EUNSUPPORTED_DIM;
}
template <>
int nx_lqt_fwd<NXc>(int nu, int N, int batch, const double* A, const double* B, const double* c, const double* Kx,
                    const double* d, const double* x0, double* u, double* x, void* ws, size_t ws_bytes,
                    cudaStream_t st) {
#define X(b) if (nu == b) return lqt_fwd_impl<NXc, b>(N, batch, A, B, c, Kx, d, x0, u, x, ws, ws_bytes, st);
    IPOC_FOR_NU(X)
#undef X
    return IPOC_EUNSUPPORTED_DIM;
}
template <>
int nx_affine_scan<NXc>(int reverse, int transpose, int N, int batch, const double* F, const double* c,
                        const double* seed, double* out, void* ws, size_t ws_bytes, cudaStream_t st) {
    return affine_scan_impl<NXc>(reverse, transpose, N, batch, F, c, seed, out, ws, ws_bytes, st);
}
template <>
int nx_newton_bwd_reduce<NXc>(int nu, int N, const double* fx, const double* fu, const double* ru, const double* Q,
                              const double* R, const double* M, const double* reg, double* carry_out, void* ws,
                              size_t ws_bytes, cudaStream_t st) {
#define X(b) if (nu == b) return newton_bwd_reduce_impl<NXc, b>(N, fx, fu, ru, Q, R, M, reg, carry_out, ws, ws_bytes, st);
    IPOC_FOR_NU(X)
#undef X
    return IPOC_EUNSUPPORTED_DIM;
}
template <>
int nx_newton_bwd_apply<NXc>(int nu, int N, int rank, int nranks, const double* fx, const double* fu,
                             const double* ru, const double* Q, const double* R, const double* M, const double* reg,
                             const double* carries, const double* ST, double* Kx, double* d, double* pred,
                             int32_t* feasible, double* fwd_carry_out, void* ws, size_t ws_bytes, cudaStream_t st) {
#define X(b) if (nu == b) return newton_bwd_apply_impl<NXc, b>(N, rank, nranks, fx, fu, ru, Q, R, M, reg, carries, ST, Kx, d, pred, feasible, fwd_carry_out, ws, ws_bytes, st);
    IPOC_FOR_NU(X)
#undef X
    return IPOC_EUNSUPPORTED_DIM;
}
template <>
int nx_newton_fwd_apply<NXc>(int nu, int N, int rank, int nranks, const double* fx, const double* fu,
                             const double* Kx, const double* d, const double* fwd_carries, double* dx, double* du,
                             void* ws, size_t ws_bytes, cudaStream_t st) {
#define X(b) if (nu == b) return newton_fwd_apply_impl<NXc, b>(N, rank, nranks, fx, fu, Kx, d, fwd_carries, dx, du, ws, ws_bytes, st);
    IPOC_FOR_NU(X)
#undef X
    return IPOC_EUNSUPPORTED_DIM;
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

}  // namespace ipoc
