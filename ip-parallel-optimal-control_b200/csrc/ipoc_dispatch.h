// Per-NX entry points: each ipoc_nx.cu translation unit (compiled with -DIPOC_NX=<n>) defines the
// specialisation for its NX; ipoc_api.cu dispatches on the runtime nx.  nu is dispatched inside.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace ipoc {

template <int NX> int nx_supported(int nu);
template <int NX> size_t nx_ws_bytes(int kind, int N, int batch, bool sharded);
template <int NX>
int nx_newton_step(int nu, int N, int batch, const double* fx, const double* fu, const double* ru, const double* Q,
                   const double* R, const double* M, const double* reg, double* dx, double* du, double* Kx, double* d,
                   double* pred, int32_t* feasible, void* ws, size_t ws_bytes, cudaStream_t st);
template <int NX>
int nx_lqt_bwd(int nu, int N, int batch, const double* A, const double* B, const double* c, const double* X,
               const double* U, const double* M, const double* q, const double* p, const double* ST, const double* vT,
               double* Kx, double* d, double* S, double* v, double* pred, int32_t* feasible, void* ws, size_t ws_bytes,
               cudaStream_t st);
template <int NX>
int nx_lqt_fwd(int nu, int N, int batch, const double* A, const double* B, const double* c, const double* Kx,
               const double* d, const double* x0, double* u, double* x, void* ws, size_t ws_bytes, cudaStream_t st);
template <int NX>
int nx_affine_scan(int reverse, int transpose, int N, int batch, const double* F, const double* c, const double* seed,
                   double* out, void* ws, size_t ws_bytes, cudaStream_t st);
template <int NX>
int nx_newton_bwd_reduce(int nu, int N, const double* fx, const double* fu, const double* ru, const double* Q,
                         const double* R, const double* M, const double* reg, double* carry_out, void* ws,
                         size_t ws_bytes, cudaStream_t st);
template <int NX>
int nx_newton_bwd_apply(int nu, int N, int rank, int nranks, const double* fx, const double* fu, const double* ru,
                        const double* Q, const double* R, const double* M, const double* reg, const double* carries,
                        const double* ST, double* Kx, double* d, double* pred, int32_t* feasible, double* fwd_carry_out,
                        void* ws, size_t ws_bytes, cudaStream_t st);
template <int NX>
int nx_newton_fwd_apply(int nu, int N, int rank, int nranks, const double* fx, const double* fu, const double* Kx,
                        const double* d, const double* fwd_carries, double* dx, double* du, void* ws, size_t ws_bytes,
                        cudaStream_t st);
template <int NX>
int nx_affine_reduce(int reverse, int transpose, int N, const double* F, const double* c, double* carry_out, void* ws,
                     size_t ws_bytes, cudaStream_t st);
template <int NX>
int nx_affine_apply(int reverse, int transpose, int N, int rank, int nranks, const double* F, const double* c,
                    const double* carries, const double* seed, double* out, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace ipoc
