// XLA FFI custom-call shim for JAX front-ends: thin handlers that forward buffers + the XLA stream to
// the C ABI of libipoc.so (include/ipoc.h).  Registered for platform "CUDA" only (no CPU fallback).
//
// NOT BUILT IN THIS IMAGE: the XLA FFI headers (xla/ffi/api/ffi.h, shipped inside jaxlib) and JAX itself
// are absent, so this translation unit compiles to nothing unless the header is found:
//   g++ -std=c++17 -fPIC -shared -I$(python -c "import jaxlib,os;print(os.path.join(os.path.dirname(jaxlib.__file__),'include'))") \
//       -Iinclude ipoc_xla_ffi.cc -L. -lipoc -o libipoc_xla.so
// The Python side is ipoc_b200/jax_ffi.py.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define IPOC_HAVE_XLA_FFI 1
#endif
#endif

#ifdef IPOC_HAVE_XLA_FFI
#include <cuda_runtime.h>
#include "xla/ffi/api/ffi.h"
#include "ipoc.h"

namespace ffi = xla::ffi;

static ffi::Error check(int rc) {
    if (rc == IPOC_OK) return ffi::Error::Success();
    return ffi::Error(ffi::ErrorCode::kInternal, ipoc_strerror(rc));
}
// XLA hands every call a fresh, uninitialised result buffer as workspace: its control block (arrival counters
// of the in-kernel scan levels, include/ipoc.h) is zeroed on the call's stream first.
#define IPOC_WS_INIT(ws)                                                                        \
    do {                                                                                        \
        if (int rc_ = ipoc_workspace_init((ws)->untyped_data(), (ws)->size_bytes(), stream)) return check(rc_); \
    } while (0)

// newton_step: (fx, fu, ru, Q, R, M, reg) -> (dx, du, Kx, d, pred, feasible, workspace)
// shapes: fx (B,N,nx,nx) ...; workspace is an extra result buffer of ipoc_workspace_bytes() bytes (uint8).
static ffi::Error NewtonStepImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> fx, ffi::Buffer<ffi::F64> fu,
                                 ffi::Buffer<ffi::F64> ru, ffi::Buffer<ffi::F64> Q, ffi::Buffer<ffi::F64> R,
                                 ffi::Buffer<ffi::F64> M, ffi::Buffer<ffi::F64> reg,
                                 ffi::ResultBuffer<ffi::F64> dx, ffi::ResultBuffer<ffi::F64> du,
                                 ffi::ResultBuffer<ffi::F64> Kx, ffi::ResultBuffer<ffi::F64> d,
                                 ffi::ResultBuffer<ffi::F64> pred, ffi::ResultBuffer<ffi::S32> feasible,
                                 ffi::ResultBuffer<ffi::U8> ws) {
    auto dims = fx.dimensions();
    if (dims.size() != 4) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "fx must be (batch, N, nx, nx)");
    const int batch = (int)dims[0], N = (int)dims[1], nx = (int)dims[2];
    const int nu = (int)fu.dimensions().back();
    IPOC_WS_INIT(ws);
    return check(ipoc_newton_step_f64(N, nx, nu, batch, fx.typed_data(), fu.typed_data(), ru.typed_data(),
                                      Q.typed_data(), R.typed_data(), M.typed_data(), reg.typed_data(),
                                      dx->typed_data(), du->typed_data(), Kx->typed_data(), d->typed_data(),
                                      pred->typed_data(), feasible->typed_data(), ws->untyped_data(),
                                      ws->size_bytes(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(IpocNewtonStep, NewtonStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()   // fx
                                  .Arg<ffi::Buffer<ffi::F64>>()   // fu
                                  .Arg<ffi::Buffer<ffi::F64>>()   // ru
                                  .Arg<ffi::Buffer<ffi::F64>>()   // Q
                                  .Arg<ffi::Buffer<ffi::F64>>()   // R
                                  .Arg<ffi::Buffer<ffi::F64>>()   // M
                                  .Arg<ffi::Buffer<ffi::F64>>()   // reg
                                  .Ret<ffi::Buffer<ffi::F64>>()   // dx
                                  .Ret<ffi::Buffer<ffi::F64>>()   // du
                                  .Ret<ffi::Buffer<ffi::F64>>()   // Kx
                                  .Ret<ffi::Buffer<ffi::F64>>()   // d
                                  .Ret<ffi::Buffer<ffi::F64>>()   // pred
                                  .Ret<ffi::Buffer<ffi::S32>>()   // feasible
                                  .Ret<ffi::Buffer<ffi::U8>>());  // workspace

// affine_scan (costates): (F, c, seed) + attrs(reverse, transpose) -> (out, workspace)
static ffi::Error AffineScanImpl(cudaStream_t stream, int32_t reverse, int32_t transpose, ffi::Buffer<ffi::F64> F,
                                 ffi::Buffer<ffi::F64> c, ffi::Buffer<ffi::F64> seed,
                                 ffi::ResultBuffer<ffi::F64> out, ffi::ResultBuffer<ffi::U8> ws) {
    auto dims = F.dimensions();
    if (dims.size() != 4) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "F must be (batch, N, nx, nx)");
    IPOC_WS_INIT(ws);
    return check(ipoc_affine_scan_f64(reverse, transpose, (int)dims[1], (int)dims[2], (int)dims[0], F.typed_data(),
                                      c.typed_data(), seed.typed_data(), out->typed_data(), ws->untyped_data(),
                                      ws->size_bytes(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(IpocAffineScan, AffineScanImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int32_t>("reverse")
                                  .Attr<int32_t>("transpose")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

using F64 = ffi::Buffer<ffi::F64>;
using RF64 = ffi::ResultBuffer<ffi::F64>;
using RS32 = ffi::ResultBuffer<ffi::S32>;
using RU8 = ffi::ResultBuffer<ffi::U8>;

// par_bwd_pass on effective LQT terms (ref call sites noc/par_interior_point_newton.py:120,
// examples/linear_mpc_parallel.py:68): (A, B, c, X, U, M, q, p, ST, vT) -> (Kx, d, S, v, pred, feasible, ws)
static ffi::Error LqtBwdImpl(cudaStream_t stream, F64 A, F64 B, F64 c, F64 X, F64 U, F64 M, F64 q, F64 p, F64 ST, F64 vT,
                             RF64 Kx, RF64 d, RF64 S, RF64 v, RF64 pred, RS32 feasible, RU8 ws) {
    auto dims = A.dimensions();
    if (dims.size() != 4) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "A must be (batch, N, nx, nx)");
    IPOC_WS_INIT(ws);
    return check(ipoc_lqt_bwd_f64((int)dims[1], (int)dims[2], (int)B.dimensions().back(), (int)dims[0], A.typed_data(),
                                  B.typed_data(), c.typed_data(), X.typed_data(), U.typed_data(), M.typed_data(),
                                  q.typed_data(), p.typed_data(), ST.typed_data(), vT.typed_data(), Kx->typed_data(),
                                  d->typed_data(), S->typed_data(), v->typed_data(), pred->typed_data(),
                                  feasible->typed_data(), ws->untyped_data(), ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(IpocLqtBwd, LqtBwdImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

// par_fwd_pass (ref :121-123, examples/linear_mpc_parallel.py:69): (A, B, c, Kx, d, x0) -> (u, x, ws)
static ffi::Error LqtFwdImpl(cudaStream_t stream, F64 A, F64 B, F64 c, F64 Kx, F64 d, F64 x0, RF64 u, RF64 x, RU8 ws) {
    auto dims = A.dimensions();
    if (dims.size() != 4) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "A must be (batch, N, nx, nx)");
    IPOC_WS_INIT(ws);
    return check(ipoc_lqt_fwd_f64((int)dims[1], (int)dims[2], (int)B.dimensions().back(), (int)dims[0], A.typed_data(),
                                  B.typed_data(), c.typed_data(), Kx.typed_data(), d.typed_data(), x0.typed_data(),
                                  u->typed_data(), x->typed_data(), ws->untyped_data(), ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(IpocLqtFwd, LqtFwdImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::U8>>());

// K4 (ref :45-47, :116, :158): (ru, cu, cons) -> (hu_norm, cu_norm, traj_feasible, ws)
static ffi::Error ReductionsImpl(cudaStream_t stream, F64 ru, F64 cu, F64 cons, RF64 hu, RF64 cn, RS32 feas, RU8 ws) {
    auto dims = ru.dimensions();
    if (dims.size() != 3) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "ru must be (batch, N, nu)");
    return check(ipoc_reductions_f64((int)dims[1], (int)dims[2], (int)cons.dimensions().back(), (int)dims[0],
                                     ru.typed_data(), cu.typed_data(), cons.typed_data(), hu->typed_data(),
                                     cn->typed_data(), feas->typed_data(), nullptr, nullptr, ws->untyped_data(),
                                     ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(IpocReductions, ReductionsImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::U8>>());

// A8 (ref :159-173), functional form for XLA: (cost, new_cost, traj_feasible, pred, bwd_feasible, rp, r_inc)
//   -> (rp', r_inc', success, gain_ratio)
static ffi::Error AcceptImpl(cudaStream_t stream, F64 cost, F64 new_cost, ffi::Buffer<ffi::S32> tf, F64 pred,
                             ffi::Buffer<ffi::S32> bf, F64 rp, F64 r_inc, RF64 rp_out, RF64 r_inc_out, RS32 success,
                             RF64 gain) {
    const int batch = (int)rp.element_count();
    if (cudaMemcpyAsync(rp_out->typed_data(), rp.typed_data(), sizeof(double) * batch, cudaMemcpyDeviceToDevice, stream) !=
            cudaSuccess ||
        cudaMemcpyAsync(r_inc_out->typed_data(), r_inc.typed_data(), sizeof(double) * batch, cudaMemcpyDeviceToDevice,
                        stream) != cudaSuccess)
        return check(IPOC_ECUDA);
    return check(ipoc_accept_update_f64(batch, cost.typed_data(), new_cost.typed_data(), tf.typed_data(),
                                        pred.typed_data(), bf.typed_data(), nullptr, rp_out->typed_data(),
                                        r_inc_out->typed_data(), success->typed_data(), gain->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(IpocAcceptUpdate, AcceptImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F64>().Arg<F64>().Arg<ffi::Buffer<ffi::S32>>().Arg<F64>().Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::S32>>().Ret<F64>());

// fused K1 + ||cu|| (ref noc/costates.py:34-40 + :116): (fx, cx, lamT, cu) -> (lam, cu_norm, ws)
static ffi::Error CostatesImpl(cudaStream_t stream, F64 fx, F64 cx, F64 lamT, F64 cu, RF64 lam, RF64 cn, RU8 ws) {
    auto dims = fx.dimensions();
    if (dims.size() != 4) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "fx must be (batch, N, nx, nx)");
    IPOC_WS_INIT(ws);
    return check(ipoc_costates_f64((int)dims[1], (int)dims[2], (int)cu.dimensions().back(), (int)dims[0], fx.typed_data(),
                                   cx.typed_data(), lamT.typed_data(), cu.typed_data(), lam->typed_data(),
                                   cn->typed_data(), ws->untyped_data(), ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(IpocCostates, CostatesImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::U8>>());

// fused attempt (ref :117, :153-158): (fx, fu, ru, Q, R, M, rp, cu_norm, x, u)
//   -> (dx, du, Kx, d, pred, feasible, hu, tx, tu, ws)
static ffi::Error AttemptImpl(cudaStream_t stream, F64 fx, F64 fu, F64 ru, F64 Q, F64 R, F64 M, F64 rp, F64 cn, F64 x,
                              F64 u, RF64 dx, RF64 du, RF64 Kx, RF64 d, RF64 pred, RS32 feasible, RF64 hu, RF64 tx,
                              RF64 tu, RU8 ws) {
    auto dims = fx.dimensions();
    if (dims.size() != 4) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "fx must be (batch, N, nx, nx)");
    IPOC_WS_INIT(ws);
    return check(ipoc_newton_attempt_f64(
        (int)dims[1], (int)dims[2], (int)fu.dimensions().back(), 1, (int)dims[0], fx.typed_data(), fu.typed_data(),
        ru.typed_data(), Q.typed_data(), R.typed_data(), M.typed_data(), rp.typed_data(), cn.typed_data(), dx->typed_data(),
        du->typed_data(), Kx->typed_data(), d->typed_data(), pred->typed_data(), feasible->typed_data(), hu->typed_data(),
        x.typed_data(), u.typed_data(), tx->typed_data(), tu->typed_data(), nullptr, nullptr, nullptr, nullptr, nullptr,
        nullptr, nullptr, nullptr, nullptr, nullptr, ws->untyped_data(), ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(IpocNewtonAttempt, AttemptImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::U8>>());
#endif  // IPOC_HAVE_XLA_FFI
