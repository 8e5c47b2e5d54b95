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
#endif  // IPOC_HAVE_XLA_FFI
