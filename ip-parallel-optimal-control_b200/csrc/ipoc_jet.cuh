// Second-order forward-mode automatic differentiation in registers ("jets"): a value, its gradient and
// its (packed symmetric) Hessian with respect to NV input variables.  Used by the optional built-in
// plants (ipoc_plants.cu) to produce the `Derivatives` of ref noc/optimal_control_problem.py:13-23 —
// what the reference obtains with vmapped jax.grad / hessian / jacrev∘jacrev
// (ref noc/par_interior_point_newton.py:13-28) — in one fused pass per time step.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace ipoc {

template <int NV>
struct Jet {
    static constexpr int NH = NV * (NV + 1) / 2;
    double v;
    double g[NV];
    double h[NH];
    __device__ __forceinline__ static constexpr int at(int i, int j) {
        return (i <= j) ? (i * (2 * NV - i + 1)) / 2 + (j - i) : (j * (2 * NV - j + 1)) / 2 + (i - j);
    }
    __device__ __forceinline__ Jet() {}
    __device__ __forceinline__ Jet(double c) : v(c) {
#pragma unroll
        for (int i = 0; i < NV; ++i) g[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NH; ++i) h[i] = 0.0;
    }
    // independent variable number k with value c
    __device__ __forceinline__ static Jet var(double c, int k) {
        Jet r(c);
#pragma unroll
        for (int i = 0; i < NV; ++i) r.g[i] = (i == k) ? 1.0 : 0.0;
        return r;
    }
    __device__ __forceinline__ double hess(int i, int j) const { return h[at(i, j)]; }
};

template <int NV>
__device__ __forceinline__ Jet<NV> operator+(const Jet<NV>& a, const Jet<NV>& b) {
    Jet<NV> r;
    r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < NV; ++i) r.g[i] = a.g[i] + b.g[i];
#pragma unroll
    for (int i = 0; i < Jet<NV>::NH; ++i) r.h[i] = a.h[i] + b.h[i];
    return r;
}
template <int NV>
__device__ __forceinline__ Jet<NV> operator-(const Jet<NV>& a, const Jet<NV>& b) {
    Jet<NV> r;
    r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < NV; ++i) r.g[i] = a.g[i] - b.g[i];
#pragma unroll
    for (int i = 0; i < Jet<NV>::NH; ++i) r.h[i] = a.h[i] - b.h[i];
    return r;
}
template <int NV>
__device__ __forceinline__ Jet<NV> operator-(const Jet<NV>& a) {
    Jet<NV> r;
    r.v = -a.v;
#pragma unroll
    for (int i = 0; i < NV; ++i) r.g[i] = -a.g[i];
#pragma unroll
    for (int i = 0; i < Jet<NV>::NH; ++i) r.h[i] = -a.h[i];
    return r;
}
template <int NV>
__device__ __forceinline__ Jet<NV> operator*(const Jet<NV>& a, const Jet<NV>& b) {
    Jet<NV> r;
    r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < NV; ++i) r.g[i] = a.g[i] * b.v + a.v * b.g[i];
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = i; j < NV; ++j)
            r.h[Jet<NV>::at(i, j)] = a.h[Jet<NV>::at(i, j)] * b.v + a.v * b.h[Jet<NV>::at(i, j)] + a.g[i] * b.g[j] +
                                     a.g[j] * b.g[i];
    return r;
}
// f(a) given f(a.v), f'(a.v), f''(a.v)
template <int NV>
__device__ __forceinline__ Jet<NV> chain(const Jet<NV>& a, double f, double f1, double f2) {
    Jet<NV> r;
    r.v = f;
#pragma unroll
    for (int i = 0; i < NV; ++i) r.g[i] = f1 * a.g[i];
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = i; j < NV; ++j)
            r.h[Jet<NV>::at(i, j)] = f1 * a.h[Jet<NV>::at(i, j)] + f2 * a.g[i] * a.g[j];
    return r;
}
template <int NV>
__device__ __forceinline__ Jet<NV> recip(const Jet<NV>& a) {
    const double r = 1.0 / a.v;
    return chain(a, r, -r * r, 2.0 * r * r * r);
}
template <int NV>
__device__ __forceinline__ Jet<NV> operator/(const Jet<NV>& a, const Jet<NV>& b) { return a * recip(b); }
template <int NV>
__device__ __forceinline__ Jet<NV> jsin(const Jet<NV>& a) {
    double s, c;
    sincos(a.v, &s, &c);
    return chain(a, s, c, -s);
}
template <int NV>
__device__ __forceinline__ Jet<NV> jcos(const Jet<NV>& a) {
    double s, c;
    sincos(a.v, &s, &c);
    return chain(a, c, -s, -c);
}
template <int NV>
__device__ __forceinline__ Jet<NV> jlog(const Jet<NV>& a) {
    const double r = 1.0 / a.v;
    return chain(a, log(a.v), r, -r * r);
}
// x mod 2*pi in [0, 2*pi): derivative 1 almost everywhere (ref noc/utils.py:8-10)
template <int NV>
__device__ __forceinline__ Jet<NV> jwrap(const Jet<NV>& a) {
    Jet<NV> r = a;
    const double two_pi = 6.283185307179586;
    r.v = a.v - two_pi * floor(a.v / two_pi);
    return r;
}
// Operations with a plain scalar.  A scalar is a jet whose derivatives are structurally zero; writing the products
// out (instead of promoting the scalar to a jet) drops the 0 * x terms the compiler must keep under IEEE rules —
// about 3/4 of the arithmetic of a jet product — and gives the same bits for finite operands (x + 0 * y == x).
template <int NV>
__device__ __forceinline__ Jet<NV> operator*(double c, const Jet<NV>& a) {
    Jet<NV> r;
    r.v = c * a.v;
#pragma unroll
    for (int i = 0; i < NV; ++i) r.g[i] = c * a.g[i];
#pragma unroll
    for (int i = 0; i < Jet<NV>::NH; ++i) r.h[i] = c * a.h[i];
    return r;
}
template <int NV>
__device__ __forceinline__ Jet<NV> operator*(const Jet<NV>& a, double c) {
    Jet<NV> r;
    r.v = a.v * c;
#pragma unroll
    for (int i = 0; i < NV; ++i) r.g[i] = a.g[i] * c;
#pragma unroll
    for (int i = 0; i < Jet<NV>::NH; ++i) r.h[i] = a.h[i] * c;
    return r;
}
template <int NV>
__device__ __forceinline__ Jet<NV> operator+(const Jet<NV>& a, double c) {
    Jet<NV> r = a;
    r.v = a.v + c;
    return r;
}
template <int NV>
__device__ __forceinline__ Jet<NV> operator+(double c, const Jet<NV>& a) {
    Jet<NV> r = a;
    r.v = c + a.v;
    return r;
}
template <int NV>
__device__ __forceinline__ Jet<NV> operator-(const Jet<NV>& a, double c) {
    Jet<NV> r = a;
    r.v = a.v - c;
    return r;
}
template <int NV>
__device__ __forceinline__ Jet<NV> operator-(double c, const Jet<NV>& a) {
    Jet<NV> r = -a;
    r.v = c - a.v;
    return r;
}
template <int NV>
__device__ __forceinline__ Jet<NV> operator/(const Jet<NV>& a, double c) {   // true division, like the host framework
    Jet<NV> r;
    r.v = a.v / c;
#pragma unroll
    for (int i = 0; i < NV; ++i) r.g[i] = a.g[i] / c;
#pragma unroll
    for (int i = 0; i < Jet<NV>::NH; ++i) r.h[i] = a.h[i] / c;
    return r;
}

// plain-double versions with the same names so that plant code is written once
__device__ __forceinline__ double jsin(double a) { return sin(a); }
__device__ __forceinline__ double jcos(double a) { return cos(a); }
__device__ __forceinline__ double jlog(double a) { return log(a); }
__device__ __forceinline__ double jwrap(double a) {
    const double two_pi = 6.283185307179586;
    return a - two_pi * floor(a / two_pi);
}

}  // namespace ipoc
