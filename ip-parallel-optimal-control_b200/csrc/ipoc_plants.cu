// Optional fast path for the reference's two example plants (SURVEY.md §8f "next" #4): the per-time-step
// `Derivatives` (ref noc/optimal_control_problem.py:13-23), the total cost / feasibility of a trajectory
// and the serial rollout, as fused kernels.  Derivatives come from second-order forward-mode autodiff in
// registers (ipoc_jet.cuh) applied to the plant written ONCE as a template — the same expressions, in the
// same order, as the example scripts (ref examples/pendulum_runtime.py:19-72, cartpole_runtime.py:18-81).
// User-defined OCPs keep going through the host framework's autodiff; this path only serves OCPs built by
// ipoc_b200.problems.make_pendulum / make_cartpole and is cross-checked against torch.func in the tests.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include "../../include/ipoc.h"
#include "ipoc_jet.cuh"
#include "ipoc_accept.cuh"

namespace ipoc {

extern unsigned long long g_launches;
void prof_mark(const char* name, cudaStream_t st);

struct PlantParams {
    double Ts, bound;
};

// ---------------------------------------------------------------- pendulum (nx=2, nu=1, nc=2)
struct Pendulum {
    static constexpr int NX = 2, NU = 1, NC = 2;
    template <class T>
    __device__ __forceinline__ static void ode(const T (&x)[NX], const T (&u)[NU], T (&out)[NX]) {
        const double gravity = 9.81, length = 1.0, mass = 1.0, damping = 1e-3;
        out[0] = x[1];
        out[1] = (-gravity / length) * jsin(x[0]) + (u[0] - damping * x[1]) / (mass * length * length);
    }
    template <class T>
    __device__ __forceinline__ static T state_cost(const T (&x)[NX]) {
        const double pi = 3.141592653589793;
        const T e0 = jwrap(x[0]) - pi, e1 = x[1] - 0.0;
        return (0.5 * e0) * 1e0 * e0 + (0.5 * e1) * 1e-1 * e1;
    }
};

// ---------------------------------------------------------------- cartpole (nx=4, nu=1, nc=2)
struct Cartpole {
    static constexpr int NX = 4, NU = 1, NC = 2;
    template <class T>
    __device__ __forceinline__ static void ode(const T (&x)[NX], const T (&u)[NU], T (&out)[NX]) {
        const double gravity = 9.81, pole_length = 0.5, cart_mass = 10.0, pole_mass = 1.0;
        const double total_mass = cart_mass + pole_mass;
        const T sth = jsin(x[1]), cth = jcos(x[1]);
        const T w2 = x[3] * x[3];
        out[0] = x[2];
        out[1] = x[3];
        out[2] = (u[0] + (pole_mass * sth) * (pole_length * w2 + gravity * cth)) / (cart_mass + pole_mass * (sth * sth));
        out[3] = (-u[0] * cth - (((pole_mass * pole_length) * w2) * cth) * sth - (total_mass * gravity) * sth) /
                 (pole_length * cart_mass + (pole_length * pole_mass) * (sth * sth));
    }
    template <class T>
    __device__ __forceinline__ static T state_cost(const T (&x)[NX]) {
        const double pi = 3.141592653589793;
        const T e0 = x[0] - 0.0, e1 = jwrap(x[1]) - pi, e2 = x[2] - 0.0, e3 = x[3] - 0.0;
        return (0.5 * e0) * 1e0 * e0 + (0.5 * e1) * 1e1 * e1 + (0.5 * e2) * 1e-1 * e2 + (0.5 * e3) * 1e-1 * e3;
    }
};

// stage cost = state cost + 1/2 u' (1e-3) u - bp * sum log(-constraints),  constraints = (u - ub, -u - ub)
template <class P, class T>
__device__ __forceinline__ T stage_cost(const T (&x)[P::NX], const T (&u)[P::NU], double bp, double ub) {
    T c = P::state_cost(x) + (0.5 * u[0]) * 1e-3 * u[0];
    const T m0 = -(u[0] - ub), m1 = -(-u[0] - ub);
    return c - bp * (jlog(m0) + jlog(m1));
}

// ---------------------------------------------------------------- derivatives
// one thread per (problem, time step); variables z = (x, u)
template <class P>
__global__ void __launch_bounds__(128)
k_plant_derivs(PlantParams pp, const double* __restrict__ bp_ptr, int N, int batch,
               const double* __restrict__ X, const double* __restrict__ U,
               double* __restrict__ cx, double* __restrict__ cu, double* __restrict__ cxx, double* __restrict__ cuu,
               double* __restrict__ cxu, double* __restrict__ fx, double* __restrict__ fu, double* __restrict__ fxx,
               double* __restrict__ fuu, double* __restrict__ fxu, double* __restrict__ lamT) {
    constexpr int NX = P::NX, NU = P::NU, NV = NX + NU;
    using J = Jet<NV>;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)batch * N) return;
    const int b = (int)(g / N), k = (int)(g % N);
    const double bp = *bp_ptr;
    const double* xp = X + ((size_t)b * (N + 1) + k) * NX;
    const double* up = U + (size_t)g * NU;
    J x[NX], u[NU];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = J::var(xp[i], i);
#pragma unroll
    for (int a = 0; a < NU; ++a) u[a] = J::var(up[a], NX + a);
    {   // dynamics: euler(ode, Ts)  (ref noc/utils.py:50-54)
        J o[NX];
        P::ode(x, u, o);
#pragma unroll
        for (int r = 0; r < NX; ++r) {
            const J f = x[r] + pp.Ts * o[r];
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                fx[((size_t)g * NX + r) * NX + i] = f.g[i];
#pragma unroll
                for (int j = 0; j < NX; ++j) fxx[(((size_t)g * NX + r) * NX + i) * NX + j] = f.hess(i, j);
#pragma unroll
                for (int a = 0; a < NU; ++a) fxu[(((size_t)g * NX + r) * NX + i) * NU + a] = f.hess(i, NX + a);
            }
#pragma unroll
            for (int a = 0; a < NU; ++a) {
                fu[((size_t)g * NX + r) * NU + a] = f.g[NX + a];
#pragma unroll
                for (int c = 0; c < NU; ++c) fuu[(((size_t)g * NX + r) * NU + a) * NU + c] = f.hess(NX + a, NX + c);
            }
        }
    }
    {   // stage cost
        const J c = stage_cost<P, J>(x, u, bp, pp.bound);
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            cx[(size_t)g * NX + i] = c.g[i];
#pragma unroll
            for (int j = 0; j < NX; ++j) cxx[((size_t)g * NX + i) * NX + j] = c.hess(i, j);
#pragma unroll
            for (int a = 0; a < NU; ++a) cxu[((size_t)g * NX + i) * NU + a] = c.hess(i, NX + a);
        }
#pragma unroll
        for (int a = 0; a < NU; ++a) {
            cu[(size_t)g * NU + a] = c.g[NX + a];
#pragma unroll
            for (int cc = 0; cc < NU; ++cc) cuu[((size_t)g * NU + a) * NU + cc] = c.hess(NX + a, NX + cc);
        }
    }
    if (k == N - 1 && lamT != nullptr) {   // lambda_N = grad final_cost(x_N)  (ref noc/costates.py:35)
        const double* xn = X + ((size_t)b * (N + 1) + N) * NX;
        J xe[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) xe[i] = J::var(xn[i], i);
        const J c = P::state_cost(xe);
#pragma unroll
        for (int i = 0; i < NX; ++i) lamT[(size_t)b * NX + i] = c.g[i];
    }
}

// ---------------------------------------------------------------- linearisation + Hamiltonian (fused A1+A3)
// What the Newton step actually consumes is fx, fu, ru, Q, R, M (+ cx for the costates, cu for the
// regularisation scale), and  Q = cxx + sum_o l_o fxx[o]  etc. (ref noc/par_interior_point_newton.py:31-42)
// are exactly the second derivatives of the Hamiltonian  H(x,u) = stage_cost(x,u) + l' f(x,u)  with
// l = lambda_{k+1}.  So two small passes replace the 130-double `Derivatives` record and the contraction:
//   pass 1 (before the costate scan): fx, fu, cx, cu, lambda_N            (25 doubles per step)
//   pass 2 (after it):                ru = H_u, Q = H_xx, R = H_uu, M = H_xu   (22 doubles per step)
template <int CNT>
__device__ __forceinline__ void store_row(double* __restrict__ p, const double* v) {
    if constexpr (CNT % 2 == 0) {
        double2* p2 = reinterpret_cast<double2*>(p);
#pragma unroll
        for (int i = 0; i < CNT / 2; ++i) p2[i] = make_double2(v[2 * i], v[2 * i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < CNT; ++i) p[i] = v[i];
    }
}

template <class P>
__global__ void __launch_bounds__(128)
k_plant_linearize(PlantParams pp, const double* __restrict__ bp_ptr, int N, int batch,
                  const double* __restrict__ X, const double* __restrict__ U, double* __restrict__ fx,
                  double* __restrict__ fu, double* __restrict__ cx, double* __restrict__ cu,
                  double* __restrict__ lamT, const int32_t* __restrict__ fresh, const double* __restrict__ TX,
                  const double* __restrict__ TU, double* Xw, double* Uw) {
    // TX / TU (may be NULL): the iterate is first TAKEN from there — x <- tx, u <- tu for the members this call
    // evaluates (fresh != 0), written through Xw / Uw (the same arrays as X / U) — i.e. the masked copy that takes an
    // accepted step (ref :184) rides along with the first kernel that reads the new iterate.
    constexpr int NX = P::NX, NU = P::NU, NV = NX + NU;
    using J = Jet<NV>;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)batch * N) return;
    const int b = (int)(g / N), k = (int)(g % N);
    if (fresh != nullptr && fresh[b] == 0) return;   // iterate unchanged since the last evaluation: outputs still valid
    const double bp = *bp_ptr;
    const size_t xrow = ((size_t)b * (N + 1) + k) * NX;
    const double* xp = (TX != nullptr ? TX : X) + xrow;
    const double* up = (TU != nullptr ? TU : U) + (size_t)g * NU;
    J x[NX], u[NU];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = J::var(xp[i], i);
#pragma unroll
    for (int a = 0; a < NU; ++a) u[a] = J::var(up[a], NX + a);
    if (TX != nullptr) {
        double xv[NX], uv[NU];
#pragma unroll
        for (int i = 0; i < NX; ++i) xv[i] = x[i].v;
#pragma unroll
        for (int a = 0; a < NU; ++a) uv[a] = u[a].v;
        store_row<NX>(Xw + xrow, xv);
        store_row<NU>(Uw + (size_t)g * NU, uv);
    }
    J o[NX];
    P::ode(x, u, o);
    double fxv[NX * NX], fuv[NX * NU], cxv[NX], cuv[NU];
#pragma unroll
    for (int r = 0; r < NX; ++r) {
        const J f = x[r] + pp.Ts * o[r];
#pragma unroll
        for (int i = 0; i < NX; ++i) fxv[r * NX + i] = f.g[i];
#pragma unroll
        for (int a = 0; a < NU; ++a) fuv[r * NU + a] = f.g[NX + a];
    }
    const J c = stage_cost<P, J>(x, u, bp, pp.bound);
#pragma unroll
    for (int i = 0; i < NX; ++i) cxv[i] = c.g[i];
#pragma unroll
    for (int a = 0; a < NU; ++a) cuv[a] = c.g[NX + a];
    store_row<NX * NX>(fx + (size_t)g * NX * NX, fxv);
    store_row<NX * NU>(fu + (size_t)g * NX * NU, fuv);
    store_row<NX>(cx + (size_t)g * NX, cxv);
    store_row<NU>(cu + (size_t)g * NU, cuv);
    if (k == N - 1 && TX != nullptr) {   // the terminal state travels with the last step's thread
        double xv[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) xv[i] = TX[xrow + NX + i];
        store_row<NX>(Xw + xrow + NX, xv);
    }
    if (k == N - 1 && lamT != nullptr) {
        const double* xn = (TX != nullptr ? TX : X) + ((size_t)b * (N + 1) + N) * NX;
        J xe[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) xe[i] = J::var(xn[i], i);
        const J ce = P::state_cost(xe);
#pragma unroll
        for (int i = 0; i < NX; ++i) lamT[(size_t)b * NX + i] = ce.g[i];
    }
}

template <class P>
__global__ void __launch_bounds__(128)
k_plant_hamiltonian(PlantParams pp, const double* __restrict__ bp_ptr, int N, int batch,
                    const double* __restrict__ X, const double* __restrict__ U, const double* __restrict__ lam,
                    double* __restrict__ ru, double* __restrict__ Q, double* __restrict__ R,
                    double* __restrict__ M, const int32_t* __restrict__ fresh) {
    constexpr int NX = P::NX, NU = P::NU, NV = NX + NU;
    using J = Jet<NV>;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)batch * N) return;
    const int b = (int)(g / N), k = (int)(g % N);
    if (fresh != nullptr && fresh[b] == 0) return;
    const double bp = *bp_ptr;
    const double* xp = X + ((size_t)b * (N + 1) + k) * NX;
    const double* lp = lam + ((size_t)b * (N + 1) + k + 1) * NX;
    J x[NX], u[NU];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = J::var(xp[i], i);
#pragma unroll
    for (int a = 0; a < NU; ++a) u[a] = J::var(U[(size_t)g * NU + a], NX + a);
    J o[NX];
    P::ode(x, u, o);
    J H = stage_cost<P, J>(x, u, bp, pp.bound);
#pragma unroll
    for (int r = 0; r < NX; ++r) H = H + lp[r] * (x[r] + pp.Ts * o[r]);
    double ruv[NU], Qv[NX * NX], Rv[NU * NU], Mv[NX * NU];
#pragma unroll
    for (int a = 0; a < NU; ++a) {
        ruv[a] = H.g[NX + a];
#pragma unroll
        for (int c = 0; c < NU; ++c) Rv[a * NU + c] = H.hess(NX + a, NX + c);
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
        for (int j = 0; j < NX; ++j) Qv[i * NX + j] = H.hess(i, j);
#pragma unroll
        for (int a = 0; a < NU; ++a) Mv[i * NU + a] = H.hess(i, NX + a);
    }
    store_row<NU>(ru + (size_t)g * NU, ruv);
    store_row<NX * NX>(Q + (size_t)g * NX * NX, Qv);
    store_row<NU * NU>(R + (size_t)g * NU * NU, Rv);
    store_row<NX * NU>(M + (size_t)g * NX * NU, Mv);
}

// ---------------------------------------------------------------- total cost + feasibility
// total_cost = final_cost(x_N) + sum_k stage_cost(x_k, u_k, bp); feasible = all(constraints <= 0).
// One CTA of 1024 threads per problem; fixed summation order (thread-strided partials, shuffle trees).
// caller-provided scratch of the grid form of the cost kernel (see ipoc_plant_cost_workspace_bytes)
struct CostScratch {
    int G;            // CTAs per problem (<= 1: not used)
    unsigned* cnt;    // [batch] arrival counters
    double* part;     // [batch][G]
    int* ok;          // [batch][G]
};
static int cost_grid_ctas(int N, int batch) {   // 0 = the cluster / single-CTA forms serve this shape
    if (batch >= 32 || N <= 8 * 1024) return 0;
    const int g = (N + 2047) / 2048;
    return g > 256 ? 256 : g;
}
static size_t cost_scratch_bytes(int N, int batch) {
    const int G = cost_grid_ctas(N, batch);
    if (G == 0) return 0;
    return 256 + ((size_t)batch * 4 + 255) / 256 * 256 + (size_t)batch * G * (sizeof(double) + sizeof(int));
}
static CostScratch carve_cost_scratch(int N, int batch, void* ws, size_t ws_bytes) {
    CostScratch c{0, nullptr, nullptr, nullptr};
    const int G = cost_grid_ctas(N, batch);
    if (G == 0 || ws == nullptr || ws_bytes < cost_scratch_bytes(N, batch)) return c;
    char* p = (char*)(((uintptr_t)ws + 255) / 256 * 256);
    c.G = G;
    c.cnt = (unsigned*)p;
    p += ((size_t)batch * 4 + 255) / 256 * 256;
    c.part = (double*)p;
    p += (size_t)batch * G * sizeof(double);
    c.ok = (int*)p;
    return c;
}

template <class P>
__global__ void __launch_bounds__(1024)
k_plant_cost(PlantParams pp, const double* __restrict__ bp_ptr, int N, const double* __restrict__ X,
             const double* __restrict__ U, double* __restrict__ total, int32_t* __restrict__ feasible, int finish,
             FinishIO fin, const int32_t* __restrict__ fresh, int csize, CostScratch gs) {
    // Several CTAs per problem take interleaved blocks of 1024 steps and reduce them in a fixed order.
    //  * cluster form (no scratch given): one thread-block CLUSTER of `csize` <= 8 CTAs per problem (csize = 1: a
    //    plain launch); every CTA leaves its partial in shared memory and CTA 0 adds them in rank order through
    //    distributed shared memory — deterministic, no global scratch, no second launch;
    //  * grid form (gs.G > 1, caller's scratch): up to 256 CTAs per problem, partials in global memory, the CTA that
    //    arrives last on the problem's counter (zero-initialised, wraps back to zero) folds them in a fixed order —
    //    the stage cost is FP64-throughput-bound (two logs, wrap, plant: ~300 FP64 operations per step), so a
    //    horizon of 1e5 ... 1e6 steps needs the whole chip, not 8 SMs.
    namespace cg = cooperative_groups;
    constexpr int NX = P::NX, NU = P::NU;
    __shared__ double s_sum[32];
    __shared__ int s_ok[32];
    __shared__ double s_part;
    __shared__ int s_pok;
    const int per = gs.G > 1 ? gs.G : csize;   // CTAs per problem
    const int b = blockIdx.x / per, rank = blockIdx.x % per;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (fresh != nullptr && fresh[b] == 0) return;    // uniform over the cluster
    if (finish && !fin.active[b]) {   // frozen member of a device-resident loop: nothing to evaluate
        if (t == 0 && rank == 0) {
            fin.advanced[b] = 0;
            if (fin.need_cost != nullptr) fin.need_cost[b] = 0;
        }
        return;
    }
    const double bp = *bp_ptr;
    double acc = 0.0;
    int ok = 1;
    for (int k = rank * 1024 + t; k < N; k += 1024 * per) {
        double x[NX], u[NU];
        const double* xp = X + ((size_t)b * (N + 1) + k) * NX;
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = xp[i];
#pragma unroll
        for (int a = 0; a < NU; ++a) u[a] = U[((size_t)b * N + k) * NU + a];
        acc += stage_cost<P, double>(x, u, bp, pp.bound);
        ok &= ((u[0] - pp.bound) <= 0.0 && (-u[0] - pp.bound) <= 0.0) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        ok &= __shfl_xor_sync(0xffffffffu, ok, o);
    }
    if (lane == 0) {
        s_sum[w] = acc;
        s_ok[w] = ok;
    }
    __syncthreads();
    if (w == 0) {
        acc = s_sum[lane];
        ok = s_ok[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
            ok &= __shfl_xor_sync(0xffffffffu, ok, o);
        }
        if (lane == 0) {
            s_part = acc;
            s_pok = ok;
        }
    }
    auto conclude = [&](double sum, int all_ok) {
        double xn[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) xn[i] = X[((size_t)b * (N + 1) + N) * NX + i];
        const double tot = P::state_cost(xn) + sum;
        total[b] = tot;
        feasible[b] = all_ok;
        if (finish) attempt_finish_rule(fin, b, tot, all_ok);   // accept + loop bookkeeping (ref :159-202)
    };
    if (gs.G > 1) {
        __shared__ int s_last;
        if (t == 0) {
            gs.part[(size_t)b * gs.G + rank] = s_part;
            gs.ok[(size_t)b * gs.G + rank] = s_pok;
            unsigned old;
            asm volatile("atom.acq_rel.gpu.global.inc.u32 %0, [%1], %2;" : "=r"(old) : "l"(gs.cnt + b), "r"((unsigned)gs.G - 1u) : "memory");
            s_last = (old == (unsigned)gs.G - 1u) ? 1 : 0;
        }
        __syncthreads();
        if (!s_last || w != 0) return;
        double a = 0.0;
        int o = 1;
        for (int r = lane; r < gs.G; r += 32) {   // lane l: ranks l, l + 32, ... in order; then a fixed butterfly
            a += __ldcg(gs.part + (size_t)b * gs.G + r);
            o &= __ldcg(gs.ok + (size_t)b * gs.G + r);
        }
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, sh);
            o &= __shfl_xor_sync(0xffffffffu, o, sh);
        }
        if (lane == 0) conclude(a, o);
        return;
    }
    if (csize > 1) cg::this_cluster().sync();
    if (rank == 0 && t == 0) {
        acc = s_part;
        ok = s_pok;
        for (int r = 1; r < csize; ++r) {
            acc += *cg::this_cluster().map_shared_rank(&s_part, r);
            ok &= *cg::this_cluster().map_shared_rank(&s_pok, r);
        }
        conclude(acc, ok);
    }
    if (csize > 1) cg::this_cluster().sync();   // the peers' shared memory stays alive until CTA 0 has read it
}

// CTAs per problem for the cost kernel: one per 1024 steps up to a portable cluster of 8, only when the batch
// leaves SMs idle (batched solves keep the single-CTA arithmetic).
static int cost_cluster(int N, int batch) {
    if (batch >= 32) return 1;
    int c = 1;
    while (c < 8 && c * 1024 < N) c *= 2;
    return c;
}
template <class P>
static int launch_cost(PlantParams pp, const double* bp, int N, int batch, const double* X, const double* U,
                       double* total, int32_t* feasible, int finish, const FinishIO& fin, const int32_t* fresh,
                       void* ws, size_t ws_bytes, cudaStream_t st) {
    const CostScratch gs = carve_cost_scratch(N, batch, ws, ws_bytes);
    if (gs.G > 1) {
        k_plant_cost<P><<<batch * gs.G, 1024, 0, st>>>(pp, bp, N, X, U, total, feasible, finish, fin, fresh, 1, gs);
        return IPOC_OK;
    }
    const int c = cost_cluster(N, batch);
    if (c == 1) {
        k_plant_cost<P><<<batch, 1024, 0, st>>>(pp, bp, N, X, U, total, feasible, finish, fin, fresh, 1, gs);
        return IPOC_OK;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(batch * c));
    cfg.blockDim = dim3(1024);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)c;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_plant_cost<P>, pp, bp, N, X, U, total, feasible, finish, fin, fresh, c, gs) == cudaSuccess
               ? IPOC_OK : IPOC_ECUDA;
}

// ---------------------------------------------------------------- serial rollout, one thread per problem
template <class P>
__global__ void k_plant_rollout(PlantParams pp, int N, int batch, const double* __restrict__ x0,
                                const double* __restrict__ U, double* __restrict__ X) {
    constexpr int NX = P::NX, NU = P::NU;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    double x[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = x0[(size_t)b * NX + i];
    double* xo = X + (size_t)b * (N + 1) * NX;
#pragma unroll
    for (int i = 0; i < NX; ++i) xo[i] = x[i];
    for (int k = 0; k < N; ++k) {
        double u[NU], o[NX];
#pragma unroll
        for (int a = 0; a < NU; ++a) u[a] = U[((size_t)b * N + k) * NU + a];
        P::ode(x, u, o);
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            x[i] = x[i] + pp.Ts * o[i];
            xo[(size_t)(k + 1) * NX + i] = x[i];
        }
    }
}

// ---------------------------------------------------------------- parallel-in-time rollout: one Newton iteration's inputs
// Linearisation of the rollout equations x_{k+1} = f(x_k, u_k) along a GUESS of the whole trajectory, one thread per
// (problem, step):  F_k = df/dx,  c_k = f(x_k, u_k) - F_k x_k  (the caller solves the affine recursion for all k at
// once with ipoc_affine_scan_f64),  fv_k = f(x_k, u_k) in the arithmetic of the serial rollout, and per problem
// stats[2b] = max_k |x_{k+1} - fv_k| (the guess's defect; NaN wins), stats[2b+1] = max |fv|.
template <class P>
__global__ void __launch_bounds__(128)
k_plant_rollout_lin(PlantParams pp, int N, int batch, const double* __restrict__ X, const double* __restrict__ U,
                    double* __restrict__ F, double* __restrict__ c, double* __restrict__ fv,
                    unsigned long long* __restrict__ stats) {
    constexpr int NX = P::NX, NU = P::NU;
    using J = Jet<NX>;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = g < (long long)batch * N;
    const int b = valid ? (int)(g / N) : -1, k = valid ? (int)(g % N) : 0;
    double defect = 0.0, amax = 0.0;
    if (valid) {
        const double* xp = X + ((size_t)b * (N + 1) + k) * NX;
        double xv[NX], uv[NU], ov[NX];
        J x[NX], u[NU], o[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            xv[i] = xp[i];
            x[i] = J::var(xv[i], i);
        }
#pragma unroll
        for (int a = 0; a < NU; ++a) {
            uv[a] = U[(size_t)g * NU + a];
            u[a] = J(uv[a]);
        }
        P::ode(x, u, o);
        P::ode(xv, uv, ov);
        double Fv[NX * NX], cv[NX], fvv[NX];
#pragma unroll
        for (int r = 0; r < NX; ++r) {
            fvv[r] = xv[r] + pp.Ts * ov[r];   // exactly the serial rollout's update
            double acc = fvv[r];
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                const double fri = ((i == r) ? 1.0 : 0.0) + pp.Ts * o[r].g[i];
                Fv[r * NX + i] = fri;
                acc -= fri * xv[i];
            }
            cv[r] = acc;
            const double dd = fabs(xp[NX + r] - fvv[r]);
            defect = (dd != dd || defect != defect) ? __longlong_as_double(0x7ff8000000000000LL) : fmax(defect, dd);
            amax = fmax(amax, fabs(fvv[r]));
        }
        store_row<NX * NX>(F + (size_t)g * NX * NX, Fv);
        store_row<NX>(c + (size_t)g * NX, cv);
        store_row<NX>(fv + (size_t)g * NX, fvv);
    }
    // Non-negative doubles (and NaN above +inf) order like their bit patterns.  One atomic pair per WARP when all its
    // lanes belong to one problem (every thread hitting the same two words serialised 2e6 atomics into 1.3 ms at
    // N = 1e6, profiles/r02_ncu_full_new_kernels_summary.txt).
    unsigned long long db = (unsigned long long)__double_as_longlong(defect) & 0x7fffffffffffffffULL;
    unsigned long long ab = (unsigned long long)__double_as_longlong(amax) & 0x7fffffffffffffffULL;
    const int b0 = __shfl_sync(0xffffffffu, b, 0);
    if (__all_sync(0xffffffffu, !valid || b == b0)) {
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) {
            const unsigned long long d2 = __shfl_xor_sync(0xffffffffu, db, sh), a2 = __shfl_xor_sync(0xffffffffu, ab, sh);
            db = d2 > db ? d2 : db;
            ab = a2 > ab ? a2 : ab;
        }
        if ((threadIdx.x & 31) == 0 && b0 >= 0) {
            atomicMax(stats + 2 * (size_t)b0, db);
            atomicMax(stats + 2 * (size_t)b0 + 1, ab);
        }
    } else if (valid) {
        atomicMax(stats + 2 * (size_t)b, db);
        atomicMax(stats + 2 * (size_t)b + 1, ab);
    }
}

#define PLANT_CHECK(st)                                              \
    do {                                                             \
        ++g_launches;                                                \
        prof_mark(__func__, st);                                     \
        if (cudaPeekAtLastError() != cudaSuccess) return IPOC_ECUDA; \
    } while (0)

template <class P>
static int derivs_impl(PlantParams pp, const double* bp, int N, int batch, const double* X, const double* U, double* cx,
                       double* cu, double* cxx, double* cuu, double* cxu, double* fx, double* fu, double* fxx,
                       double* fuu, double* fxu, double* lamT, cudaStream_t st) {
    const long long n = (long long)N * batch;
    k_plant_derivs<P><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(pp, bp, N, batch, X, U, cx, cu, cxx, cuu, cxu, fx, fu,
                                                                 fxx, fuu, fxu, lamT);
    PLANT_CHECK(st);
    return IPOC_OK;
}
template <class P>
static int linearize_impl(PlantParams pp, const double* bp, int N, int batch, const double* X, const double* U,
                          double* fx, double* fu, double* cx, double* cu, double* lamT, const int32_t* fresh,
                          const double* TX, const double* TU, double* Xw, double* Uw, cudaStream_t st) {
    const long long n = (long long)N * batch;
    k_plant_linearize<P><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(pp, bp, N, batch, X, U, fx, fu, cx, cu, lamT, fresh,
                                                                    TX, TU, Xw, Uw);
    PLANT_CHECK(st);
    return IPOC_OK;
}
template <class P>
static int hamiltonian_impl(PlantParams pp, const double* bp, int N, int batch, const double* X, const double* U,
                            const double* lam, double* ru, double* Q, double* R, double* M, const int32_t* fresh,
                            cudaStream_t st) {
    const long long n = (long long)N * batch;
    k_plant_hamiltonian<P><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(pp, bp, N, batch, X, U, lam, ru, Q, R, M, fresh);
    PLANT_CHECK(st);
    return IPOC_OK;
}
template <class P>
static int cost_impl(PlantParams pp, const double* bp, int N, int batch, const double* X, const double* U,
                     double* total, int32_t* feasible, const int32_t* fresh, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (int rc = launch_cost<P>(pp, bp, N, batch, X, U, total, feasible, 0, FinishIO{}, fresh, ws, ws_bytes, st)) return rc;
    PLANT_CHECK(st);
    return IPOC_OK;
}
template <class P>
static int cost_finish_impl(PlantParams pp, const double* bp, int N, int batch, const double* X, const double* U,
                            double* total, int32_t* feasible, const FinishIO& fin, void* ws, size_t ws_bytes,
                            cudaStream_t st) {
    if (int rc = launch_cost<P>(pp, bp, N, batch, X, U, total, feasible, 1, fin, nullptr, ws, ws_bytes, st)) return rc;
    PLANT_CHECK(st);
    return IPOC_OK;
}
template <class P>
static int rollout_impl(PlantParams pp, int N, int batch, const double* x0, const double* U, double* X,
                        cudaStream_t st) {
    k_plant_rollout<P><<<(batch + 31) / 32, 32, 0, st>>>(pp, N, batch, x0, U, X);
    PLANT_CHECK(st);
    return IPOC_OK;
}

template <class P>
static int rollout_lin_impl(PlantParams pp, int N, int batch, const double* X, const double* U, double* F, double* c,
                            double* fv, double* stats, cudaStream_t st) {
    if (cudaMemsetAsync(stats, 0, 2 * sizeof(double) * (size_t)batch, st) != cudaSuccess) return IPOC_ECUDA;
    const long long n = (long long)N * batch;
    k_plant_rollout_lin<P><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(pp, N, batch, X, U, F, c, fv,
                                                                      reinterpret_cast<unsigned long long*>(stats));
    PLANT_CHECK(st);
    return IPOC_OK;
}

}  // namespace ipoc

using namespace ipoc;

extern "C" {

int ipoc_plant_dims(int plant, int* nx, int* nu, int* nc) {
    if (plant == IPOC_PLANT_PENDULUM) { *nx = Pendulum::NX; *nu = Pendulum::NU; *nc = Pendulum::NC; return IPOC_OK; }
    if (plant == IPOC_PLANT_CARTPOLE) { *nx = Cartpole::NX; *nu = Cartpole::NU; *nc = Cartpole::NC; return IPOC_OK; }
    return IPOC_EINVAL;
}

int ipoc_plant_derivatives_f64(int plant, int N, int batch, double Ts, double bound, const double* bp, const double* x,
                               const double* u, double* cx, double* cu, double* cxx, double* cuu, double* cxu,
                               double* fx, double* fu, double* fxx, double* fuu, double* fxu, double* lamT,
                               ipoc_stream_t stream) {
    if (N < 1 || batch < 1 || !bp || !x || !u || !cx || !cu || !cxx || !cuu || !cxu || !fx || !fu || !fxx || !fuu || !fxu)
        return IPOC_EINVAL;
    const PlantParams pp{Ts, bound};
    cudaStream_t st = (cudaStream_t)stream;
    if (plant == IPOC_PLANT_PENDULUM)
        return derivs_impl<Pendulum>(pp, bp, N, batch, x, u, cx, cu, cxx, cuu, cxu, fx, fu, fxx, fuu, fxu, lamT, st);
    if (plant == IPOC_PLANT_CARTPOLE)
        return derivs_impl<Cartpole>(pp, bp, N, batch, x, u, cx, cu, cxx, cuu, cxu, fx, fu, fxx, fuu, fxu, lamT, st);
    return IPOC_EINVAL;
}

int ipoc_plant_linearize_f64(int plant, int N, int batch, double Ts, double bound, const double* bp, const double* x,
                             const double* u, double* fx, double* fu, double* cx, double* cu, double* lamT,
                             const int32_t* fresh, ipoc_stream_t stream) {
    if (N < 1 || batch < 1 || !bp || !x || !u || !fx || !fu || !cx || !cu) return IPOC_EINVAL;
    const PlantParams pp{Ts, bound};
    cudaStream_t st = (cudaStream_t)stream;
    if (plant == IPOC_PLANT_PENDULUM)
        return linearize_impl<Pendulum>(pp, bp, N, batch, x, u, fx, fu, cx, cu, lamT, fresh, nullptr, nullptr, nullptr, nullptr, st);
    if (plant == IPOC_PLANT_CARTPOLE)
        return linearize_impl<Cartpole>(pp, bp, N, batch, x, u, fx, fu, cx, cu, lamT, fresh, nullptr, nullptr, nullptr, nullptr, st);
    return IPOC_EINVAL;
}

int ipoc_plant_take_linearize_f64(int plant, int N, int batch, double Ts, double bound, const double* bp, const double* tx,
                                  const double* tu, double* x, double* u, double* fx, double* fu, double* cx, double* cu,
                                  double* lamT, const int32_t* fresh, ipoc_stream_t stream) {
    if (N < 1 || batch < 1 || !bp || !tx || !tu || !x || !u || !fx || !fu || !cx || !cu) return IPOC_EINVAL;
    const PlantParams pp{Ts, bound};
    cudaStream_t st = (cudaStream_t)stream;
    if (plant == IPOC_PLANT_PENDULUM)
        return linearize_impl<Pendulum>(pp, bp, N, batch, x, u, fx, fu, cx, cu, lamT, fresh, tx, tu, x, u, st);
    if (plant == IPOC_PLANT_CARTPOLE)
        return linearize_impl<Cartpole>(pp, bp, N, batch, x, u, fx, fu, cx, cu, lamT, fresh, tx, tu, x, u, st);
    return IPOC_EINVAL;
}

int ipoc_plant_hamiltonian_f64(int plant, int N, int batch, double Ts, double bound, const double* bp,
                               const double* x, const double* u, const double* lam, double* ru, double* Q,
                               double* R, double* M, const int32_t* fresh, ipoc_stream_t stream) {
    if (N < 1 || batch < 1 || !bp || !x || !u || !lam || !ru || !Q || !R || !M) return IPOC_EINVAL;
    const PlantParams pp{Ts, bound};
    cudaStream_t st = (cudaStream_t)stream;
    if (plant == IPOC_PLANT_PENDULUM) return hamiltonian_impl<Pendulum>(pp, bp, N, batch, x, u, lam, ru, Q, R, M, fresh, st);
    if (plant == IPOC_PLANT_CARTPOLE) return hamiltonian_impl<Cartpole>(pp, bp, N, batch, x, u, lam, ru, Q, R, M, fresh, st);
    return IPOC_EINVAL;
}

int ipoc_plant_cost_f64(int plant, int N, int batch, double Ts, double bound, const double* bp, const double* x,
                        const double* u, double* total_cost, int32_t* feasible, const int32_t* fresh, void* ws,
                        size_t ws_bytes, ipoc_stream_t stream) {
    if (N < 1 || batch < 1 || !bp || !x || !u || !total_cost || !feasible) return IPOC_EINVAL;
    const PlantParams pp{Ts, bound};
    cudaStream_t st = (cudaStream_t)stream;
    if (plant == IPOC_PLANT_PENDULUM)
        return cost_impl<Pendulum>(pp, bp, N, batch, x, u, total_cost, feasible, fresh, ws, ws_bytes, st);
    if (plant == IPOC_PLANT_CARTPOLE)
        return cost_impl<Cartpole>(pp, bp, N, batch, x, u, total_cost, feasible, fresh, ws, ws_bytes, st);
    return IPOC_EINVAL;
}

size_t ipoc_plant_cost_workspace_bytes(int N, int batch) {
    return (N < 1 || batch < 1) ? 0 : cost_scratch_bytes(N, batch);
}

int ipoc_plant_attempt_finish_f64(int plant, int N, int batch, double Ts, double bound, const double* bp,
                                  const double* tx, const double* tu, double* new_cost, int32_t* traj_feasible,
                                  const double* cost, const double* pred, const int32_t* bwd_feasible, const double* hu,
                                  int32_t* active, double* rp, double* r_inc, int32_t* success, double* gain_ratio,
                                  int64_t* inner, int64_t* iteration, uint8_t* outer_done, int32_t* advanced,
                                  double hu_tol, int max_attempts, int max_iterations, double* cost_carry,
                                  int32_t* need_cost, void* ws, size_t ws_bytes, ipoc_stream_t stream) {
    if (N < 1 || batch < 1 || !bp || !tx || !tu || !new_cost || !traj_feasible || !cost || !pred || !bwd_feasible || !hu ||
        !active || !rp || !r_inc || !success || !inner || !iteration || !outer_done || !advanced)
        return IPOC_EINVAL;
    const PlantParams pp{Ts, bound};
    cudaStream_t st = (cudaStream_t)stream;
    const FinishIO fin{AcceptIO{cost, pred, bwd_feasible, rp, r_inc, success, gain_ratio}, hu, active, (long long*)inner,
                       (long long*)iteration, outer_done, advanced, hu_tol, max_attempts, max_iterations, cost_carry,
                       need_cost};
    if (plant == IPOC_PLANT_PENDULUM)
        return cost_finish_impl<Pendulum>(pp, bp, N, batch, tx, tu, new_cost, traj_feasible, fin, ws, ws_bytes, st);
    if (plant == IPOC_PLANT_CARTPOLE)
        return cost_finish_impl<Cartpole>(pp, bp, N, batch, tx, tu, new_cost, traj_feasible, fin, ws, ws_bytes, st);
    return IPOC_EINVAL;
}

int ipoc_plant_rollout_f64(int plant, int N, int batch, double Ts, const double* x0, const double* u, double* x,
                           ipoc_stream_t stream) {
    if (N < 1 || batch < 1 || !x0 || !u || !x) return IPOC_EINVAL;
    const PlantParams pp{Ts, 0.0};
    cudaStream_t st = (cudaStream_t)stream;
    if (plant == IPOC_PLANT_PENDULUM) return rollout_impl<Pendulum>(pp, N, batch, x0, u, x, st);
    if (plant == IPOC_PLANT_CARTPOLE) return rollout_impl<Cartpole>(pp, N, batch, x0, u, x, st);
    return IPOC_EINVAL;
}

int ipoc_plant_rollout_lin_f64(int plant, int N, int batch, double Ts, const double* x, const double* u, double* F,
                               double* c, double* fv, double* stats, ipoc_stream_t stream) {
    if (N < 1 || batch < 1 || !x || !u || !F || !c || !fv || !stats) return IPOC_EINVAL;
    const PlantParams pp{Ts, 0.0};
    cudaStream_t st = (cudaStream_t)stream;
    if (plant == IPOC_PLANT_PENDULUM) return rollout_lin_impl<Pendulum>(pp, N, batch, x, u, F, c, fv, stats, st);
    if (plant == IPOC_PLANT_CARTPOLE) return rollout_lin_impl<Cartpole>(pp, N, batch, x, u, F, c, fv, stats, st);
    return IPOC_EINVAL;
}

}  // extern "C"
