// C ABI of libipoc.so (include/ipoc.h): argument checks, dispatch on the runtime state dimension to
// the per-NX translation units (ipoc_nx.cu), the dimension-independent kernels (K4 reductions, A8
// accept/update), the host-buffer wrapper and the optional per-launch CUDA-event profiler.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "../../include/ipoc.h"
#include "ipoc_dispatch.h"
#include "ipoc_accept.cuh"

namespace ipoc {

struct Tuning {
    int leaf_chunk, mid_fanin, top_max;
};
unsigned long long g_launches = 0;
Tuning g_tune = {0, 0, 0};
int g_literal_lqt = 0;
struct HierTuning {
    int enabled, group_warps, serial_top, aff_warps_per_sm;
};
HierTuning g_hier = {1, 0, 0, 0};
// ---- per-launch profiler: one CUDA event after every kernel launch, on the launching stream ----
constexpr int kMaxProf = 256;
struct Prof {
    bool armed = false, created = false;
    int n = 0;
    cudaEvent_t ev[kMaxProf];
    const char* name[kMaxProf];
};
static Prof g_prof;
void prof_mark(const char* name, cudaStream_t st) {
    if (!g_prof.armed || g_prof.n >= kMaxProf) return;
    cudaEventRecord(g_prof.ev[g_prof.n], st);
    g_prof.name[g_prof.n++] = name;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct Bump {
    char* base;
    size_t off, cap;
    bool dry;
    template <class T>
    T* take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T* r = dry ? nullptr : reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return r;
    }
};

#define IPOC_API_LAUNCH_CHECK(st)                                   \
    do {                                                            \
        ++g_launches;                                               \
        prof_mark(__func__, st);                                    \
        if (cudaPeekAtLastError() != cudaSuccess) return IPOC_ECUDA; \
    } while (0)

// ------------------------------------------------------------------ K4 reductions + A8 update
// K4, stage 1: grid (nblk, batch); block j of problem b reduces its contiguous slice of the three
// arrays to (max|ru| [NaN-propagating], sum cu^2, all(cons<=0)) -> partials[(b*nblk + j)*3 ..].
// Stage 2 (one block per problem) folds the nblk partials in index order -> bit-reproducible.
constexpr int kRedThreads = 256;
static __device__ __forceinline__ double nan_max(double a, double c) {
    return (a != a || c != c) ? __longlong_as_double(0x7ff8000000000000LL) : fmax(a, c);
}
static __global__ void __launch_bounds__(kRedThreads)
k_reduce_partial(const double* __restrict__ ru, const double* __restrict__ cu, const double* __restrict__ cons,
                 int N, int nu, int nc, int nblk, double* __restrict__ partials) {
    __shared__ double s_max[kRedThreads];
    __shared__ double s_sq[kRedThreads];
    __shared__ int s_ok[kRedThreads];
    const int b = blockIdx.x / nblk, j = blockIdx.x % nblk, t = threadIdx.x;   // 1-D grid: batch may exceed 65535
    double mx = 0.0, sq = 0.0;
    int ok = 1;
    auto slice = [&](long long total, long long& lo, long long& hi) {
        const long long per = (total + nblk - 1) / nblk;
        lo = (long long)j * per;
        hi = lo + per < total ? lo + per : total;
    };
    long long lo, hi;
    if (ru != nullptr) {
        const double* p = ru + (size_t)b * N * nu;
        slice((long long)N * nu, lo, hi);
#pragma unroll 4
        for (long long i = lo + t; i < hi; i += kRedThreads) mx = nan_max(mx, fabs(p[i]));
    }
    if (cu != nullptr) {
        const double* p = cu + (size_t)b * N * nu;
        slice((long long)N * nu, lo, hi);
#pragma unroll 4
        for (long long i = lo + t; i < hi; i += kRedThreads) sq += p[i] * p[i];
    }
    if (cons != nullptr) {
        const double* p = cons + (size_t)b * N * nc;
        slice((long long)N * nc, lo, hi);
#pragma unroll 4
        for (long long i = lo + t; i < hi; i += kRedThreads) ok &= (p[i] <= 0.0) ? 1 : 0;
    }
    s_max[t] = mx;
    s_sq[t] = sq;
    s_ok[t] = ok;
    __syncthreads();
    for (int o = kRedThreads / 2; o > 0; o >>= 1) {
        if (t < o) {
            s_max[t] = nan_max(s_max[t], s_max[t + o]);
            s_sq[t] += s_sq[t + o];
            s_ok[t] &= s_ok[t + o];
        }
        __syncthreads();
    }
    if (t == 0) {
        double* o = partials + ((size_t)b * nblk + j) * 3;
        o[0] = s_max[0];
        o[1] = s_sq[0];
        o[2] = (double)s_ok[0];
    }
}

static __global__ void __launch_bounds__(kRedThreads)
k_reduce_final(const double* __restrict__ partials, int nblk, int has_ru, int has_cu, int has_cons,
               double* __restrict__ hu_norm, double* __restrict__ cu_norm, int32_t* __restrict__ traj_feasible,
               const double* __restrict__ rp, double* __restrict__ reg) {
    __shared__ double s_max[kRedThreads];
    __shared__ double s_sq[kRedThreads];
    __shared__ int s_ok[kRedThreads];
    const int b = blockIdx.x, t = threadIdx.x;
    double mx = 0.0, sq = 0.0;
    int ok = 1;
    for (int j = t; j < nblk; j += kRedThreads) {
        const double* p = partials + ((size_t)b * nblk + j) * 3;
        mx = nan_max(mx, p[0]);
        sq += p[1];
        ok &= (p[2] != 0.0) ? 1 : 0;
    }
    s_max[t] = mx;
    s_sq[t] = sq;
    s_ok[t] = ok;
    __syncthreads();
    for (int o = kRedThreads / 2; o > 0; o >>= 1) {
        if (t < o) {
            s_max[t] = nan_max(s_max[t], s_max[t + o]);
            s_sq[t] += s_sq[t + o];
            s_ok[t] &= s_ok[t + o];
        }
        __syncthreads();
    }
    if (t == 0) {
        if (has_ru) hu_norm[b] = s_max[0];
        if (has_cu) {
            const double nrm = sqrt(s_sq[0]);
            cu_norm[b] = nrm;
            if (rp != nullptr && reg != nullptr) reg[b] = rp[b] * nrm;   // ref :117
        }
        if (has_cons) traj_feasible[b] = s_ok[0];
    }
}

// Small problems: one CTA of 1024 threads per problem does the slice reduction and the finalisation
// in a single launch (N*width <= 65536); the combination order is fixed, so the result is
// bit-reproducible.
constexpr int kRedSingleThreads = 1024;
static __global__ void __launch_bounds__(kRedSingleThreads)
k_reduce_single(const double* __restrict__ ru, const double* __restrict__ cu, const double* __restrict__ cons,
                int N, int nu, int nc, double* __restrict__ hu_norm, double* __restrict__ cu_norm,
                int32_t* __restrict__ traj_feasible, const double* __restrict__ rp, double* __restrict__ reg) {
    __shared__ double s_max[32];
    __shared__ double s_sq[32];
    __shared__ int s_ok[32];
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    constexpr int NT = kRedSingleThreads;
    double mx = 0.0, sq = 0.0;
    int ok = 1;
    // All loads of a batch of U strided entries (of BOTH ru and cu) are issued before the first use, so a
    // thread pays one memory round trip per batch instead of one per entry (N = 1e4: 10 entries per thread
    // and array -> 2 round trips instead of 6); accumulation order is fixed (4 interleaved accumulators).
    constexpr int U = 8;
    {
        const double* pr = ru != nullptr ? ru + (size_t)b * N * nu : nullptr;
        const double* pc = cu != nullptr ? cu + (size_t)b * N * nu : nullptr;
        const int n = (pr != nullptr || pc != nullptr) ? N * nu : 0;
        double m[4] = {0.0, 0.0, 0.0, 0.0}, a[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i0 = t; i0 < n; i0 += U * NT) {
            double r[U], c[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int idx = i0 + k * NT;
                r[k] = (pr != nullptr && idx < n) ? pr[idx] : 0.0;
                c[k] = (pc != nullptr && idx < n) ? pc[idx] : 0.0;
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                m[k & 3] = nan_max(m[k & 3], fabs(r[k]));
                a[k & 3] += c[k] * c[k];
            }
        }
        mx = nan_max(nan_max(m[0], m[1]), nan_max(m[2], m[3]));
        sq = (a[0] + a[1]) + (a[2] + a[3]);
    }
    if (cons != nullptr) {
        const double* p = cons + (size_t)b * N * nc;
        const int n = N * nc;
        for (int i0 = t; i0 < n; i0 += U * NT) {
            double c[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int idx = i0 + k * NT;
                c[k] = idx < n ? p[idx] : 0.0;
            }
#pragma unroll
            for (int k = 0; k < U; ++k) ok &= (c[k] <= 0.0) ? 1 : 0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        ok &= __shfl_xor_sync(0xffffffffu, ok, o);
    }
    if (lane == 0) {
        s_max[w] = mx;
        s_sq[w] = sq;
        s_ok[w] = ok;
    }
    __syncthreads();
    if (w == 0) {
        mx = s_max[lane];
        sq = s_sq[lane];
        ok = s_ok[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
            ok &= __shfl_xor_sync(0xffffffffu, ok, o);
        }
        if (lane == 0) {
            if (ru != nullptr) hu_norm[b] = mx;
            if (cu != nullptr) {
                const double nrm = sqrt(sq);
                cu_norm[b] = nrm;
                if (rp != nullptr && reg != nullptr) reg[b] = rp[b] * nrm;   // ref :117
            }
            if (cons != nullptr) traj_feasible[b] = ok;
        }
    }
}

static int reduce_blocks(int N, int width, int batch) {
    if ((long long)N * width <= 65536) return 1;
    long long per_problem = ((long long)N * width + 2047) / 2048;   // >= 2048 entries per block
    long long cap = (148LL * 8 + batch - 1) / batch;                // fill the chip, not more
    long long n = per_problem < cap ? per_problem : cap;
    return (int)(n < 1 ? 1 : n);
}

// A3: ru = cu + fu' l, Q = cxx + sum_o l_o fxx[o], R = cuu + sum_o l_o fuu[o], M = cxu + sum_o l_o fxu[o]
// with l = lambda_{k+1} (ref noc/par_interior_point_newton.py:31-42; tensordot contracts the OUTPUT index).
// One thread per (problem, time step); purely streaming (about 1.4 KB per step at nx = 4).
static __global__ void __launch_bounds__(128)
k_lqr_params(int N, int nx, int nu, int batch, const double* __restrict__ lam, const double* __restrict__ cu,
             const double* __restrict__ cxx, const double* __restrict__ cuu, const double* __restrict__ cxu,
             const double* __restrict__ fu, const double* __restrict__ fxx, const double* __restrict__ fuu,
             const double* __restrict__ fxu, double* __restrict__ ru, double* __restrict__ Q,
             double* __restrict__ R, double* __restrict__ M) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)batch * N) return;
    const int b = (int)(g / N), k = (int)(g % N);
    const double* l = lam + ((size_t)b * (N + 1) + k + 1) * nx;
    double lv[8];
    for (int o = 0; o < nx; ++o) lv[o] = l[o];
    for (int a = 0; a < nu; ++a) {
        double v = cu[(size_t)g * nu + a];
        for (int o = 0; o < nx; ++o) v += fu[((size_t)g * nx + o) * nu + a] * lv[o];
        ru[(size_t)g * nu + a] = v;
    }
    const int nxx = nx * nx, nuu = nu * nu, nxu = nx * nu;
    for (int i = 0; i < nxx; ++i) {
        double v = cxx[(size_t)g * nxx + i];
        for (int o = 0; o < nx; ++o) v += lv[o] * fxx[((size_t)g * nx + o) * nxx + i];
        Q[(size_t)g * nxx + i] = v;
    }
    for (int i = 0; i < nuu; ++i) {
        double v = cuu[(size_t)g * nuu + i];
        for (int o = 0; o < nx; ++o) v += lv[o] * fuu[((size_t)g * nx + o) * nuu + i];
        R[(size_t)g * nuu + i] = v;
    }
    for (int i = 0; i < nxu; ++i) {
        double v = cxu[(size_t)g * nxu + i];
        for (int o = 0; o < nx; ++o) v += lv[o] * fxu[((size_t)g * nx + o) * nxu + i];
        M[(size_t)g * nxu + i] = v;
    }
}

static __global__ void k_accept_update(int batch, const double* __restrict__ cost, const double* __restrict__ new_cost,
                                const int32_t* __restrict__ traj_feasible, const double* __restrict__ pred,
                                const int32_t* __restrict__ bwd_feasible, const int32_t* __restrict__ active,
                                double* __restrict__ rp, double* __restrict__ r_inc,
                                int32_t* __restrict__ success, double* __restrict__ gain_ratio) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    if (active != nullptr && !active[b]) return;
    accept_rule_core(AcceptIO{cost, pred, bwd_feasible, rp, r_inc, success, gain_ratio}, b, new_cost[b], traj_feasible[b]);
}

static __global__ void k_attempt_finish(int batch, FinishIO f, const double* __restrict__ new_cost,
                                        const int32_t* __restrict__ traj_feasible) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) attempt_finish_rule(f, b, new_cost[b], traj_feasible[b]);
}

// ---- attempt-loop glue (see include/ipoc.h) ---------------------------------------------------
static __global__ void k_attempt_begin(int batch, const uint8_t* __restrict__ done, const double* __restrict__ rp,
                                       const double* __restrict__ cu_norm, int32_t* __restrict__ active,
                                       double* __restrict__ reg) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    active[b] = (done == nullptr || done[b] == 0) ? 1 : 0;
    reg[b] = rp[b] * cu_norm[b];
}

// mask (may be NULL): int32 per problem, members with 0 keep their tx / tu (per_x, per_u = doubles per problem)
static __global__ void k_trial_point(long long nxs, long long nus, long long per_x, long long per_u,
                                     const int32_t* __restrict__ mask, const double* __restrict__ x,
                                     const double* __restrict__ dx, const double* __restrict__ u,
                                     const double* __restrict__ du, double* __restrict__ tx, double* __restrict__ tu) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nxs) {
        if (mask == nullptr || mask[i / per_x] != 0) tx[i] = x[i] + dx[i];
    } else if (i < nxs + nus) {
        const long long j = i - nxs;
        if (mask == nullptr || mask[j / per_u] != 0) tu[j] = u[j] + du[j];
    }
}
static int trial_point_masked(int N, int nx, int nu, int batch, const int32_t* mask, const double* x, const double* dx,
                              const double* u, const double* du, double* tx, double* tu, cudaStream_t st_) {
    const long long per_x = (long long)(N + 1) * nx, per_u = (long long)N * nu;
    const long long nxs = batch * per_x, nus = batch * per_u;
    k_trial_point<<<(unsigned)((nxs + nus + 255) / 256), 256, 0, st_>>>(nxs, nus, per_x, per_u, mask, x, dx, u, du, tx, tu);
    IPOC_API_LAUNCH_CHECK(st_);
    return IPOC_OK;
}

constexpr int kCommitThreads = 256;
static __global__ void k_attempt_commit(long long per_x, long long per_u, int chunks, const int32_t* __restrict__ active,
                                        const int32_t* __restrict__ success, const double* __restrict__ tx,
                                        const double* __restrict__ tu, double* __restrict__ keep_x,
                                        double* __restrict__ keep_u, long long* __restrict__ inner,
                                        uint8_t* __restrict__ done, int max_attempts) {
    const int b = blockIdx.x / chunks, ch = blockIdx.x % chunks;
    if (!active[b]) return;                      // `active` is not written here: no block can see a changed flag
    const long long per = per_x + per_u;
    for (long long i = (long long)ch * kCommitThreads + threadIdx.x; i < per; i += (long long)chunks * kCommitThreads) {
        if (i < per_x) keep_x[b * per_x + i] = tx[b * per_x + i];
        else keep_u[b * per_u + (i - per_x)] = tu[b * per_u + (i - per_x)];
    }
    if (ch == 0 && threadIdx.x == 0) {
        const long long n = inner[b] + 1;
        inner[b] = n;
        if (success[b] != 0 || n > max_attempts) done[b] = 1;
    }
}

static __global__ void k_newton_advance_flags(int batch, const double* __restrict__ hu, uint8_t* __restrict__ inner_done,
                                              uint8_t* __restrict__ outer_done, long long* __restrict__ inner,
                                              long long* __restrict__ iteration, int32_t* __restrict__ adv,
                                              double hu_tol, int max_iter) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const int a = (outer_done[b] == 0 && inner_done[b] != 0) ? 1 : 0;
    adv[b] = a;
    if (a) {
        const long long it = iteration[b] + 1;
        iteration[b] = it;
        inner[b] = 0;
        inner_done[b] = 0;
        if (hu[b] < hu_tol || it > max_iter) outer_done[b] = 1;
    }
}

static __global__ void k_masked_copy2(long long per_x, long long per_u, int chunks, const int32_t* __restrict__ mask,
                                      const double* __restrict__ sx, const double* __restrict__ su,
                                      double* __restrict__ dx_, double* __restrict__ du_) {
    const int b = blockIdx.x / chunks, ch = blockIdx.x % chunks;
    if (!mask[b]) return;
    const long long per = per_x + per_u;
    for (long long i = (long long)ch * kCommitThreads + threadIdx.x; i < per; i += (long long)chunks * kCommitThreads) {
        if (i < per_x) dx_[b * per_x + i] = sx[b * per_x + i];
        else du_[b * per_u + (i - per_x)] = su[b * per_u + (i - per_x)];
    }
}

static inline long long copy_chunks(long long per, int batch) {
    long long chunks = (per + 4 * kCommitThreads - 1) / (4 * kCommitThreads);   // ~4 entries per thread
    const long long cap = (long long)148 * 16 / batch;
    if (chunks > cap) chunks = cap;
    return chunks < 1 ? 1 : chunks;
}

}  // namespace ipoc

// =================================================================== C ABI
using namespace ipoc;


extern "C" {

const char* ipoc_strerror(int code) {
    switch (code) {
        case IPOC_OK: return "ok";
        case IPOC_EUNSUPPORTED_DIM: return "unsupported (nx, nu): no kernel instantiated and there is no CPU fallback";
        case IPOC_EWORKSPACE: return "workspace too small (see ipoc_workspace_bytes)";
        case IPOC_ECUDA: return "CUDA error at kernel launch";
        case IPOC_ENCCL: return "NCCL error";
        case IPOC_EINVAL: return "invalid argument";
        case IPOC_EALIGN: return "pointer not 16-byte aligned";
        default: return "unknown ipoc error";
    }
}

int ipoc_version(void) { return 100; }

int ipoc_supported(int nx, int nu) {
#define X(a, b) if (nx == a && nu == b) return 1;
    IPOC_FOR_PAIRS(X)
#undef X
    return 0;
}

void ipoc_set_tuning(int leaf_chunk, int mid_fanin, int top_max) {
    g_tune.leaf_chunk = leaf_chunk;
    g_tune.mid_fanin = mid_fanin;
    g_tune.top_max = top_max;
}

unsigned long long ipoc_launch_count(void) { return g_launches; }

void ipoc_set_literal_lqt(int on) { g_literal_lqt = on ? 1 : 0; }

int ipoc_workspace_init(void* ws, size_t ws_bytes, ipoc_stream_t stream) {
    if (ws == nullptr) return IPOC_EINVAL;
    const size_t n = ws_bytes < (size_t)IPOC_WS_CONTROL_BYTES ? ws_bytes : (size_t)IPOC_WS_CONTROL_BYTES;
    return cudaMemsetAsync(ws, 0, n, (cudaStream_t)stream) == cudaSuccess ? IPOC_OK : IPOC_ECUDA;
}

void ipoc_set_affine_occupancy(int warps_per_sm) { g_hier.aff_warps_per_sm = warps_per_sm; }

void ipoc_set_hier(int enabled, int group_warps, int serial_top) {
    g_hier.enabled = enabled < 0 ? 0 : (enabled > 4 ? 4 : enabled);
    g_hier.group_warps = group_warps;
    g_hier.serial_top = serial_top;
}

int ipoc_carry_doubles(int kind, int nx) {
    if (nx < 1 || nx > 8) return 0;
    const int sy = nx * (nx + 1) / 2;
    return kind == IPOC_CARRY_RICCATI ? nx * nx + 2 * nx + 2 * sy : nx * nx + nx;
}

size_t ipoc_workspace_bytes(int kind, int N, int nx, int nu, int batch) {
    if (N < 1 || batch < 1) return 0;
    if (kind == IPOC_WS_REDUCTIONS)   // here `nx` carries max(nu, nc)
        return (size_t)batch * reduce_blocks(N, nx > nu ? nx : nu, batch) * 3 * sizeof(double) + 256;
    if (kind == IPOC_WS_COSTATES || kind == IPOC_WS_NEWTON_ATTEMPT) {   // scan + room for the stand-alone reductions
        const size_t a = ipoc_workspace_bytes(kind == IPOC_WS_COSTATES ? IPOC_WS_AFFINE_SCAN : IPOC_WS_NEWTON_STEP, N, nx,
                                              nu, batch);
        return a == 0 ? 0 : a + ipoc_workspace_bytes(IPOC_WS_REDUCTIONS, N, nu, nu, batch);
    }
    // the sharded entry points use the same formula with batch = 1 and forced chunking; take the max
#define X(a) if (nx == a) { size_t s1 = nx_ws_bytes<a>(kind, N, batch, false); \
                            size_t s2 = batch == 1 ? nx_ws_bytes<a>(kind, N, 1, true) : 0; return s1 > s2 ? s1 : s2; }
    IPOC_FOR_NX(X)
#undef X
    return 0;
}

#define CHECK_ARGS(cond) do { if (!(cond)) return IPOC_EINVAL; } while (0)
#define CHECK_ALIGN(p) do { if ((p) != nullptr && !aligned16(p)) return IPOC_EALIGN; } while (0)

int ipoc_newton_step_f64(int N, int nx, int nu, int batch, const double* fx, const double* fu, const double* ru,
                         const double* Q, const double* R, const double* M, const double* reg, double* dx, double* du,
                         double* Kx, double* d, double* pred, int32_t* feasible, void* ws, size_t ws_bytes,
                         ipoc_stream_t stream) {
    CHECK_ARGS(N >= 1 && batch >= 1 && fx && fu && ru && Q && R && M && reg && dx && du && Kx && d && pred && feasible && ws);
    CHECK_ALIGN(fx); CHECK_ALIGN(fu); CHECK_ALIGN(ru); CHECK_ALIGN(Q); CHECK_ALIGN(R); CHECK_ALIGN(M);
    CHECK_ALIGN(dx); CHECK_ALIGN(du); CHECK_ALIGN(Kx); CHECK_ALIGN(d); CHECK_ALIGN(ws);
#define X(a, b) if (nx == a && nu == b) return nxu_newton_step<a, b>(N, batch, fx, fu, ru, Q, R, M, reg, dx, du, Kx, d, pred, feasible, ws, ws_bytes, (cudaStream_t)stream, nullptr);
    IPOC_FOR_PAIRS(X)
#undef X
    return IPOC_EUNSUPPORTED_DIM;
}

int ipoc_lqt_bwd_f64(int N, int nx, int nu, int batch, const double* A, const double* B, const double* c,
                     const double* Xm, const double* U, const double* M, const double* q, const double* p,
                     const double* ST, const double* vT, double* Kx, double* d, double* S, double* v, double* pred,
                     int32_t* feasible, void* ws, size_t ws_bytes, ipoc_stream_t stream) {
    CHECK_ARGS(N >= 1 && batch >= 1 && A && B && Xm && U && M && q && p && ST && Kx && d && pred && feasible && ws);
    CHECK_ARGS((S == nullptr) == (v == nullptr));
    CHECK_ALIGN(A); CHECK_ALIGN(B); CHECK_ALIGN(c); CHECK_ALIGN(Xm); CHECK_ALIGN(U); CHECK_ALIGN(M); CHECK_ALIGN(q);
    CHECK_ALIGN(p); CHECK_ALIGN(Kx); CHECK_ALIGN(d); CHECK_ALIGN(ws);
#define X(a, b) if (nx == a && nu == b) return nxu_lqt_bwd<a, b>(N, batch, A, B, c, Xm, U, M, q, p, ST, vT, Kx, d, S, v, pred, feasible, ws, ws_bytes, (cudaStream_t)stream);
    IPOC_FOR_PAIRS(X)
#undef X
    return IPOC_EUNSUPPORTED_DIM;
}

int ipoc_lqt_fwd_f64(int N, int nx, int nu, int batch, const double* A, const double* B, const double* c,
                     const double* Kx, const double* d, const double* x0, double* u, double* x, void* ws,
                     size_t ws_bytes, ipoc_stream_t stream) {
    CHECK_ARGS(N >= 1 && batch >= 1 && A && B && Kx && d && u && x && ws);
    CHECK_ALIGN(A); CHECK_ALIGN(B); CHECK_ALIGN(c); CHECK_ALIGN(Kx); CHECK_ALIGN(d); CHECK_ALIGN(u); CHECK_ALIGN(x);
    CHECK_ALIGN(ws);
#define X(a, b) if (nx == a && nu == b) return nxu_lqt_fwd<a, b>(N, batch, A, B, c, Kx, d, x0, u, x, ws, ws_bytes, (cudaStream_t)stream);
    IPOC_FOR_PAIRS(X)
#undef X
    return IPOC_EUNSUPPORTED_DIM;
}

int ipoc_affine_scan_f64(int reverse, int transpose, int N, int nx, int batch, const double* F, const double* c,
                         const double* seed, double* out, void* ws, size_t ws_bytes, ipoc_stream_t stream) {
    CHECK_ARGS(N >= 1 && batch >= 1 && F && c && out && ws);
    CHECK_ALIGN(F); CHECK_ALIGN(c); CHECK_ALIGN(out); CHECK_ALIGN(ws);
#define X(a) if (nx == a) return nx_affine_scan<a>(reverse, transpose, N, batch, F, c, seed, out, ws, ws_bytes, (cudaStream_t)stream, nullptr, 0, nullptr, nullptr, nullptr);
    IPOC_FOR_NX(X)
#undef X
    return IPOC_EUNSUPPORTED_DIM;
}

int ipoc_reductions_f64(int N, int nu, int nc, int batch, const double* ru, const double* cu, const double* cons,
                        double* hu_norm, double* cu_norm, int32_t* traj_feasible, const double* rp, double* reg,
                        void* ws, size_t ws_bytes, ipoc_stream_t stream) {
    CHECK_ARGS(N >= 1 && batch >= 1 && ws);
    CHECK_ARGS((ru == nullptr || hu_norm) && (cu == nullptr || cu_norm) && (cons == nullptr || traj_feasible));
    const int width = nu > nc ? nu : nc;
    const int nblk = reduce_blocks(N, width, batch);
    if (ws_bytes < (size_t)batch * nblk * 3 * sizeof(double)) return IPOC_EWORKSPACE;
    cudaStream_t st_ = (cudaStream_t)stream;
    double* partials = (double*)ws;
    if (nblk == 1) {   // small problem: one launch does slice reduction and finalisation
        k_reduce_single<<<batch, kRedSingleThreads, 0, st_>>>(ru, cu, cons, N, nu, nc, hu_norm, cu_norm, traj_feasible, rp,
                                                       reg);
        IPOC_API_LAUNCH_CHECK(st_);
        return IPOC_OK;
    }
    k_reduce_partial<<<(unsigned)((long long)nblk * batch), kRedThreads, 0, st_>>>(ru, cu, cons, N, nu, nc, nblk, partials);
    IPOC_API_LAUNCH_CHECK(st_);
    k_reduce_final<<<batch, kRedThreads, 0, st_>>>(partials, nblk, ru != nullptr, cu != nullptr, cons != nullptr,
                                                   hu_norm, cu_norm, traj_feasible, rp, reg);
    IPOC_API_LAUNCH_CHECK(st_);
    return IPOC_OK;
}

int ipoc_lqr_params_f64(int N, int nx, int nu, int batch, const double* lam, const double* cu, const double* cxx,
                        const double* cuu, const double* cxu, const double* fu, const double* fxx, const double* fuu,
                        const double* fxu, double* ru, double* Q, double* R, double* M, ipoc_stream_t stream) {
    CHECK_ARGS(N >= 1 && batch >= 1 && nx >= 1 && nx <= 8 && nu >= 1 && lam && cu && cxx && cuu && cxu && fu && fxx && fuu && fxu && ru && Q && R && M);
    cudaStream_t st_ = (cudaStream_t)stream;
    const long long n = (long long)N * batch;
    k_lqr_params<<<(unsigned)((n + 127) / 128), 128, 0, st_>>>(N, nx, nu, batch, lam, cu, cxx, cuu, cxu, fu, fxx, fuu, fxu,
                                                              ru, Q, R, M);
    IPOC_API_LAUNCH_CHECK(st_);
    return IPOC_OK;
}

int ipoc_accept_update_f64(int batch, const double* cost, const double* new_cost, const int32_t* traj_feasible,
                           const double* pred, const int32_t* bwd_feasible, const int32_t* active, double* rp,
                           double* r_inc, int32_t* success, double* gain_ratio, ipoc_stream_t stream) {
    CHECK_ARGS(batch >= 1 && cost && new_cost && traj_feasible && pred && bwd_feasible && rp && r_inc && success);
    cudaStream_t st_ = (cudaStream_t)stream;
    k_accept_update<<<(batch + 127) / 128, 128, 0, st_>>>(batch, cost, new_cost, traj_feasible, pred,
                                                                            bwd_feasible, active, rp, r_inc, success,
                                                                            gain_ratio);
    IPOC_API_LAUNCH_CHECK(st_);
    return IPOC_OK;
}

int ipoc_attempt_begin_f64(int batch, const uint8_t* done, const double* rp, const double* cu_norm, int32_t* active,
                           double* reg, ipoc_stream_t stream) {
    CHECK_ARGS(batch >= 1 && rp && cu_norm && active && reg);
    cudaStream_t st_ = (cudaStream_t)stream;
    k_attempt_begin<<<(batch + 127) / 128, 128, 0, st_>>>(batch, done, rp, cu_norm, active, reg);
    IPOC_API_LAUNCH_CHECK(st_);
    return IPOC_OK;
}

int ipoc_trial_point_f64(int N, int nx, int nu, int batch, const double* x, const double* dx, const double* u,
                         const double* du, double* tx, double* tu, ipoc_stream_t stream) {
    CHECK_ARGS(N >= 1 && nx >= 1 && nu >= 1 && batch >= 1 && x && dx && u && du && tx && tu);
    return trial_point_masked(N, nx, nu, batch, nullptr, x, dx, u, du, tx, tu, (cudaStream_t)stream);
}

int ipoc_attempt_commit_f64(int N, int nx, int nu, int batch, const int32_t* active, const int32_t* success,
                            const double* tx, const double* tu, double* keep_x, double* keep_u, int64_t* inner,
                            uint8_t* done, int max_attempts, ipoc_stream_t stream) {
    CHECK_ARGS(N >= 1 && nx >= 1 && nu >= 1 && batch >= 1 && active && success && tx && tu && keep_x && keep_u && inner && done);
    cudaStream_t st_ = (cudaStream_t)stream;
    const long long per_x = (long long)(N + 1) * nx, per_u = (long long)N * nu;
    const long long chunks = copy_chunks(per_x + per_u, batch);
    k_attempt_commit<<<(unsigned)(chunks * batch), kCommitThreads, 0, st_>>>(
        per_x, per_u, (int)chunks, active, success, tx, tu, keep_x, keep_u, (long long*)inner, done, max_attempts);
    IPOC_API_LAUNCH_CHECK(st_);
    return IPOC_OK;
}

int ipoc_newton_advance_f64(int N, int nx, int nu, int batch, const double* hu, uint8_t* inner_done,
                            uint8_t* outer_done, int64_t* inner, int64_t* iteration, int32_t* advanced,
                            const double* tx, const double* tu, double* x, double* u, double hu_tol,
                            int max_iterations, ipoc_stream_t stream) {
    CHECK_ARGS(N >= 1 && nx >= 1 && nu >= 1 && batch >= 1 && hu && inner_done && outer_done && inner && iteration && advanced && tx && tu && x && u);
    cudaStream_t st_ = (cudaStream_t)stream;
    k_newton_advance_flags<<<(batch + 127) / 128, 128, 0, st_>>>(batch, hu, inner_done, outer_done, (long long*)inner,
                                                               (long long*)iteration, advanced, hu_tol,
                                                               max_iterations);
    IPOC_API_LAUNCH_CHECK(st_);
    const long long per_x = (long long)(N + 1) * nx, per_u = (long long)N * nu;
    const long long chunks = copy_chunks(per_x + per_u, batch);
    k_masked_copy2<<<(unsigned)(chunks * batch), kCommitThreads, 0, st_>>>(per_x, per_u, (int)chunks, advanced, tx, tu, x,
                                                                       u);
    IPOC_API_LAUNCH_CHECK(st_);
    return IPOC_OK;
}

int ipoc_attempt_finish_f64(int batch, const double* cost, const double* new_cost, const int32_t* traj_feasible,
                            const double* pred, const int32_t* bwd_feasible, const double* hu, int32_t* active,
                            double* rp, double* r_inc, int32_t* success, double* gain_ratio, int64_t* inner,
                            int64_t* iteration, uint8_t* outer_done, int32_t* advanced, double hu_tol,
                            int max_attempts, int max_iterations, ipoc_stream_t stream) {
    CHECK_ARGS(batch >= 1 && cost && new_cost && traj_feasible && pred && bwd_feasible && hu && active && rp && r_inc && success && inner && iteration && outer_done && advanced);
    cudaStream_t st_ = (cudaStream_t)stream;
    const FinishIO f{AcceptIO{cost, pred, bwd_feasible, rp, r_inc, success, gain_ratio}, hu, active, (long long*)inner,
                     (long long*)iteration, outer_done, advanced, hu_tol, max_attempts, max_iterations};
    k_attempt_finish<<<(batch + 127) / 128, 128, 0, st_>>>(batch, f, new_cost, traj_feasible);
    IPOC_API_LAUNCH_CHECK(st_);
    return IPOC_OK;
}

int ipoc_masked_copy_f64(int N, int nx, int nu, int batch, const int32_t* mask, const double* sx, const double* su,
                         double* dx_, double* du_, ipoc_stream_t stream) {
    CHECK_ARGS(N >= 1 && nx >= 1 && nu >= 1 && batch >= 1 && mask && sx && su && dx_ && du_);
    cudaStream_t st_ = (cudaStream_t)stream;
    const long long per_x = (long long)(N + 1) * nx, per_u = (long long)N * nu;
    const long long chunks = copy_chunks(per_x + per_u, batch);
    k_masked_copy2<<<(unsigned)(chunks * batch), kCommitThreads, 0, st_>>>(per_x, per_u, (int)chunks, mask, sx, su, dx_, du_);
    IPOC_API_LAUNCH_CHECK(st_);
    return IPOC_OK;
}

int ipoc_newton_bwd_reduce_f64(int N, int nx, int nu, const double* fx, const double* fu, const double* ru,
                               const double* Q, const double* R, const double* M, const double* reg,
                               double* carry_out, void* ws, size_t ws_bytes, ipoc_stream_t stream) {
    CHECK_ARGS(N >= 1 && fx && fu && ru && Q && R && M && reg && carry_out && ws);
#define X(a, b) if (nx == a && nu == b) return nxu_newton_bwd_reduce<a, b>(N, fx, fu, ru, Q, R, M, reg, carry_out, ws, ws_bytes, (cudaStream_t)stream);
    IPOC_FOR_PAIRS(X)
#undef X
    return IPOC_EUNSUPPORTED_DIM;
}

int ipoc_newton_bwd_apply_f64(int N, int nx, int nu, int rank, int nranks, const double* fx, const double* fu,
                              const double* ru, const double* Q, const double* R, const double* M, const double* reg,
                              const double* carries, const double* ST, double* Kx, double* d, double* pred,
                              int32_t* feasible, double* fwd_carry_out, void* ws, size_t ws_bytes,
                              ipoc_stream_t stream) {
    CHECK_ARGS(N >= 1 && nranks >= 1 && rank >= 0 && rank < nranks && carries && ST && Kx && d && pred && feasible && fwd_carry_out && ws);
#define X(a, b) if (nx == a && nu == b) return nxu_newton_bwd_apply<a, b>(N, rank, nranks, fx, fu, ru, Q, R, M, reg, carries, ST, Kx, d, pred, feasible, fwd_carry_out, ws, ws_bytes, (cudaStream_t)stream);
    IPOC_FOR_PAIRS(X)
#undef X
    return IPOC_EUNSUPPORTED_DIM;
}

int ipoc_newton_fwd_apply_f64(int N, int nx, int nu, int rank, int nranks, const double* fx, const double* fu,
                              const double* Kx, const double* d, const double* fwd_carries, double* dx, double* du,
                              void* ws, size_t ws_bytes, ipoc_stream_t stream) {
    CHECK_ARGS(N >= 1 && nranks >= 1 && rank >= 0 && rank < nranks && fx && fu && Kx && d && fwd_carries && dx && du && ws);
#define X(a, b) if (nx == a && nu == b) return nxu_newton_fwd_apply<a, b>(N, rank, nranks, fx, fu, Kx, d, fwd_carries, dx, du, ws, ws_bytes, (cudaStream_t)stream);
    IPOC_FOR_PAIRS(X)
#undef X
    return IPOC_EUNSUPPORTED_DIM;
}

int ipoc_affine_reduce_f64(int reverse, int transpose, int N, int nx, const double* F, const double* c,
                           double* carry_out, void* ws, size_t ws_bytes, ipoc_stream_t stream) {
    CHECK_ARGS(N >= 1 && F && c && carry_out && ws);
#define X(a) if (nx == a) return nx_affine_reduce<a>(reverse, transpose, N, F, c, carry_out, ws, ws_bytes, (cudaStream_t)stream);
    IPOC_FOR_NX(X)
#undef X
    return IPOC_EUNSUPPORTED_DIM;
}

int ipoc_affine_apply_f64(int reverse, int transpose, int N, int nx, int rank, int nranks, const double* F,
                          const double* c, const double* carries, const double* seed, double* out, void* ws,
                          size_t ws_bytes, ipoc_stream_t stream) {
    CHECK_ARGS(N >= 1 && nranks >= 1 && rank >= 0 && rank < nranks && F && c && carries && seed && out && ws);
#define X(a) if (nx == a) return nx_affine_apply<a>(reverse, transpose, N, rank, nranks, F, c, carries, seed, out, ws, ws_bytes, (cudaStream_t)stream);
    IPOC_FOR_NX(X)
#undef X
    return IPOC_EUNSUPPORTED_DIM;
}

// ---- fused entry points: the same phases with their neighbouring reductions / glue as side jobs ----------
int ipoc_costates_f64(int N, int nx, int nu, int batch, const double* fx, const double* cx, const double* lamT,
                      const double* cu, double* lam, double* cu_norm, const int32_t* fresh, void* ws, size_t ws_bytes,
                      ipoc_stream_t stream) {
    CHECK_ARGS(N >= 1 && batch >= 1 && nu >= 1 && fx && cx && lam && ws && ((cu == nullptr) == (cu_norm == nullptr)));
    CHECK_ALIGN(fx); CHECK_ALIGN(cx); CHECK_ALIGN(lam); CHECK_ALIGN(ws);
    cudaStream_t st_ = (cudaStream_t)stream;
    int handled = 0, rc = IPOC_EUNSUPPORTED_DIM;
    const size_t scan_bytes = ipoc_workspace_bytes(IPOC_WS_AFFINE_SCAN, N, nx, nu, batch);
    if (scan_bytes == 0) return IPOC_EUNSUPPORTED_DIM;
    if (ws_bytes < scan_bytes) return IPOC_EWORKSPACE;
#define X(a) if (nx == a) rc = nx_affine_scan<a>(1, 1, N, batch, fx, cx, lamT, lam, ws, ws_bytes, st_, cu, nu, cu_norm, &handled, fresh);
    IPOC_FOR_NX(X)
#undef X
    if (rc != IPOC_OK) return rc;
    if (cu != nullptr && !(handled & IPOC_X_NORM)) {   // plan without in-kernel completion: the stand-alone reduction
        const size_t red = ipoc_workspace_bytes(IPOC_WS_REDUCTIONS, N, nu, nu, batch);
        if (ws_bytes < scan_bytes + red) return IPOC_EWORKSPACE;
        return ipoc_reductions_f64(N, nu, 1, batch, nullptr, cu, nullptr, nullptr, cu_norm, nullptr, nullptr, nullptr,
                                   (char*)ws + scan_bytes, ws_bytes - scan_bytes, stream);
    }
    return IPOC_OK;
}

int ipoc_newton_attempt_f64(int N, int nx, int nu, int nc, int batch, const double* fx, const double* fu,
                            const double* ru, const double* Q, const double* R, const double* M, const double* rp,
                            const double* cu_norm, double* dx, double* du, double* Kx, double* d, double* pred,
                            int32_t* feasible, double* hu, const double* x, const double* u, double* tx, double* tu,
                            const double* cons, int32_t* traj_feasible, const double* cost, const double* new_cost,
                            const int32_t* traj_feas_in, const int32_t* active, double* rp_out, double* r_inc,
                            int32_t* success, double* gain_ratio, void* ws, size_t ws_bytes, ipoc_stream_t stream) {
    CHECK_ARGS(N >= 1 && batch >= 1 && fx && fu && ru && Q && R && M && rp && dx && du && Kx && d && pred && feasible && ws);
    CHECK_ARGS((tx == nullptr) == (tu == nullptr) && (tx == nullptr || (x && u)));
    CHECK_ARGS(cons == nullptr || (traj_feasible != nullptr && nc >= 1));
    CHECK_ARGS(rp_out == nullptr || (cost && new_cost && r_inc && success && (cons || traj_feas_in)));
    CHECK_ALIGN(fx); CHECK_ALIGN(fu); CHECK_ALIGN(ru); CHECK_ALIGN(Q); CHECK_ALIGN(R); CHECK_ALIGN(M);
    CHECK_ALIGN(dx); CHECK_ALIGN(du); CHECK_ALIGN(Kx); CHECK_ALIGN(d); CHECK_ALIGN(ws);
    CHECK_ALIGN(x); CHECK_ALIGN(u); CHECK_ALIGN(tx); CHECK_ALIGN(tu);
    cudaStream_t st_ = (cudaStream_t)stream;
    const size_t step_bytes = ipoc_workspace_bytes(IPOC_WS_NEWTON_STEP, N, nx, nu, batch);
    if (step_bytes == 0) return IPOC_EUNSUPPORTED_DIM;
    if (ws_bytes < step_bytes) return IPOC_EWORKSPACE;
    AttemptExtras xt{};
    xt.reg_scale = cu_norm;
    xt.hu = hu;
    xt.x = x; xt.u = u; xt.tx = tx; xt.tu = tu;
    xt.cons = cons; xt.nc = nc; xt.traj_feasible = traj_feasible;
    xt.cost = cost; xt.new_cost = new_cost; xt.traj_feas_in = traj_feas_in; xt.active = active;
    xt.rp = rp_out; xt.r_inc = r_inc; xt.success = success; xt.gain = gain_ratio;
    int rc = IPOC_EUNSUPPORTED_DIM;
#define X(a, b) if (nx == a && nu == b) rc = nxu_newton_step<a, b>(N, batch, fx, fu, ru, Q, R, M, rp, dx, du, Kx, d, pred, feasible, ws, step_bytes, st_, &xt);
    IPOC_FOR_PAIRS(X)
#undef X
    if (rc != IPOC_OK) return rc;
    // whatever the scan plan could not carry as a side job runs as the stand-alone kernels (same results)
    char* rws = (char*)ws + step_bytes;
    const size_t rbytes = ws_bytes - step_bytes;
    if (hu != nullptr && !(xt.handled & IPOC_X_HU))
        if (int r2 = ipoc_reductions_f64(N, nu, 1, batch, ru, nullptr, nullptr, hu, nullptr, nullptr, nullptr, nullptr, rws,
                                         rbytes, stream)) return r2;
    if (tx != nullptr && !(xt.handled & IPOC_X_TRIAL))   // big problems: one coalesced pass beats the leaf kernel's per-lane rows
        if (int r2 = trial_point_masked(N, nx, nu, batch, active, x, dx, u, du, tx, tu, (cudaStream_t)stream)) return r2;
    if (cons != nullptr && !(xt.handled & IPOC_X_CONS))
        if (int r2 = ipoc_reductions_f64(N, nu, nc, batch, nullptr, nullptr, cons, nullptr, nullptr, traj_feasible, nullptr,
                                         nullptr, rws, rbytes, stream)) return r2;
    if (rp_out != nullptr && !(xt.handled & IPOC_X_ACCEPT))
        if (int r2 = ipoc_accept_update_f64(batch, cost, new_cost, cons != nullptr ? traj_feasible : traj_feas_in, pred,
                                            feasible, active, rp_out, r_inc, success, gain_ratio, stream)) return r2;
    return IPOC_OK;
}

size_t ipoc_newton_step_host_scratch_bytes(int N, int nx, int nu, int batch) {
    const size_t per = (size_t)nx * nx * 2 + (size_t)nx * nu * 2 + (size_t)nu * nu + nu;   // inputs
    const size_t outs = (size_t)nx + nu + (size_t)nu * nx + nu;
    size_t b = ((size_t)N * per + (size_t)(N + 1) * outs) * batch * sizeof(double) + 64 * 256;
    b += ipoc_workspace_bytes(IPOC_WS_NEWTON_STEP, N, nx, nu, batch) + 4096 * (size_t)batch;
    return b;
}

int ipoc_newton_step_host_f64(int N, int nx, int nu, int batch, const double* fx, const double* fu, const double* ru,
                              const double* Q, const double* R, const double* M, const double* reg, double* dx,
                              double* du, double* pred, int32_t* feasible, void* dws, size_t dws_bytes,
                              ipoc_stream_t stream) {
    CHECK_ARGS(N >= 1 && batch >= 1 && dws);
    if (dws_bytes < ipoc_newton_step_host_scratch_bytes(N, nx, nu, batch)) return IPOC_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    Bump bp{(char*)dws, 0, dws_bytes, false};
    const size_t T = (size_t)N * batch;
    double* dfx = bp.take<double>(T * nx * nx);
    double* dfu = bp.take<double>(T * nx * nu);
    double* dru = bp.take<double>(T * nu);
    double* dQ = bp.take<double>(T * nx * nx);
    double* dR = bp.take<double>(T * nu * nu);
    double* dM = bp.take<double>(T * nx * nu);
    double* dreg = bp.take<double>(batch);
    double* ddx = bp.take<double>((size_t)(N + 1) * batch * nx);
    double* ddu = bp.take<double>(T * nu);
    double* dKx = bp.take<double>(T * nu * nx);
    double* dd = bp.take<double>(T * nu);
    double* dpred = bp.take<double>(batch);
    int32_t* dfeas = bp.take<int32_t>(batch);
    const size_t wsb = ipoc_workspace_bytes(IPOC_WS_NEWTON_STEP, N, nx, nu, batch);
    void* ws = bp.take<char>(wsb);
    if (bp.off > dws_bytes) return IPOC_EWORKSPACE;
    if (cudaMemsetAsync(ws, 0, IPOC_WS_CONTROL_BYTES, st) != cudaSuccess) return IPOC_ECUDA;   // control block of a fresh carve
    const size_t D = sizeof(double);
    bool ok = true;
    ok &= cudaMemcpyAsync(dfx, fx, T * nx * nx * D, cudaMemcpyHostToDevice, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(dfu, fu, T * nx * nu * D, cudaMemcpyHostToDevice, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(dru, ru, T * nu * D, cudaMemcpyHostToDevice, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(dQ, Q, T * nx * nx * D, cudaMemcpyHostToDevice, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(dR, R, T * nu * nu * D, cudaMemcpyHostToDevice, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(dM, M, T * nx * nu * D, cudaMemcpyHostToDevice, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(dreg, reg, batch * D, cudaMemcpyHostToDevice, st) == cudaSuccess;
    if (!ok) return IPOC_ECUDA;
    int rc = ipoc_newton_step_f64(N, nx, nu, batch, dfx, dfu, dru, dQ, dR, dM, dreg, ddx, ddu, dKx, dd, dpred, dfeas, ws,
                                  wsb, stream);
    if (rc) return rc;
    ok &= cudaMemcpyAsync(dx, ddx, (size_t)(N + 1) * batch * nx * D, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(du, ddu, T * nu * D, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(pred, dpred, batch * D, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(feasible, dfeas, batch * sizeof(int32_t), cudaMemcpyDeviceToHost, st) == cudaSuccess;
    return ok ? IPOC_OK : IPOC_ECUDA;
}


int ipoc_profile_begin(ipoc_stream_t stream) {
    if (!g_prof.created) {
        for (int i = 0; i < kMaxProf; ++i)
            if (cudaEventCreate(&g_prof.ev[i]) != cudaSuccess) return IPOC_ECUDA;
        g_prof.created = true;
    }
    g_prof.n = 0;
    g_prof.armed = true;
    prof_mark("begin", (cudaStream_t)stream);
    return IPOC_OK;
}

int ipoc_profile_end(char* names, size_t names_len, float* ms, int max_entries) {
    g_prof.armed = false;
    if (g_prof.n < 1) return 0;
    if (cudaEventSynchronize(g_prof.ev[g_prof.n - 1]) != cudaSuccess) return IPOC_ECUDA;
    int cnt = 0;
    size_t pos = 0;
    if (names && names_len) names[0] = 0;
    for (int i = 1; i < g_prof.n && cnt < max_entries; ++i, ++cnt) {
        float t = 0.f;
        cudaEventElapsedTime(&t, g_prof.ev[i - 1], g_prof.ev[i]);
        ms[cnt] = t;
        if (names) {
            const size_t l = strlen(g_prof.name[i]);
            if (pos + l + 2 < names_len) {
                memcpy(names + pos, g_prof.name[i], l);
                pos += l;
                names[pos++] = ',';
                names[pos] = 0;
            }
        }
    }
    return cnt;
}

}  // extern "C"
