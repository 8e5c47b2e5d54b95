#pragma once
// Hand-written sm_100a FP64 kernels for the par IP-Newton hot path (templated on NX, NU).
// Instantiated once per NX by ipoc_nx.cu; the C ABI lives in ipoc_api.cu.
//
// Scan organisation (all three scans: K1 costates, K2 Riccati, K3 forward): hierarchical
// reduce / seeded re-scan.
//   leaf-up   : a thread folds `T0` consecutive time steps sequentially (work-optimal); the 32
//               thread aggregates of a warp are then scanned (Kogge-Stone, operands exchanged
//               through the warp's shared scratch) inside the same kernel, where the latency hides
//               behind the other resident warps; every thread stores its in-warp inclusive
//               aggregate, every warp its total.
//   levels    : warp totals are folded by mid kernels only when there are more than `top_max` of
//               them; a single CTA per sequence scans the rest (in-warp scans + a short serial
//               chain of `apply` across warps).
//   leaf-down : only VALUES travel down ((S, v) for K2, a state vector for K1/K3): a thread applies
//               its neighbour's stored in-warp aggregate to the value entering its warp and re-walks
//               its chunk with the cheap seeded recursion, emitting outputs.
// Global loads of the leaf kernels are warp-cooperative cp.async copies into a per-warp
// shared-memory stage (rows padded to an odd number of 16-byte units -> conflict-free LDS.128):
// each lane needs ITS OWN chunk's time step, i.e. a stride-T0 gather; fetching it with per-lane
// loads costs 32 L1 wavefronts per instruction and made the first version L1-bound (profiles/r01a).
// Aggregates live in SoA planes (component-major).  There is no spin-waiting between CTAs, so
// every call is graph-capturable and cannot hang.  With enough independent problems (`batch`) the
// plan degenerates to one sequence per lane: a single pass, no up-sweep at all.
//
// K2's down-sweep emits K3's leaf aggregates for free (same chunks), so one Newton step reads
// fx, fu, Q, R, M, ru twice (the second time from L2 when the working set fits) and Kx, d once.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "ipoc_math.cuh"
#include "../../include/ipoc.h"
#include "ipoc_dispatch.h"
#include "ipoc_accept.cuh"

namespace ipoc {

struct Tuning {
    int leaf_chunk, mid_fanin, top_max;
};
// shared, defined in ipoc_api.cu
extern unsigned long long g_launches;
extern Tuning g_tune;
extern int g_literal_lqt;   // 0: q = 0, p = ru (default); 1: literal noc_to_lqt arithmetic
struct HierTuning {
    int enabled;      // 1 (default): levels above the warps are completed inside the leaf kernels (last arriver)
    int group_warps;  // warps per group (0 = 32)
    int serial_top;   // up to this many groups the top is a serial chain of `apply` (0 = 8)
    int aff_warps_per_sm;   // leaf warps per SM the stand-alone affine scans (K1) are planned for (0 = default)
};
extern HierTuning g_hier;
void prof_mark(const char* name, cudaStream_t st);   // no-op unless profiling is armed

constexpr int kLeafThreads = 128;
#ifndef IPOC_NS_RIC_UP
#define IPOC_NS_RIC_UP 1
#endif
#ifndef IPOC_NS_RIC_DOWN
#define IPOC_NS_RIC_DOWN 1
#endif
#ifndef IPOC_NS_LIGHT
#define IPOC_NS_LIGHT 3
#endif
#ifndef IPOC_RIC_MINB
#define IPOC_RIC_MINB 1
#endif
constexpr int kMidThreads = 128;
constexpr int kTopThreads = 256;
constexpr int kTargetThreads = 148 * 256;       // K2/K3: 8 warps per SM (register-limited)
constexpr int kAffTargetThreads = 148 * 256;    // K1: light kernels are bytes-in-flight limited
static inline int aff_target_threads() {
    return g_hier.aff_warps_per_sm > 0 ? 148 * 32 * g_hier.aff_warps_per_sm : kAffTargetThreads;
}

// ------------------------------------------------------------------ SoA helpers
template <class T>
IPOC_DEV void soa_load(T& t, const double* __restrict__ base, size_t stride, size_t idx) {
    constexpr int SZ = sizeof(T) / sizeof(double);
#pragma unroll
    for (int c = 0; c < SZ; ++c) t.r[c] = base[(size_t)c * stride + idx];
}
// same, through L2 (data another SM may have written during this kernel)
template <class T>
IPOC_DEV void soa_load_cg(T& t, const double* base, size_t stride, size_t idx) {
    constexpr int SZ = sizeof(T) / sizeof(double);
#pragma unroll
    for (int c = 0; c < SZ; ++c) t.r[c] = __ldcg(base + (size_t)c * stride + idx);
}
template <class T>
IPOC_DEV void soa_store(const T& t, double* __restrict__ base, size_t stride, size_t idx) {
    constexpr int SZ = sizeof(T) / sizeof(double);
#pragma unroll
    for (int c = 0; c < SZ; ++c) base[(size_t)c * stride + idx] = t.r[c];
}
// contiguous per-step loads: CNT doubles at p (16-byte aligned when CNT is even)
template <int CNT>
IPOC_DEV void ld_vec(double* dst, const double* __restrict__ p) {
    if constexpr (CNT % 2 == 0) {
        const double2* p2 = reinterpret_cast<const double2*>(p);
#pragma unroll
        for (int i = 0; i < CNT / 2; ++i) {
            const double2 v = __ldg(p2 + i);
            dst[2 * i] = v.x;
            dst[2 * i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < CNT; ++i) dst[i] = __ldg(p + i);
    }
}
template <int CNT>
IPOC_DEV void st_vec(double* __restrict__ p, const double* src) {
    if constexpr (CNT % 2 == 0) {
        double2* p2 = reinterpret_cast<double2*>(p);
#pragma unroll
        for (int i = 0; i < CNT / 2; ++i) p2[i] = make_double2(src[2 * i], src[2 * i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < CNT; ++i) p[i] = src[i];
    }
}

// In-warp inclusive scan of one aggregate per lane, in SCAN order, through a per-warp shared
// scratch `ws` of 2 x ESZ x 32 doubles (two component-major buffers, ping-pong): both operands of
// every combine are read from shared memory and the result goes back to shared memory, so no
// aggregate has to live in registers during the scan.  Returns the buffer holding the result.
// reverse = true : scan order runs from lane 31 down to lane 0 (backward-in-time scans);
//                  afterwards lane l holds the composition of lanes 31..l, lane 0 the warp total.
// reverse = false: lane l holds the composition of lanes 0..l, lane 31 the warp total.
template <class Op>
constexpr size_t scan_scratch_bytes() { return 2 * sizeof(typename Op::Elem) * 32; }

// `used` = number of leading lanes that hold real aggregates (the rest are identities, forward scans
// only): rounds with delta >= used would only copy, so they are skipped — a top scan over 8 warp totals
// (N = 1000) takes 3 combine rounds instead of 5.  Must be the same in every warp of a CTA (buffer parity).
template <class Op>
IPOC_DEV double* warp_scan_mem(double* ws, int lane, bool reverse, int used = 32) {
    constexpr int ESZ = sizeof(typename Op::Elem) / sizeof(double);
    int cur = 0;
#pragma unroll 1
    for (int delta = 1; delta < used; delta <<= 1) {
        const double* src = ws + cur * (ESZ * 32);
        double* dst = ws + (cur ^ 1) * (ESZ * 32) + lane;
        const int partner = reverse ? lane + delta : lane - delta;
        if (partner >= 0 && partner < 32) {
            Op::compose_mm(dst, 32, src + partner, 32, src + lane, 32);
        } else {
#pragma unroll
            for (int c = 0; c < ESZ; ++c) dst[c * 32] = src[c * 32 + lane];
        }
        __syncwarp();
        cur ^= 1;
    }
    return ws + cur * (ESZ * 32);
}
// Register-resident variant for the leaf kernels (many warps per SM, code stays hot): the combine
// is inlined, both operands are read from one shared buffer, the result stays in registers.
template <class Op>
IPOC_DEV void warp_scan(typename Op::Elem& a, double* ws, int lane, bool reverse, int used = 32) {
    constexpr int ESZ = sizeof(typename Op::Elem) / sizeof(double);
    using View = typename Op::View;
#pragma unroll 1
    for (int delta = 1; delta < used; delta <<= 1) {
#pragma unroll
        for (int c = 0; c < ESZ; ++c) ws[c * 32 + lane] = a.r[c];
        __syncwarp();
        const int partner = reverse ? lane + delta : lane - delta;
        if (partner >= 0 && partner < 32) Op::compose_t(a, View{ws + partner, 32}, View{ws + lane, 32});
        __syncwarp();
    }
}

// ------------------------------------------------------------------ generic mid / top kernels
template <class Op>
__global__ void __launch_bounds__(kMidThreads)
k_mid_up(const double* __restrict__ in, size_t istride, int n_in,
         double* __restrict__ out, size_t ostride, int n_out, int T, int batch) {
    constexpr int ESZ = sizeof(typename Op::Elem) / sizeof(double);
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)batch * n_out) return;
    const int b = (int)(g / n_out), j2 = (int)(g % n_out);
    const int j0 = j2 * T, j1 = min(n_in, j0 + T);
    const size_t base = (size_t)b * n_in;
    double acc[ESZ];   // running aggregate (thread-local memory, read through a stride-1 view)
#pragma unroll
    for (int c = 0; c < ESZ; ++c) acc[c] = in[(size_t)c * istride + base + j0];
    for (int j = j0 + 1; j < j1; ++j) Op::compose_mm(acc, 1, acc, 1, in + base + j, (int)istride);
#pragma unroll
    for (int c = 0; c < ESZ; ++c) out[(size_t)c * ostride + (size_t)g] = acc[c];
}

template <class Op>
__global__ void __launch_bounds__(kMidThreads)
k_mid_down(const double* __restrict__ agg, size_t astride, int n_in,
           double* __restrict__ vals_in, size_t vistride,
           const double* __restrict__ vals_out, size_t vostride, int n_out, int T, int batch) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)batch * n_out) return;
    const int b = (int)(g / n_out), j2 = (int)(g % n_out);
    const int j0 = j2 * T, j1 = min(n_in, j0 + T);
    const size_t base = (size_t)b * n_in;
    typename Op::Val v;
    soa_load(v, vals_out, vostride, (size_t)g);
    for (int j = j0; j < j1; ++j) {
        soa_store(v, vals_in, vistride, base + j);
        if (j + 1 < j1) Op::apply_mm(v, agg + base + j, (int)astride);
    }
}

// fixed-order folds of per-warp partials, done by the warp that completes a sequence
struct SideJobs {
    int n;                        // partials per sequence (= nW)
    const double* pred_part;      // K2 down-sweep: sum d'Gd, AND (G > 0)       -> pred, feasible
    const int* feas_part;
    double* pred;
    int32_t* feasible;
    const double* sq_part;        // K1 up-sweep: sum cu^2                      -> ||cu||_F
    double* cu_norm;
    const double* mx_part;        // K2 up-sweep: max|ru| (NaN-propagating)     -> hu
    double* hu;
};
IPOC_DEV double nan_max_d(double a, double c) {
    return (a != a || c != c) ? __longlong_as_double(0x7ff8000000000000LL) : fmax(a, c);
}
// The partials come from L2 (other SMs wrote them): loads are issued eight at a time before they are consumed, so
// a fold over n partials costs n/256 memory round trips instead of n/32; the accumulation order is fixed.
template <class T, class F>
IPOC_DEV void fold_partials(const T* p, int n, int lane, F&& consume) {
    constexpr int U = 8;
    for (int j0 = lane; j0 < n; j0 += 32 * U) {
        T v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int j = j0 + 32 * k;
            if (j < n) v[k] = __ldcg(p + j);
        }
#pragma unroll
        for (int k = 0; k < U; ++k)
            if (j0 + 32 * k < n) consume(v[k]);
    }
}
IPOC_DEV void side_finish(const SideJobs& sj, int b, int lane) {   // one full warp
    if (sj.pred != nullptr) {
        double acc = 0.0;
        int f = 1;
        fold_partials(sj.pred_part + (size_t)b * sj.n, sj.n, lane, [&](double v) { acc += v; });
        fold_partials(sj.feas_part + (size_t)b * sj.n, sj.n, lane, [&](int v) { f &= v; });
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
            f &= __shfl_xor_sync(0xffffffffu, f, o);
        }
        if (lane == 0) {
            sj.pred[b] = -0.5 * acc;
            sj.feasible[b] = f;
        }
    }
    if (sj.cu_norm != nullptr) {
        double acc = 0.0;
        fold_partials(sj.sq_part + (size_t)b * sj.n, sj.n, lane, [&](double v) { acc += v; });
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) sj.cu_norm[b] = sqrt(acc);
    }
    if (sj.hu != nullptr) {
        double m = 0.0;
        fold_partials(sj.mx_part + (size_t)b * sj.n, sj.n, lane, [&](double v) { m = nan_max_d(m, v); });
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = nan_max_d(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) sj.hu[b] = m;
    }
}

// One CTA per sequence.  vals[j] = value ENTERING aggregate j (i.e. after aggregates 0..j-1 were
// applied to the seed).  Optionally writes the composition of all n aggregates to `total`
// (component-major with stride `batch`) and/or skips the value pass (reduce_only).
// Dynamic shared memory: per warp 2 x ESZ x 32 doubles of scan scratch (see warp_scan_mem); the
// buffer holding the warp's inclusive aggregates stays valid afterwards, so warp totals and
// neighbours are read from there.
template <class Op>
__global__ void __launch_bounds__(kTopThreads)
k_top(const double* __restrict__ agg, size_t astride, int n, int batch,
      const double* __restrict__ seed, double* __restrict__ vals, size_t vstride,
      double* __restrict__ total, int reduce_only, SideJobs sj) {
    using Val = typename Op::Val;
    constexpr int ESZ = sizeof(typename Op::Elem) / sizeof(double);
    constexpr int VSZ = sizeof(Val) / sizeof(double);
    constexpr int MAXW = kTopThreads / 32;
    constexpr int WSZ = 2 * ESZ * 32;
    extern __shared__ __align__(16) double s_scan[];   // [nw][2][ESZ][32]
    __shared__ double s_v[MAXW][VSZ];
    __shared__ double s_tot[ESZ];

    const int b = blockIdx.x, t = threadIdx.x, nt = blockDim.x;
    const int lane = t & 31, w = t >> 5, nw = nt >> 5;
    const int q = (n + nt - 1) / nt;
    const size_t base = (size_t)b * n;
    const int j0 = t * q, j1 = min(n, j0 + q);
    double* ws = s_scan + (size_t)w * WSZ;
    if (w == nw - 1) side_finish(sj, b, lane);   // side jobs of the sequence (saves a launch each)

    // thread-level fold of its q aggregates, accumulated in place in the scan scratch
    if (j0 < n) {
#pragma unroll
        for (int c = 0; c < ESZ; ++c) ws[c * 32 + lane] = agg[(size_t)c * astride + base + j0];
        for (int j = j0 + 1; j < j1; ++j)
            Op::compose_mm(ws + lane, 32, ws + lane, 32, agg + base + j, (int)astride);
    } else {
        typename Op::Elem id;
        Op::identity(id);
#pragma unroll
        for (int c = 0; c < ESZ; ++c) ws[c * 32 + lane] = id.r[c];
    }
    __syncwarp();
    // this warp's inclusive aggregates; with fewer than 32 aggregates in the whole CTA the late rounds are skipped
    const int used = (n + q - 1) / q;
    const double* inc = warp_scan_mem<Op>(ws, lane, false, used < 32 ? used : 32);
    const int inc_off = (int)(inc - ws);                      // same buffer parity in every warp
    __syncthreads();
    if (w == 0 && lane == 0) {
        if (total != nullptr) {
            // composition of all warp totals (time-sharded reduce phase)
#pragma unroll
            for (int c = 0; c < ESZ; ++c) s_tot[c] = s_scan[inc_off + c * 32 + (used < 32 ? used - 1 : 31)];
            for (int ww = 1; ww < nw; ++ww)
                Op::compose_mm(s_tot, 1, s_tot, 1, s_scan + (size_t)ww * WSZ + inc_off + 31, 32);
#pragma unroll
            for (int c = 0; c < ESZ; ++c) total[(size_t)c * batch + b] = s_tot[c];
        }
        if (!reduce_only) {
            // value entering every warp: a short serial chain of `apply` (about half a combine each)
            Val v;
            soa_load(v, seed, (size_t)batch, (size_t)b);
            for (int ww = 1; ww < nw; ++ww) {
                Op::apply_mm(v, s_scan + (size_t)(ww - 1) * WSZ + inc_off + 31, 32);
#pragma unroll
                for (int c = 0; c < VSZ; ++c) s_v[ww - 1][c] = v.r[c];
            }
        }
    }
    if (reduce_only) return;
    __syncthreads();
    Val v;
    if (w == 0) {
        soa_load(v, seed, (size_t)batch, (size_t)b);
    } else {
#pragma unroll
        for (int c = 0; c < VSZ; ++c) v.r[c] = s_v[w - 1][c];
    }
    if (lane > 0) Op::apply_mm(v, inc + lane - 1, 32);   // exclusive prefix = previous lane's inclusive
    if (j0 < n) {
        for (int j = j0; j < j1; ++j) {
            soa_store(v, vals, vstride, base + j);
            if (j + 1 < j1) Op::apply_mm(v, agg + base + j, (int)astride);
        }
    }
}

// Sequentially push a seed through gathered segment aggregates (time-sharded mode):
// carries are rank-major AoS: carry[r * ESZ + c].  first = index of the first aggregate to
// apply, step = +1/-1, count = how many.
template <class Op>
__global__ void k_chain_seed(const double* __restrict__ carries, int first, int step, int count,
                             const double* __restrict__ seed_in, double* __restrict__ seed_out) {
    using Val = typename Op::Val;
    constexpr int ESZ = sizeof(typename Op::Elem) / sizeof(double);
    constexpr int VSZ = sizeof(Val) / sizeof(double);
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Val v;
#pragma unroll
    for (int c = 0; c < VSZ; ++c) v.r[c] = seed_in[c];
    for (int i = 0, r = first; i < count; ++i, r += step) Op::apply_mm(v, carries + (size_t)r * ESZ, 1);
#pragma unroll
    for (int c = 0; c < VSZ; ++c) seed_out[c] = v.r[c];
}


// ------------------------------------------------------------------ levels inside the leaf kernels
// "Last arriver works" completion of a hierarchical scan, WITHOUT spin-waiting (a launch can never hang
// and stays graph-capturable): the warps of a sequence are grouped by `gw` consecutive warps in scan
// order.  A warp that has stored its total arrives on its group's counter; the last arriver of a group
// scans the group's (<= 32) warp totals in place of a mid-level kernel and arrives on the sequence's
// counter; the last group scans the group totals in place of the single-CTA top kernel and finishes the
// side jobs (fixed-order folds of per-warp partials).  Group scans of early groups overlap with the leaf
// work of later ones.  The seeded way down needs no kernel of its own either: every warp of the
// down-sweep re-derives the value entering it with at most two `apply`s (hier_enter).
//
// Arrival counters are 32-bit words of the workspace's CONTROL BLOCK (its first kCtrlBytes): the caller
// zero-fills it once (ipoc_workspace_init) and every launch leaves it zero again — the count is an atomic
// increment that wraps to 0 on the last arrival — so replays of captured graphs and calls with other plans
// on the same workspace always find zeros.  One acq_rel atomic per warp: its release side publishes what the
// warp's lanes wrote before the __syncwarp (cumulativity), its acquire side orders the last arriver's reads.
constexpr int kCtrlBytes = 64 * 1024;
constexpr int kCtrlRegionWords = 4096;     // three regions: backward/primary scan, forward scan, tail job
IPOC_DEV bool warp_arrive_last(unsigned* w, unsigned expected, int lane) {
    __syncwarp();
    unsigned old = 0;
    if (lane == 0)
        asm volatile("atom.acq_rel.gpu.global.inc.u32 %0, [%1], %2;" : "=r"(old) : "l"(w), "r"(expected - 1u) : "memory");
    old = __shfl_sync(0xffffffffu, old, 0);
    __syncwarp();
    return old == expected - 1u;
}

struct Hier {
    unsigned* cnt;             // NULL = levels run as separate kernels
    int gw, ngroups, serial_top;
    double* incl1;             // in-group inclusive aggregates of the warp totals, index b*nW + s   (stride s1)
    size_t s1;
    double* agg2;              // group totals, index b*ngroups + g                                  (stride s2)
    double* incl2;             // inclusive aggregates over the groups (split-phase / wide tops)     (stride s2)
    size_t s2;
    double* val2;              // value entering each group, index b*ngroups + g                     (stride s2)
    const double* seed;        // SoA, stride batch; NULL = reduce only (time-sharded phase 1)
    double* total;             // composition of the whole sequence, SoA stride batch; may be NULL
};

// Called by every warp of a leaf kernel after it has stored its total at agg0[b*nW + s] (s = index of
// the warp in scan order).  `scr` = the warp's shared scratch, 2 x ESZ x 32 doubles.
template <class Op>
IPOC_DEV void hier_finish(const Hier& h, int nW, int batch, const double* agg0, size_t a0stride, int b, int s, int lane,
                          double* scr, const SideJobs& sj) {
    using Val = typename Op::Val;
    constexpr int ESZ = sizeof(typename Op::Elem) / sizeof(double);
    const int grp = s / h.gw;
    const int g0 = grp * h.gw;
    const int gsize = min(h.gw, nW - g0);
    if (gsize > 1 && !warp_arrive_last(h.cnt + (size_t)b * h.ngroups + grp, (unsigned)gsize, lane)) return;
    // ---- level 2: inclusive scan of this group's warp totals
    {
        typename Op::Elem id;
        Op::identity(id);
#pragma unroll
        for (int c = 0; c < ESZ; ++c)
            scr[c * 32 + lane] = (lane < gsize) ? __ldcg(agg0 + (size_t)c * a0stride + (size_t)b * nW + g0 + lane) : id.r[c];
        __syncwarp();
        const double* inc = warp_scan_mem<Op>(scr, lane, false, gsize);
        if (lane < gsize) {
#pragma unroll
            for (int c = 0; c < ESZ; ++c) h.incl1[(size_t)c * h.s1 + (size_t)b * nW + g0 + lane] = inc[c * 32 + lane];
        }
        if (lane == gsize - 1) {
#pragma unroll
            for (int c = 0; c < ESZ; ++c) h.agg2[(size_t)c * h.s2 + (size_t)b * h.ngroups + grp] = inc[c * 32 + lane];
        }
        __syncwarp();
    }
    if (h.ngroups > 1)
        if (!warp_arrive_last(h.cnt + (size_t)batch * h.ngroups + b, (unsigned)h.ngroups, lane)) return;
    // ---- level 3: this warp completes the sequence
    const size_t gb = (size_t)b * h.ngroups;
    const bool wide = h.ngroups > h.serial_top || h.total != nullptr || h.seed == nullptr;
    if (!wide) {
        // few groups: values travel through a short serial chain of `apply` (about half a combine each);
        // every lane does the same arithmetic, lane 0 stores.  All loads (group totals and seed through L2:
        // other SMs wrote them) are issued before the side jobs' folds so that their round trips overlap.
        double tot[ESZ];
        if (lane < h.ngroups) {
#pragma unroll
            for (int c = 0; c < ESZ; ++c) tot[c] = __ldcg(h.agg2 + (size_t)c * h.s2 + gb + lane);
        }
        Val v;
        soa_load_cg(v, h.seed, (size_t)batch, (size_t)b);
        side_finish(sj, b, lane);
        if (lane < h.ngroups) {
#pragma unroll
            for (int c = 0; c < ESZ; ++c) scr[c * 32 + lane] = tot[c];
        }
        __syncwarp();
        for (int g = 0; g < h.ngroups; ++g) {
            if (lane == 0) soa_store(v, h.val2, h.s2, gb + g);
            if (g + 1 < h.ngroups) Op::apply_mm(v, scr + g, 32);
        }
        return;
    }
    side_finish(sj, b, lane);
    // many groups (or the composition of everything is wanted): fold q consecutive group totals per lane,
    // scan the lanes, then walk the q groups again with the value entering the lane
    const int q = (h.ngroups + 31) / 32;
    const int j0 = lane * q, j1 = min(h.ngroups, j0 + q);
    if (j0 < h.ngroups) {
#pragma unroll
        for (int c = 0; c < ESZ; ++c) scr[c * 32 + lane] = __ldcg(h.agg2 + (size_t)c * h.s2 + gb + j0);
        for (int j = j0 + 1; j < j1; ++j) Op::compose_mm(scr + lane, 32, scr + lane, 32, h.agg2 + gb + j, (int)h.s2);
    } else {
        typename Op::Elem id;
        Op::identity(id);
#pragma unroll
        for (int c = 0; c < ESZ; ++c) scr[c * 32 + lane] = id.r[c];
    }
    __syncwarp();
    const int used = (h.ngroups + q - 1) / q;
    const double* inc = warp_scan_mem<Op>(scr, lane, false, used);
    if (q == 1 && h.incl2 != nullptr && lane < h.ngroups) {
#pragma unroll
        for (int c = 0; c < ESZ; ++c) h.incl2[(size_t)c * h.s2 + gb + lane] = inc[c * 32 + lane];
    }
    if (h.total != nullptr && lane == used - 1) {
#pragma unroll
        for (int c = 0; c < ESZ; ++c) h.total[(size_t)c * batch + b] = inc[c * 32 + lane];
    }
    if (h.seed == nullptr) return;
    Val v;
    soa_load_cg(v, h.seed, (size_t)batch, (size_t)b);
    if (lane > 0 && lane < used) Op::apply_mm(v, inc + lane - 1, 32);
    if (j0 < h.ngroups) {
        for (int j = j0; j < j1; ++j) {
            soa_store(v, h.val2, h.s2, gb + j);
            if (j + 1 < j1) Op::apply_mm(v, h.agg2 + gb + j, (int)h.s2);
        }
    }
}

// ------------------------------------------------------------------ lane-cooperative Riccati levels
// The levels of the Riccati scan (warp totals -> value entering every warp) when the level scans do NOT run
// inside the up-sweep: one CTA per group of <= 32 warp totals, every combine of a Kogge-Stone round done by
// coop_gs<NX>() cooperating lanes (ric_compose_coop), the groups of a sequence on different SMs; the CTA
// that arrives last on the sequence's counter scans the (<= 32) group totals the same way and pushes the
// seed through them.  Writes the same arrays as hier_finish (incl1, agg2, val2), so the down-sweep enters
// through hier_enter.  Replaces the single-CTA k_top<RicOp> whose rounds cost one full single-thread combine.
template <int NX>
__host__ __device__ constexpr int coop_gs() { return NX <= 1 ? 1 : (NX <= 2 ? 2 : (NX <= 4 ? 4 : 8)); }

// Inclusive scan (scan order = slot order) of n <= 32 aggregates held component-major in buf[0] of a
// 2 x ESZ x 32 ping-pong; all 32 * GS threads of the CTA call it.  Returns the buffer holding the result.
template <int NX>
static __device__ __noinline__ int coop_scan(double* buf, int slot, int g, int n) {   // one copy: cold code is fetched once
    constexpr int ESZ = RicElem<NX>::ESZ, GS = coop_gs<NX>();
    int cur = 0;
#pragma unroll 1
    for (int delta = 1; delta < n; delta <<= 1) {
        const double* src = buf + cur * (ESZ * 32);
        double* dst = buf + (cur ^ 1) * (ESZ * 32);
        if (slot >= delta && slot < n) {
            if (g < NX) ric_compose_coop<NX>(dst + slot, src + slot - delta, src + slot, g);
        } else {
            for (int c = g; c < ESZ; c += GS) dst[c * 32 + slot] = src[c * 32 + slot];
        }
        __syncthreads();
        cur ^= 1;
    }
    return cur;
}

template <int NX>
__global__ void __launch_bounds__(32 * coop_gs<NX>())
k_ric_top_coop(const double* __restrict__ agg0, size_t a0stride, int nW, int batch, Hier h, SideJobs sj) {
    using Op = RicOp<NX>;
    using Val = typename Op::Val;
    constexpr int ESZ = RicElem<NX>::ESZ, GS = coop_gs<NX>();
    extern __shared__ __align__(16) double s_buf[];   // [2][ESZ][32]
    __shared__ int s_last;
    const int b = blockIdx.x / h.ngroups, grp = blockIdx.x % h.ngroups;
    const int t = threadIdx.x, slot = t / GS, g = t % GS;
    const int g0 = grp * h.gw, gsize = min(h.gw, nW - g0);
    // component c of the identity element (A = I, the rest zero) without a runtime-indexed local array
    auto idc = [](int c) { return (c < NX * NX && c / NX == c % NX) ? 1.0 : 0.0; };
    for (int c = g; c < ESZ; c += GS)
        s_buf[c * 32 + slot] = (slot < gsize) ? agg0[(size_t)c * a0stride + (size_t)b * nW + g0 + slot] : idc(c);
    __syncthreads();
    {
        const double* inc = s_buf + coop_scan<NX>(s_buf, slot, g, gsize) * (ESZ * 32);
        if (slot < gsize)
            for (int c = g; c < ESZ; c += GS) h.incl1[(size_t)c * h.s1 + (size_t)b * nW + g0 + slot] = inc[c * 32 + slot];
        if (slot == gsize - 1) {
            for (int c = g; c < ESZ; c += GS) h.agg2[(size_t)c * h.s2 + (size_t)b * h.ngroups + grp] = inc[c * 32 + slot];
            if (h.total != nullptr && h.ngroups == 1)   // a single group: its total is the sequence's
                for (int c = g; c < ESZ; c += GS) h.total[(size_t)c * batch + b] = inc[c * 32 + slot];
        }
    }
    const size_t gb = (size_t)b * h.ngroups;
    if (h.ngroups > 1) {
        // the CTA's stores happen-before the barrier, the barrier before thread 0's release increment
        __syncthreads();
        if (t == 0) {
            unsigned old = 0;
            asm volatile("atom.acq_rel.gpu.global.inc.u32 %0, [%1], %2;"
                         : "=r"(old) : "l"(h.cnt + (size_t)batch * h.ngroups + b), "r"((unsigned)h.ngroups - 1u) : "memory");
            s_last = (old == (unsigned)h.ngroups - 1u) ? 1 : 0;
        }
        __syncthreads();
        if (!s_last) return;
    }
    if (t < 32) side_finish(sj, b, t);
    if (h.ngroups == 1) {
        if (t == 0 && h.seed != nullptr) {
            Val v;
            soa_load_cg(v, h.seed, (size_t)batch, (size_t)b);
            soa_store(v, h.val2, h.s2, gb);
        }
        return;
    }
    for (int c = g; c < ESZ; c += GS)
        s_buf[c * 32 + slot] = (slot < h.ngroups) ? __ldcg(h.agg2 + (size_t)c * h.s2 + gb + slot) : idc(c);
    __syncthreads();
    const double* inc = s_buf + coop_scan<NX>(s_buf, slot, g, h.ngroups) * (ESZ * 32);
    if (h.incl2 != nullptr && slot < h.ngroups)
        for (int c = g; c < ESZ; c += GS) h.incl2[(size_t)c * h.s2 + gb + slot] = inc[c * 32 + slot];
    if (h.total != nullptr && slot == h.ngroups - 1)   // composition of the whole sequence (time-sharded reduce phase)
        for (int c = g; c < ESZ; c += GS) h.total[(size_t)c * batch + b] = inc[c * 32 + slot];
    if (h.seed == nullptr) return;                     // reduce only: the values come with the apply phase
    if (g == 0 && slot < h.ngroups) {   // value entering group `slot`: the seed through the groups before it
        Val v;
        soa_load_cg(v, h.seed, (size_t)batch, (size_t)b);
        if (slot > 0) Op::apply_mm(v, inc + slot - 1, 32);
        soa_store(v, h.val2, h.s2, gb + slot);
    }
}

// Way down: the value entering warp s of sequence b (every lane computes the same thing).
//   chain (time-sharded phase 2): the seed of the whole horizon is first pushed through the gathered
//   aggregates of the other ranks (rank-major AoS), then through this rank's inclusive group aggregates.
struct HierIn {
    int on;                    // 0 = values come from the legacy level kernels / the seed
    int gw, ngroups;
    const double* incl1;
    size_t s1;
    const double* val2;        // NULL in chain mode
    const double* incl2;
    size_t s2;
    const double* carries;     // chain mode: gathered rank aggregates [r * ESZ + c]
    int chain_first, chain_step, chain_count;
    const double* seed0;       // chain mode: packed Val of the horizon's end (device, VSZ doubles); NULL = zeros
};
template <class Op>
IPOC_DEV void hier_enter(typename Op::Val& v, const HierIn& h, int nW, int b, int s) {
    constexpr int ESZ = sizeof(typename Op::Elem) / sizeof(double);
    constexpr int VSZ = sizeof(typename Op::Val) / sizeof(double);
    const int grp = s / h.gw, r = s - grp * h.gw;
    if (h.val2 != nullptr) {
        soa_load(v, h.val2, h.s2, (size_t)b * h.ngroups + grp);
    } else {
#pragma unroll
        for (int c = 0; c < VSZ; ++c) v.r[c] = (h.seed0 != nullptr) ? h.seed0[c] : 0.0;
        for (int i = 0, rk = h.chain_first; i < h.chain_count; ++i, rk += h.chain_step)
            Op::apply_mm(v, h.carries + (size_t)rk * ESZ, 1);
        if (grp > 0) Op::apply_mm(v, h.incl2 + (size_t)b * h.ngroups + grp - 1, (int)h.s2);
    }
    if (r > 0) Op::apply_mm(v, h.incl1 + (size_t)b * nW + s - 1, (int)h.s1);
}

// ------------------------------------------------------------------ leaf geometry
// mode A (per_lane = 0): sequence b is cut into n1 chunks of T0 steps; warp `wi` of the sequence
//   owns chunks wi*32 .. wi*32+31 (warps never straddle sequences), nW = ceil(n1/32).
// mode B (per_lane = 1): one whole sequence per lane (T0 = N), no scan at all.
struct Geom {
    int N, T0, n1, nW, batch, per_lane;
    int pw_bytes;   // per-warp shared-memory pitch of the current launch (set by the launcher)
};
struct Lane {
    int b, wi, len;        // sequence, warp-in-sequence, number of valid steps of this lane's chunk
    long long t0;          // global step index (b*N + k0) of the chunk's first step
    long long slot;        // per-thread slot in the SoA planes
    int k0;
};
__host__ __device__ inline long long total_warps(const Geom& g) {
    return g.per_lane ? ((long long)g.batch + 31) / 32 : (long long)g.batch * g.nW;
}
IPOC_DEV Lane lane_info(const Geom& g, long long wg, int lane) {
    Lane L;
    if (g.per_lane) {
        const long long b = wg * 32 + lane;
        L.b = (int)(b < g.batch ? b : g.batch - 1);
        L.wi = 0;
        L.k0 = 0;
        L.len = b < g.batch ? g.N : 0;
    } else {
        L.b = (int)(wg / g.nW);
        L.wi = (int)(wg % g.nW);
        const long long c = (long long)L.wi * 32 + lane;
        const long long k0 = c * g.T0;
        L.k0 = (int)(k0 < g.N ? k0 : g.N);
        const long long rem = (long long)g.N - k0;
        L.len = (int)(rem <= 0 ? 0 : (rem < g.T0 ? rem : g.T0));
    }
    L.t0 = (long long)L.b * g.N + L.k0;
    L.slot = wg * 32 + lane;
    return L;
}

// ------------------------------------------------------------------ staged (cp.async) row loads
__host__ __device__ constexpr int row_gran(int cnt) { return (cnt % 2 == 0) ? 16 : 8; }
__host__ __device__ constexpr int row_cpr(int cnt) { return cnt * 8 / row_gran(cnt); }
__host__ __device__ constexpr int row_pitch(int cnt) { return (row_cpr(cnt) % 2 == 1) ? cnt * 8 : cnt * 8 + row_gran(cnt); }
__host__ __device__ constexpr int arr_bytes(int cnt) { return (32 * row_pitch(cnt) + 15) / 16 * 16; }

IPOC_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int NKEEP>
IPOC_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(NKEEP) : "memory"); }

// Rows of a warp are affine in the lane index: row r starts at global step tb + r*S (mode A: S = T0,
// consecutive chunks of one sequence; mode B: S = N, consecutive sequences).  No per-row table.
struct RowMap {
    long long tb;    // global step index of row 0
    int remc;        // min(steps left from row 0's start, 32*S): row r, step j is valid iff r*S + j < remc
    int S, T0;
};

// ---- TMA (bulk async copy) staging -----------------------------------------------------------------------
// A row whose size is a multiple of 16 bytes travels as ONE `cp.async.bulk` (TMA, SASS: UBLKCP) issued by the
// lane that owns it — source and destination are 16-byte aligned by construction — and completes on the
// stage's mbarrier (transaction bytes; SASS: SYNCS).  Rows of 8 bytes (R and ru when nu = 1) cannot be bulk
// copies (16-byte granularity) and keep the per-granule cp.async path and its wait_group.  Compared with the
// 16-byte cp.async granules this is one copy instruction per row instead of eight for an nx = 4 matrix row.
#ifndef IPOC_TMA
#define IPOC_TMA 0
#endif
IPOC_DEV void mbar_init(unsigned mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(mbar), "r"(count) : "memory");
}
IPOC_DEV void mbar_expect_tx(unsigned mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar), "r"(bytes) : "memory");
}
IPOC_DEV void mbar_wait(unsigned mbar, unsigned parity) {
    unsigned done = 0;
    int spins = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(mbar), "r"(parity)
            : "memory");
        if (done) break;
        if (++spins > (1 << 24)) __trap();   // a lost transaction must fail the launch, never hang the GPU
    }
}
template <int CNT>
__host__ __device__ constexpr bool row_is_bulk() { return IPOC_TMA != 0 && (CNT * 8) % 16 == 0; }
template <int CNT>
__host__ __device__ constexpr int row_bulk_bytes() { return row_is_bulk<CNT>() ? CNT * 8 : 0; }

// Copy step j of every lane-row r of one input array into the stage.  Bulk rows: lane r copies its own row.
// Other rows: the warp cooperates, `CPR` granules per row, consecutive lanes on consecutive granules of the
// same row.  32-bit index arithmetic relative to the warp's first row, one wide multiply-add per copy.
template <int CNT>
IPOC_DEV void issue_rows(unsigned dst_arr, unsigned mbar, const double* __restrict__ g, const RowMap& m, int j, int lane) {
    constexpr int G = row_gran(CNT), CPR = row_cpr(CNT), PITCH = row_pitch(CNT);
    const char* gbase = reinterpret_cast<const char*>(g + m.tb * CNT);
    if constexpr (row_is_bulk<CNT>()) {
        const int k = lane * m.S + j;
        if (k < m.remc) {
            const char* src = gbase + (long long)k * (CNT * 8);
            const unsigned dst = dst_arr + lane * PITCH;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                         "l"(src), "n"(CNT * 8), "r"(mbar)
                         : "memory");
        }
    } else {
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
            const int idx = lane + 32 * i;
            const int r = idx / CPR, part = idx % CPR;
            const int k = r * m.S + j;
            if (k < m.remc) {
                const char* src = gbase + (long long)k * (CNT * 8) + part * G;
                const unsigned dst = dst_arr + r * PITCH + part * G;
                if constexpr (G == 16)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
                else
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src) : "memory");
            }
        }
    }
}
template <int CNT>
IPOC_DEV void read_row(double* dst, const char* src_arr, int lane) {
    const char* p = src_arr + lane * row_pitch(CNT);
    if constexpr (CNT % 2 == 0) {
#pragma unroll
        for (int i = 0; i < CNT / 2; ++i) {
            const double2 v = *reinterpret_cast<const double2*>(p + 16 * i);
            dst[2 * i] = v.x;
            dst[2 * i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < CNT; ++i) dst[i] = *reinterpret_cast<const double*>(p + 8 * i);
    }
}

// Per-warp shared memory: [stage 0 | stage 1] (+ scan scratch overlaid after the walk)
struct WarpSmem {
    RowMap map;
    char* stage0;
    unsigned stage0_s;   // the same address in the shared window (computed once, not per copy)
    unsigned mbar_s;     // kMbarBytes at the end of the warp's area: one mbarrier per stage of the ring
};
constexpr int kMbarBytes = 64;   // up to 8 stages
IPOC_DEV WarpSmem warp_smem(char* smem, const Geom& g, int warp_in_block, long long wg) {
    WarpSmem w;
    w.stage0 = smem + (size_t)warp_in_block * g.pw_bytes;
    w.stage0_s = (unsigned)__cvta_generic_to_shared(w.stage0);
    w.mbar_s = w.stage0_s + (unsigned)g.pw_bytes - (unsigned)kMbarBytes;
    long long rem;
    if (g.per_lane) {
        w.map.tb = wg * 32 * (long long)g.N;
        rem = ((long long)g.batch - wg * 32) * g.N;
        w.map.S = g.N;
        w.map.T0 = g.N;
    } else {
        const long long b = wg / g.nW, wi = wg % g.nW;
        w.map.tb = b * g.N + wi * 32 * g.T0;
        rem = (long long)g.N - wi * 32 * g.T0;
        w.map.S = g.T0;
        w.map.T0 = g.T0;
    }
    const long long cap = 32LL * w.map.S;
    w.map.remc = (int)(rem < cap ? rem : cap);
    return w;
}

// Walk over the T steps of a chunk (uniform trip count; lanes with shorter chunks idle) through a ring
// of NS shared stages per warp: fetch(stage) copies the lane's row into registers, after which that
// stage is free again, so the copies of step it+NS are issued before compute(j) runs.  NS = 1 already
// overlaps one step of arithmetic with the copies; the memory-bound kernels (K1, K3, K2's down-sweep)
// use deeper rings to keep more bytes in flight (they were waiting on the cp.async group 40 % of the
// time with NS = 1, profiles/r01).
template <int NS, class Ld, class F1, class F2>
IPOC_DEV void staged_walk(const Ld& ld, const WarpSmem& w, int T, int len, int lane, bool reverse, F1&& fetch,
                          F2&& compute) {
    const int bulk_row = ld.bulk_row_bytes();   // bytes per row that travel as bulk copies (0: cp.async only)
    if (IPOC_TMA) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < NS; ++s) mbar_init(w.mbar_s + 8 * s, 1);
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        __syncwarp();
    }
    // all copies of step j into ring slot `slot`; lane 0 arms the slot's mbarrier with the bytes of the valid rows
    auto issue = [&](int slot, int j) {
        const unsigned mb = w.mbar_s + 8 * slot;
        if (IPOC_TMA && lane == 0) {
            const int nvalid = (j < w.map.remc) ? min(32, (w.map.remc - j - 1) / w.map.S + 1) : 0;
            mbar_expect_tx(mb, (unsigned)(nvalid * bulk_row));
        }
        ld.issue(w.stage0_s + slot * Ld::STAGE_BYTES, mb, w.map, j, lane);
    };
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        if (s < T) issue(s, reverse ? T - 1 - s : s);
        cp_async_commit();
    }
    int slot = 0;
    unsigned parity = 0;
    for (int it = 0; it < T; ++it) {
        const int j = reverse ? T - 1 - it : it;
        cp_async_wait<NS - 1>();
        if (IPOC_TMA) mbar_wait(w.mbar_s + 8 * slot, parity);
        __syncwarp();
        if (j < len) fetch(w.stage0 + slot * Ld::STAGE_BYTES);
        __syncwarp();
        if (it + NS < T) issue(slot, reverse ? j - NS : j + NS);
        cp_async_commit();
        if (j < len) compute(j);
        if (slot + 1 == NS) {
            slot = 0;
            parity ^= 1u;
        } else {
            ++slot;
        }
    }
}

// ------------------------------------------------------------------ K2 loaders
// Newton mode: the LQT terms of `noc_to_lqt` (ref noc/par_interior_point_newton.py:50-84) on the fly.
// LITERAL = true follows the reference operation by operation: U = R + reg I (:118); X^-1 M (:63);
//   s = -(U - M'X^-1M)^-1 ru (:64); r = -X^-1 M s (:65); then the tracking references are folded
//   back into linear cost terms q = -(X r + M s), p = -(U s + M' r)   (H = Z = I, c = 0, :72-80).
// LITERAL = false (default) uses what those lines evaluate to in exact arithmetic, q = 0 and
//   p = ru (the identities -X r - M s = 0, -U s - M'r = ru of :62-66), skipping the nx x nx solve
//   per step; it differs from the literal path by rounding of order eps * cond(Q) only and does
//   not break down for singular Q.  Both are tested against the oracle; ipoc_set_literal_lqt(1)
//   selects the literal one at run time.
template <int NX, int NU, bool LITERAL = false>
struct NewtonLoader {
    const double *fx, *fu, *ru, *Q, *R, *M, *reg;
    const double* reg_scale;   // optional second device scalar per problem: reg_eff = reg * reg_scale (rp * ||cu||, ref :117)
    static constexpr int O_FX = 0, O_FU = O_FX + arr_bytes(NX * NX), O_Q = O_FU + arr_bytes(NX * NU),
                         O_R = O_Q + arr_bytes(NX * NX), O_M = O_R + arr_bytes(NU * NU),
                         O_RU = O_M + arr_bytes(NX * NU), STAGE_BYTES = O_RU + arr_bytes(NU);
    IPOC_DEV int bulk_row_bytes() const {
        return 2 * row_bulk_bytes<NX * NX>() + 2 * row_bulk_bytes<NX * NU>() + row_bulk_bytes<NU * NU>() + row_bulk_bytes<NU>();
    }
    IPOC_DEV void issue(unsigned st, unsigned mb, const RowMap& m, int j, int lane) const {
        issue_rows<NX * NX>(st + O_FX, mb, fx, m, j, lane);
        issue_rows<NX * NU>(st + O_FU, mb, fu, m, j, lane);
        issue_rows<NX * NX>(st + O_Q, mb, Q, m, j, lane);
        issue_rows<NU * NU>(st + O_R, mb, R, m, j, lane);
        issue_rows<NX * NU>(st + O_M, mb, M, m, j, lane);
        issue_rows<NU>(st + O_RU, mb, ru, m, j, lane);
    }
    // per-lane constant fetched ONCE before the walk (a global load inside the step loop would expose
    // its full latency every step: 16 % of the stall samples in profiles/r01)
    IPOC_DEV double aux(int b) const { return reg_scale != nullptr ? __ldg(reg + b) * __ldg(reg_scale + b) : __ldg(reg + b); }
    // max |ru| of the staged step (side job of the up-sweep: ref :158)
    IPOC_DEV double ru_absmax(const char* st, int lane) const {
        double ruv[NU];
        read_row<NU>(ruv, st + O_RU, lane);
        double m = 0.0;
#pragma unroll
        for (int a = 0; a < NU; ++a) m = nan_max_d(m, fabs(ruv[a]));
        return m;
    }
    IPOC_DEV void read(StepLQ<NX, NU>& s, const char* st, int lane, double rg) const {
        double Qf[NX][NX], Rf[NU][NU], ruv[NU];
        read_row<NX * NX>(&s.A[0][0], st + O_FX, lane);
        read_row<NX * NU>(&s.B[0][0], st + O_FU, lane);
        read_row<NX * NX>(&Qf[0][0], st + O_Q, lane);
        read_row<NU * NU>(&Rf[0][0], st + O_R, lane);
        read_row<NX * NU>(&s.M[0][0], st + O_M, lane);
        read_row<NU>(ruv, st + O_RU, lane);
#pragma unroll
        for (int a = 0; a < NU; ++a)
#pragma unroll
            for (int c = 0; c < NU; ++c) s.U[a][c] = Rf[a][c] + ((a == c) ? rg : 0.0);
        if constexpr (!LITERAL) {
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                s.q[i] = 0.0;
                s.c[i] = 0.0;
#pragma unroll
                for (int j = i; j < NX; ++j) s.X[Sym<NX>::at(i, j)] = 0.5 * (Qf[i][j] + Qf[j][i]);
            }
#pragma unroll
            for (int a = 0; a < NU; ++a) s.p[a] = ruv[a];
        } else {
        // X^-1 M
        double W[NX][NX], XiM[NX][NU];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
            for (int j = 0; j < NX; ++j) W[i][j] = Qf[i][j];
#pragma unroll
            for (int a = 0; a < NU; ++a) XiM[i][a] = s.M[i][a];
        }
        lu_solve<NX, NU>(W, XiM);
        double Sm[NU][NU], sv[NU][1];
#pragma unroll
        for (int a = 0; a < NU; ++a) {
#pragma unroll
            for (int c = 0; c < NU; ++c) {
                double v = s.U[a][c];
#pragma unroll
                for (int i = 0; i < NX; ++i) v -= s.M[i][a] * XiM[i][c];
                Sm[a][c] = v;
            }
            sv[a][0] = ruv[a];
        }
        small_solve<NU, 1>(Sm, sv);   // sv = (U - M'X^-1M)^-1 ru  = -s
        double rr[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double v = 0.0;
#pragma unroll
            for (int a = 0; a < NU; ++a) v += XiM[i][a] * sv[a][0];   // r = -XiM s = XiM sv
            rr[i] = v;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double v = 0.0;
#pragma unroll
            for (int j = 0; j < NX; ++j) v += Qf[i][j] * rr[j];
#pragma unroll
            for (int a = 0; a < NU; ++a) v -= s.M[i][a] * sv[a][0];   // + M s
            s.q[i] = -v;
            s.c[i] = 0.0;
#pragma unroll
            for (int j = i; j < NX; ++j) s.X[Sym<NX>::at(i, j)] = 0.5 * (Qf[i][j] + Qf[j][i]);
        }
#pragma unroll
        for (int a = 0; a < NU; ++a) {
            double v = 0.0;
#pragma unroll
            for (int c = 0; c < NU; ++c) v -= s.U[a][c] * sv[c][0];   // U s
#pragma unroll
            for (int i = 0; i < NX; ++i) v += s.M[i][a] * rr[i];     // + M' r
            s.p[a] = -v;
        }
        }
    }
};

// LQT mode: effective terms given directly (c may be NULL = 0).
template <int NX, int NU>
struct LqtLoader {
    const double *A, *B, *c, *X, *U, *M, *q, *p;
    static constexpr int O_A = 0, O_B = O_A + arr_bytes(NX * NX), O_C = O_B + arr_bytes(NX * NU),
                         O_X = O_C + arr_bytes(NX), O_U = O_X + arr_bytes(NX * NX), O_M = O_U + arr_bytes(NU * NU),
                         O_Q = O_M + arr_bytes(NX * NU), O_P = O_Q + arr_bytes(NX), STAGE_BYTES = O_P + arr_bytes(NU);
    IPOC_DEV int bulk_row_bytes() const {
        return 2 * row_bulk_bytes<NX * NX>() + 2 * row_bulk_bytes<NX * NU>() + row_bulk_bytes<NU * NU>() +
               (c != nullptr ? 2 : 1) * row_bulk_bytes<NX>() + row_bulk_bytes<NU>();
    }
    IPOC_DEV void issue(unsigned st, unsigned mb, const RowMap& m, int j, int lane) const {
        issue_rows<NX * NX>(st + O_A, mb, A, m, j, lane);
        issue_rows<NX * NU>(st + O_B, mb, B, m, j, lane);
        if (c != nullptr) issue_rows<NX>(st + O_C, mb, c, m, j, lane);
        issue_rows<NX * NX>(st + O_X, mb, X, m, j, lane);
        issue_rows<NU * NU>(st + O_U, mb, U, m, j, lane);
        issue_rows<NX * NU>(st + O_M, mb, M, m, j, lane);
        issue_rows<NX>(st + O_Q, mb, q, m, j, lane);
        issue_rows<NU>(st + O_P, mb, p, m, j, lane);
    }
    IPOC_DEV double aux(int) const { return 0.0; }
    IPOC_DEV double ru_absmax(const char*, int) const { return 0.0; }
    IPOC_DEV void read(StepLQ<NX, NU>& s, const char* st, int lane, double) const {
        double Xf[NX][NX], Uf[NU][NU];
        read_row<NX * NX>(&s.A[0][0], st + O_A, lane);
        read_row<NX * NU>(&s.B[0][0], st + O_B, lane);
        read_row<NX * NX>(&Xf[0][0], st + O_X, lane);
        read_row<NU * NU>(&Uf[0][0], st + O_U, lane);
        read_row<NX * NU>(&s.M[0][0], st + O_M, lane);
        read_row<NX>(s.q, st + O_Q, lane);
        read_row<NU>(s.p, st + O_P, lane);
        if (c != nullptr) {
            read_row<NX>(s.c, st + O_C, lane);
        } else {
#pragma unroll
            for (int i = 0; i < NX; ++i) s.c[i] = 0.0;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i)
#pragma unroll
            for (int j = i; j < NX; ++j) s.X[Sym<NX>::at(i, j)] = 0.5 * (Xf[i][j] + Xf[j][i]);
#pragma unroll
        for (int a = 0; a < NU; ++a)
#pragma unroll
            for (int cidx = 0; cidx < NU; ++cidx) s.U[a][cidx] = 0.5 * (Uf[a][cidx] + Uf[cidx][a]);
    }
};

// closed-loop step data of K3: A, B, c (may be NULL), Kx, d
template <int NX, int NU>
struct FwdLoader {
    const double *A, *B, *c, *Kx, *d;
    static constexpr int O_A = 0, O_B = O_A + arr_bytes(NX * NX), O_C = O_B + arr_bytes(NX * NU),
                         O_K = O_C + arr_bytes(NX), O_D = O_K + arr_bytes(NU * NX), STAGE_BYTES = O_D + arr_bytes(NU);
    IPOC_DEV int bulk_row_bytes() const {
        return row_bulk_bytes<NX * NX>() + row_bulk_bytes<NX * NU>() + (c != nullptr ? row_bulk_bytes<NX>() : 0) +
               row_bulk_bytes<NU * NX>() + row_bulk_bytes<NU>();
    }
    IPOC_DEV void issue(unsigned st, unsigned mb, const RowMap& m, int j, int lane) const {
        issue_rows<NX * NX>(st + O_A, mb, A, m, j, lane);
        issue_rows<NX * NU>(st + O_B, mb, B, m, j, lane);
        if (c != nullptr) issue_rows<NX>(st + O_C, mb, c, m, j, lane);
        issue_rows<NU * NX>(st + O_K, mb, Kx, m, j, lane);
        issue_rows<NU>(st + O_D, mb, d, m, j, lane);
    }
};

// one affine array pair: F (NX x NX), c (NX)
template <int NX>
struct AffLoader {
    const double *F, *c;
    static constexpr int O_F = 0, O_C = O_F + arr_bytes(NX * NX), STAGE_BYTES = O_C + arr_bytes(NX);
    IPOC_DEV int bulk_row_bytes() const { return row_bulk_bytes<NX * NX>() + row_bulk_bytes<NX>(); }
    IPOC_DEV void issue(unsigned st, unsigned mb, const RowMap& m, int j, int lane) const {
        issue_rows<NX * NX>(st + O_F, mb, F, m, j, lane);
        issue_rows<NX>(st + O_C, mb, c, m, j, lane);
    }
    IPOC_DEV void read(AffElem<NX>& e, const char* st, int lane, int transpose) const {
        double Fm[NX][NX], cv[NX];
        read_row<NX * NX>(&Fm[0][0], st + O_F, lane);
        read_row<NX>(cv, st + O_C, lane);
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
            for (int j = 0; j < NX; ++j) e.F(i, j) = transpose ? Fm[j][i] : Fm[i][j];
            e.c(i) = cv[i];
        }
    }
};

// Seed construction as a side job of the first CTA of the up-sweep kernel (saves a launch): the seeds
// are only consumed by kernels that run after the up-sweep.
struct SeedJob {
    const double* ST;      // NULL = no job
    size_t st_stride;
    const double* vT;
    int batch;
    double* seed;
    double* zero_aff_seed;
};
template <int NX>
IPOC_DEV void make_ric_seed(const SeedJob& sj, int b) {
    if (sj.zero_aff_seed != nullptr) {
#pragma unroll
        for (int i = 0; i < NX; ++i) sj.zero_aff_seed[(size_t)i * sj.batch + b] = 0.0;
    }
    RicVal<NX> v;
    const double* s = sj.ST + (size_t)b * sj.st_stride;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
        for (int j = i; j < NX; ++j) v.S(i, j) = 0.5 * (s[i * NX + j] + s[j * NX + i]);
        v.v(i) = (sj.vT != nullptr) ? sj.vT[(size_t)b * NX + i] : 0.0;
    }
    soa_store(v, sj.seed, (size_t)sj.batch, (size_t)b);
}

// Terminal value function per problem -> SoA seed (stride = batch).
// ST: full (nx,nx) matrix at ST + b*st_stride (symmetrised), vT at vT + b*nx (NULL = 0).
template <int NX>
__global__ void k_ric_seed(SeedJob sj) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < sj.batch) make_ric_seed<NX>(sj, b);
}

// ------------------------------------------------------------------ K2 leaf kernels
// Up-sweep: each lane folds its chunk backwards in time into one element (starting from the
// identity, so the first prepend reproduces the single-step element exactly), then the warp scans.
//   incl  [slot]               : in-warp inclusive aggregate of every lane          (SoA, stride istride)
//   agg1  [b*nW + (nW-1-wi)]   : warp totals in scan order (end of horizon first)   (SoA, stride a1stride)
template <int NX, int NU, class Loader>
__global__ void __launch_bounds__(kLeafThreads, IPOC_RIC_MINB)
k_ric_leaf_up(Loader ld, Geom g, double* __restrict__ incl, size_t istride, double* __restrict__ agg1,
              size_t a1stride, SeedJob sj, Hier h, SideJobs side, double* __restrict__ mx_part) {
    extern __shared__ __align__(16) char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (wg >= total_warps(g)) return;
    const Lane L = lane_info(g, wg, lane);
    const WarpSmem w = warp_smem(smem, g, wib, wg);
    // seeds of a sequence are written by its first warp, i.e. before that warp's arrival on the sequence's
    // counters: whoever completes the levels in this kernel, and every later kernel, sees them
    if (sj.ST != nullptr && L.wi == 0 && lane == 0) make_ric_seed<NX>(sj, L.b);
    RicElem<NX> a;
    RicOp<NX>::identity(a);
    StepElem<NX, NU> e;   // formed in the fetch phase: smaller than the raw step, lives across the copy issue
    const double aux = ld.aux(L.b);
    double mx = 0.0;
    staged_walk<IPOC_NS_RIC_UP>(ld, w, g.T0, L.len, lane, true,
                [&](const char* st) {
                    StepLQ<NX, NU> s;
                    ld.read(s, st, lane, aux);
                    if (mx_part != nullptr) mx = nan_max_d(mx, ld.ru_absmax(st, lane));
                    make_step_elem(e, s);
                },
                [&](int) { ric_prepend_step(a, e); });
    warp_scan<RicOp<NX>>(a, reinterpret_cast<double*>(w.stage0), lane, true);
    soa_store(a, incl, istride, (size_t)L.slot);
    const int sidx = g.nW - 1 - L.wi;   // index of this warp in scan order (end of the horizon first)
    if (lane == 0) soa_store(a, agg1, a1stride, (size_t)L.b * g.nW + sidx);
    if (mx_part != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = nan_max_d(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) mx_part[(size_t)L.b * g.nW + sidx] = mx;
    }
    if (h.cnt != nullptr) {
        __syncwarp();
        hier_finish<RicOp<NX>>(h, g.nW, g.batch, agg1, a1stride, L.b, sidx, lane, reinterpret_cast<double*>(w.stage0), side);
    } else if (g.nW == 1 && side.hu != nullptr && lane == 0) {
        side.hu[L.b] = mx;   // one warp per sequence: nothing to fold
    }
}

// Down-sweep: seeded Riccati recursion over the chunk, gains out, pred/feasibility partials, and
// the in-warp scan of the chunk's forward (closed-loop) affine aggregates for K3.
//   wvals [b*nW + (nW-1-wi)] : value function entering each warp (from the levels / the seed)
//   mode B (per_lane): wvals = seed (stride batch), no neighbours, final pred/feasible written here.
template <int NX, int NU, class Loader>
__global__ void __launch_bounds__(kLeafThreads, IPOC_RIC_MINB)
k_ric_leaf_down(Loader ld, Geom g, const double* __restrict__ incl, size_t istride,
                const double* __restrict__ wvals, size_t wvstride,
                double* __restrict__ Kx, double* __restrict__ d,
                double* __restrict__ S_out, double* __restrict__ v_out,
                double* __restrict__ pred_part, int* __restrict__ feas_part,
                double* __restrict__ pred, int32_t* __restrict__ feasible,
                double* __restrict__ fincl, size_t fistride, double* __restrict__ fagg1, size_t fa1stride,
                HierIn hin, Hier hf, SideJobs side) {
    extern __shared__ __align__(16) char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // the grid runs over the warps in the REVERSE order of the up-sweep: the first CTAs re-read what the
    // up-sweep fetched last, i.e. what is still in L2
    const long long wg = total_warps(g) - 1 - ((long long)blockIdx.x * (blockDim.x >> 5) + wib);
    if (wg < 0) return;
    const Lane L = lane_info(g, wg, lane);
    const WarpSmem w = warp_smem(smem, g, wib, wg);
    const int N = g.N;
    RicVal<NX> val;
    if (g.per_lane) {
        soa_load(val, wvals, wvstride, (size_t)L.b);
    } else {
        // exclusive prefix (scan order) = inclusive aggregate of the next lane; requested BEFORE the value entering
        // the warp is derived (hier_enter's `apply` is a call the load could not be moved across by the compiler)
        RicElem<NX> ex;
        if (lane < 31) soa_load(ex, incl, istride, (size_t)L.slot + 1);
        if (hin.on) hier_enter<RicOp<NX>>(val, hin, g.nW, L.b, g.nW - 1 - L.wi);
        else soa_load(val, wvals, wvstride, (size_t)L.b * g.nW + (g.nW - 1 - L.wi));
        if (lane < 31) RicOp<NX>::apply(val, ex, val);
    }
    AffElem<NX> fa;
    AffOp<NX>::identity(fa);
    double predsum = 0.0;
    bool feas = true;
    auto write_Sv = [&](int k) {
        double* Sp = S_out + ((size_t)L.b * (N + 1) + k) * NX * NX;
        double* vp = v_out + ((size_t)L.b * (N + 1) + k) * NX;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
            for (int j = 0; j < NX; ++j) Sp[i * NX + j] = val.S(i, j);
            vp[i] = val.v(i);
        }
    };
    if (S_out != nullptr && L.len > 0 && L.k0 + L.len == N) write_Sv(N);
    StepLQ<NX, NU> s;
    const double aux = ld.aux(L.b);
    staged_walk<IPOC_NS_RIC_DOWN>(ld, w, g.T0, L.len, lane, true, [&](const char* st) { ld.read(s, st, lane, aux); }, [&](int j) {
        StepGain<NX, NU> gn;
        ric_step_back(val, gn, s);
        const size_t t = (size_t)(L.t0 + j);
        st_vec<NU * NX>(Kx + t * NU * NX, &gn.Kx[0][0]);
        st_vec<NU>(d + t * NU, gn.d);
        predsum += gn.dGd;
        feas = feas && gn.pd;
        if (S_out != nullptr) write_Sv(L.k0 + j);
        if (fincl != nullptr || fagg1 != nullptr) {
            // fa <- fa o step_k :  P <- P Fcl,  q <- P ccl + q
            AffElem<NX> se;
#pragma unroll
            for (int i = 0; i < NX; ++i) {
#pragma unroll
                for (int jj = 0; jj < NX; ++jj) se.F(i, jj) = gn.Fcl[i][jj];
                se.c(i) = gn.ccl[i];
            }
            AffOp<NX>::compose(fa, se, fa);
        }
    });
    if (g.per_lane) {
        if (L.len > 0) {
            pred[L.b] = -0.5 * predsum;
            feasible[L.b] = feas ? 1 : 0;
        }
        return;
    }
    // deterministic in-warp reduction of the pred / feasibility partials (xor butterfly)
    int fi = feas ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        predsum += __shfl_xor_sync(0xffffffffu, predsum, o);
        fi &= __shfl_xor_sync(0xffffffffu, fi, o);
    }
    if (lane == 0) {
        pred_part[wg] = predsum;
        feas_part[wg] = fi;
    }
    if (fincl != nullptr) {
        warp_scan<AffOp<NX>>(fa, reinterpret_cast<double*>(w.stage0), lane, false);
        soa_store(fa, fincl, fistride, (size_t)L.slot);
        if (lane == 31) soa_store(fa, fagg1, fa1stride, (size_t)L.b * g.nW + L.wi);
        if (hf.cnt != nullptr) {   // K3's levels (and the pred / feasibility fold) completed by the last arrivers
            __syncwarp();
            hier_finish<AffOp<NX>>(hf, g.nW, g.batch, fagg1, fa1stride, L.b, L.wi, lane, reinterpret_cast<double*>(w.stage0),
                                   side);
        }
    }
}

// pred / feasible finalisation as a stand-alone launch (only when no K3 top scan follows that could
// carry it as a side job).
static __global__ void __launch_bounds__(32) k_finalize_pred(SideJobs sj) { side_finish(sj, blockIdx.x, threadIdx.x); }

// Tail of an accept/reject attempt as a side job of K3's leaf kernel (saves three launches per attempt):
//   trial point   tx = x + dx, tu = u + du                                  (ref :156-157)
//   constraints   traj_feasible = all(cons <= 0) of a GIVEN constraint array (ref :45-47)
//   accept        gain ratio, success, rp / r_inc update                    (ref :159-173) by the last arriver
struct TailJob {
    unsigned* cnt;             // NULL = no last-arriver work (the trial point alone needs none)
    const double *x, *u;       // trial point (all four NULL = off)
    double *tx, *tu;
    const int32_t* trial_mask; // members whose trial point is written (NULL = all): frozen members keep theirs
    const double* cons;        // (batch, N, nc); NULL = off
    int nc;
    int* cons_part;            // per-warp AND
    int32_t* traj_feasible;
    const double *cost, *new_cost;   // accept (rp == NULL = off)
    const int32_t* traj_feas_in;     // used when cons == NULL
    const double* pred;
    const int32_t *bwd_feasible, *active;
    double *rp, *r_inc;
    int32_t* success;
    double* gain;
};
IPOC_DEV void accept_rule(const TailJob& t, int b, int traj_ok) {   // same arithmetic as k_accept_update
    if (t.active != nullptr && !t.active[b]) return;
    accept_rule_core(AcceptIO{t.cost, t.pred, t.bwd_feasible, t.rp, t.r_inc, t.success, t.gain}, b, t.new_cost[b], traj_ok);
}

// ------------------------------------------------------------------ K3 leaves
template <int NX, int NU>
IPOC_DEV void read_fwd_step(const FwdLoader<NX, NU>& ld, const char* st, int lane, double (&Am)[NX][NX],
                            double (&Bm)[NX][NU], double (&Km)[NU][NX], double (&dv)[NU], double (&cv)[NX]) {
    using FL = FwdLoader<NX, NU>;
    read_row<NX * NX>(&Am[0][0], st + FL::O_A, lane);
    read_row<NX * NU>(&Bm[0][0], st + FL::O_B, lane);
    read_row<NU * NX>(&Km[0][0], st + FL::O_K, lane);
    read_row<NU>(dv, st + FL::O_D, lane);
    if (ld.c != nullptr) {
        read_row<NX>(cv, st + FL::O_C, lane);
    } else {
#pragma unroll
        for (int i = 0; i < NX; ++i) cv[i] = 0.0;
    }
}

//   xw [b*nW + wi] : state entering each warp;  fincl[slot] : in-warp inclusive forward aggregates
template <int NX, int NU>
__global__ void __launch_bounds__(kLeafThreads)
k_fwd_leaf_down(FwdLoader<NX, NU> ld, Geom g, const double* __restrict__ fincl, size_t fistride,
                const double* __restrict__ xw, size_t xwstride,
                double* __restrict__ x_out, double* __restrict__ u_out, HierIn hin, TailJob tail) {
    extern __shared__ __align__(16) char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (wg >= total_warps(g)) return;
    const Lane L = lane_info(g, wg, lane);
    const WarpSmem w = warp_smem(smem, g, wib, wg);
    const int N = g.N;
    AffVal<NX> xv;
    if (g.per_lane) {
        soa_load(xv, xw, xwstride, (size_t)L.b);
    } else {
        if (hin.on) hier_enter<AffOp<NX>>(xv, hin, g.nW, L.b, L.wi);
        else soa_load(xv, xw, xwstride, (size_t)L.b * g.nW + L.wi);
        if (lane > 0) {
            AffElem<NX> ex;
            soa_load(ex, fincl, fistride, (size_t)L.slot - 1);
            AffOp<NX>::apply(xv, ex, xv);
        }
    }
    double x[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = xv.r[i];
    double xnom[NX], unom[NU];   // nominal trajectory rows of the trial-point side job, one step ahead
    const bool do_trial = tail.tx != nullptr && (tail.trial_mask == nullptr || tail.trial_mask[L.b] != 0);
    if (do_trial && L.len > 0) {
        ld_vec<NX>(xnom, tail.x + ((size_t)L.b * (N + 1) + L.k0) * NX);
        ld_vec<NU>(unom, tail.u + (size_t)L.t0 * NU);
    }
    double Am[NX][NX], Bm[NX][NU], Km[NU][NX], dv[NU], cv[NX];
    staged_walk<IPOC_NS_LIGHT>(ld, w, g.T0, L.len, lane, false,
                [&](const char* st) { read_fwd_step<NX, NU>(ld, st, lane, Am, Bm, Km, dv, cv); }, [&](int j) {
        double u[NU], xn[NX];
#pragma unroll
        for (int a = 0; a < NU; ++a) {
            double v = dv[a];
#pragma unroll
            for (int jj = 0; jj < NX; ++jj) v -= Km[a][jj] * x[jj];
            u[a] = v;
        }
        const size_t t = (size_t)(L.t0 + j);
        const size_t xrow = ((size_t)L.b * (N + 1) + L.k0 + j) * NX;
        st_vec<NX>(x_out + xrow, x);
        st_vec<NU>(u_out + t * NU, u);
        if (do_trial) {   // trial point (ref :156-157); the nominal rows were prefetched one step ahead
            double tv[NX], tw[NU];
#pragma unroll
            for (int i = 0; i < NX; ++i) tv[i] = xnom[i] + x[i];
#pragma unroll
            for (int a = 0; a < NU; ++a) tw[a] = unom[a] + u[a];
            st_vec<NX>(tail.tx + xrow, tv);
            st_vec<NU>(tail.tu + t * NU, tw);
            if (j + 1 < L.len) {
                ld_vec<NX>(xnom, tail.x + xrow + NX);
                ld_vec<NU>(unom, tail.u + (t + 1) * NU);
            }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double v = cv[i];
#pragma unroll
            for (int jj = 0; jj < NX; ++jj) v += Am[i][jj] * x[jj];
#pragma unroll
            for (int a = 0; a < NU; ++a) v += Bm[i][a] * u[a];
            xn[i] = v;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = xn[i];
    });
    if (L.len > 0 && L.k0 + L.len == N) {
        const size_t xrow = ((size_t)L.b * (N + 1) + N) * NX;
        st_vec<NX>(x_out + xrow, x);
        if (do_trial) {
            double tv[NX];
            ld_vec<NX>(tv, tail.x + xrow);
#pragma unroll
            for (int i = 0; i < NX; ++i) tv[i] += x[i];
            st_vec<NX>(tail.tx + xrow, tv);
        }
    }
    if (tail.cnt == nullptr || g.per_lane) return;
    // ---- constraint feasibility of a given array + accept / regularisation update by the last arriver
    int ok = 1;
    if (tail.cons != nullptr) {
        // the warp's steps are contiguous: coalesced, sixteen loads in flight per lane (NaN compares false -> infeasible)
        const double* cp = tail.cons + (size_t)w.map.tb * tail.nc;
        const int n = w.map.remc * tail.nc;
        constexpr int U = 16;
        for (int i0 = lane; i0 < n; i0 += 32 * U) {
            double v[U];
#pragma unroll
            for (int k = 0; k < U; ++k) v[k] = (i0 + 32 * k < n) ? __ldg(cp + i0 + 32 * k) : 0.0;
#pragma unroll
            for (int k = 0; k < U; ++k) ok &= (v[k] <= 0.0) ? 1 : 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ok &= __shfl_xor_sync(0xffffffffu, ok, o);
        if (lane == 0) tail.cons_part[wg] = ok;
    }
    if (g.nW > 1)
        if (!warp_arrive_last(tail.cnt + L.b, (unsigned)g.nW, lane)) return;
    int traj_ok = 1;
    if (tail.cons != nullptr) {
        fold_partials(tail.cons_part + (size_t)L.b * g.nW, g.nW, lane, [&](int v) { traj_ok &= v; });
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) traj_ok &= __shfl_xor_sync(0xffffffffu, traj_ok, o);
        if (lane == 0) tail.traj_feasible[L.b] = traj_ok;
    } else if (tail.traj_feas_in != nullptr) {
        traj_ok = tail.traj_feas_in[L.b];
    }
    if (tail.rp != nullptr && lane == 0) accept_rule(tail, L.b, traj_ok);
}

// K3 leaf up (only for the stand-alone par_fwd_pass API; the Newton step gets these aggregates
// from K2's down-sweep).
template <int NX, int NU>
__global__ void __launch_bounds__(kLeafThreads)
k_fwd_leaf_up(FwdLoader<NX, NU> ld, Geom g, double* __restrict__ fincl, size_t fistride,
              double* __restrict__ fagg1, size_t fa1stride, Hier h) {
    extern __shared__ __align__(16) char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (wg >= total_warps(g)) return;
    const Lane L = lane_info(g, wg, lane);
    const WarpSmem w = warp_smem(smem, g, wib, wg);
    AffElem<NX> fa;
    AffOp<NX>::identity(fa);
    double Am[NX][NX], Bm[NX][NU], Km[NU][NX], dv[NU], cv[NX];
    staged_walk<IPOC_NS_LIGHT>(ld, w, g.T0, L.len, lane, false,
                [&](const char* st) { read_fwd_step<NX, NU>(ld, st, lane, Am, Bm, Km, dv, cv); }, [&](int) {
        AffElem<NX> se;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
            for (int jj = 0; jj < NX; ++jj) {
                double v = Am[i][jj];
#pragma unroll
                for (int a = 0; a < NU; ++a) v -= Bm[i][a] * Km[a][jj];
                se.F(i, jj) = v;
            }
            double v = cv[i];
#pragma unroll
            for (int a = 0; a < NU; ++a) v += Bm[i][a] * dv[a];
            se.c(i) = v;
        }
        AffOp<NX>::compose(fa, fa, se);
    });
    warp_scan<AffOp<NX>>(fa, reinterpret_cast<double*>(w.stage0), lane, false);
    soa_store(fa, fincl, fistride, (size_t)L.slot);
    if (lane == 31) soa_store(fa, fagg1, fa1stride, (size_t)L.b * g.nW + L.wi);
    if (h.cnt != nullptr) {
        __syncwarp();
        hier_finish<AffOp<NX>>(h, g.nW, g.batch, fagg1, fa1stride, L.b, L.wi, lane, reinterpret_cast<double*>(w.stage0),
                               SideJobs{});
    }
}

// ------------------------------------------------------------------ K1 (generic affine scan) leaves
template <int NX>
__global__ void __launch_bounds__(kLeafThreads)
k_aff_leaf_up(AffLoader<NX> ld, int reverse, int transpose, Geom g, double* __restrict__ incl, size_t istride,
              double* __restrict__ agg1, size_t a1stride, Hier h, SideJobs side, const double* __restrict__ sq_src,
              int sq_width, double* __restrict__ sq_part, const int32_t* __restrict__ fresh) {
    extern __shared__ __align__(16) char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (wg >= total_warps(g)) return;
    const Lane L = lane_info(g, wg, lane);
    // sequences whose inputs did not change since the last scan keep their stored results (no warp of a skipped
    // sequence arrives anywhere, so its counters stay at zero)
    if (fresh != nullptr && !g.per_lane && fresh[L.b] == 0) return;
    const WarpSmem w = warp_smem(smem, g, wib, wg);
    // side job ||cu||_F: the warp's rows of `sq_src` are contiguous; the first eight values per lane are requested
    // before the walk (consumed after it: no exposed latency up to 256 values per warp), the rest in batches
    constexpr int SQP = 8;
    double sqv[SQP];
    const double* sq_cp = nullptr;
    int sq_n = 0;
    if (sq_part != nullptr) {
        sq_cp = sq_src + (size_t)w.map.tb * sq_width;
        sq_n = w.map.remc * sq_width;
#pragma unroll
        for (int k = 0; k < SQP; ++k) sqv[k] = (lane + 32 * k < sq_n) ? __ldg(sq_cp + lane + 32 * k) : 0.0;
    }
    AffElem<NX> a;
    AffOp<NX>::identity(a);
    AffElem<NX> se;
    staged_walk<IPOC_NS_LIGHT>(ld, w, g.T0, L.len, lane, reverse != 0,
                               [&](const char* st) { ld.read(se, st, lane, transpose); },
                               [&](int) { AffOp<NX>::compose(a, a, se); });
    warp_scan<AffOp<NX>>(a, reinterpret_cast<double*>(w.stage0), lane, reverse != 0);
    soa_store(a, incl, istride, (size_t)L.slot);
    const int sidx = reverse ? g.nW - 1 - L.wi : L.wi;   // index of this warp in scan order
    if (lane == (reverse ? 0 : 31)) soa_store(a, agg1, a1stride, (size_t)L.b * g.nW + sidx);
    if (sq_part != nullptr) {   // side job: sum of squares of this warp's rows of `sq_src` (||cu||_F, ref :116)
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < SQP; ++k) acc += sqv[k] * sqv[k];
        constexpr int U = 16;
        for (int i0 = lane + 32 * SQP; i0 < sq_n; i0 += 32 * U) {
            double v[U];
#pragma unroll
            for (int k = 0; k < U; ++k) v[k] = (i0 + 32 * k < sq_n) ? __ldg(sq_cp + i0 + 32 * k) : 0.0;
#pragma unroll
            for (int k = 0; k < U; ++k) acc += v[k] * v[k];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) sq_part[(size_t)L.b * g.nW + sidx] = acc;
        if (h.cnt == nullptr && g.nW == 1 && lane == 0) side.cu_norm[L.b] = sqrt(acc);
    }
    if (h.cnt != nullptr) {
        __syncwarp();
        hier_finish<AffOp<NX>>(h, g.nW, g.batch, agg1, a1stride, L.b, sidx, lane, reinterpret_cast<double*>(w.stage0), side);
    }
}

template <int NX>
__global__ void __launch_bounds__(kLeafThreads)
k_aff_leaf_down(AffLoader<NX> ld, int reverse, int transpose, Geom g, const double* __restrict__ incl,
                size_t istride, const double* __restrict__ wvals, size_t wvstride, double* __restrict__ out, HierIn hin,
                const int32_t* __restrict__ fresh) {
    extern __shared__ __align__(16) char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // reverse grid order relative to the up-sweep: start with what it left in L2
    const long long wg = total_warps(g) - 1 - ((long long)blockIdx.x * (blockDim.x >> 5) + wib);
    if (wg < 0) return;
    const Lane L = lane_info(g, wg, lane);
    if (fresh != nullptr && !g.per_lane && fresh[L.b] == 0) return;
    const WarpSmem w = warp_smem(smem, g, wib, wg);
    const int N = g.N;
    AffVal<NX> x;
    if (g.per_lane) {
        soa_load(x, wvals, wvstride, (size_t)L.b);
    } else if (reverse) {
        if (hin.on) hier_enter<AffOp<NX>>(x, hin, g.nW, L.b, g.nW - 1 - L.wi);
        else soa_load(x, wvals, wvstride, (size_t)L.b * g.nW + (g.nW - 1 - L.wi));
        if (lane < 31) {
            AffElem<NX> ex;
            soa_load(ex, incl, istride, (size_t)L.slot + 1);
            AffOp<NX>::apply(x, ex, x);
        }
    } else {
        if (hin.on) hier_enter<AffOp<NX>>(x, hin, g.nW, L.b, L.wi);
        else soa_load(x, wvals, wvstride, (size_t)L.b * g.nW + L.wi);
        if (lane > 0) {
            AffElem<NX> ex;
            soa_load(ex, incl, istride, (size_t)L.slot - 1);
            AffOp<NX>::apply(x, ex, x);
        }
    }
    double* ob = out + (size_t)L.b * (N + 1) * NX;
    if (reverse) {
        if (L.len > 0 && L.k0 + L.len == N) st_vec<NX>(ob + (size_t)N * NX, x.r);
    } else {
        if (L.len > 0 && L.k0 == 0) st_vec<NX>(ob, x.r);
    }
    AffElem<NX> se;
    staged_walk<IPOC_NS_LIGHT>(ld, w, g.T0, L.len, lane, reverse != 0,
                               [&](const char* st) { ld.read(se, st, lane, transpose); }, [&](int j) {
                    AffOp<NX>::apply(x, se, x);
                    st_vec<NX>(ob + (size_t)(L.k0 + j + (reverse ? 0 : 1)) * NX, x.r);
                });
}

// AoS (batch, nx) -> SoA seed planes (stride batch); NULL source = zeros.
template <int NX>
__global__ void k_aff_seed(const double* __restrict__ src, int batch, double* __restrict__ seed) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
#pragma unroll
    for (int i = 0; i < NX; ++i) seed[(size_t)i * batch + b] = (src != nullptr) ? src[(size_t)b * NX + i] : 0.0;
}

// SoA planes (stride `stride`, index idx) -> AoS carry (time-sharded mode)
static __global__ void k_soa_to_aos(const double* __restrict__ soa, size_t stride, size_t idx, int sz,
                                    double* __restrict__ aos) {
    const int c = threadIdx.x;
    if (c < sz) aos[c] = soa[(size_t)c * stride + idx];
}

// =================================================================== host side
constexpr int MAXLEV = 8;
struct Plan {
    Geom g;
    int nlev;          // aggregate levels above the warps (0: the seed enters every warp directly)
    int n[MAXLEV];     // aggregates per sequence at level l (n[0] = nW)
    int T[MAXLEV];     // fan-in from level l to l+1
    long long warps, slots;
    int hier;          // 1: the levels CAN be completed inside the leaf kernels (last arriver), no level kernels
    int ric_coop;      // 1: Riccati levels by the lane-cooperative level kernel (k_ric_top_coop)
    int hier_ric;      // 1: ... also those of the Riccati up-sweep (pays only for long horizons, see make_plan)
    int gw, ngroups, serial_top;
};

// force_scan: time-sharded mode always wants the segment total, hence at least one level.
// lat_heuristic: the Riccati plans trade leaf length against the number of warp totals their levels must scan
// (measured optimum below); the stand-alone affine scans complete their cheap levels inside the leaf kernels and
// are planned for occupancy alone.
static Plan make_plan(int N, int batch, bool force_scan = false, int target_threads = kTargetThreads,
                      bool lat_heuristic = true) {
    Plan p{};
    Geom& g = p.g;
    g.N = N;
    g.batch = batch;
    const int top_max = g_tune.top_max > 0 ? g_tune.top_max : 2048;
    const int mid = g_tune.mid_fanin > 1 ? g_tune.mid_fanin : 8;
    int T0 = g_tune.leaf_chunk;
    g.per_lane = 0;
    if (T0 <= 0) {
        if (!force_scan && (long long)batch * 2 >= target_threads) {
            T0 = N;   // enough independent problems: one sequence per lane, single pass, no scan
        } else {
            long long want = (target_threads + batch - 1) / batch;   // chunks per sequence to fill the chip
            want = ((want + 31) / 32) * 32;                          // whole warps
            if (want <= 64) want = 32;   // one warp per sequence: no level scan at all beats two half-filled warps
            long long t = ((long long)N + want - 1) / want;          // one resident wave of leaf warps
            // below one wave the pass is latency-bound: leaf time grows with T0, the top scan with
            // N / (32 T0) — measured optimum T0 ~ sqrt(N) / 32 (4 @1e4, 8 @1e5, 16 @3e5, 32 @1e6)
            // chunks of a multiple of 8 steps keep the staged rows sector-aligned (measurably faster)
            const long long t_lat = (long long)(sqrt((double)N) / 256.0 + 0.5) * 8;
            if (t > 4) t = ((t + 7) / 8) * 8;
            if (lat_heuristic && t < t_lat) t = t_lat;
            // shortest chunk: 4 steps, 2 while the whole horizon then still fits ONE group of 32 warps (N <= 2048: no
            // third level; measured N = 1e3: 75.8 us per pass with T0 = 2 against 81.9 with 4, r02_chunk_sweep.log)
            const long long tmin = (N <= 2048) ? 2 : 4;
            T0 = (int)(t < tmin ? tmin : t);
        }
    }
    if (T0 > N) T0 = N;
    if (T0 < 1) T0 = 1;
    g.T0 = T0;
    g.n1 = (N + T0 - 1) / T0;
    if (g.n1 == 1 && !force_scan && batch >= 32) g.per_lane = 1;
    g.nW = g.per_lane ? 1 : (g.n1 + 31) / 32;
    p.warps = g.per_lane ? ((long long)batch + 31) / 32 : (long long)batch * g.nW;
    p.slots = p.warps * 32;
    p.nlev = 0;
    if (!g.per_lane && (g.nW > 1 || force_scan)) {
        p.n[0] = g.nW;
        p.nlev = 1;
        while (p.n[p.nlev - 1] > top_max && p.nlev < MAXLEV) {
            p.T[p.nlev - 1] = mid;
            p.n[p.nlev] = (p.n[p.nlev - 1] + mid - 1) / mid;
            p.nlev++;
        }
        // default: no level kernels at all.  The legacy multi-kernel levels remain for explicitly tuned
        // mid/top shapes (tests) and for more groups than one warp can fold (N beyond ~1e8).
        p.gw = (g_hier.group_warps > 0 && g_hier.group_warps <= 32) ? g_hier.group_warps : 32;
        p.ngroups = (g.nW + p.gw - 1) / p.gw;
        p.serial_top = (g_hier.serial_top > 0) ? (g_hier.serial_top > 32 ? 32 : g_hier.serial_top) : 8;
        p.hier = (g_hier.enabled && g_tune.mid_fanin <= 0 && g_tune.top_max <= 0 && p.ngroups <= 32 * 16 &&
                  (long long)batch * (p.ngroups + 1) <= kCtrlRegionWords) ? 1 : 0;
        // Measured (profiles/r02_variants*.log, cartpole nx = 4): for the cheap affine operator the in-kernel
        // levels win at every size (K1 -6 us at N = 1e4); for the Riccati operator a single-CTA top kernel is
        // faster until the group scans can hide behind the streaming of later groups (N = 1e5: +13 us,
        // N = 1e6: -8 us), i.e. from about two dozen groups.  enabled = 2 forces them everywhere (tests).
        p.hier_ric = (p.hier && (g_hier.enabled == 2 || p.ngroups >= 24)) ? 1 : 0;
        // Lane-cooperative level kernel (k_ric_top_coop): one launch like k_top, but its combine rounds run on
        // coop_gs lanes each (measured 2.2 us per round against 3.4) and the groups on separate SMs, at the price
        // of one more `apply` on the way into the down-sweep (hier_enter).  Measured (profiles/r02_variants_coop.log,
        // r02_chunk_sweep.log): 3 groups (N = 1e4) +5 us, 13 groups (1e5) -9 us, 31 groups (1e6) -25 us, so it is
        // the default from 6 groups up to the 32 its top scan takes; the in-kernel Riccati levels remain beyond.
        // enabled = 4 uses it for any number of groups <= 32 (tests), 2 / 3 never.
        p.ric_coop = (p.hier && g_hier.enabled != 3 && g_hier.enabled != 2 && p.ngroups <= 32 && p.gw == 32 &&
                      (p.ngroups >= 6 || g_hier.enabled == 4)) ? 1 : 0;
        if (p.ric_coop) p.hier_ric = 0;
    }
    return p;
}

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

struct ScanWs {   // workspace of one hierarchical scan
    double* incl;          // per-thread in-warp inclusive aggregates (SoA, stride = slots)
    double* agg[MAXLEV];   // agg[0] = warp totals
    double* val[MAXLEV];   // val[0] = value entering each warp
    double* seed;
    double* total;
    // hier mode
    double *incl1, *agg2, *incl2, *val2;
    unsigned* cnt;             // control-block region: [batch*ngroups] group counters, [batch] sequence counters
    double* part;              // per-warp partials of a side job (batch*nW)
    int* ipart;
};

static void carve_scan(Bump& bp, const Plan& p, int esz, int vsz, ScanWs& w, int ctrl_region = 0) {
    w.incl = bp.take<double>((size_t)esz * p.slots);
    for (int l = 0; l < p.nlev; ++l) {
        w.agg[l] = bp.take<double>((size_t)esz * p.g.batch * p.n[l]);
        w.val[l] = bp.take<double>((size_t)vsz * p.g.batch * p.n[l]);
    }
    if (p.nlev == 0) {   // still need a place for the (unused) warp totals of the leaf-up kernels
        w.agg[0] = bp.take<double>((size_t)esz * p.g.batch * p.g.nW);
        w.val[0] = nullptr;
    }
    w.seed = bp.take<double>((size_t)vsz * p.g.batch);
    w.total = bp.take<double>((size_t)esz * p.g.batch);
    w.incl1 = w.agg2 = w.incl2 = w.val2 = w.part = nullptr;
    w.cnt = nullptr;
    w.ipart = nullptr;
    if (p.nlev > 0) {   // carved whether or not the plan uses them: the size must not depend on the knobs' history
        const size_t ng = (size_t)p.g.batch * p.ngroups;
        w.incl1 = bp.take<double>((size_t)esz * p.g.batch * p.g.nW);
        w.agg2 = bp.take<double>((size_t)esz * ng);
        w.incl2 = bp.take<double>((size_t)esz * ng);
        w.val2 = bp.take<double>((size_t)vsz * ng);
    }
    w.cnt = bp.dry ? nullptr : reinterpret_cast<unsigned*>(bp.base) + (size_t)ctrl_region * kCtrlRegionWords;
    w.part = bp.take<double>((size_t)p.g.batch * p.g.nW);
    w.ipart = bp.take<int>((size_t)p.g.batch * p.g.nW);
}

template <class Op>
static Hier make_hier(const Plan& p, const ScanWs& w, const double* seed, double* total) {
    Hier h{};
    if (!p.hier) return h;
    h.cnt = w.cnt;
    h.gw = p.gw;
    h.ngroups = p.ngroups;
    h.serial_top = p.serial_top;
    h.incl1 = w.incl1;
    h.s1 = (size_t)p.g.batch * p.g.nW;
    h.agg2 = w.agg2;
    h.incl2 = w.incl2;
    h.val2 = w.val2;
    h.s2 = (size_t)p.g.batch * p.ngroups;
    h.seed = seed;
    h.total = total;
    return h;
}
static HierIn make_hier_in(const Plan& p, const ScanWs& w) {
    HierIn h{};
    if (!p.hier) return h;
    h.on = 1;
    h.gw = p.gw;
    h.ngroups = p.ngroups;
    h.incl1 = w.incl1;
    h.s1 = (size_t)p.g.batch * p.g.nW;
    h.val2 = w.val2;
    h.incl2 = w.incl2;
    h.s2 = (size_t)p.g.batch * p.ngroups;
    return h;
}
// time-sharded phase 2: the seed is chained through the other ranks' aggregates inside the down-sweep
static HierIn make_hier_chain(const Plan& p, const ScanWs& w, const double* carries, int first, int step, int count,
                              const double* seed0) {
    HierIn h = make_hier_in(p, w);
    h.val2 = nullptr;
    h.carries = carries;
    h.chain_first = first;
    h.chain_step = step;
    h.chain_count = count;
    h.seed0 = seed0;
    return h;
}

struct NewtonWs {
    ScanWs ric, aff;
    double* pred_part;
    int* feas_part;
    double* scratch;   // misc small device scalars
};

template <int NX>
static void carve_newton(Bump& bp, const Plan& p, NewtonWs& w) {
    carve_scan(bp, p, RicElem<NX>::ESZ, RicVal<NX>::VSZ, w.ric);
    carve_scan(bp, p, AffElem<NX>::ESZ, NX, w.aff, 1);
    w.pred_part = bp.take<double>((size_t)p.warps);
    w.feas_part = bp.take<int>((size_t)p.warps);
    w.scratch = bp.take<double>(256);
}

#define IPOC_LAUNCH_CHECK_N(name, st)                              \
    do {                                                           \
        ++g_launches;                                              \
        prof_mark(name, st);                                       \
        if (cudaPeekAtLastError() != cudaSuccess) return IPOC_ECUDA; \
    } while (0)

static inline unsigned grid_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }
// Leaf launch shape: warps per CTA limited by the double-buffered stage (two CTAs per SM should fit).
struct LeafLaunch {
    int wpc, threads;
    unsigned grid;
    size_t smem;
};
static LeafLaunch leaf_launch(const Plan& p, int stage_bytes, size_t scratch_bytes = 0, int nstages = 1) {
    // Warps of a leaf CTA never synchronise with each other, so the CTA size is free: take the
    // smallest CTA that still reaches the largest number of resident warps per SM under the
    // shared-memory limit (228 KB per SM, 1 KB reserved per CTA, at most 32 CTAs) — small CTAs
    // balance the single wave better.
    size_t body = (size_t)stage_bytes * nstages;   // stage ring; the scan scratch reuses the area after the walk
    if (body < scratch_bytes) body = scratch_bytes;
    const size_t per_warp = body + kMbarBytes;      // + the ring's mbarriers (never overlaid)
    const size_t sm_bytes = 228 * 1024;
    int best_wpc = 1, best_warps = 0;
    for (int wpc = 1; wpc <= kLeafThreads / 32; wpc *= 2) {
        const size_t cta = per_warp * wpc + 1024;
        if (per_warp * wpc > 227 * 1024) break;
        long long ctas = (long long)(sm_bytes / cta);
        if (ctas > 32) ctas = 32;
        const int warps = (int)(ctas * wpc);
        if (warps > best_warps) {
            best_warps = warps;
            best_wpc = wpc;
        }
    }
    LeafLaunch l;
    l.wpc = best_wpc;
    l.threads = best_wpc * 32;
    l.grid = (unsigned)((p.warps + best_wpc - 1) / best_wpc);
    l.smem = per_warp * best_wpc;
    return l;
}
template <class K>
static int set_smem_plain(K kernel, size_t bytes) {
    if (bytes > 48 * 1024)
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
            return IPOC_ECUDA;
    return IPOC_OK;
}
// Shared-memory set-up of a leaf kernel: opt in to > 48 KB, then ask for a carve-out that just fits the
// CTAs the register file allows — NOT the maximum: what is left over is L1, and the few spill slots of
// the 255-register Riccati kernels are re-read every time step (with the maximum carve-out those
// LDLs missed L1 and showed up as 18 % long-scoreboard stalls, profiles/r01).
template <class K>
static int set_smem(K kernel, size_t bytes, int threads) {
    if (bytes > 48 * 1024)
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
            return IPOC_ECUDA;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) !=
        cudaSuccess)
        return IPOC_ECUDA;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, bytes) != cudaSuccess) return IPOC_ECUDA;
    if (nb < 1) nb = 1;
    const size_t need = (size_t)nb * (bytes + 1024);
    int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024)) + 3;
    if (pct > 100) pct = 100;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct) != cudaSuccess)
        return IPOC_ECUDA;
    return IPOC_OK;
}

// top-scan launch shape: up to kTopThreads threads, limited by the per-warp scan scratch
template <class Op>
static inline int top_threads(int n) {
    constexpr size_t per_warp = 2 * sizeof(typename Op::Elem) * 32;
    int maxw = (int)((200 * 1024) / per_warp);
    if (maxw > kTopThreads / 32) maxw = kTopThreads / 32;
    if (maxw < 1) maxw = 1;
    int t = ((n + 31) / 32) * 32;
    return t > maxw * 32 ? maxw * 32 : t;
}
template <class Op>
static inline size_t top_smem(int threads) { return 2 * sizeof(typename Op::Elem) * 32 * (size_t)(threads / 32); }
template <class Op>
static int top_prepare(int threads) {
    return set_smem_plain(k_top<Op>, top_smem<Op>(threads));
}

// up-sweep over the aggregate levels (level 0 = warp totals, already filled), top scan, down-sweep
// to level 0.  On return w.val[0] holds the value entering every warp.  If reduce_only, only the
// total aggregate of each sequence is produced (w.total, SoA stride batch).
template <class Op>
static int run_levels(const Plan& p, const ScanWs& w, bool want_total, bool reduce_only, cudaStream_t st,
                      SideJobs pj = SideJobs{}) {
    const int L = p.nlev, batch = p.g.batch;
    for (int l = 0; l + 1 < L; ++l) {
        const long long cnt = (long long)batch * p.n[l + 1];
        k_mid_up<Op><<<grid_for(cnt, kMidThreads), kMidThreads, 0, st>>>(
            w.agg[l], (size_t)batch * p.n[l], p.n[l], w.agg[l + 1], (size_t)batch * p.n[l + 1], p.n[l + 1],
            p.T[l], batch);
        IPOC_LAUNCH_CHECK_N(Op::tag_mid_up, st);
    }
    const int tt = top_threads<Op>(p.n[L - 1]);
    if (int rc = top_prepare<Op>(tt)) return rc;
    k_top<Op><<<batch, tt, top_smem<Op>(tt), st>>>(
        w.agg[L - 1], (size_t)batch * p.n[L - 1], p.n[L - 1], batch, w.seed, w.val[L - 1],
        (size_t)batch * p.n[L - 1], (want_total || reduce_only) ? w.total : nullptr, reduce_only ? 1 : 0, pj);
    IPOC_LAUNCH_CHECK_N(Op::tag_top, st);
    if (reduce_only) return IPOC_OK;
    for (int l = L - 2; l >= 0; --l) {
        const long long cnt = (long long)batch * p.n[l + 1];
        k_mid_down<Op><<<grid_for(cnt, kMidThreads), kMidThreads, 0, st>>>(
            w.agg[l], (size_t)batch * p.n[l], p.n[l], w.val[l], (size_t)batch * p.n[l], w.val[l + 1],
            (size_t)batch * p.n[l + 1], p.n[l + 1], p.T[l], batch);
        IPOC_LAUNCH_CHECK_N(Op::tag_mid_down, st);
    }
    return IPOC_OK;
}

// values entering the warps: from the levels, or straight from the seed when there is one warp
// per sequence (or one sequence per lane)
static void leaf_values(const Plan& p, const ScanWs& w, const double*& vals, size_t& stride) {
    if (p.nlev > 0) {
        vals = w.val[0];
        stride = (size_t)p.g.batch * p.g.nW;
    } else {
        vals = w.seed;
        stride = (size_t)p.g.batch;
    }
}

// ---- K2 (+K3 aggregates) for any loader ---------------------------------------------------
template <int NX, int NU, class Loader>
static int run_bwd_up(const Plan& p, const NewtonWs& w, const Loader& ld, cudaStream_t st, SeedJob sj, const Hier& h,
                      const SideJobs& side, double* mx_part) {
    const LeafLaunch ll = leaf_launch(p, Loader::STAGE_BYTES, scan_scratch_bytes<RicOp<NX>>(), IPOC_NS_RIC_UP);
    Geom g = p.g;
    g.pw_bytes = (int)(ll.smem / ll.wpc);
    auto kern = k_ric_leaf_up<NX, NU, Loader>;
    if (int rc = set_smem(kern, ll.smem, ll.threads)) return rc;
    kern<<<ll.grid, ll.threads, ll.smem, st>>>(ld, g, w.ric.incl, (size_t)p.slots, w.ric.agg[0],
                                              (size_t)p.g.batch * p.g.nW, sj, h, side, mx_part);
    IPOC_LAUNCH_CHECK_N("k_ric_leaf_up", st);
    return IPOC_OK;
}

// hin: how the value entering each warp is obtained; hf: in-kernel levels of the forward aggregates (K3)
template <int NX, int NU, class Loader>
static int run_bwd_down(const Plan& p, const NewtonWs& w, const Loader& ld, double* Kx, double* d, double* S,
                        double* v, double* pred, int32_t* feasible, bool want_fwd_agg, cudaStream_t st,
                        bool defer_pred, const HierIn& hin, const Hier& hf) {
    const double* vals;
    size_t vstride;
    leaf_values(p, w.ric, vals, vstride);
    const LeafLaunch ll = leaf_launch(p, Loader::STAGE_BYTES, scan_scratch_bytes<AffOp<NX>>(), IPOC_NS_RIC_DOWN);
    Geom g = p.g;
    g.pw_bytes = (int)(ll.smem / ll.wpc);
    auto kern = k_ric_leaf_down<NX, NU, Loader>;
    if (int rc = set_smem(kern, ll.smem, ll.threads)) return rc;
    const bool fwd = want_fwd_agg && !p.g.per_lane;
    SideJobs side{};
    const bool fold_here = fwd && hf.cnt != nullptr;   // the warp completing K3's levels also folds pred / feasible
    if (fold_here) {
        side.n = p.g.nW;
        side.pred_part = w.pred_part;
        side.feas_part = w.feas_part;
        side.pred = pred;
        side.feasible = feasible;
    }
    kern<<<ll.grid, ll.threads, ll.smem, st>>>(ld, g, w.ric.incl, (size_t)p.slots, vals, vstride, Kx, d, S, v,
                                              w.pred_part, w.feas_part, pred, feasible, fwd ? w.aff.incl : nullptr,
                                              (size_t)p.slots, fwd ? w.aff.agg[0] : nullptr,
                                              (size_t)p.g.batch * p.g.nW, hin, hf, side);
    IPOC_LAUNCH_CHECK_N("k_ric_leaf_down", st);
    if (!p.g.per_lane && !defer_pred && !fold_here) {
        SideJobs fin{};
        fin.n = p.g.nW;
        fin.pred_part = w.pred_part;
        fin.feas_part = w.feas_part;
        fin.pred = pred;
        fin.feasible = feasible;
        k_finalize_pred<<<p.g.batch, 32, 0, st>>>(fin);
        IPOC_LAUNCH_CHECK_N("k_finalize_pred", st);
    }
    return IPOC_OK;
}

// Whole backward pass.  `xtra` (may be NULL) carries the max|ru| side job.
template <int NX, int NU, class Loader>
static int run_bwd(const Plan& p, const NewtonWs& w, const Loader& ld, double* Kx, double* d, double* S, double* v,
                   double* pred, int32_t* feasible, bool want_fwd_agg, cudaStream_t st, bool defer_pred, SeedJob sj,
                   AttemptExtras* xtra) {
    Hier hup{}, hf{};
    HierIn hin{};
    if (p.g.per_lane) {   // no up-sweep to piggy-back on
        k_ric_seed<NX><<<grid_for(sj.batch, 128), 128, 0, st>>>(sj);
        IPOC_LAUNCH_CHECK_N("k_ric_seed", st);
    } else {
        SideJobs side{};
        double* mx_part = nullptr;
        if (xtra != nullptr && xtra->hu != nullptr) {   // folded by whoever completes the levels (in-kernel or k_top)
            side.n = p.g.nW;
            side.mx_part = mx_part = w.ric.part;
            side.hu = xtra->hu;
            xtra->handled |= IPOC_X_HU;
        }
        if (p.hier_ric) {
            hup = make_hier<RicOp<NX>>(p, w.ric, w.ric.seed, nullptr);
            hin = make_hier_in(p, w.ric);
        }
        if (p.hier && want_fwd_agg) hf = make_hier<AffOp<NX>>(p, w.aff, w.aff.seed, nullptr);
        if (int rc = run_bwd_up<NX, NU>(p, w, ld, st, sj, hup, side, mx_part)) return rc;
        if (p.nlev > 0 && p.ric_coop && NX > 1) {
            const Hier hc = make_hier<RicOp<NX>>(p, w.ric, w.ric.seed, nullptr);
            constexpr size_t smem = 2 * sizeof(RicElem<NX>) * 32;
            if (int rc = set_smem_plain(k_ric_top_coop<NX>, smem)) return rc;
            k_ric_top_coop<NX><<<p.g.batch * p.ngroups, 32 * coop_gs<NX>(), smem, st>>>(
                w.ric.agg[0], (size_t)p.g.batch * p.g.nW, p.g.nW, p.g.batch, hc, side);
            IPOC_LAUNCH_CHECK_N("k_ric_top_coop", st);
            hin = make_hier_in(p, w.ric);
        } else if (p.nlev > 0 && !p.hier_ric) {
            if (int rc = run_levels<RicOp<NX>>(p, w.ric, false, false, st, side)) return rc;
        }
    }
    return run_bwd_down<NX, NU>(p, w, ld, Kx, d, S, v, pred, feasible, want_fwd_agg, st, defer_pred, hin, hf);
}

// ---- K3 given in-warp forward aggregates in w.aff.incl / w.aff.agg[0] and the seed in w.aff.seed
template <int NX, int NU>
static int run_fwd_down(const Plan& p, const NewtonWs& w, const double* A, const double* B, const double* c,
                        const double* Kx, const double* d, double* x, double* u, bool levels, cudaStream_t st,
                        SideJobs pj, const HierIn& hin, const TailJob& tail) {
    if (levels && p.nlev > 0 && !hin.on)
        if (int rc = run_levels<AffOp<NX>>(p, w.aff, false, false, st, pj)) return rc;
    const double* vals;
    size_t vstride;
    leaf_values(p, w.aff, vals, vstride);
    FwdLoader<NX, NU> ld{A, B, c, Kx, d};
    const LeafLaunch ll = leaf_launch(p, FwdLoader<NX, NU>::STAGE_BYTES, 0, IPOC_NS_LIGHT);
    Geom g = p.g;
    g.pw_bytes = (int)(ll.smem / ll.wpc);
    auto kern = k_fwd_leaf_down<NX, NU>;
    if (int rc = set_smem(kern, ll.smem, ll.threads)) return rc;
    kern<<<ll.grid, ll.threads, ll.smem, st>>>(ld, g, w.aff.incl, (size_t)p.slots, vals, vstride, x, u, hin, tail);
    IPOC_LAUNCH_CHECK_N("k_fwd_leaf_down", st);
    return IPOC_OK;
}

static TailJob make_tail(const Plan& p, const NewtonWs& w, void* ws, AttemptExtras* x, const double* pred,
                         const int32_t* bwd_feasible) {
    TailJob t{};
    if (x == nullptr) return t;
    // The trial point needs no fold (any plan), but as a side job its rows are per-lane scattered 8/32-byte accesses
    // (one L1 transaction each): measured +390 us for 0.66 GB at 8192 x N = 1000, +46 us at N = 1e6, against one
    // coalesced elementwise pass of 120 B per step.  Fused where the launch it saves matters, i.e. small problems.
    constexpr long long kTrialFuseMaxSteps = 400000;
    if (x->tx != nullptr && (long long)p.g.batch * p.g.N <= kTrialFuseMaxSteps) {
        t.x = x->x;
        t.u = x->u;
        t.tx = x->tx;
        t.tu = x->tu;
        t.trial_mask = x->active;
        x->handled |= IPOC_X_TRIAL;
    }
    if (p.g.per_lane || (p.g.nW > 1 && p.g.batch > kCtrlRegionWords)) return t;   // folds need one counter per sequence
    if (x->cons == nullptr && x->rp == nullptr) return t;
    t.cnt = reinterpret_cast<unsigned*>(ws) + 2 * (size_t)kCtrlRegionWords;
    if (x->cons != nullptr) {
        t.cons = x->cons;
        t.nc = x->nc;
        t.cons_part = w.aff.ipart;
        t.traj_feasible = x->traj_feasible;
        x->handled |= IPOC_X_CONS;
    }
    if (x->rp != nullptr) {
        t.cost = x->cost;
        t.new_cost = x->new_cost;
        t.traj_feas_in = x->traj_feas_in;
        t.pred = pred;
        t.bwd_feasible = bwd_feasible;
        t.active = x->active;
        t.rp = x->rp;
        t.r_inc = x->r_inc;
        t.success = x->success;
        t.gain = x->gain;
        x->handled |= IPOC_X_ACCEPT;
    }
    return t;
}

template <int NX, int NU>
static int newton_step_impl(int N, int batch, const double* fx, const double* fu, const double* ru, const double* Q,
                            const double* R, const double* M, const double* reg, double* dx, double* du, double* Kx,
                            double* d, double* pred, int32_t* feasible, void* ws, size_t ws_bytes, cudaStream_t st,
                            AttemptExtras* xtra) {
    const Plan p = make_plan(N, batch);
    Bump bp{(char*)ws, (size_t)kCtrlBytes, ws_bytes, false};
    NewtonWs w;
    carve_newton<NX>(bp, p, w);
    if (bp.off > ws_bytes) return IPOC_EWORKSPACE;
    // terminal value function: XT = Q[0], HT = I, rT = 0 (ref noc/par_interior_point_newton.py:73-75);
    // zero initial deviation dx_0 = 0 (:122) — both seeds as a side job of the up-sweep
    const SeedJob sj{Q, (size_t)N * NX * NX, nullptr, batch, w.ric.seed, w.aff.seed};
    // legacy levels: the pred / feasibility partials are folded by K3's top scan when there is one
    const bool defer = p.nlev > 0 && !p.hier;
    const double* reg_scale = xtra != nullptr ? xtra->reg_scale : nullptr;
    if (g_literal_lqt) {
        NewtonLoader<NX, NU, true> ld{fx, fu, ru, Q, R, M, reg, reg_scale};
        if (int rc = run_bwd<NX, NU>(p, w, ld, Kx, d, nullptr, nullptr, pred, feasible, true, st, defer, sj, xtra))
            return rc;
    } else {
        NewtonLoader<NX, NU, false> ld{fx, fu, ru, Q, R, M, reg, reg_scale};
        if (int rc = run_bwd<NX, NU>(p, w, ld, Kx, d, nullptr, nullptr, pred, feasible, true, st, defer, sj, xtra))
            return rc;
    }
    SideJobs pj{};
    if (defer) {
        pj.n = p.g.nW;
        pj.pred_part = w.pred_part;
        pj.feas_part = w.feas_part;
        pj.pred = pred;
        pj.feasible = feasible;
    }
    const HierIn hin = make_hier_in(p, w.aff);
    const TailJob tail = make_tail(p, w, ws, xtra, pred, feasible);
    return run_fwd_down<NX, NU>(p, w, fx, fu, nullptr, Kx, d, dx, du, true, st, pj, hin, tail);
}

template <int NX, int NU>
static int lqt_bwd_impl(int N, int batch, const double* A, const double* B, const double* c, const double* X,
                        const double* U, const double* M, const double* q, const double* pp, const double* ST,
                        const double* vT, double* Kx, double* d, double* S, double* v, double* pred,
                        int32_t* feasible, void* ws, size_t ws_bytes, cudaStream_t st) {
    const Plan p = make_plan(N, batch);
    Bump bp{(char*)ws, (size_t)kCtrlBytes, ws_bytes, false};
    NewtonWs w;
    carve_newton<NX>(bp, p, w);
    if (bp.off > ws_bytes) return IPOC_EWORKSPACE;
    const SeedJob sj{ST, (size_t)NX * NX, vT, batch, w.ric.seed, nullptr};
    LqtLoader<NX, NU> ld{A, B, c, X, U, M, q, pp};
    return run_bwd<NX, NU>(p, w, ld, Kx, d, S, v, pred, feasible, false, st, false, sj, nullptr);
}

template <int NX, int NU>
static int run_fwd_up(const Plan& p, const NewtonWs& w, const double* A, const double* B, const double* c,
                      const double* Kx, const double* d, cudaStream_t st, const Hier& h) {
    FwdLoader<NX, NU> ld{A, B, c, Kx, d};
    const LeafLaunch ll = leaf_launch(p, FwdLoader<NX, NU>::STAGE_BYTES, scan_scratch_bytes<AffOp<NX>>(), IPOC_NS_LIGHT);
    Geom g = p.g;
    g.pw_bytes = (int)(ll.smem / ll.wpc);
    auto kern = k_fwd_leaf_up<NX, NU>;
    if (int rc = set_smem(kern, ll.smem, ll.threads)) return rc;
    kern<<<ll.grid, ll.threads, ll.smem, st>>>(ld, g, w.aff.incl, (size_t)p.slots, w.aff.agg[0],
                                              (size_t)p.g.batch * p.g.nW, h);
    IPOC_LAUNCH_CHECK_N("k_fwd_leaf_up", st);
    return IPOC_OK;
}

template <int NX, int NU>
static int lqt_fwd_impl(int N, int batch, const double* A, const double* B, const double* c, const double* Kx,
                        const double* d, const double* x0, double* u, double* x, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
    const Plan p = make_plan(N, batch);
    Bump bp{(char*)ws, (size_t)kCtrlBytes, ws_bytes, false};
    NewtonWs w;
    carve_newton<NX>(bp, p, w);
    if (bp.off > ws_bytes) return IPOC_EWORKSPACE;
    k_aff_seed<NX><<<grid_for(batch, 128), 128, 0, st>>>(x0, batch, w.aff.seed);
    IPOC_LAUNCH_CHECK_N("k_aff_seed", st);
    const Hier h = make_hier<AffOp<NX>>(p, w.aff, w.aff.seed, nullptr);
    if (!p.g.per_lane)
        if (int rc = run_fwd_up<NX, NU>(p, w, A, B, c, Kx, d, st, h)) return rc;
    return run_fwd_down<NX, NU>(p, w, A, B, c, Kx, d, x, u, true, st, SideJobs{}, make_hier_in(p, w.aff), TailJob{});
}

// sq_src (batch, N, sq_width) -> side.cu_norm = ||.||_F per problem, folded into the up-sweep (may be NULL)
template <int NX>
static int aff_up(const Plan& p, const ScanWs& w, const double* F, const double* c, int reverse, int transpose,
                  cudaStream_t st, const Hier& h, const double* sq_src = nullptr, int sq_width = 0,
                  double* cu_norm = nullptr, const int32_t* fresh = nullptr) {
    AffLoader<NX> ld{F, c};
    const LeafLaunch ll = leaf_launch(p, AffLoader<NX>::STAGE_BYTES, scan_scratch_bytes<AffOp<NX>>(), IPOC_NS_LIGHT);
    Geom g = p.g;
    g.pw_bytes = (int)(ll.smem / ll.wpc);
    auto kern = k_aff_leaf_up<NX>;
    if (int rc = set_smem(kern, ll.smem, ll.threads)) return rc;
    SideJobs side{};
    double* sq_part = nullptr;
    if (sq_src != nullptr) {
        side.n = p.g.nW;
        side.sq_part = sq_part = w.part;
        side.cu_norm = cu_norm;
    }
    kern<<<ll.grid, ll.threads, ll.smem, st>>>(ld, reverse, transpose, g, w.incl, (size_t)p.slots, w.agg[0],
                                              (size_t)p.g.batch * p.g.nW, h, side, sq_src, sq_width, sq_part, fresh);
    IPOC_LAUNCH_CHECK_N("k_aff_leaf_up", st);
    return IPOC_OK;
}
template <int NX>
static int aff_down(const Plan& p, const ScanWs& w, const double* F, const double* c, int reverse, int transpose,
                    double* out, cudaStream_t st, const HierIn& hin, const int32_t* fresh = nullptr) {
    const double* vals;
    size_t vstride;
    leaf_values(p, w, vals, vstride);
    AffLoader<NX> ld{F, c};
    const LeafLaunch ll = leaf_launch(p, AffLoader<NX>::STAGE_BYTES, 0, IPOC_NS_LIGHT);
    Geom g = p.g;
    g.pw_bytes = (int)(ll.smem / ll.wpc);
    auto kern = k_aff_leaf_down<NX>;
    if (int rc = set_smem(kern, ll.smem, ll.threads)) return rc;
    kern<<<ll.grid, ll.threads, ll.smem, st>>>(ld, reverse, transpose, g, w.incl, (size_t)p.slots, vals, vstride,
                                              out, hin, fresh);
    IPOC_LAUNCH_CHECK_N("k_aff_leaf_down", st);
    return IPOC_OK;
}

// `sq_src`/`cu_norm`: optional ||cu||_F side job; *handled tells the caller whether the scan took it
template <int NX>
static int affine_scan_impl(int reverse, int transpose, int N, int batch, const double* F, const double* c,
                            const double* seed, double* out, void* ws, size_t ws_bytes, cudaStream_t st,
                            const double* sq_src, int sq_width, double* cu_norm, int* handled, const int32_t* fresh) {
    const Plan p = make_plan(N, batch, false, aff_target_threads(), g_hier.aff_warps_per_sm <= 0);
    Bump bp{(char*)ws, (size_t)kCtrlBytes, ws_bytes, false};
    ScanWs w;
    carve_scan(bp, p, AffElem<NX>::ESZ, NX, w);
    if (bp.off > ws_bytes) return IPOC_EWORKSPACE;
    if (batch == 1 && seed != nullptr) {
        w.seed = const_cast<double*>(seed);   // one problem: the SoA seed plane IS the caller's vector
    } else {
        k_aff_seed<NX><<<grid_for(batch, 128), 128, 0, st>>>(seed, batch, w.seed);
        IPOC_LAUNCH_CHECK_N("k_aff_seed", st);
    }
    HierIn hin{};
    if (!p.g.per_lane) {
        const Hier h = make_hier<AffOp<NX>>(p, w, w.seed, nullptr);
        hin = make_hier_in(p, w);
        const bool fold = sq_src != nullptr;
        if (fold && handled != nullptr) *handled |= IPOC_X_NORM;
        if (p.g.per_lane || (p.nlev > 0 && !p.hier)) fresh = nullptr;   // masks are carried by the in-kernel plans only
        if (int rc = aff_up<NX>(p, w, F, c, reverse, transpose, st, h, fold ? sq_src : nullptr, sq_width, cu_norm, fresh))
            return rc;
        if (p.nlev > 0 && !p.hier) {
            SideJobs side{};
            if (fold) {
                side.n = p.g.nW;
                side.sq_part = w.part;
                side.cu_norm = cu_norm;
            }
            if (int rc = run_levels<AffOp<NX>>(p, w, false, false, st, side)) return rc;
        }
    }
    return aff_down<NX>(p, w, F, c, reverse, transpose, out, st, hin, p.g.per_lane ? nullptr : fresh);
}

// ---- time-sharded split-phase implementations ----------------------------------------------
// The reduce phase leaves the in-warp aggregates and the level arrays in the workspace; the apply
// phase (same workspace, same plan) only runs the seeded way down.  With the in-kernel levels a phase is
// ONE launch: the reduce phase ends with the composition of the whole segment (written straight into
// `carry_out`), the apply phase pushes the horizon's seed through the other ranks' aggregates inside the
// down-sweep (hier_enter, chain mode).
template <int NX, int NU>
static int newton_bwd_reduce_impl(int N, const double* fx, const double* fu, const double* ru, const double* Q,
                                  const double* R, const double* M, const double* reg, double* carry_out, void* ws,
                                  size_t ws_bytes, cudaStream_t st) {
    const Plan p = make_plan(N, 1, true);
    Bump bp{(char*)ws, (size_t)kCtrlBytes, ws_bytes, false};
    NewtonWs w;
    carve_newton<NX>(bp, p, w);
    if (bp.off > ws_bytes) return IPOC_EWORKSPACE;
    const bool split_hier = p.hier && p.ngroups <= 32;
    // from two groups on the levels of the segment run in the lane-cooperative level kernel (reduce-only form: the
    // group scans on separate SMs, the last CTA composes the segment total) instead of inside the up-sweep
    const bool coop = split_hier && NX > 1 && p.gw == 32 && p.ngroups >= 2 && g_hier.enabled != 2 && g_hier.enabled != 3;
    Hier h{};
    if (split_hier) h = make_hier<RicOp<NX>>(p, w.ric, nullptr, carry_out);   // batch = 1: SoA == AoS
    const Hier hup = coop ? Hier{} : h;
    const SeedJob nosj{nullptr, 0, nullptr, 0, nullptr, nullptr};
    if (g_literal_lqt) {
        NewtonLoader<NX, NU, true> ld{fx, fu, ru, Q, R, M, reg, nullptr};
        if (int rc = run_bwd_up<NX, NU>(p, w, ld, st, nosj, hup, SideJobs{}, nullptr)) return rc;
    } else {
        NewtonLoader<NX, NU, false> ld{fx, fu, ru, Q, R, M, reg, nullptr};
        if (int rc = run_bwd_up<NX, NU>(p, w, ld, st, nosj, hup, SideJobs{}, nullptr)) return rc;
    }
    if (coop) {
        if constexpr (NX > 1) {
            constexpr size_t smem = 2 * sizeof(RicElem<NX>) * 32;
            if (int rc = set_smem_plain(k_ric_top_coop<NX>, smem)) return rc;
            k_ric_top_coop<NX><<<p.ngroups, 32 * coop_gs<NX>(), smem, st>>>(w.ric.agg[0], (size_t)p.g.nW, p.g.nW, 1, h,
                                                                            SideJobs{});
            IPOC_LAUNCH_CHECK_N("k_ric_top_coop", st);
        }
        return IPOC_OK;
    }
    if (split_hier) return IPOC_OK;
    if (int rc = run_levels<RicOp<NX>>(p, w.ric, true, true, st)) return rc;
    k_soa_to_aos<<<1, 256, 0, st>>>(w.ric.total, 1, 0, RicElem<NX>::ESZ, carry_out);
    IPOC_LAUNCH_CHECK_N("k_soa_to_aos", st);
    return IPOC_OK;
}

// levels on the way down only (aggregates already in the workspace)
template <class Op>
static int run_levels_down(const Plan& p, const ScanWs& w, cudaStream_t st) {
    const int L = p.nlev;
    const int tt = top_threads<Op>(p.n[L - 1]);
    if (int rc = top_prepare<Op>(tt)) return rc;
    k_top<Op><<<1, tt, top_smem<Op>(tt), st>>>(w.agg[L - 1], (size_t)p.n[L - 1], p.n[L - 1], 1, w.seed,
                                               w.val[L - 1], (size_t)p.n[L - 1], nullptr, 0, SideJobs{});
    IPOC_LAUNCH_CHECK_N(Op::tag_top, st);
    for (int l = L - 2; l >= 0; --l) {
        k_mid_down<Op><<<grid_for(p.n[l + 1], kMidThreads), kMidThreads, 0, st>>>(
            w.agg[l], (size_t)p.n[l], p.n[l], w.val[l], (size_t)p.n[l], w.val[l + 1], (size_t)p.n[l + 1], p.n[l + 1],
            p.T[l], 1);
        IPOC_LAUNCH_CHECK_N(Op::tag_mid_down, st);
    }
    return IPOC_OK;
}

template <int NX, int NU>
static int newton_bwd_apply_impl(int N, int rank, int nranks, const double* fx, const double* fu, const double* ru,
                                 const double* Q, const double* R, const double* M, const double* reg,
                                 const double* carries, const double* ST, double* Kx, double* d, double* pred,
                                 int32_t* feasible, double* fwd_carry_out, void* ws, size_t ws_bytes,
                                 cudaStream_t st) {
    const Plan p = make_plan(N, 1, true);
    Bump bp{(char*)ws, (size_t)kCtrlBytes, ws_bytes, false};
    NewtonWs w;
    carve_newton<NX>(bp, p, w);
    if (bp.off > ws_bytes) return IPOC_EWORKSPACE;
    // terminal seed of the whole horizon, pushed back through the later ranks' aggregates
    double* seed0 = w.scratch;   // RicVal packed
    k_ric_seed<NX><<<1, 32, 0, st>>>(SeedJob{ST, (size_t)NX * NX, nullptr, 1, seed0, nullptr});
    IPOC_LAUNCH_CHECK_N("k_ric_seed", st);
    const bool split_hier = p.hier && p.ngroups <= 32;
    HierIn hin{};
    Hier hf{};
    if (split_hier) {
        hin = make_hier_chain(p, w.ric, carries, nranks - 1, -1, nranks - 1 - rank, seed0);
        hf = make_hier<AffOp<NX>>(p, w.aff, nullptr, fwd_carry_out);
    } else {
        k_chain_seed<RicOp<NX>><<<1, 32, 0, st>>>(carries, nranks - 1, -1, nranks - 1 - rank, seed0, w.ric.seed);
        IPOC_LAUNCH_CHECK_N("k_chain_seed_ric", st);
        if (int rc = run_levels_down<RicOp<NX>>(p, w.ric, st)) return rc;
    }
    if (g_literal_lqt) {
        NewtonLoader<NX, NU, true> ld{fx, fu, ru, Q, R, M, reg, nullptr};
        if (int rc = run_bwd_down<NX, NU>(p, w, ld, Kx, d, nullptr, nullptr, pred, feasible, true, st, false, hin, hf))
            return rc;
    } else {
        NewtonLoader<NX, NU, false> ld{fx, fu, ru, Q, R, M, reg, nullptr};
        if (int rc = run_bwd_down<NX, NU>(p, w, ld, Kx, d, nullptr, nullptr, pred, feasible, true, st, false, hin, hf))
            return rc;
    }
    if (split_hier) return IPOC_OK;
    if (int rc = run_levels<AffOp<NX>>(p, w.aff, true, true, st)) return rc;
    k_soa_to_aos<<<1, 256, 0, st>>>(w.aff.total, 1, 0, AffElem<NX>::ESZ, fwd_carry_out);
    IPOC_LAUNCH_CHECK_N("k_soa_to_aos", st);
    return IPOC_OK;
}

template <int NX, int NU>
static int newton_fwd_apply_impl(int N, int rank, int nranks, const double* fx, const double* fu, const double* Kx,
                                 const double* d, const double* fwd_carries, double* dx, double* du, void* ws,
                                 size_t ws_bytes, cudaStream_t st) {
    const Plan p = make_plan(N, 1, true);
    Bump bp{(char*)ws, (size_t)kCtrlBytes, ws_bytes, false};
    NewtonWs w;
    carve_newton<NX>(bp, p, w);
    if (bp.off > ws_bytes) return IPOC_EWORKSPACE;
    const SideJobs nopj{};
    if (p.hier && p.ngroups <= 32) {   // dx_0 = 0 chained through the earlier ranks' aggregates in the leaf kernel
        const HierIn hin = make_hier_chain(p, w.aff, fwd_carries, 0, +1, rank, nullptr);
        return run_fwd_down<NX, NU>(p, w, fx, fu, nullptr, Kx, d, dx, du, false, st, nopj, hin, TailJob{});
    }
    double* seed0 = w.scratch;
    k_aff_seed<NX><<<1, 32, 0, st>>>(nullptr, 1, seed0);
    IPOC_LAUNCH_CHECK_N("k_aff_seed", st);
    k_chain_seed<AffOp<NX>><<<1, 32, 0, st>>>(fwd_carries, 0, +1, rank, seed0, w.aff.seed);
    IPOC_LAUNCH_CHECK_N("k_chain_seed_aff", st);
    if (int rc = run_levels_down<AffOp<NX>>(p, w.aff, st)) return rc;
    return run_fwd_down<NX, NU>(p, w, fx, fu, nullptr, Kx, d, dx, du, false, st, nopj, HierIn{}, TailJob{});
}

template <int NX>
static int affine_reduce_impl(int reverse, int transpose, int N, const double* F, const double* c, double* carry_out,
                              void* ws, size_t ws_bytes, cudaStream_t st) {
    const Plan p = make_plan(N, 1, true);
    Bump bp{(char*)ws, (size_t)kCtrlBytes, ws_bytes, false};
    ScanWs w;
    carve_scan(bp, p, AffElem<NX>::ESZ, NX, w);
    if (bp.off > ws_bytes) return IPOC_EWORKSPACE;
    if (p.hier && p.ngroups <= 32) {
        const Hier h = make_hier<AffOp<NX>>(p, w, nullptr, carry_out);
        return aff_up<NX>(p, w, F, c, reverse, transpose, st, h);
    }
    if (int rc = aff_up<NX>(p, w, F, c, reverse, transpose, st, Hier{})) return rc;
    if (int rc = run_levels<AffOp<NX>>(p, w, true, true, st)) return rc;
    k_soa_to_aos<<<1, 256, 0, st>>>(w.total, 1, 0, AffElem<NX>::ESZ, carry_out);
    IPOC_LAUNCH_CHECK_N("k_soa_to_aos", st);
    return IPOC_OK;
}

template <int NX>
static int affine_apply_impl(int reverse, int transpose, int N, int rank, int nranks, const double* F,
                             const double* c, const double* carries, const double* seed, double* out, void* ws,
                             size_t ws_bytes, cudaStream_t st) {
    const Plan p = make_plan(N, 1, true);
    Bump bp{(char*)ws, (size_t)kCtrlBytes, ws_bytes, false};
    ScanWs w;
    carve_scan(bp, p, AffElem<NX>::ESZ, NX, w);
    if (bp.off > ws_bytes) return IPOC_EWORKSPACE;
    using AOp = AffOp<NX>;
    if (p.hier && p.ngroups <= 32) {
        const HierIn hin = reverse ? make_hier_chain(p, w, carries, nranks - 1, -1, nranks - 1 - rank, seed)
                                   : make_hier_chain(p, w, carries, 0, +1, rank, seed);
        return aff_down<NX>(p, w, F, c, reverse, transpose, out, st, hin);
    }
    if (reverse)
        k_chain_seed<AOp><<<1, 32, 0, st>>>(carries, nranks - 1, -1, nranks - 1 - rank, seed, w.seed);
    else
        k_chain_seed<AOp><<<1, 32, 0, st>>>(carries, 0, +1, rank, seed, w.seed);
    IPOC_LAUNCH_CHECK_N("k_chain_seed_aff", st);
    if (int rc = run_levels_down<AOp>(p, w, st)) return rc;
    return aff_down<NX>(p, w, F, c, reverse, transpose, out, st, HierIn{});
}

template <int NX>
static size_t ws_bytes_impl(int kind, int N, int batch, bool sharded) {
    const bool aff = kind == IPOC_WS_AFFINE_SCAN && !sharded;
    const Plan p = make_plan(N, sharded ? 1 : batch, sharded, aff ? aff_target_threads() : kTargetThreads,
                             !(aff && g_hier.aff_warps_per_sm > 0));
    Bump bp{nullptr, (size_t)kCtrlBytes, 0, true};
    if (kind == IPOC_WS_AFFINE_SCAN) {
        ScanWs w;
        carve_scan(bp, p, AffElem<NX>::ESZ, NX, w);
    } else {
        NewtonWs w;
        carve_newton<NX>(bp, p, w);
    }
    return bp.off + 256;
}

}  // namespace ipoc
