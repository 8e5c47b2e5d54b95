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

namespace ipoc {

struct Tuning {
    int leaf_chunk, mid_fanin, top_max;
};
// shared, defined in ipoc_api.cu
extern unsigned long long g_launches;
extern Tuning g_tune;
extern int g_literal_lqt;   // 0: q = 0, p = ru (default); 1: literal noc_to_lqt arithmetic
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

// ------------------------------------------------------------------ SoA helpers
template <class T>
IPOC_DEV void soa_load(T& t, const double* __restrict__ base, size_t stride, size_t idx) {
    constexpr int SZ = sizeof(T) / sizeof(double);
#pragma unroll
    for (int c = 0; c < SZ; ++c) t.r[c] = base[(size_t)c * stride + idx];
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
IPOC_DEV void warp_scan(typename Op::Elem& a, double* ws, int lane, bool reverse) {
    constexpr int ESZ = sizeof(typename Op::Elem) / sizeof(double);
    using View = typename Op::View;
#pragma unroll 1
    for (int delta = 1; delta < 32; delta <<= 1) {
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

// Side job a top kernel can do for its sequence (saves a launch): fixed-order reduction of the
// per-warp pred / feasibility partials of K2 -> pred = -1/2 sum d'Gd, feasible = AND (G > 0).
struct PredJob {
    const double* pred_part;
    const int* feas_part;
    int n;
    double* pred;
    int32_t* feasible;
};
IPOC_DEV void pred_reduce(const PredJob& pj, int b, int lane) {   // one full warp; fixed order
    double acc = 0.0;
    int f = 1;
    for (int j = lane; j < pj.n; j += 32) {
        acc += pj.pred_part[(size_t)b * pj.n + j];
        f &= pj.feas_part[(size_t)b * pj.n + j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        f &= __shfl_xor_sync(0xffffffffu, f, o);
    }
    if (lane == 0) {
        pj.pred[b] = -0.5 * acc;
        pj.feasible[b] = f;
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
      double* __restrict__ total, int reduce_only, PredJob pj) {
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
    if (pj.pred != nullptr && w == nw - 1) pred_reduce(pj, b, lane);

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

// Copy step j of every lane-row r of one input array into the stage: the warp cooperates, `CPR`
// granules per row, consecutive lanes on consecutive granules of the same row.  32-bit index
// arithmetic relative to the warp's first row, one wide multiply-add per copy.
template <int CNT>
IPOC_DEV void issue_rows(unsigned dst_arr, const double* __restrict__ g, const RowMap& m, int j, int lane) {
    constexpr int G = row_gran(CNT), CPR = row_cpr(CNT), PITCH = row_pitch(CNT);
    const char* gbase = reinterpret_cast<const char*>(g + m.tb * CNT);
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
};
IPOC_DEV WarpSmem warp_smem(char* smem, const Geom& g, int warp_in_block, long long wg) {
    WarpSmem w;
    w.stage0 = smem + (size_t)warp_in_block * g.pw_bytes;
    w.stage0_s = (unsigned)__cvta_generic_to_shared(w.stage0);
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
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        if (s < T) ld.issue(w.stage0_s + s * Ld::STAGE_BYTES, w.map, reverse ? T - 1 - s : s, lane);
        cp_async_commit();
    }
    int slot = 0;
    for (int it = 0; it < T; ++it) {
        const int j = reverse ? T - 1 - it : it;
        cp_async_wait<NS - 1>();
        __syncwarp();
        if (j < len) fetch(w.stage0 + slot * Ld::STAGE_BYTES);
        __syncwarp();
        if (it + NS < T)
            ld.issue(w.stage0_s + slot * Ld::STAGE_BYTES, w.map, reverse ? j - NS : j + NS, lane);
        cp_async_commit();
        if (j < len) compute(j);
        slot = (slot + 1 == NS) ? 0 : slot + 1;
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
    static constexpr int O_FX = 0, O_FU = O_FX + arr_bytes(NX * NX), O_Q = O_FU + arr_bytes(NX * NU),
                         O_R = O_Q + arr_bytes(NX * NX), O_M = O_R + arr_bytes(NU * NU),
                         O_RU = O_M + arr_bytes(NX * NU), STAGE_BYTES = O_RU + arr_bytes(NU);
    IPOC_DEV void issue(unsigned st, const RowMap& m, int j, int lane) const {
        issue_rows<NX * NX>(st + O_FX, fx, m, j, lane);
        issue_rows<NX * NU>(st + O_FU, fu, m, j, lane);
        issue_rows<NX * NX>(st + O_Q, Q, m, j, lane);
        issue_rows<NU * NU>(st + O_R, R, m, j, lane);
        issue_rows<NX * NU>(st + O_M, M, m, j, lane);
        issue_rows<NU>(st + O_RU, ru, m, j, lane);
    }
    // per-lane constant fetched ONCE before the walk (a global load inside the step loop would expose
    // its full latency every step: 16 % of the stall samples in profiles/r01)
    IPOC_DEV double aux(int b) const { return __ldg(reg + b); }
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
    IPOC_DEV void issue(unsigned st, const RowMap& m, int j, int lane) const {
        issue_rows<NX * NX>(st + O_A, A, m, j, lane);
        issue_rows<NX * NU>(st + O_B, B, m, j, lane);
        if (c != nullptr) issue_rows<NX>(st + O_C, c, m, j, lane);
        issue_rows<NX * NX>(st + O_X, X, m, j, lane);
        issue_rows<NU * NU>(st + O_U, U, m, j, lane);
        issue_rows<NX * NU>(st + O_M, M, m, j, lane);
        issue_rows<NX>(st + O_Q, q, m, j, lane);
        issue_rows<NU>(st + O_P, p, m, j, lane);
    }
    IPOC_DEV double aux(int) const { return 0.0; }
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
    IPOC_DEV void issue(unsigned st, const RowMap& m, int j, int lane) const {
        issue_rows<NX * NX>(st + O_A, A, m, j, lane);
        issue_rows<NX * NU>(st + O_B, B, m, j, lane);
        if (c != nullptr) issue_rows<NX>(st + O_C, c, m, j, lane);
        issue_rows<NU * NX>(st + O_K, Kx, m, j, lane);
        issue_rows<NU>(st + O_D, d, m, j, lane);
    }
};

// one affine array pair: F (NX x NX), c (NX)
template <int NX>
struct AffLoader {
    const double *F, *c;
    static constexpr int O_F = 0, O_C = O_F + arr_bytes(NX * NX), STAGE_BYTES = O_C + arr_bytes(NX);
    IPOC_DEV void issue(unsigned st, const RowMap& m, int j, int lane) const {
        issue_rows<NX * NX>(st + O_F, F, m, j, lane);
        issue_rows<NX>(st + O_C, c, m, j, lane);
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
              size_t a1stride, SeedJob sj) {
    extern __shared__ __align__(16) char smem[];
    if (sj.ST != nullptr && blockIdx.x == 0)
        for (int b = threadIdx.x; b < sj.batch; b += blockDim.x) make_ric_seed<NX>(sj, b);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (wg >= total_warps(g)) return;
    const Lane L = lane_info(g, wg, lane);
    const WarpSmem w = warp_smem(smem, g, wib, wg);
    RicElem<NX> a;
    RicOp<NX>::identity(a);
    StepElem<NX, NU> e;   // formed in the fetch phase: smaller than the raw step, lives across the copy issue
    const double aux = ld.aux(L.b);
    staged_walk<IPOC_NS_RIC_UP>(ld, w, g.T0, L.len, lane, true,
                [&](const char* st) {
                    StepLQ<NX, NU> s;
                    ld.read(s, st, lane, aux);
                    make_step_elem(e, s);
                },
                [&](int) { ric_prepend_step(a, e); });
    warp_scan<RicOp<NX>>(a, reinterpret_cast<double*>(w.stage0), lane, true);
    soa_store(a, incl, istride, (size_t)L.slot);
    if (lane == 0) soa_store(a, agg1, a1stride, (size_t)L.b * g.nW + (g.nW - 1 - L.wi));
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
                double* __restrict__ fincl, size_t fistride, double* __restrict__ fagg1, size_t fa1stride) {
    extern __shared__ __align__(16) char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (wg >= total_warps(g)) return;
    const Lane L = lane_info(g, wg, lane);
    const WarpSmem w = warp_smem(smem, g, wib, wg);
    const int N = g.N;
    RicVal<NX> val;
    if (g.per_lane) {
        soa_load(val, wvals, wvstride, (size_t)L.b);
    } else {
        soa_load(val, wvals, wvstride, (size_t)L.b * g.nW + (g.nW - 1 - L.wi));
        if (lane < 31) {   // exclusive prefix (scan order) = inclusive aggregate of the next lane
            RicElem<NX> ex;
            soa_load(ex, incl, istride, (size_t)L.slot + 1);
            RicOp<NX>::apply(val, ex, val);
        }
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
    }
}

// pred / feasible finalisation as a stand-alone launch (only when no K3 top scan follows that could
// carry it as a side job).
static __global__ void __launch_bounds__(32) k_finalize_pred(PredJob pj) { pred_reduce(pj, blockIdx.x, threadIdx.x); }

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
                double* __restrict__ x_out, double* __restrict__ u_out) {
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
        soa_load(xv, xw, xwstride, (size_t)L.b * g.nW + L.wi);
        if (lane > 0) {
            AffElem<NX> ex;
            soa_load(ex, fincl, fistride, (size_t)L.slot - 1);
            AffOp<NX>::apply(xv, ex, xv);
        }
    }
    double x[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = xv.r[i];
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
        st_vec<NX>(x_out + ((size_t)L.b * (N + 1) + L.k0 + j) * NX, x);
        st_vec<NU>(u_out + t * NU, u);
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
    if (L.len > 0 && L.k0 + L.len == N) st_vec<NX>(x_out + ((size_t)L.b * (N + 1) + N) * NX, x);
}

// K3 leaf up (only for the stand-alone par_fwd_pass API; the Newton step gets these aggregates
// from K2's down-sweep).
template <int NX, int NU>
__global__ void __launch_bounds__(kLeafThreads)
k_fwd_leaf_up(FwdLoader<NX, NU> ld, Geom g, double* __restrict__ fincl, size_t fistride,
              double* __restrict__ fagg1, size_t fa1stride) {
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
}

// ------------------------------------------------------------------ K1 (generic affine scan) leaves
template <int NX>
__global__ void __launch_bounds__(kLeafThreads)
k_aff_leaf_up(AffLoader<NX> ld, int reverse, int transpose, Geom g, double* __restrict__ incl, size_t istride,
              double* __restrict__ agg1, size_t a1stride) {
    extern __shared__ __align__(16) char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (wg >= total_warps(g)) return;
    const Lane L = lane_info(g, wg, lane);
    const WarpSmem w = warp_smem(smem, g, wib, wg);
    AffElem<NX> a;
    AffOp<NX>::identity(a);
    AffElem<NX> se;
    staged_walk<IPOC_NS_LIGHT>(ld, w, g.T0, L.len, lane, reverse != 0,
                               [&](const char* st) { ld.read(se, st, lane, transpose); },
                               [&](int) { AffOp<NX>::compose(a, a, se); });
    warp_scan<AffOp<NX>>(a, reinterpret_cast<double*>(w.stage0), lane, reverse != 0);
    soa_store(a, incl, istride, (size_t)L.slot);
    if (reverse) {
        if (lane == 0) soa_store(a, agg1, a1stride, (size_t)L.b * g.nW + (g.nW - 1 - L.wi));
    } else {
        if (lane == 31) soa_store(a, agg1, a1stride, (size_t)L.b * g.nW + L.wi);
    }
}

template <int NX>
__global__ void __launch_bounds__(kLeafThreads)
k_aff_leaf_down(AffLoader<NX> ld, int reverse, int transpose, Geom g, const double* __restrict__ incl,
                size_t istride, const double* __restrict__ wvals, size_t wvstride, double* __restrict__ out) {
    extern __shared__ __align__(16) char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (wg >= total_warps(g)) return;
    const Lane L = lane_info(g, wg, lane);
    const WarpSmem w = warp_smem(smem, g, wib, wg);
    const int N = g.N;
    AffVal<NX> x;
    if (g.per_lane) {
        soa_load(x, wvals, wvstride, (size_t)L.b);
    } else if (reverse) {
        soa_load(x, wvals, wvstride, (size_t)L.b * g.nW + (g.nW - 1 - L.wi));
        if (lane < 31) {
            AffElem<NX> ex;
            soa_load(ex, incl, istride, (size_t)L.slot + 1);
            AffOp<NX>::apply(x, ex, x);
        }
    } else {
        soa_load(x, wvals, wvstride, (size_t)L.b * g.nW + L.wi);
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
};

// force_scan: time-sharded mode always wants the segment total, hence at least one level.
static Plan make_plan(int N, int batch, bool force_scan = false, int target_threads = kTargetThreads) {
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
            if (t < t_lat) t = t_lat;
            T0 = (int)(t < 4 ? 4 : t);
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
};

static void carve_scan(Bump& bp, const Plan& p, int esz, int vsz, ScanWs& w) {
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
    carve_scan(bp, p, AffElem<NX>::ESZ, NX, w.aff);
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
    const size_t per_warp = body;
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
                      PredJob pj = PredJob{nullptr, nullptr, 0, nullptr, nullptr}) {
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
static int run_bwd_up(const Plan& p, const NewtonWs& w, const Loader& ld, cudaStream_t st,
                      SeedJob sj = SeedJob{nullptr, 0, nullptr, 0, nullptr, nullptr}) {
    const LeafLaunch ll = leaf_launch(p, Loader::STAGE_BYTES, scan_scratch_bytes<RicOp<NX>>(), IPOC_NS_RIC_UP);
    Geom g = p.g;
    g.pw_bytes = (int)(ll.smem / ll.wpc);
    auto kern = k_ric_leaf_up<NX, NU, Loader>;
    if (int rc = set_smem(kern, ll.smem, ll.threads)) return rc;
    kern<<<ll.grid, ll.threads, ll.smem, st>>>(ld, g, w.ric.incl, (size_t)p.slots, w.ric.agg[0],
                                              (size_t)p.g.batch * p.g.nW, sj);
    IPOC_LAUNCH_CHECK_N("k_ric_leaf_up", st);
    return IPOC_OK;
}

template <int NX, int NU, class Loader>
static int run_bwd_down(const Plan& p, const NewtonWs& w, const Loader& ld, double* Kx, double* d, double* S,
                        double* v, double* pred, int32_t* feasible, bool want_fwd_agg, cudaStream_t st,
                        bool defer_pred = false) {
    const double* vals;
    size_t vstride;
    leaf_values(p, w.ric, vals, vstride);
    const LeafLaunch ll = leaf_launch(p, Loader::STAGE_BYTES, scan_scratch_bytes<AffOp<NX>>(), IPOC_NS_RIC_DOWN);
    Geom g = p.g;
    g.pw_bytes = (int)(ll.smem / ll.wpc);
    auto kern = k_ric_leaf_down<NX, NU, Loader>;
    if (int rc = set_smem(kern, ll.smem, ll.threads)) return rc;
    const bool fwd = want_fwd_agg && !p.g.per_lane;
    kern<<<ll.grid, ll.threads, ll.smem, st>>>(ld, g, w.ric.incl, (size_t)p.slots, vals, vstride, Kx, d, S, v,
                                              w.pred_part, w.feas_part, pred, feasible, fwd ? w.aff.incl : nullptr,
                                              (size_t)p.slots, fwd ? w.aff.agg[0] : nullptr,
                                              (size_t)p.g.batch * p.g.nW);
    IPOC_LAUNCH_CHECK_N("k_ric_leaf_down", st);
    if (!p.g.per_lane && !defer_pred) {
        k_finalize_pred<<<p.g.batch, 32, 0, st>>>(PredJob{w.pred_part, w.feas_part, p.g.nW, pred, feasible});
        IPOC_LAUNCH_CHECK_N("k_finalize_pred", st);
    }
    return IPOC_OK;
}

template <int NX, int NU, class Loader>
static int run_bwd(const Plan& p, const NewtonWs& w, const Loader& ld, double* Kx, double* d, double* S, double* v,
                   double* pred, int32_t* feasible, bool want_fwd_agg, cudaStream_t st, bool defer_pred, SeedJob sj) {
    if (p.g.per_lane) {   // no up-sweep to piggy-back on
        k_ric_seed<NX><<<grid_for(sj.batch, 128), 128, 0, st>>>(sj);
        IPOC_LAUNCH_CHECK_N("k_ric_seed", st);
    } else {
        if (int rc = run_bwd_up<NX, NU>(p, w, ld, st, sj)) return rc;
        if (p.nlev > 0)
            if (int rc = run_levels<RicOp<NX>>(p, w.ric, false, false, st)) return rc;
    }
    return run_bwd_down<NX, NU>(p, w, ld, Kx, d, S, v, pred, feasible, want_fwd_agg, st, defer_pred);
}

// ---- K3 given in-warp forward aggregates in w.aff.incl / w.aff.agg[0] and the seed in w.aff.seed
template <int NX, int NU>
static int run_fwd_down(const Plan& p, const NewtonWs& w, const double* A, const double* B, const double* c,
                        const double* Kx, const double* d, double* x, double* u, bool levels, cudaStream_t st,
                        PredJob pj = PredJob{nullptr, nullptr, 0, nullptr, nullptr}) {
    if (levels && p.nlev > 0)
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
    kern<<<ll.grid, ll.threads, ll.smem, st>>>(ld, g, w.aff.incl, (size_t)p.slots, vals, vstride, x, u);
    IPOC_LAUNCH_CHECK_N("k_fwd_leaf_down", st);
    return IPOC_OK;
}

template <int NX, int NU>
static int newton_step_impl(int N, int batch, const double* fx, const double* fu, const double* ru, const double* Q,
                            const double* R, const double* M, const double* reg, double* dx, double* du, double* Kx,
                            double* d, double* pred, int32_t* feasible, void* ws, size_t ws_bytes, cudaStream_t st) {
    const Plan p = make_plan(N, batch);
    Bump bp{(char*)ws, 0, ws_bytes, false};
    NewtonWs w;
    carve_newton<NX>(bp, p, w);
    if (bp.off > ws_bytes) return IPOC_EWORKSPACE;
    // terminal value function: XT = Q[0], HT = I, rT = 0 (ref noc/par_interior_point_newton.py:73-75);
    // zero initial deviation dx_0 = 0 (:122) — both seeds in one launch
    const SeedJob sj{Q, (size_t)N * NX * NX, nullptr, batch, w.ric.seed, w.aff.seed};
    // the pred / feasibility partials are folded by K3's top scan when there is one
    const bool defer = p.nlev > 0;
    if (g_literal_lqt) {
        NewtonLoader<NX, NU, true> ld{fx, fu, ru, Q, R, M, reg};
        if (int rc = run_bwd<NX, NU>(p, w, ld, Kx, d, nullptr, nullptr, pred, feasible, true, st, defer, sj)) return rc;
    } else {
        NewtonLoader<NX, NU, false> ld{fx, fu, ru, Q, R, M, reg};
        if (int rc = run_bwd<NX, NU>(p, w, ld, Kx, d, nullptr, nullptr, pred, feasible, true, st, defer, sj)) return rc;
    }
    PredJob pj{nullptr, nullptr, 0, nullptr, nullptr};
    if (defer) pj = PredJob{w.pred_part, w.feas_part, p.g.nW, pred, feasible};
    return run_fwd_down<NX, NU>(p, w, fx, fu, nullptr, Kx, d, dx, du, true, st, pj);
}

template <int NX, int NU>
static int lqt_bwd_impl(int N, int batch, const double* A, const double* B, const double* c, const double* X,
                        const double* U, const double* M, const double* q, const double* pp, const double* ST,
                        const double* vT, double* Kx, double* d, double* S, double* v, double* pred,
                        int32_t* feasible, void* ws, size_t ws_bytes, cudaStream_t st) {
    const Plan p = make_plan(N, batch);
    Bump bp{(char*)ws, 0, ws_bytes, false};
    NewtonWs w;
    carve_newton<NX>(bp, p, w);
    if (bp.off > ws_bytes) return IPOC_EWORKSPACE;
    const SeedJob sj{ST, (size_t)NX * NX, vT, batch, w.ric.seed, nullptr};
    LqtLoader<NX, NU> ld{A, B, c, X, U, M, q, pp};
    return run_bwd<NX, NU>(p, w, ld, Kx, d, S, v, pred, feasible, false, st, false, sj);
}

template <int NX, int NU>
static int run_fwd_up(const Plan& p, const NewtonWs& w, const double* A, const double* B, const double* c,
                      const double* Kx, const double* d, cudaStream_t st) {
    FwdLoader<NX, NU> ld{A, B, c, Kx, d};
    const LeafLaunch ll = leaf_launch(p, FwdLoader<NX, NU>::STAGE_BYTES, scan_scratch_bytes<AffOp<NX>>(), IPOC_NS_LIGHT);
    Geom g = p.g;
    g.pw_bytes = (int)(ll.smem / ll.wpc);
    auto kern = k_fwd_leaf_up<NX, NU>;
    if (int rc = set_smem(kern, ll.smem, ll.threads)) return rc;
    kern<<<ll.grid, ll.threads, ll.smem, st>>>(ld, g, w.aff.incl, (size_t)p.slots, w.aff.agg[0],
                                              (size_t)p.g.batch * p.g.nW);
    IPOC_LAUNCH_CHECK_N("k_fwd_leaf_up", st);
    return IPOC_OK;
}

template <int NX, int NU>
static int lqt_fwd_impl(int N, int batch, const double* A, const double* B, const double* c, const double* Kx,
                        const double* d, const double* x0, double* u, double* x, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
    const Plan p = make_plan(N, batch);
    Bump bp{(char*)ws, 0, ws_bytes, false};
    NewtonWs w;
    carve_newton<NX>(bp, p, w);
    if (bp.off > ws_bytes) return IPOC_EWORKSPACE;
    k_aff_seed<NX><<<grid_for(batch, 128), 128, 0, st>>>(x0, batch, w.aff.seed);
    IPOC_LAUNCH_CHECK_N("k_aff_seed", st);
    if (!p.g.per_lane)
        if (int rc = run_fwd_up<NX, NU>(p, w, A, B, c, Kx, d, st)) return rc;
    return run_fwd_down<NX, NU>(p, w, A, B, c, Kx, d, x, u, true, st);
}

template <int NX>
static int aff_up(const Plan& p, const ScanWs& w, const double* F, const double* c, int reverse, int transpose,
                  cudaStream_t st) {
    AffLoader<NX> ld{F, c};
    const LeafLaunch ll = leaf_launch(p, AffLoader<NX>::STAGE_BYTES, scan_scratch_bytes<AffOp<NX>>(), IPOC_NS_LIGHT);
    Geom g = p.g;
    g.pw_bytes = (int)(ll.smem / ll.wpc);
    auto kern = k_aff_leaf_up<NX>;
    if (int rc = set_smem(kern, ll.smem, ll.threads)) return rc;
    kern<<<ll.grid, ll.threads, ll.smem, st>>>(ld, reverse, transpose, g, w.incl, (size_t)p.slots, w.agg[0],
                                              (size_t)p.g.batch * p.g.nW);
    IPOC_LAUNCH_CHECK_N("k_aff_leaf_up", st);
    return IPOC_OK;
}
template <int NX>
static int aff_down(const Plan& p, const ScanWs& w, const double* F, const double* c, int reverse, int transpose,
                    double* out, cudaStream_t st) {
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
                                              out);
    IPOC_LAUNCH_CHECK_N("k_aff_leaf_down", st);
    return IPOC_OK;
}

template <int NX>
static int affine_scan_impl(int reverse, int transpose, int N, int batch, const double* F, const double* c,
                            const double* seed, double* out, void* ws, size_t ws_bytes, cudaStream_t st) {
    const Plan p = make_plan(N, batch, false, kAffTargetThreads);
    Bump bp{(char*)ws, 0, ws_bytes, false};
    ScanWs w;
    carve_scan(bp, p, AffElem<NX>::ESZ, NX, w);
    if (bp.off > ws_bytes) return IPOC_EWORKSPACE;
    if (batch == 1 && seed != nullptr) {
        w.seed = const_cast<double*>(seed);   // one problem: the SoA seed plane IS the caller's vector
    } else {
        k_aff_seed<NX><<<grid_for(batch, 128), 128, 0, st>>>(seed, batch, w.seed);
        IPOC_LAUNCH_CHECK_N("k_aff_seed", st);
    }
    if (!p.g.per_lane) {
        if (int rc = aff_up<NX>(p, w, F, c, reverse, transpose, st)) return rc;
        if (p.nlev > 0)
            if (int rc = run_levels<AffOp<NX>>(p, w, false, false, st)) return rc;
    }
    return aff_down<NX>(p, w, F, c, reverse, transpose, out, st);
}

// ---- time-sharded split-phase implementations ----------------------------------------------
// The reduce phase leaves the in-warp aggregates and the level arrays in the workspace; the apply
// phase (same workspace, same plan) only runs the seeded way down.
template <int NX, int NU>
static int newton_bwd_reduce_impl(int N, const double* fx, const double* fu, const double* ru, const double* Q,
                                  const double* R, const double* M, const double* reg, double* carry_out, void* ws,
                                  size_t ws_bytes, cudaStream_t st) {
    const Plan p = make_plan(N, 1, true);
    Bump bp{(char*)ws, 0, ws_bytes, false};
    NewtonWs w;
    carve_newton<NX>(bp, p, w);
    if (bp.off > ws_bytes) return IPOC_EWORKSPACE;
    if (g_literal_lqt) {
        NewtonLoader<NX, NU, true> ld{fx, fu, ru, Q, R, M, reg};
        if (int rc = run_bwd_up<NX, NU>(p, w, ld, st)) return rc;
    } else {
        NewtonLoader<NX, NU, false> ld{fx, fu, ru, Q, R, M, reg};
        if (int rc = run_bwd_up<NX, NU>(p, w, ld, st)) return rc;
    }
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
                                               w.val[L - 1], (size_t)p.n[L - 1], nullptr, 0,
                                               PredJob{nullptr, nullptr, 0, nullptr, nullptr});
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
    Bump bp{(char*)ws, 0, ws_bytes, false};
    NewtonWs w;
    carve_newton<NX>(bp, p, w);
    if (bp.off > ws_bytes) return IPOC_EWORKSPACE;
    // terminal seed of the whole horizon, pushed back through the later ranks' aggregates
    double* seed0 = w.scratch;   // RicVal packed
    k_ric_seed<NX><<<1, 32, 0, st>>>(SeedJob{ST, (size_t)NX * NX, nullptr, 1, seed0, nullptr});
    IPOC_LAUNCH_CHECK_N("k_ric_seed", st);
    k_chain_seed<RicOp<NX>><<<1, 32, 0, st>>>(carries, nranks - 1, -1, nranks - 1 - rank, seed0, w.ric.seed);
    IPOC_LAUNCH_CHECK_N("k_chain_seed_ric", st);
    if (int rc = run_levels_down<RicOp<NX>>(p, w.ric, st)) return rc;
    if (g_literal_lqt) {
        NewtonLoader<NX, NU, true> ld{fx, fu, ru, Q, R, M, reg};
        if (int rc = run_bwd_down<NX, NU>(p, w, ld, Kx, d, nullptr, nullptr, pred, feasible, true, st)) return rc;
    } else {
        NewtonLoader<NX, NU, false> ld{fx, fu, ru, Q, R, M, reg};
        if (int rc = run_bwd_down<NX, NU>(p, w, ld, Kx, d, nullptr, nullptr, pred, feasible, true, st)) return rc;
    }
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
    Bump bp{(char*)ws, 0, ws_bytes, false};
    NewtonWs w;
    carve_newton<NX>(bp, p, w);
    if (bp.off > ws_bytes) return IPOC_EWORKSPACE;
    double* seed0 = w.scratch;
    k_aff_seed<NX><<<1, 32, 0, st>>>(nullptr, 1, seed0);
    IPOC_LAUNCH_CHECK_N("k_aff_seed", st);
    k_chain_seed<AffOp<NX>><<<1, 32, 0, st>>>(fwd_carries, 0, +1, rank, seed0, w.aff.seed);
    IPOC_LAUNCH_CHECK_N("k_chain_seed_aff", st);
    if (int rc = run_levels_down<AffOp<NX>>(p, w.aff, st)) return rc;
    return run_fwd_down<NX, NU>(p, w, fx, fu, nullptr, Kx, d, dx, du, false, st);
}

template <int NX>
static int affine_reduce_impl(int reverse, int transpose, int N, const double* F, const double* c, double* carry_out,
                              void* ws, size_t ws_bytes, cudaStream_t st) {
    const Plan p = make_plan(N, 1, true);
    Bump bp{(char*)ws, 0, ws_bytes, false};
    ScanWs w;
    carve_scan(bp, p, AffElem<NX>::ESZ, NX, w);
    if (bp.off > ws_bytes) return IPOC_EWORKSPACE;
    if (int rc = aff_up<NX>(p, w, F, c, reverse, transpose, st)) return rc;
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
    Bump bp{(char*)ws, 0, ws_bytes, false};
    ScanWs w;
    carve_scan(bp, p, AffElem<NX>::ESZ, NX, w);
    if (bp.off > ws_bytes) return IPOC_EWORKSPACE;
    using AOp = AffOp<NX>;
    if (reverse)
        k_chain_seed<AOp><<<1, 32, 0, st>>>(carries, nranks - 1, -1, nranks - 1 - rank, seed, w.seed);
    else
        k_chain_seed<AOp><<<1, 32, 0, st>>>(carries, 0, +1, rank, seed, w.seed);
    IPOC_LAUNCH_CHECK_N("k_chain_seed_aff", st);
    if (int rc = run_levels_down<AOp>(p, w, st)) return rc;
    return aff_down<NX>(p, w, F, c, reverse, transpose, out, st);
}

template <int NX>
static size_t ws_bytes_impl(int kind, int N, int batch, bool sharded) {
    const Plan p = make_plan(N, sharded ? 1 : batch, sharded,
                             (kind == IPOC_WS_AFFINE_SCAN && !sharded) ? kAffTargetThreads : kTargetThreads);
    Bump bp{nullptr, 0, 0, true};
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
