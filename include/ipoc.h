/* ipoc.h — C ABI of the B200-native (sm_100a, FP64) scans behind the parallel-in-time
 * interior-point Newton step of casiacob/ip-parallel-optimal-control.
 *
 * This is the drop-in boundary.  The reference has NO native interface for this path: it is
 * pure Python/JAX (`jax.lax.associative_scan` + `vmap`), so each entry point below cites the
 * reference *Python* interface it replaces; INTEGRATION.md shows the `jax.ffi` / ctypes
 * bindings a maintainer of the reference would add on top of these symbols.
 *
 * Conventions
 *   - All tensors are float64, row-major, contiguous, 16-byte aligned, in DEVICE memory
 *     (the *_host_* wrappers at the end take HOST memory and copy inside).
 *   - `batch` independent problems are stacked on a leading axis (batch = 1 for a single OCP);
 *     time is the next axis.  Shapes are given per problem.
 *   - Every call only ENQUEUES work on `stream` (no allocation, no implicit synchronisation,
 *     no host reads) and is therefore CUDA-graph capturable.  The caller owns all buffers
 *     including the workspace (`ipoc_workspace_bytes`).  The library keeps no mutable state
 *     apart from the optional tuning knobs below.
 *   - WORKSPACE CONTROL BLOCK: the first IPOC_WS_CONTROL_BYTES of a scan workspace (every kind except
 *     IPOC_WS_REDUCTIONS) hold the arrival counters of the in-kernel scan levels.  They must be ZERO when
 *     the workspace is first used — call ipoc_workspace_init once after allocating (or cudaMemset) — and
 *     every call leaves them zero again (each counter wraps to 0 on its last arrival), so replays of
 *     captured graphs and calls with other problem sizes on the same workspace need nothing.  A
 *     workspace must not be used by two calls at the same time.
 *   - Return value: 0 on success, a negative IPOC_E* code otherwise.  Numerical failure (NaN,
 *     non-positive-definite G) is DATA (`feasible = 0`), never an error — it is an ordinary
 *     branch of the algorithm (ref noc/par_interior_point_newton.py:166).
 *   - There is no CPU fallback: unsupported (nx, nu) -> IPOC_EUNSUPPORTED_DIM.
 *     Supported: nx in {1..8}, nu in {1..min(nx, 4)} (see ipoc_supported).
 */
#ifndef IPOC_H
#define IPOC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ipoc_stream_t; /* cudaStream_t */

enum {
    IPOC_OK = 0,
    IPOC_EUNSUPPORTED_DIM = -1,
    IPOC_EWORKSPACE = -2,
    IPOC_ECUDA = -3,
    IPOC_ENCCL = -4,
    IPOC_EINVAL = -5,
    IPOC_EALIGN = -6
};

enum { /* `kind` of ipoc_workspace_bytes */
    IPOC_WS_NEWTON_STEP = 0,
    IPOC_WS_LQT_BWD = 1,
    IPOC_WS_LQT_FWD = 2,
    IPOC_WS_AFFINE_SCAN = 3,
    IPOC_WS_REDUCTIONS = 4,
    IPOC_WS_COSTATES = 5,       /* ipoc_costates_f64 */
    IPOC_WS_NEWTON_ATTEMPT = 6  /* ipoc_newton_attempt_f64: pass max(nu, nc) in place of nu when cons is given */
};

#define IPOC_WS_CONTROL_BYTES 65536
/* Zero the control block of a freshly allocated workspace (stream-ordered; once per allocation). */
int ipoc_workspace_init(void* ws, size_t ws_bytes, ipoc_stream_t stream);

const char* ipoc_strerror(int code);
int ipoc_version(void);
/* 1 if kernels for (nx, nu) are compiled in, else 0. */
int ipoc_supported(int nx, int nu);
size_t ipoc_workspace_bytes(int kind, int N, int nx, int nu, int batch);

/* Optional tuning knobs (0 = keep default): leaf chunk length (steps folded per thread),
 * mid-level fan-in, maximum number of aggregates handled by the single-CTA top scan. */
void ipoc_set_tuning(int leaf_chunk, int mid_fanin, int top_max);

/* Newton-step entry points only: 0 (default) = use the closed form of `noc_to_lqt`'s references
 * (q = 0, p = ru: what ref noc/par_interior_point_newton.py:62-66 evaluates to in exact arithmetic);
 * 1 = follow those lines operation by operation (X^-1 M by pivoted LU, s, r, then fold back).
 * The two differ by rounding of order eps*cond(Q); both are tested against the oracle. */
void ipoc_set_literal_lqt(int on);

/* Scan organisation knob (tests / experiments): by default the levels above the warp scans are completed
 * inside the leaf kernels by their last-arriving warps ("hierarchical" plans: no top / mid kernels, no
 * spin-waiting) for the affine scans and for Riccati scans of more than 32 groups of warps; Riccati scans
 * of 6 ... 32 groups run their levels in ONE lane-cooperative level kernel (every combine of a round on up
 * to 8 lanes, the groups on separate SMs), shorter ones in a single-CTA top kernel.  enabled = 0 selects the
 * separate single-thread level kernels everywhere, 2 the in-kernel levels everywhere, 3 the mix without
 * the cooperative kernel (single-CTA top kernel under 24 groups), 4 the cooperative kernel from 1 group; group_warps (<= 32, 0 = 32) and serial_top (<= 32, 0 = 8) shape the hierarchy.  ipoc_set_tuning with a non-zero mid_fanin or top_max
 * also selects the separate level kernels. */
void ipoc_set_hier(int enabled, int group_warps, int serial_top);
/* Leaf warps per SM the stand-alone affine scans are planned for (0 = default); experiment knob. */
void ipoc_set_affine_occupancy(int warps_per_sm);

/* ---- K2 + K3: one Newton step ----------------------------------------------------------
 * Replaces `par_Newton` (ref noc/par_interior_point_newton.py:107-124) from the regularisation
 * add onwards: R + reg*I (:118), `noc_to_lqt` (:50-84, r/s by two small solves per step,
 * XT = Q[0], H = Z = I, c = 0), `paroc.par_bwd_pass` (:120: reverse associative scan over
 * (A, b, C, eta, J), gains, pred_reduction, feasibility) and `paroc.par_fwd_pass` with zero
 * initial deviation (:121-123).
 *   in : fx (N,nx,nx) fu (N,nx,nu) ru (N,nu) Q (N,nx,nx) R (N,nu,nu) M (N,nx,nu)
 *        reg (batch) device scalars = reg_param * ||cu||_F (:116-117)
 *   out: dx (N+1,nx) du (N,nu) Kx (N,nu,nx) d (N,nu) pred (batch) feasible (batch, int32)
 */
int ipoc_newton_step_f64(int N, int nx, int nu, int batch,
                         const double* fx, const double* fu, const double* ru,
                         const double* Q, const double* R, const double* M,
                         const double* reg,
                         double* dx, double* du, double* Kx, double* d,
                         double* pred, int32_t* feasible,
                         void* ws, size_t ws_bytes, ipoc_stream_t stream);

/* ---- K2 alone: raw `paroc.par_bwd_pass(lqt)` --------------------------------------------
 * (call sites ref noc/par_interior_point_newton.py:120, examples/linear_mpc_parallel.py:68)
 * for an LQT problem already reduced to effective terms (H, Z folded in by the caller):
 *   x+ = A x + B u + c,  stage cost 1/2 x'Xx + 1/2 u'Uu + x'Mu + q'x + p'u,
 *   terminal value S_N = ST, v_N = vT   (V(x) = 1/2 x'Sx - v'x).
 *   in : A (N,nx,nx) B (N,nx,nu) c (N,nx) X (N,nx,nx) U (N,nu,nu) M (N,nx,nu) q (N,nx) p (N,nu)
 *        ST (nx,nx) vT (nx)  per problem
 *   out: Kx (N,nu,nx) d (N,nu) [S (N+1,nx,nx) v (N+1,nx) — may be NULL] pred, feasible
 */
int ipoc_lqt_bwd_f64(int N, int nx, int nu, int batch,
                     const double* A, const double* B, const double* c,
                     const double* X, const double* U, const double* M,
                     const double* q, const double* p,
                     const double* ST, const double* vT,
                     double* Kx, double* d, double* S, double* v,
                     double* pred, int32_t* feasible,
                     void* ws, size_t ws_bytes, ipoc_stream_t stream);

/* ---- K3 alone: raw `paroc.par_fwd_pass(lqt, x0, Kx, d)` ---------------------------------
 * (call sites ref noc/par_interior_point_newton.py:121-123, examples/linear_mpc_parallel.py:69)
 *   in : A, B, c (c may be NULL = 0), Kx (N,nu,nx), d (N,nu), x0 (nx) per problem
 *   out: u (N,nu), x (N+1,nx)          [u_k = -Kx_k x_k + d_k]
 */
int ipoc_lqt_fwd_f64(int N, int nx, int nu, int batch,
                     const double* A, const double* B, const double* c,
                     const double* Kx, const double* d, const double* x0,
                     double* u, double* x,
                     void* ws, size_t ws_bytes, ipoc_stream_t stream);

/* ---- K1: affine-map scan -----------------------------------------------------------------
 * Replaces `par_costates` / `par_scan` / `combine_fc` (ref noc/costates.py:6-40).
 *   reverse = 0: out[0] = seed,  out[k+1] = F_k    out[k]   + c_k
 *   reverse = 1: out[N] = seed,  out[k]   = F_k(') out[k+1] + c_k    (transpose = 1 uses F_k')
 * Costates: reverse = 1, transpose = 1, F = fx, c = cx, seed = grad final_cost(x_N).
 *   in : F (N,nx,nx) c (N,nx) seed (nx) per problem;   out: (N+1,nx)
 */
int ipoc_affine_scan_f64(int reverse, int transpose, int N, int nx, int batch,
                         const double* F, const double* c, const double* seed, double* out,
                         void* ws, size_t ws_bytes, ipoc_stream_t stream);

/* ---- K1 + ||cu||: `par_costates` with the norm of ref :116 as a side job --------------------------------
 * lam = costates (ref noc/costates.py:34-40: reverse scan of lam_k = cx_k + fx_k' lam_{k+1}, lam_N = lamT)
 * and, if cu (N,nu) is given, cu_norm (batch) = ||cu||_F (ref noc/par_interior_point_newton.py:116) folded
 * into the up-sweep: per-warp partial sums, fixed-order fold by the warp that completes the scan.
 * Workspace: ipoc_workspace_bytes(IPOC_WS_COSTATES, N, nx, nu, batch). */
int ipoc_costates_f64(int N, int nx, int nu, int batch,
                      const double* fx, const double* cx, const double* lamT, const double* cu,
                      double* lam, double* cu_norm, const int32_t* fresh /* may be NULL, see ipoc_plant_* */,
                      void* ws, size_t ws_bytes, ipoc_stream_t stream);

/* ---- K2 + K3 + the glue of one accept/reject attempt, fused (ref :151-175) --------------------------------
 * ipoc_newton_step_f64 with reg = rp * cu_norm formed in the kernel (:117) and, each optional (NULL = off),
 *   hu            = max|ru| (:158), folded into the up-sweep;
 *   tx, tu        = x + dx, u + du (:156-157), written by K3's leaf kernel next to dx, du;
 *   traj_feasible = all(cons <= 0) of a GIVEN array cons (batch,N,nc) (:45-47 — when the constraints depend on
 *                   the trial point the caller evaluates them afterwards and passes traj_feas_in instead);
 *   accept        : rho = (new_cost - cost)/pred, success, rp_out / r_inc update (:159-173, the rule of
 *                   ipoc_accept_update_f64) by the warp that completes K3.  rp_out may alias rp.
 * Five launches in all for one problem; plans that cannot carry a side job (one sequence per lane, explicitly
 * tuned level kernels) run the stand-alone kernels for it, with identical results.
 * Workspace: ipoc_workspace_bytes(IPOC_WS_NEWTON_ATTEMPT, N, nx, max(nu, nc), batch). */
int ipoc_newton_attempt_f64(int N, int nx, int nu, int nc, int batch,
                            const double* fx, const double* fu, const double* ru,
                            const double* Q, const double* R, const double* M,
                            const double* rp, const double* cu_norm,
                            double* dx, double* du, double* Kx, double* d, double* pred, int32_t* feasible,
                            double* hu,
                            const double* x, const double* u, double* tx, double* tu,
                            const double* cons, int32_t* traj_feasible,
                            const double* cost, const double* new_cost, const int32_t* traj_feas_in,
                            const int32_t* active, double* rp_out, double* r_inc, int32_t* success,
                            double* gain_ratio,
                            void* ws, size_t ws_bytes, ipoc_stream_t stream);

/* ---- K4: reductions of the accept/reject test ---------------------------------------------
 * (ref noc/par_interior_point_newton.py:45-47 `all(cons <= 0)`, :116 `norm(d.cu)`, :158
 * `max|ru|`).  Any input pointer may be NULL (its output is then left untouched).
 *   in : ru (N,nu) cu (N,nu) cons (N,nc);  out per problem: hu_norm, cu_norm, traj_feasible
 * If `rp` and `reg` are given (with cu), also reg = rp * ||cu||_F (:117) — the device scalar that
 * ipoc_newton_step_f64 consumes, so the regularisation never visits the host.
 * Two launches (per-slice partials, then a fixed-order fold) -> bit-reproducible run to run.
 * Workspace: ipoc_workspace_bytes(IPOC_WS_REDUCTIONS, N, max(nu, nc), nu, batch).
 */
int ipoc_reductions_f64(int N, int nu, int nc, int batch,
                        const double* ru, const double* cu, const double* cons,
                        double* hu_norm, double* cu_norm, int32_t* traj_feasible,
                        const double* rp, double* reg,
                        void* ws, size_t ws_bytes, ipoc_stream_t stream);

/* ---- A3: LQ parameters of the Newton step (optional fusion of a host-framework step) --------------
 * Replaces `compute_lqr_params` (ref noc/par_interior_point_newton.py:31-42): with l = lam[k+1],
 *   ru = cu + fu' l,  Q = cxx + sum_o l_o fxx[o],  R = cuu + sum_o l_o fuu[o],  M = cxu + sum_o l_o fxu[o].
 *   in : lam (N+1,nx), cu (N,nu), cxx (N,nx,nx), cuu (N,nu,nu), cxu (N,nx,nu), fu (N,nx,nu),
 *        fxx (N,nx,nx,nx), fuu (N,nx,nu,nu), fxu (N,nx,nx,nu)     out: ru, Q, R, M
 */
int ipoc_lqr_params_f64(int N, int nx, int nu, int batch, const double* lam,
                        const double* cu, const double* cxx, const double* cuu, const double* cxu,
                        const double* fu, const double* fxx, const double* fuu, const double* fxu,
                        double* ru, double* Q, double* R, double* M, ipoc_stream_t stream);

/* ---- A8: scalar accept / regularisation update, on device --------------------------------
 * (ref noc/par_interior_point_newton.py:159-173) for `batch` independent problems:
 *   new_cost = traj_feasible ? new_cost : inf;  rho = (new_cost - cost) / pred;
 *   success = rho > 0 && bwd_feasible;
 *   rp <- clip(success ? rp*max(1/3, 1-(2 rho-1)^3) : rp*r_inc, 1e-16, 1e16);
 *   r_inc <- success ? 2 : 2*r_inc
 * `active` (may be NULL) masks problems that must be left untouched.
 */
int ipoc_accept_update_f64(int batch, const double* cost, const double* new_cost,
                           const int32_t* traj_feasible, const double* pred,
                           const int32_t* bwd_feasible, const int32_t* active,
                           double* rp, double* r_inc, int32_t* success, double* gain_ratio,
                           ipoc_stream_t stream);

/* ---- attempt-loop glue of `while_inner_loop` (ref noc/par_interior_point_newton.py:151-182) for `batch`
 * independent problems whose loops run on the device (frozen-when-done, like a vmapped lax.while_loop):
 *   begin : active = !done (all active if done == NULL);  reg = rp * cu_norm            (ref :117, :177-182)
 *   trial : tx = x + dx ((N+1)*nx per problem),  tu = u + du (N*nu)                      (ref :156-157)
 *   commit: for active problems  keep_x <- tx, keep_u <- tu (ref :175, kept whether or not the attempt
 *           succeeded), inner += 1 (:174), done |= success || inner > max_attempts (:180-181)
 * `done` is one byte per problem (0/1), `inner` int64, `active`/`success` int32. */
int ipoc_attempt_begin_f64(int batch, const uint8_t* done, const double* rp, const double* cu_norm,
                           int32_t* active, double* reg, ipoc_stream_t stream);
int ipoc_trial_point_f64(int N, int nx, int nu, int batch, const double* x, const double* dx, const double* u,
                         const double* du, double* tx, double* tu, ipoc_stream_t stream);
int ipoc_attempt_commit_f64(int N, int nx, int nu, int batch, const int32_t* active, const int32_t* success,
                            const double* tx, const double* tu, double* keep_x, double* keep_u,
                            int64_t* inner, uint8_t* done, int max_attempts, ipoc_stream_t stream);

/* ---- Newton-loop glue of `newton_oc` (ref noc/par_interior_point_newton.py:184-202) for device-resident loops:
 * for every problem that is not finished (`outer_done` == 0) and whose attempt loop has just ended
 * (`inner_done` != 0):  x <- tx, u <- tu (:184, taken even if no attempt succeeded), iteration += 1 (:194),
 * inner = 0, inner_done = 0, and outer_done = 1 if hu < hu_tol or iteration > max_iterations (:199-202; `hu` is
 * max|ru| of the iterate BEFORE the step, as in the reference).  `advanced` (int32 scratch, one per problem)
 * receives which problems moved on.  Two launches (flags, masked copy). */
int ipoc_newton_advance_f64(int N, int nx, int nu, int batch, const double* hu, uint8_t* inner_done,
                            uint8_t* outer_done, int64_t* inner, int64_t* iteration, int32_t* advanced,
                            const double* tx, const double* tu, double* x, double* u, double hu_tol,
                            int max_iterations, ipoc_stream_t stream);

/* ---- device-resident loops, one launch for everything after the trial cost (ref :159-202) ----------------
 * For every member with active != 0: the accept rule of ipoc_accept_update_f64; inner += 1 (:174); if the attempt
 * loop ends (success or inner > max_attempts, :180-181) the Newton iteration ends with it: advanced = 1 (the
 * caller's next ipoc_masked_copy_f64 takes the step x <- tx, :184), iteration += 1 (:194), inner = 0, and
 * outer_done = 1, active = 0 if hu < hu_tol or iteration > max_iterations (:199-202; hu = max|ru| of the iterate
 * BEFORE the step).  Members with active == 0 are left untouched (advanced = 0): the `select` semantics of a
 * vmapped lax.while_loop.  Replaces accept_update + attempt_commit + the flag half of newton_advance. */
int ipoc_attempt_finish_f64(int batch, const double* cost, const double* new_cost, const int32_t* traj_feasible,
                            const double* pred, const int32_t* bwd_feasible, const double* hu, int32_t* active,
                            double* rp, double* r_inc, int32_t* success, double* gain_ratio, int64_t* inner,
                            int64_t* iteration, uint8_t* outer_done, int32_t* advanced, double hu_tol,
                            int max_attempts, int max_iterations, ipoc_stream_t stream);
/* dst_x <- src_x ((N+1)*nx per member), dst_u <- src_u (N*nu) for the members with mask != 0. */
int ipoc_masked_copy_f64(int N, int nx, int nu, int batch, const int32_t* mask, const double* src_x,
                         const double* src_u, double* dst_x, double* dst_u, ipoc_stream_t stream);

/* ---- time-sharded (multi-GPU) split-phase variants -----------------------------------------
 * A horizon of P*N steps is cut into P contiguous segments, one per rank (no reference
 * counterpart — the reference is single-device).  Each scan is: local reduce -> exchange of
 * the P segment aggregates (caller: NCCL all-gather) -> local seeded scan.
 * Carry sizes in doubles: ipoc_carry_doubles(kind, nx).
 */
enum { IPOC_CARRY_RICCATI = 0, IPOC_CARRY_AFFINE = 1 };
int ipoc_carry_doubles(int kind, int nx);

/* Phase 1 of K2 on this rank's segment: writes the segment aggregate (A,b,C,eta,J packed) to
 * `carry_out`.  Newton-step inputs as in ipoc_newton_step_f64 (batch = 1). */
int ipoc_newton_bwd_reduce_f64(int N, int nx, int nu,
                               const double* fx, const double* fu, const double* ru,
                               const double* Q, const double* R, const double* M,
                               const double* reg, double* carry_out,
                               void* ws, size_t ws_bytes, ipoc_stream_t stream);
/* Phase 2 of K2: `carries` = the gathered aggregates of all `nranks` segments (rank-major);
 * `ST` (nx,nx) = terminal weight of the WHOLE horizon (Q[0] of rank 0's segment in the
 * reference's convention).  Produces this segment's Kx, d, partial pred / feasible, and the
 * forward-scan aggregate of the segment in `fwd_carry_out`. */
int ipoc_newton_bwd_apply_f64(int N, int nx, int nu, int rank, int nranks,
                              const double* fx, const double* fu, const double* ru,
                              const double* Q, const double* R, const double* M,
                              const double* reg, const double* carries, const double* ST,
                              double* Kx, double* d, double* pred, int32_t* feasible,
                              double* fwd_carry_out,
                              void* ws, size_t ws_bytes, ipoc_stream_t stream);
/* K3 on this rank's segment given the gathered forward aggregates; rank 0 starts from dx = 0.
 * dx has N+1 rows (row N duplicates the next rank's row 0). */
int ipoc_newton_fwd_apply_f64(int N, int nx, int nu, int rank, int nranks,
                              const double* fx, const double* fu,
                              const double* Kx, const double* d, const double* fwd_carries,
                              double* dx, double* du,
                              void* ws, size_t ws_bytes, ipoc_stream_t stream);
/* K1, time-sharded: reduce and apply phases of the reverse/transposed affine scan. */
int ipoc_affine_reduce_f64(int reverse, int transpose, int N, int nx,
                           const double* F, const double* c, double* carry_out,
                           void* ws, size_t ws_bytes, ipoc_stream_t stream);
int ipoc_affine_apply_f64(int reverse, int transpose, int N, int nx, int rank, int nranks,
                          const double* F, const double* c, const double* carries,
                          const double* seed, double* out,
                          void* ws, size_t ws_bytes, ipoc_stream_t stream);

/* ---- host-buffer convenience wrapper (end-to-end timing, simple embedding) -----------------
 * Same as ipoc_newton_step_f64 but every pointer is HOST memory (pinned for best speed);
 * `dws` is a DEVICE scratch buffer of at least ipoc_newton_step_host_scratch_bytes().  Copies
 * H2D, runs, copies D2H, all on `stream`; the caller synchronises the stream. */
size_t ipoc_newton_step_host_scratch_bytes(int N, int nx, int nu, int batch);
int ipoc_newton_step_host_f64(int N, int nx, int nu, int batch,
                              const double* fx, const double* fu, const double* ru,
                              const double* Q, const double* R, const double* M,
                              const double* reg,
                              double* dx, double* du, double* pred, int32_t* feasible,
                              void* dws, size_t dws_bytes, ipoc_stream_t stream);

/* ---- optional built-in plants (SURVEY §8f "next" #4) -------------------------------------------
 * Fused evaluation of the user functions of the reference's two example problems
 * (ref examples/pendulum_runtime.py:19-72, examples/cartpole_runtime.py:18-81; dynamics =
 * euler(ode, Ts), stage cost with log barrier on |u| <= bound), as a fast path BESIDE the host
 * framework's autodiff (which remains the general path for user-defined OCPs):
 *   ipoc_plant_derivatives_f64: the ten `Derivatives` tensors of ref noc/optimal_control_problem.py:13-23
 *       (what ref noc/par_interior_point_newton.py:13-28 gets from vmapped grad/hessian/jacrev) via
 *       second-order forward-mode autodiff in registers, plus lamT = grad final_cost(x_N) (may be NULL).
 *       x (batch,N+1,nx), u (batch,N,nu), bp = device scalar.
 *   ipoc_plant_cost_f64: total_cost (ref examples/cartpole_runtime.py:48-51) and all(constraints<=0)
 *       (ref noc/par_interior_point_newton.py:45-47) per problem, fixed summation order.
 *   ipoc_plant_rollout_f64: serial rollout (ref noc/utils.py:57-63), one thread per problem.
 *   ipoc_plant_rollout_lin_f64: one iteration's inputs of the parallel-in-time rollout: along a
 *       guess x (batch,N+1,nx) of the trajectory, F_k = df/dx and c_k = f(x_k,u_k) - F_k x_k (feed them to
 *       ipoc_affine_scan_f64 with seed x_0 for the next guess), fv_k = f(x_k,u_k) in the serial rollout's arithmetic,
 *       stats (batch,2) = (max_k |x_{k+1} - fv_k|, max |fv|) per problem (a NaN defect stays NaN).  Iterated to a
 *       defect at rounding level, (x_0, fv) IS the serial rollout up to rounding; host loop: ipoc_b200/plants.py.
 */
enum { IPOC_PLANT_PENDULUM = 1, IPOC_PLANT_CARTPOLE = 2 };
int ipoc_plant_dims(int plant, int* nx, int* nu, int* nc);
int ipoc_plant_derivatives_f64(int plant, int N, int batch, double Ts, double bound, const double* bp,
                               const double* x, const double* u,
                               double* cx, double* cu, double* cxx, double* cuu, double* cxu,
                               double* fx, double* fu, double* fxx, double* fuu, double* fxu, double* lamT,
                               ipoc_stream_t stream);
/* Fused A1 + A3 for the built-in plants: ru, Q, R, M of ref noc/par_interior_point_newton.py:31-42 are the
 * first/second derivatives of the Hamiltonian H = stage_cost + lam[k+1]' f(x,u), so the 130-double
 * `Derivatives` record never has to exist:  linearize (before the costate scan) -> fx, fu, cx, cu, lamT;
 * hamiltonian (after it, lam (batch,N+1,nx)) -> ru, Q, R, M. */
/* `fresh` (int32 per problem, may be NULL = all): problems whose flag is 0 are skipped and keep their previous
 * outputs — device-resident loops re-run an attempt on an UNCHANGED iterate after a rejection (ref :177-182) and
 * must not pay for re-evaluating it. */
int ipoc_plant_linearize_f64(int plant, int N, int batch, double Ts, double bound, const double* bp,
                             const double* x, const double* u, double* fx, double* fu, double* cx, double* cu,
                             double* lamT, const int32_t* fresh, ipoc_stream_t stream);
/* ipoc_masked_copy_f64(mask = fresh: x <- tx, u <- tu) and ipoc_plant_linearize_f64 in ONE launch: members with
 * fresh != 0 take their iterate from (tx, tu) — written to x (batch,N+1,nx) and u (batch,N,nu) — and are linearised
 * there; the others are left alone, as in ipoc_plant_linearize_f64.  (Device-resident loops: fresh = `advanced`.) */
int ipoc_plant_take_linearize_f64(int plant, int N, int batch, double Ts, double bound, const double* bp,
                                  const double* tx, const double* tu, double* x, double* u, double* fx, double* fu,
                                  double* cx, double* cu, double* lamT, const int32_t* fresh, ipoc_stream_t stream);
int ipoc_plant_hamiltonian_f64(int plant, int N, int batch, double Ts, double bound, const double* bp,
                               const double* x, const double* u, const double* lam,
                               double* ru, double* Q, double* R, double* M, const int32_t* fresh,
                               ipoc_stream_t stream);
/* `ws` (may be NULL): scratch of ipoc_plant_cost_workspace_bytes(N, batch) bytes, ZERO-FILLED once by the caller (its
 * arrival counters return to zero after every launch).  With it, small batches of long horizons (batch < 32,
 * N > 8192) spread the FP64-bound stage-cost sum over up to 256 CTAs per problem whose last arriver folds the
 * partials in a fixed order; without it the kernel runs as one thread-block cluster of <= 8 CTAs per problem
 * (partials through distributed shared memory).  Both are deterministic; they differ in summation order. */
size_t ipoc_plant_cost_workspace_bytes(int N, int batch);   /* 0: this shape does not use scratch */
int ipoc_plant_cost_f64(int plant, int N, int batch, double Ts, double bound, const double* bp,
                        const double* x, const double* u, double* total_cost, int32_t* feasible,
                        const int32_t* fresh, void* ws, size_t ws_bytes, ipoc_stream_t stream);
int ipoc_plant_rollout_f64(int plant, int N, int batch, double Ts, const double* x0, const double* u,
                           double* x, ipoc_stream_t stream);
int ipoc_plant_rollout_lin_f64(int plant, int N, int batch, double Ts, const double* x, const double* u,
                               double* F, double* c, double* fv, double* stats, ipoc_stream_t stream);
/* ipoc_plant_cost_f64 of the trial point (tx, tu) followed, in the same launch, by ipoc_attempt_finish_f64 with
 * the cost / feasibility just computed (members with active == 0 are skipped altogether).  cost_carry / need_cost
 * (both may be NULL): when a member's step is taken its trial cost is also stored to cost_carry[b] — it IS the cost of
 * the next iterate, bit for bit — and need_cost[b] is cleared for every member, so a loop that evaluates
 * ipoc_plant_cost_f64(x, u, fresh = need_cost) pays that evaluation only after need_cost was set by whoever loaded
 * a new iterate (start of a barrier stage). */
int ipoc_plant_attempt_finish_f64(int plant, int N, int batch, double Ts, double bound, const double* bp,
                                  const double* tx, const double* tu, double* new_cost, int32_t* traj_feasible,
                                  const double* cost, const double* pred, const int32_t* bwd_feasible, const double* hu,
                                  int32_t* active, double* rp, double* r_inc, int32_t* success, double* gain_ratio,
                                  int64_t* inner, int64_t* iteration, uint8_t* outer_done, int32_t* advanced,
                                  double hu_tol, int max_attempts, int max_iterations, double* cost_carry, int32_t* need_cost, void* ws, size_t ws_bytes, ipoc_stream_t stream);

/* Number of kernels the library has launched since load (for launch accounting in bench.py). */
unsigned long long ipoc_launch_count(void);

/* Optional per-launch profiler (bench.py's roofline): after ipoc_profile_begin, a CUDA event is
 * recorded on the launching stream after every kernel launch of this library;
 * ipoc_profile_end synchronises the last event and returns the number of launches seen, their
 * device durations in ms (event-to-event) and a comma-separated list of kernel names.
 * Not thread-safe; never armed unless asked. */
int ipoc_profile_begin(ipoc_stream_t stream);
int ipoc_profile_end(char* names, size_t names_len, float* ms, int max_entries);

#ifdef __cplusplus
}
#endif
#endif /* IPOC_H */
