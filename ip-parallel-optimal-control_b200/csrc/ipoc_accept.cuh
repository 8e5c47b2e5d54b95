// Scalar control of one accept/reject attempt (ref noc/par_interior_point_newton.py:159-202), shared by the
// stand-alone glue kernels (ipoc_api.cu), K3's tail job (ipoc_impl.cuh) and the plant cost kernel (ipoc_plants.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ipoc {

// gain ratio, success, regularisation update (ref :159-173)
struct AcceptIO {
    const double *cost, *pred;
    const int32_t* bwd_feasible;
    double *rp, *r_inc;
    int32_t* success;
    double* gain;      // may be NULL
};
__device__ __forceinline__ bool accept_rule_core(const AcceptIO& a, int b, double new_cost, int traj_ok) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double nc = traj_ok ? new_cost : inf;                       // :161
    const double rho = (nc - a.cost[b]) / a.pred[b];                  // :164
    const bool ok = (rho > 0.0) && (a.bwd_feasible[b] != 0);          // :166
    double r = a.rp[b], ri = a.r_inc[b];
    if (ok) {
        const double tq = 2.0 * rho - 1.0;
        r = r * fmax(1.0 / 3.0, 1.0 - tq * tq * tq);                  // :168
        ri = 2.0;
    } else {
        r = r * ri;                                                   // :171
        ri = 2.0 * ri;
    }
    r = fmin(fmax(r, 1e-16), 1e16);                                   // :173
    a.rp[b] = r;
    a.r_inc[b] = ri;
    a.success[b] = ok ? 1 : 0;
    if (a.gain != nullptr) a.gain[b] = rho;
    return ok;
}

// Everything that follows the trial cost of an attempt, for loops that live on the device
// (frozen-when-done, the `select` semantics of a vmapped lax.while_loop):
//   accept rule; inner += 1 (:174); if the attempt loop ends (success or inner > max_attempts, :180-181) the
//   Newton iteration ends too: advanced = 1 (the caller's next masked copy takes the step x <- tx, :184, kept
//   whether or not an attempt ever succeeded), iteration += 1 (:194), inner = 0 and the exit test on max|ru| of
//   the iterate BEFORE the step (:199-202); active = !outer_done for the next attempt.
struct FinishIO {
    AcceptIO acc;
    const double* hu;
    int32_t* active;          // in: this attempt ran for the member; out: it takes part in the next one
    long long *inner, *iteration;
    uint8_t* outer_done;
    int32_t* advanced;
    double hu_tol;
    int max_attempts, max_iterations;
    // Optional: when the step is taken, the trial point BECOMES the iterate, so its total cost IS the next
    // iteration's `cost` (:142) — same kernel, same data, same summation order, hence the same bits.  cost_carry
    // (may alias acc.cost) receives it; need_cost (per member) is cleared by every attempt and set by whoever loads
    // a new iterate, so the loop's cost evaluation runs once per stage instead of once per iteration.
    double* cost_carry;
    int32_t* need_cost;
};
__device__ __forceinline__ void attempt_finish_rule(const FinishIO& f, int b, double new_cost, int traj_ok) {
    if (f.need_cost != nullptr) f.need_cost[b] = 0;
    if (!f.active[b]) {
        f.advanced[b] = 0;
        return;
    }
    const bool ok = accept_rule_core(f.acc, b, new_cost, traj_ok);
    const long long n = f.inner[b] + 1;
    if (ok || n > f.max_attempts) {
        const long long it = f.iteration[b] + 1;
        f.iteration[b] = it;
        f.inner[b] = 0;
        f.advanced[b] = 1;
        if (f.cost_carry != nullptr) f.cost_carry[b] = new_cost;
        if (f.hu[b] < f.hu_tol || it > f.max_iterations) {
            f.outer_done[b] = 1;
            f.active[b] = 0;
        }
    } else {
        f.inner[b] = n;
        f.advanced[b] = 0;
    }
}

}  // namespace ipoc
