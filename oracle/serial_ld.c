/* serial_ld.c — extended-precision (x87 `long double`, 64-bit mantissa) SERIAL restatement of the
 * reference's in-tree sequential Newton step.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): arbitrates CUDA-vs-oracle differences at the
 * large end of BASELINE config 4 (N = 1e6, Ts = 1e-6), where two float64 association orders may
 * legitimately differ by more than the 1e-9 tolerance and neither can be called "right".
 *
 * Follows, line by line:
 *   ref noc/seq_interior_point_newton.py:42-75  bwd_pass  (Riccati recursion, gains, dV, convexity)
 *   ref noc/seq_interior_point_newton.py:78-90  fwd_pass  (x_0 = 0, u = k + K x, x+ = fx x + fu u)
 * with the two documented differences of the par path neutralised by the caller
 * (SURVEY.md section 3.4): terminal Hessian VxxN := Q[0] (par: XT = Q[0], ref
 * noc/par_interior_point_newton.py:73) and rp := reg_param * ||cu|| (:116-117).
 *
 * Inputs/outputs are float64 arrays (row-major, the Newton-step layout); all arithmetic in between
 * is long double.  The same body is also instantiated in plain double (`*_f64`): the serial O(N) CPU
 * comparator "B2" of BASELINE.md section 3 that bench.py times next to the tree-scan port.
 * Build (done by __graft_entry__.build() and, if missing, by oracle/serial_ld.py):
 *   gcc -O2 -shared -fPIC -o oracle/_build/libserial_ld.so oracle/serial_ld.c -lm
 */
#include <math.h>
#include <string.h>

#define MAXN 8

#define REAL long double
#define FN(name) name##_ld
#define ABS(x) fabsl(x)
#include "serial_body.inc"
#undef REAL
#undef FN
#undef ABS

#define REAL double
#define FN(name) name##_f64
#define ABS(x) fabs(x)
#include "serial_body.inc"
#undef REAL
#undef FN
#undef ABS
