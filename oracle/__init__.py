"""CPU oracle for the par IP-Newton hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under `oracle/` is product code.  Only `tests/`, `__graft_entry__.smoke()`
and the CPU legs of `bench.py` (`cpu_baseline`, `--impl reference`) may import it,
and there only as the checker / the timed CPU baseline.  The product package
(`ip-parallel-optimal-control_b200/ipoc_b200`) never imports it and has no CPU
fallback.

PARITY STATUS: "parity unpinned" in the strict sense — the reference
(casiacob/ip-parallel-optimal-control) ships no tests, golden vectors or
known-answer fixtures, and its Riccati/forward scans live in the un-vendored,
un-pinned dependency `paroc` (github.com/casiacob/parallel-optimal-control,
HEAD; not on this machine).  What pins this oracle instead:
  * the reference's OWN source files (`/root/reference/noc/*.py`) executed in
    this container on a small jax->torch shim (`oracle/jaxshim`), producing the
    fixtures in `tests/golden/*.npz` (generator: tests/golden/gen_golden.py);
  * in particular the reference's in-tree sequential Newton step
    (noc/seq_interior_point_newton.py:42-90), an independent formulation that
    needs no `paroc`, against which the restated `paroc` scans are checked.
"""
