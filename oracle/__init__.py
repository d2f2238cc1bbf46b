"""CPU oracle for the TSCoDe conformer-ensemble hot path — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (tscode_b200/) never does.

  oracle.oracle_c  — ctypes front-end of oracle.c (fast, OpenMP)
  oracle.oracle_np — numpy restatement (LAPACK SVD like the reference; small cases)

Parity pinning: checked against outputs of the live reference frozen under tests/golden/ by
oracle/gen_golden.py (the reference has no golden vectors of its own for this path).
"""
