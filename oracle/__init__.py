"""CPU oracle for the DG-SQP hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy restatement of the reference solver
(zhu-edward/DGSQP, ``DGSQP/solvers/DGSQP.py`` v1 ``solve()`` and the game
definitions in ``scripts/DGSQP_*monte_carlo*.py``).  It exists so the CUDA path
can be checked against an independent implementation; it is **never** on the
product path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

PARITY UNPINNED.  The reference ships no tests, golden vectors or stored
results, and its two native dependencies (CasADi, OSQP) are not importable in
the authoring container, so the reference itself could not be run.  The oracle
is pinned instead by (see ``tests/test_oracle_*.py``):

* closed-form known answers of the track geometry (chicane key points),
* ``torch.autograd`` (float64) derivatives of the *whole* rollout, which is the
  same check as the reference's unused ``f_Duu_L`` / ``f_Du_L`` functions
  (``DGSQP.py:937-941``) against its dynamic-programming Hessian ``f_Q``,
* NumPy ``eigh`` and SciPy ``lsqr`` -- the very routines the reference calls,
* an independent OSQP-style ADMM + polish solver for the QP sub-problem.

What cannot be reproduced: OSQP iterates when polish fails or max-iter hits
(v1 never checks ``success``), and OSQP's wall-clock adaptive rho.  The oracle
solves every QP exactly (Goldfarb-Idnani dual active set).
"""
