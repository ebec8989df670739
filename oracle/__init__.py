"""CPU oracle for the ALIBY per-object extraction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``aliby_b200/`` imports this package;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may use it, and only as the checker or
as the timed CPU arm — never as the thing shipped.

* ``oracle.port``  — faithful NumPy/SciPy restatement of the reference algorithm
  (one full-plane pass per (object, instruction), same cost model Θ(obj·instr·Y·X)).
* ``oracle.fast``  — O(Y·X log) sort-by-label implementation of the same numbers,
  proven equal to ``oracle.port`` on small inputs so that full-size fields are
  tractable.

Parity pinning: ``oracle/make_golden.py`` runs the *real* reference
(``/root/reference/src`` + two import shims) in the build container and commits
its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks both
oracles against them.  The cp_measure-backed features are
not part of this oracle: parity for them is unpinned and they are not built.
"""
