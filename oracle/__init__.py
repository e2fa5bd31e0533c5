"""CPU oracle for the OS2D dense correlation-and-alignment head.

TEST INFRASTRUCTURE ONLY.  Nothing under ``os2d_b200/`` may import this package.
It is used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` as the checker / reported baseline, never as the
product path.

Parity pinning: the reference repository ships no tests, golden vectors or known-answer
fixtures for this path, so the oracle is pinned against the reference
itself executed in the build container: ``tests/golden/make_golden.py`` imports
``/root/reference/os2d`` (read-only), runs it on seeded inputs and commits the small
input/output fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this
restatement against them on any machine, ``tests/test_oracle_vs_reference.py`` checks it
against the live reference when ``/root/reference`` exists.
"""
