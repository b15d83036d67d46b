"""ORACLE - TEST INFRASTRUCTURE ONLY.  Nothing under ``elimrec_b200/`` may import this package.

CPU restatement (torch-CPU / numpy / plain C) of the reference algorithm for the hot path
(SURVEY.md section 8a), each function citing the reference ``file:line`` it follows.  Allowed
importers: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs - there only as the checker or as the timed CPU baseline ("port"),
never as the thing shipped.

Parity pinning: the reference has no tests or golden vectors of its own, so the
oracle is pinned against outputs of the reference ITSELF, executed in the build container by
``tests/golden/make_golden.py`` (unmodified reference + import shims) and committed as
``tests/golden/{generic,kwai}.npz``; ``tests/test_oracle_golden.py`` checks every oracle function
against them.  The C++ metric/top-K code of the reference additionally compiles here straight
from ``/root/reference`` (``oracle/Makefile`` -> ``oracle/_ref/libref_eval.so``) and is used as a
second checker when present.
"""
