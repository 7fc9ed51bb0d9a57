"""CPU oracle for the NW-head hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in ``nwhead_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may use it, and there only as the checker or as the reported CPU baseline.

Parity status: PINNED.  The restatements in ``nw_oracle.py`` are checked against the
unmodified reference (imported from /root/reference through ``ref_import.py``) by
``oracle/gen_golden.py``, which also writes the committed fixtures in ``tests/golden/``.
The reference itself ships no tests / golden vectors, so the fixtures are
outputs of the reference run in the build container.
"""
