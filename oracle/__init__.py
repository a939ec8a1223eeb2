"""ORACLE -- test infrastructure only.

CPU restatement of the reference hot path (Group Matching env, REFIL/QMIX-attention learner).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  Nothing under ``refil_b200/`` imports it.
"""
