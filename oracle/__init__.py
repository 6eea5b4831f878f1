"""ORACLE -- test infrastructure only.

CPU restatement of the reference's hot path (MAICoS v0.11.2 under ``/root/reference``).
Nothing under ``maicos_b200/`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` do, and there only as the checker / the reported CPU baseline.

Parity status: PINNED.  ``oracle/_ref`` holds the reference's own Cython kernel compiled
from ``/root/reference/src/maicos/lib/_cmath.pyx`` (recipe: ``oracle/Makefile``); the C port
``oracle/sfactor_port.c`` and the NumPy restatement ``oracle/restatement.py`` are checked
against it and against the reference's golden vectors in ``tests/test_oracle.py``.
"""
