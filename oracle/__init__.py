"""CPU oracle for the matrix-correction hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is imported by the product package
(``hicexplorer_b200``).  It may be imported only by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` -- as the checker or the reported CPU baseline, never as
the thing shipped.
"""
