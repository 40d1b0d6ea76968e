"""CPU oracle for the batched Poisson hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is imported by the product package
(``superresolution_for_pdes_b200``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
there only as the checker or as the timed CPU baseline.
"""
