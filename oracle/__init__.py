"""TEST INFRASTRUCTURE ONLY: CPU oracle for the pyBMC inference path.

Importable from ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` -- never from ``pybmc_b200``.
"""
