"""CPU oracle for the leanfe estimation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline -- never as a fallback for the CUDA path.

Parity status: the reference (jorgenhost/leanfe) ships no golden vectors for
this path and its arithmetic lives in Polars 1.36.1, which is absent here, so
parity is pinned by (a) this NumPy restatement, (b) the reference's own
NumPy/SciPy functions executed unmodified under stubbed ``polars``/``duckdb``
imports (``oracle/ref_loader.py``; fixtures in ``tests/golden/``) and (c) an
independent dense dummy-variable OLS cross-check.  See DESIGN.md.
"""
