"""TEST INFRASTRUCTURE ONLY - CPU restatements of the reference's render + log-mel path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it, and only as the checker (or the CPU arm being timed), never
as the thing shipped.  The product (``adt_str_b200``) does not import it and
fails loudly when its CUDA library is missing.

Pinning: the reference has no tests or golden vectors for this path
(SURVEY §4, §8c), so parity is pinned on outputs of the *unmodified reference
code run in the build container* (``oracle/ref_harness.py``), frozen as fixtures
under ``tests/golden/`` by ``oracle/make_golden.py``.
"""
