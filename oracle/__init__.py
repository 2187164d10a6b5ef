"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the FastGRNN recurrence path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker / the timed CPU
baseline.  The product path (``kws_b200``) never imports this package and
fails loudly when its CUDA library is missing.

Parity status: **parity unpinned by the reference's own tests** -- the
reference ships no tests, golden vectors or fixtures for this path (SURVEY.md
section 8c).  The oracle is therefore pinned against *outputs of the reference
itself run in the build container*: ``oracle/ref_shim.py`` imports the
unmodified ``/root/reference/rnn.py`` (with two test-side shims for the
reference's D1/D2 call-site defects), ``tests/test_oracle_vs_reference.py``
asserts the restatement is bit-identical to it, and ``oracle/make_golden.py``
mints the committed fixtures under ``tests/golden/`` from it.
"""
