"""CPU oracle for the nndepth stereo-correlation hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain-numpy restatement of the reference's algorithm (anhtu293/nndepth,
``/root/reference`` in the build container), one function per row of SURVEY.md section 8(a), each
citing the reference file:line it follows.  It exists to *check* the CUDA product path:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
  reference`` legs may import it;
* nothing under ``nndepth_b200/`` imports it -- the product path has no CPU fallback and raises when
  the CUDA library is missing.

Parity pinning: the reference ships no golden vectors or numeric tests for this path
(``tests/models/test_stereo.py:7-36`` only asserts ``isinstance(outputs, list)``), so the oracle is
pinned against outputs of the *unmodified reference itself*, executed in the build container by
``tests/golden/make_goldens.py`` and committed as ``tests/golden/*.npz``.  ``tests/test_oracle_*.py``
re-check the oracle against those fixtures on every CPU run.

All arithmetic is IEEE fp32 (numpy ``float32``), in the reference's operation order wherever that
order decides an integer (window indices) -- see ``corr1d.sampler_indices``.
"""
