"""CPU oracle for the pyremap weight-application path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and there only as the checker or as the timed CPU
baseline -- never as the thing shipped.  ``pyremap_b200`` never imports it.

Parity status: PINNED.  The restatement in ``oracle/remap_oracle.py`` is checked
bit-for-bit (values) and mask-for-mask against outputs of the reference module
``/root/reference/pyremap/remapper/remap_numpy.py`` itself, executed in the
authoring container by ``tests/golden/make_golden.py`` (fixtures committed under
``tests/golden/``), and live against that module whenever ``/root/reference``
is present (``tests/test_oracle_vs_reference.py``).
"""
