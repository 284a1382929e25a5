"""CPU oracle for the Graph WaveNet hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker (or the CPU arm being timed beside
the GPU number).  ``multimodal_outage_b200`` never imports it.

Parity status: PINNED.  The restatement in :mod:`oracle.gwnet_oracle` is checked
by ``tests/test_oracle_golden.py`` against golden vectors produced by executing the
reference's own ``nconv/linear/gcn/gwnet`` classes (``models/graph_wavenet.py:60-256``)
and ``asym_adj`` (``utils.py:152-158``); the generating script is
``tests/golden/make_golden.py``.  The reference ships no tests or golden vectors of
its own.
"""
