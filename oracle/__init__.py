"""CPU oracle for the V2CE hot path -- TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement of the reference's algorithms
(``/root/reference``: ``scripts/LDATI.py``, ``v2ce.py``, ``scripts/v2ce_3d.py`` ...)
used as the *checker* for the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  The product package ``v2ce_toolbox_b200`` never
does: it fails loudly when its CUDA library is missing.

Parity status: pinned.  ``tests/golden/make_golden.py`` (run in the build
container, where ``/root/reference`` is mounted) executes the unmodified
reference on CPU and stores its outputs as fixtures under ``tests/golden/``;
``tests/test_oracle_vs_golden.py`` checks every oracle function against those
fixtures, and against the one known-answer vector the reference repo carries
(``train/scripts/stage2/vis_stage2.ipynb`` cells 1-2).
"""
