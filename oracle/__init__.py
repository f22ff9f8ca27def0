"""CPU oracle for the MFT / SFT / B3-MFN fusion-model hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as
the checker / the CPU arm.  The product path (``multimodal-transformer_b200``)
never imports this package and fails loudly when its CUDA library is missing.

Parity pinning: the reference (frankaging/Multimodal-Transformer) ships no
tests and no golden vectors for the model path, so this restatement is pinned
against outputs of the *imported reference classes run in the build container*
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``) and against the one
known-answer the reference does publish (``eval_ccc`` on ``PredSave/*.csv``
== ``PerfSave/*.csv``; see ``tests/golden/ccc_kat.json``).
"""
