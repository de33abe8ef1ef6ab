"""CPU oracle for the MSMD speech-to-face hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker.  The product
path (``ubisoft-laforge-msmd_b200``) never imports this package and fails
loudly when its CUDA library is missing.

Every function restates, in plain numpy / torch-CPU, the algorithm of the
reference file:line it cites.  Parity status ("pinning"):

* The reference ships no tests, golden vectors or fixtures of its own
  (SURVEY.md section 4), so the oracle is pinned against *outputs of the
  reference itself*: ``tests/test_oracle_vs_reference.py`` imports the
  unmodified modules from ``/root/reference`` (build container only) and
  ``tests/golden/*.npz`` holds vectors produced by ``oracle/make_golden.py``
  from those modules (the script is committed next to the vectors).
* Third-party arithmetic (HF ``transformers`` Hubert/Wav2Vec2, pinned 4.44.2
  by the reference, 5.5.0 installed) is restated in ``oracle/audio.py`` and
  pinned against the installed ``transformers`` classes.
"""
