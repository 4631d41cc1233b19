"""CPU oracle for the ConvVAE hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.

PARITY UNPINNED: the reference (JeremyCCHsu/vae-npvc) ships no tests, golden vectors or
fixtures, and its arithmetic lives in TensorFlow 1.2.1 (``requirements.txt:1``, unpinned,
not vendored, not installable here).  The oracle is therefore a restatement of
``model/vae.py`` + ``util/layers.py`` + ``trainer/vae.py`` + TF-1.x op semantics, pinned
only by two independent formulations agreeing with each other (``convvae_ref`` = torch
library convolutions + autograd, ``convvae_loops`` = explicit numpy tap loops) and by
finite-difference gradient checks.
"""
