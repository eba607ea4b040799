"""CPU oracle for the SR-WaveNet hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``sr-wavenet_b200/``) never imports it and has no CPU fallback.

Parity status: **pinned by executing the reference's own source** -- ``/root/reference/ops.py`` and
``model.py`` are imported unmodified on top of a NumPy stand-in for TensorFlow 1.x (``tests/tf_shim``)
and every function of ``srwn_oracle.py`` is compared with them at 1e-11
(``tests/test_reference_shim.py``); the reference's printed known answers (``ops.py:243-254``) and the
fixtures written from that code path (``tests/golden/reference_*.npz``) hold the oracle on machines
without the reference tree.  TensorFlow itself cannot be installed here: the semantics of single
``tf.*`` operations are the stand-in's restatement.  See DESIGN.md section "Oracle".
"""
