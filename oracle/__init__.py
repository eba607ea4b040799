"""CPU oracle for the SR-WaveNet hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``sr-wavenet_b200/``) never imports it and has no CPU fallback.

Parity status: the dilated-conv restatement is pinned by the reference's own
known-answer prints (``ops.py:243-254``).  Everything else (gated block,
decoder, flows, MoL loss / sampler) is **parity unpinned**: the reference is
TensorFlow 1.x, which cannot be installed or run here, and it holds no golden
vectors for those functions.  See DESIGN.md section "Oracle".
"""
