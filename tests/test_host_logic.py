"""Host-side logic that needs no GPU: variable naming, synthetic inputs, sharding."""
import numpy as np

import sr_wavenet_b200 as srwn
from sr_wavenet_b200 import synth, shard


def test_teacher_variable_names_and_shapes():
    w = synth.make_teacher_weights()
    p = synth.TEACHER_PREFIX
    assert w[p + "causal_conv_Kernel"].shape == (2, 1, 32)
    assert w[p + "conv1d/kernel"].shape == (1, 32, 32)             # layer-0 conditioning conv
    assert w[p + "conv1d_3/kernel"].shape == (1, 32, 32)           # layer-1 conditioning conv
    assert w[p + "conv1d_1/kernel"].shape == (1, 32, 32)           # layer-0 residual 1x1
    assert w[p + "conv1d_2/kernel"].shape == (1, 32, 128)          # layer-0 skip 1x1
    assert w[p + "dilated_conv_29_filter/dilated_conv_29_Kernel"].shape == (2, 32, 32)
    assert w[p + "dilated_conv_29_gate/dilated_conv_29_Kernel"].shape == (2, 32, 32)   # dead
    assert w[p + "conv1d_90/kernel"].shape == (1, 128, 128)
    assert w[p + "conv1d_91/kernel"].shape == (1, 128, 20)
    live = sum(v.size for k, v in w.items() if "_gate/" not in k)
    assert live == 271668


def test_student_variable_names():
    w = synth.make_student_weights(num_flows=4)
    assert w["ParallelWaveNet/Flow3/Flow3/conv1d_90/kernel"].shape == (1, 32, 2)
    assert "ParallelWaveNet/Flow0/Flow0/conv1d_2/kernel" in w       # dead skip conv exists
    live = sum(v.size for k, v in w.items()
               if k.startswith("ParallelWaveNet/Flow0/") and "_gate/" not in k
               and not any(k.endswith("conv1d_%d/%s" % (3 * i + 2, s)) for i in range(30) for s in ("kernel", "bias")))
    assert live == 125922


def test_weights_are_deterministic_and_glorot_bounded():
    a, b = synth.make_teacher_weights(seed=1), synth.make_teacher_weights(seed=1)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    k = synth.TEACHER_PREFIX + "conv1d_2/kernel"
    assert np.abs(a[k]).max() <= np.sqrt(6.0 / (32 + 128))


def test_synthetic_inputs():
    x = synth.synthetic_audio(3, 4096)
    assert x.shape == (3, 4096) and x.min() == -1.0 and x.max() == 1.0
    u1, u2 = synth.sampler_uniforms(2, 16)
    assert u1.min() >= 1e-5 and u1.max() <= 1 - 1e-5 and u2.shape == (2, 16)


def test_shard_batch_covers_everything():
    for B in (1, 7, 32, 256):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_batch(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a1 >= a0
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
