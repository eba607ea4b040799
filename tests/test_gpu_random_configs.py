"""-m gpu: differential check of the 16-bit tensor-core kernels against the fp32-grade path on random configurations
(dilation schedules with non-powers of two, short and long stacks, several utterance lengths and mixture counts): the
work partition, the warm-up dependency cone and the barrier phase bookkeeping of the fused kernel, and the layer /
queue bookkeeping of the generation kernel, all depend on these."""
import numpy as np
import pytest
import torch

from sr_wavenet_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def srwn(lib):
    import sr_wavenet_b200
    assert torch.cuda.is_available()
    return sr_wavenet_b200


def _config(seed):
    rng = np.random.default_rng(seed)
    L = int(rng.integers(3, 13))
    pool = [1, 2, 3, 4, 5, 8, 16, 27, 64, 100, 128, 256, 300, 512]
    dil = [int(rng.choice(pool)) for _ in range(L)]
    M = int(rng.choice([3, 5]))
    B = int(rng.integers(1, 6))
    T = 128 * int(rng.integers(3, 40))
    return dil, M, B, T


@pytest.mark.parametrize("seed", range(8))
def test_fused_teacher_and_student_vs_fp32(srwn, seed):
    dil, M, B, T = _config(seed)
    t = srwn.WaveNetAutoEncoder(input_size=T, condition_size=0, num_mixtures=M, dilations=dil, skip_channels=128,
                                latent_channels=32, pool_stride=128)
    t.set_weights(synth.make_teacher_weights(dil, num_mixtures=M, seed=100 + seed))
    if "fp16" not in t.available_precisions():
        pytest.skip("fused path not available for this configuration")
    x = synth.synthetic_audio(B, T, seed=seed)
    enc = synth.synthetic_encoding(B, T // 128, seed=seed)
    ref = t.get_logits(x, enc, precision="fp32")
    got = t.get_logits(x, enc, precision="fp16")
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() <= 1e-2 * max(1.0, np.abs(ref).max()), (dil, M, B, T)
    n32, n16 = t.nll(x, enc, precision="fp32"), t.nll(x, enc, precision="fp16")
    assert abs(n16 - n32) <= 2e-3 * abs(n32)
    s = srwn.ParallelWaveNet(input_size=T, condition_size=0, dilations=dil, teacher=None, num_flows=2, skip_channels=128,
                             latent_channels=32, pool_stride=128)
    s.set_weights(synth.make_student_weights(dil, num_flows=2, seed=200 + seed))
    z = synth.logistic_noise(B, T, seed=seed)
    o32 = s.generate(None, z, enc, precision="fp32")
    o16 = s.generate(None, z, enc, precision="fp16")
    assert np.abs(o16 - o32).max() <= 2e-2, (dil, B, T)


@pytest.mark.parametrize("seed", range(4))
def test_generation_fp16_vs_fp32_kernel(srwn, seed):
    """Both generation kernels consume the same injected noise; until a near-tie flips a mixture choice they produce the
    same audio (fp16 operand bound)."""
    dil, M, B, _ = _config(50 + seed)
    T = 256
    t = srwn.WaveNetAutoEncoder(input_size=T, condition_size=0, num_mixtures=M, dilations=dil, skip_channels=128,
                                latent_channels=32, pool_stride=128)
    t.set_weights(synth.make_teacher_weights(dil, num_mixtures=M, seed=300 + seed))
    enc = synth.synthetic_encoding(B + 7, T // 128, seed=seed)              # more than one CTA of 8 utterances
    u1, u2 = synth.sampler_uniforms(B + 7, T, M, seed=seed)
    x32, l32 = t.generate(enc, u1=u1, u2=u2, return_logits=True, precision="fp32")
    try:
        x16, l16 = t.generate(enc, u1=u1, u2=u2, return_logits=True, precision="fp16")
    except RuntimeError:
        pytest.skip("fp16 generation kernel not available for this configuration")
    k32 = np.argmax(l32[..., :M] - np.log(-np.log(u1)), axis=-1)
    k16 = np.argmax(l16[..., :M] - np.log(-np.log(u1)), axis=-1)
    for b in range(B + 7):
        flips = np.nonzero(k32[b] != k16[b])[0]
        n = int(flips[0]) if flips.size else T                              # trajectories agree up to the first flipped choice
        assert n >= 8, (dil, b, n)
        assert np.abs(l16[b, :n] - l32[b, :n]).max() <= 2e-2
        assert np.abs(x16[b, :n] - x32[b, :n]).max() <= 2e-2
