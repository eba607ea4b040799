"""-m gpu: on-device noise (csrc/philox.cuh, csrc/random.cu) and the student's fused sampling path.

The reference draws the student's input on the host (student.py:104 / :172, np.random.logistic(0, 1)) and the sampler's
uniforms inside the graph (ops.py:187, 196, tf.random_uniform(1e-5, 1 - 1e-5)); neither stream is reproducible, so what
is checked here is the LAW of the draws (Kolmogorov-Smirnov), their independence across streams / seeds, and that the
fused flow kernel, which evaluates the draws in place instead of reading a noise tensor, gives bit-identical results to
the same kernel fed with those draws as a tensor."""
import ctypes

import numpy as np
import pytest
import torch
from scipy import stats

from sr_wavenet_b200 import synth, _lib

pytestmark = pytest.mark.gpu
DIL = synth.DEFAULT_DILATIONS


@pytest.fixture(scope="module")
def srwn(lib):
    import sr_wavenet_b200
    assert torch.cuda.is_available()
    return sr_wavenet_b200


def _logistic(lib, n, seed, stream):
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    _lib.check(lib.srwn_random_logistic(out.data_ptr(), n, seed, stream, 0))
    torch.cuda.synchronize()
    return out.cpu().numpy().astype(np.float64)


def test_logistic_draws_follow_the_logistic_law(lib):
    z = _logistic(lib, 1 << 20, 12345, 1)
    ks = stats.kstest(z, "logistic")
    print("KS logistic: D = %.5f, p = %.3f, mean %.4f, var %.4f (pi^2/3 = %.4f)" % (ks.statistic, ks.pvalue, z.mean(), z.var(), np.pi ** 2 / 3))
    assert ks.statistic < 2.5e-3 and ks.pvalue > 1e-3
    assert abs(z.mean()) < 1e-2 and abs(z.var() - np.pi ** 2 / 3) < 3e-2
    assert np.isfinite(z).all()
    # serial correlation, and different streams / seeds are unrelated
    assert abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 5e-3
    z2, z3 = _logistic(lib, 1 << 20, 12345, 2), _logistic(lib, 1 << 20, 12346, 1)
    assert abs(np.corrcoef(z, z2)[0, 1]) < 5e-3 and abs(np.corrcoef(z, z3)[0, 1]) < 5e-3
    np.testing.assert_array_equal(z, _logistic(lib, 1 << 20, 12345, 1))            # a pure function of (seed, stream, index)
    np.testing.assert_array_equal(z[:1001], _logistic(lib, 1001, 12345, 1))        # unaligned tail


def test_uniform_draws_cover_the_samplers_interval(lib):
    n, lo, hi = 1 << 20, 1e-5, 1.0 - 1e-5
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    _lib.check(lib.srwn_random_uniform(out.data_ptr(), n, 7, 3, lo, hi, 0))
    u = out.cpu().numpy().astype(np.float64)
    assert u.min() >= lo and u.max() <= hi                                          # ops.py:187: minval / maxval
    ks = stats.kstest((u - lo) / (hi - lo), "uniform")
    assert ks.statistic < 2.5e-3 and ks.pvalue > 1e-3


@pytest.mark.parametrize("prec", ["fp16", "fp32"])
def test_student_sampling_in_the_flow_kernel_equals_feeding_the_draws(srwn, prec):
    """generate(sess, None, encoding): noise drawn inside the flow kernel (fp16 path) == the same call fed with z."""
    B, T = 3, 8192
    s = srwn.ParallelWaveNet(input_size=T, condition_size=0, dilations=DIL, teacher=None, num_flows=4, skip_channels=128,
                             latent_channels=32, pool_stride=128)
    s.set_weights(synth.make_student_weights(DIL, 4))
    s.seed = 99
    enc = torch.from_numpy(synth.synthetic_encoding(B, T // 128)).cuda()
    r = s.forward_all(None, enc, precision=prec)
    z = r["z"]
    ks = stats.kstest(z.cpu().numpy().ravel().astype(np.float64), "logistic")
    assert ks.pvalue > 1e-3
    r2 = s.forward_all(z, enc, precision=prec)
    for k in ("out", "s_tot", "mu_tot", "x_last"):
        assert torch.equal(r[k], r2[k]), k
    # next call: a fresh stream, different noise; same seed + same call index: the same noise
    r3 = s.forward_all(None, enc, precision=prec)
    assert not torch.equal(r3["z"], z)
    s._noise_calls = 0
    assert torch.equal(s.forward_all(None, enc, precision=prec)["z"], z)
    out = s.generate(None, None, enc.cpu().numpy(), precision=prec)                 # host boundary: only the encoding goes up
    assert out.shape == (B, T, 1) and np.isfinite(out).all() and np.abs(out).max() <= 1.0


def test_teacher_sampling_draws_its_own_uniforms(srwn):
    t = srwn.WaveNetAutoEncoder(input_size=1024, condition_size=0, num_mixtures=5, dilations=DIL, skip_channels=128,
                                latent_channels=32, pool_stride=128)
    t.set_weights(synth.make_teacher_weights(DIL))
    x, enc = synth.synthetic_audio(2, 1024), synth.synthetic_encoding(2, 8)
    a = t.reconstruct_with_encoding(x, enc)
    b = t.reconstruct_with_encoding(x, enc)
    assert a.shape == (2, 1024) and np.abs(a).max() <= 1 and not np.array_equal(a, b)
    g1 = t.generate(enc[:, :2], precision="fp32")
    g2 = t.generate(enc[:, :2], precision="fp32")
    assert g1.shape == (2, 256) and not np.array_equal(g1, g2)
