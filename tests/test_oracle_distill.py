"""CPU: the differentiable (torch, float64) restatement of the distillation graph used as the gradient
oracle is pinned against the NumPy oracle's forward and against central finite differences."""
import numpy as np
import pytest

from oracle import srwn_oracle as orc
from oracle import distill_torch as dt
from sr_wavenet_b200 import synth

DIL, F, P, C, M = [1, 2, 4], 2, 128, 8, 3


def _case(B=2, T=768, seed=5):
    rng = np.random.default_rng(seed)
    sw = {k: v.astype(np.float64) for k, v in synth.make_student_weights(DIL, F, latent_channels=C, seed=11).items()}
    z = rng.logistic(0, 1, size=(B, T))
    truth = synth.synthetic_audio(B, T).astype(np.float64)
    enc = rng.normal(0, 1, size=(B, T // P, C))
    tl = rng.normal(0, 1, size=(B, T, 4 * M)) * 0.5
    return sw, z, truth, enc, tl


def test_torch_forward_matches_numpy_oracle():
    import torch
    sw, z, truth, enc, tl = _case()
    net = orc.student_network(sw, z, enc, DIL, P, F)
    W = {k: torch.tensor(v) for k, v in sw.items()}
    out, s_tot, mu_tot = dt.student_forward(W, torch.tensor(z), torch.tensor(enc), DIL, P, F)
    np.testing.assert_allclose(out.numpy(), net["out"][:, :, 0], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(s_tot.numpy(), net["s_tot"][:, :, 0], rtol=1e-10)
    np.testing.assert_allclose(mu_tot.numpy(), net["mu_tot"][:, :, 0], rtol=1e-10, atol=1e-12)
    # loss pieces: mixture NLL, entropy, power loss against the NumPy formulas (model.py:356-379)
    loss, power, ent, _ = dt.loss_and_grads(sw, z, truth, enc, tl, DIL, P, F, alpha=0.25, beta=1.0, gamma=1.0)
    ent_np = np.sum(np.log(net["s_tot"]) + 2.0)
    pow_np = np.sum((orc.stft_power(truth) - orc.stft_power(net["out"][:, :, 0])) ** 2)
    ce_np = orc.discretized_mix_logistic_loss(np.clip(net["out"], -1, 1), tl, True)
    np.testing.assert_allclose(ent, ent_np, rtol=1e-10)
    np.testing.assert_allclose(power, pow_np, rtol=1e-9)
    np.testing.assert_allclose(loss, (ce_np - 0.25 * ent_np + pow_np) / z.shape[0], rtol=1e-9)


def test_gradients_match_finite_differences():
    sw, z, truth, enc, tl = _case(B=1, T=640)
    args = (z, truth, enc, tl, DIL, P, F)
    _, _, _, g = dt.loss_and_grads(sw, *args, alpha=0.25, beta=1.0, gamma=0.5)
    rng = np.random.default_rng(0)
    names = ["ParallelWaveNet/Flow0/Flow0/causal_conv_Kernel", "ParallelWaveNet/Flow0/Flow0/conv1d/kernel",
             "ParallelWaveNet/Flow1/Flow1/dilated_conv_1_filter/dilated_conv_1_Kernel",
             "ParallelWaveNet/Flow0/Flow0/conv1d_4/kernel", "ParallelWaveNet/Flow1/Flow1/conv1d_9/bias",
             "ParallelWaveNet/Flow0/Flow0/dilated_conv_2_filter/dilated_conv_2_Bias"]
    for name in names:
        idx = tuple(rng.integers(0, s) for s in sw[name].shape)
        eps = 1e-6
        vals = []
        for sgn in (+1, -1):
            w2 = dict(sw)
            w2[name] = sw[name].copy()
            w2[name][idx] += sgn * eps
            vals.append(dt.loss_and_grads(w2, *args, alpha=0.25, beta=1.0, gamma=0.5)[0])
        fd = (vals[0] - vals[1]) / (2 * eps)
        assert abs(fd - g[name][idx]) <= 1e-5 * max(1.0, abs(fd)), (name, idx, fd, g[name][idx])
    # dead variables (gate conv ops.py:31-33, student skip conv model.py:438-454) get no gradient
    assert not np.any(g["ParallelWaveNet/Flow0/Flow0/dilated_conv_0_gate/dilated_conv_0_Kernel"])
    assert not np.any(g["ParallelWaveNet/Flow0/Flow0/conv1d_2/kernel"])


def test_adam_reference_first_step_is_sign_step():
    w, g = [np.array([1.0, -2.0])], [np.array([0.3, -0.4])]     # |g| = 0.5 < clip: unscaled
    (w1, m1, v1), = dt.adam_reference(w, g, [np.zeros(2)], [np.zeros(2)], step=1, lr=1e-3)
    np.testing.assert_allclose(w1, w[0] - 1e-3 * np.sign(g[0]), rtol=1e-6)
    (w2, _, _), = dt.adam_reference(w, [np.array([30.0, -40.0])], [np.zeros(2)], [np.zeros(2)], step=1, lr=1e-3)
    np.testing.assert_allclose(w2, w1, rtol=1e-6)               # clipped to norm 1: same direction, same first step


def test_stft_power_known_answer():
    """tf.contrib.signal.stft(x, 512, 256) semantics (model.py:360-367) on a closed-form case: a cosine on bin k0 under
    the periodic Hann window has |X[k0]| = N/4 and |X[k0 +- 1]| = N/8 in every frame, nothing elsewhere; 1 + (T-N)//step
    frames, the tail that does not fill a frame is dropped."""
    N, step, k0, T = 512, 256, 37, 512 + 256 * 5 + 100
    n = np.arange(T)
    x = np.cos(2 * np.pi * k0 * n / N)[None, :]
    p = orc.stft_power(x, N, step)
    assert p.shape == (1, N // 2 + 1)
    expect = np.zeros(N // 2 + 1)
    expect[k0] = (N / 4) ** 2
    expect[k0 - 1] = expect[k0 + 1] = (N / 8) ** 2
    np.testing.assert_allclose(p[0], expect, atol=1e-6)
    x2 = x.copy()
    x2[:, 512 + 256 * 5:] = 7.0               # past the last full frame: ignored (pad_end=False)
    np.testing.assert_allclose(orc.stft_power(x2, N, step), p, atol=0)
