"""Deterministic synthetic weights and inputs for the hot path (host logic, NumPy only).

Weights are keyed by the TF1 variable names the reference graph creates
(``model.py:173-196`` for the teacher decoder, ``model.py:415-454``
for the student flows) and keep TF's ``[K, Cin, Cout]`` kernel layout, so a real
checkpoint converted offline to a ``name -> ndarray`` dict drops in unchanged.

Initialisation follows the reference: Glorot/Xavier-uniform kernels
(``ops.py:15``; ``tf.layers.conv1d`` default), limit = sqrt(6 / (K*Cin + K*Cout)).
Biases are N(0, 0.01) instead of TF's zeros so that every bias path is exercised by
parity tests.
"""
import numpy as np

DEFAULT_DILATIONS = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 3   # teacher.py:55-57

TEACHER_PREFIX = 'WaveNetAutoEncoder/Decoder/'


def student_prefix(flow):
    """Variables of flow f live under Flow{f}/Flow{f}/ (model.py:469 + model.py:416)."""
    return 'ParallelWaveNet/Flow%d/Flow%d/' % (flow, flow)


def cond_conv_name(i):
    """i-th layer's 1x1 conditioning conv (model.py:180): tf.layers default naming."""
    return 'conv1d' if i == 0 else 'conv1d_%d' % (3 * i)


def _glorot(rng, shape, dtype):
    K, cin, cout = shape
    limit = np.sqrt(6.0 / (K * cin + K * cout))
    return rng.uniform(-limit, limit, size=shape).astype(dtype)


def _bias(rng, shape, dtype):
    return rng.normal(0.0, 0.01, size=shape).astype(dtype)


def _stack_weights(rng, prefix, n_layers, K, R, S, C, head_shapes, dtype, dead_vars):
    w = {}
    w[prefix + 'causal_conv_Kernel'] = _glorot(rng, (K, 1, R), dtype)
    w[prefix + 'causal_conv_Bias'] = _bias(rng, (1, 1, R), dtype)
    for i in range(n_layers):
        name = 'dilated_conv_%d' % i
        c = prefix + cond_conv_name(i)
        w[c + '/kernel'] = _glorot(rng, (1, C, R), dtype)
        w[c + '/bias'] = _bias(rng, (R,), dtype)
        w['%s%s_filter/%s_Kernel' % (prefix, name, name)] = _glorot(rng, (K, R, R), dtype)
        w['%s%s_filter/%s_Bias' % (prefix, name, name)] = _bias(rng, (1, 1, R), dtype)
        if dead_vars:   # created by the reference, no forward contribution (ops.py:31-33)
            w['%s%s_gate/%s_Kernel' % (prefix, name, name)] = _glorot(rng, (K, R, R), dtype)
            w['%s%s_gate/%s_Bias' % (prefix, name, name)] = _bias(rng, (1, 1, R), dtype)
        w['%sconv1d_%d/kernel' % (prefix, 3 * i + 1)] = _glorot(rng, (1, R, R), dtype)
        w['%sconv1d_%d/bias' % (prefix, 3 * i + 1)] = _bias(rng, (R,), dtype)
        w['%sconv1d_%d/kernel' % (prefix, 3 * i + 2)] = _glorot(rng, (1, R, S), dtype)
        w['%sconv1d_%d/bias' % (prefix, 3 * i + 2)] = _bias(rng, (S,), dtype)
    for j, (cin, cout) in enumerate(head_shapes):
        w['%sconv1d_%d/kernel' % (prefix, 3 * n_layers + j)] = _glorot(rng, (1, cin, cout), dtype)
        w['%sconv1d_%d/bias' % (prefix, 3 * n_layers + j)] = _bias(rng, (cout,), dtype)
    return w


def make_teacher_weights(dilations=DEFAULT_DILATIONS, filter_width=2, dilation_channels=32,
                         skip_channels=128, latent_channels=32, num_mixtures=5, seed=42,
                         dtype=np.float32, dead_vars=True):
    """Teacher decoder variables (model.py:158-196)."""
    rng = np.random.default_rng(seed)
    return _stack_weights(rng, TEACHER_PREFIX, len(dilations), filter_width, dilation_channels,
                          skip_channels, latent_channels,
                          [(skip_channels, skip_channels), (skip_channels, 4 * num_mixtures)],
                          dtype, dead_vars)


def make_student_weights(dilations=DEFAULT_DILATIONS, num_flows=4, filter_width=2,
                         dilation_channels=32, skip_channels=128, latent_channels=32, seed=43,
                         dtype=np.float32, dead_vars=True):
    """Student flow variables (model.py:415-454); the skip conv exists but is dead."""
    rng = np.random.default_rng(seed)
    w = {}
    for f in range(num_flows):
        w.update(_stack_weights(rng, student_prefix(f), len(dilations), filter_width,
                                dilation_channels, skip_channels, latent_channels,
                                [(dilation_channels, 2)], dtype, dead_vars))
    return w


ENCODER_PREFIX = 'WaveNetAutoEncoder/Encoder/'


def encoder_conv_name(idx):
    """idx-th tf.layers.conv1d created directly in the Encoder scope (ops.py:54-55, model.py:152)."""
    return 'conv1d' if idx == 0 else 'conv1d_%d' % idx


def make_encoder_weights(n_layers=len(DEFAULT_DILATIONS), filter_width=2, encoder_channels=128,
                         skip_channels=128, latent_channels=32, seed=44, dtype=np.float32,
                         dead_vars=True, gain=2.0):
    """Teacher encoder variables (model.py:137-155, ops.py:48-57).  Layer j (0 = ``nc_conv``,
    i+1 = ``dilated_conv_i``): ``<name>_NC/conv1d`` [K,Cin,E], ``conv1d_{2j}`` residual [1,E,E],
    ``conv1d_{2j+1}`` skip [1,E,S] (dead for nc_conv: model.py:141 discards it); latent conv
    ``conv1d_{2(n+1)}`` [1,S,latent].  The residual kernels are Glorot x ``gain``: the encoder has no
    skip connection around its two relus per layer, so with plain Glorot the signal halves per layer
    and the encoding would be bias noise, useless as a parity probe."""
    rng = np.random.default_rng(seed)
    p, E, S = ENCODER_PREFIX, encoder_channels, skip_channels
    w = {}
    for j in range(n_layers + 1):
        name = 'nc_conv' if j == 0 else 'dilated_conv_%d' % (j - 1)
        cin = 1 if j == 0 else E
        w['%s%s_NC/conv1d/kernel' % (p, name)] = _glorot(rng, (filter_width, cin, E), dtype)
        w['%s%s_NC/conv1d/bias' % (p, name)] = _bias(rng, (E,), dtype)
        w[p + encoder_conv_name(2 * j) + '/kernel'] = (_glorot(rng, (1, E, E), dtype) * gain).astype(dtype)
        w[p + encoder_conv_name(2 * j) + '/bias'] = _bias(rng, (E,), dtype)
        if j > 0 or dead_vars:
            w[p + encoder_conv_name(2 * j + 1) + '/kernel'] = _glorot(rng, (1, E, S), dtype)
            w[p + encoder_conv_name(2 * j + 1) + '/bias'] = _bias(rng, (S,), dtype)
    lat = p + encoder_conv_name(2 * (n_layers + 1))
    w[lat + '/kernel'] = _glorot(rng, (1, S, latent_channels), dtype)
    w[lat + '/bias'] = _bias(rng, (latent_channels,), dtype)
    return w


def synthetic_audio(batch, length, seed=1234, dtype=np.float32):
    """NSynth-shaped synthetic clips in the style of simple_audio.py:40-61: one of
    sine/square/saw/triangle at a random frequency + N(0, 0.05) noise, min-max
    normalised to [-1, 1].  Utterance b uses default_rng(seed + b)."""
    out = np.empty((batch, length), dtype=dtype)
    tt = np.arange(length, dtype=np.float64) / length
    for b in range(batch):
        rng = np.random.default_rng(seed + b)
        freq = float(rng.integers(18) + 22) * max(1, length // 4096)
        kind = int(rng.integers(4))
        ph = 2.0 * np.pi * freq * tt
        if kind == 0:
            wave = np.sin(ph)
        elif kind == 1:
            wave = np.sign(np.sin(ph))
        elif kind == 2:
            wave = 2.0 * ((freq * tt) % 1.0) - 1.0
        else:
            wave = 2.0 * np.abs(2.0 * ((freq * tt) % 1.0) - 1.0) - 1.0
        wave = wave + rng.normal(0.0, 0.05, size=length)
        lo, hi = wave.min(), wave.max()
        out[b] = ((wave - lo) / (hi - lo) * 2.0 - 1.0).astype(dtype)
    return out


def synthetic_encoding(batch, frames, channels=32, seed=4321, dtype=np.float32):
    """Latent conditioning [B, T/P, C] ~ N(0,1) (the encoder is outside the hot path)."""
    return np.random.default_rng(seed).normal(0, 1, size=(batch, frames, channels)).astype(dtype)


def logistic_noise(batch, length, seed=777, dtype=np.float32):
    """Student input noise z ~ Logistic(0,1), host-sampled like student.py:104."""
    return np.random.default_rng(seed).logistic(0, 1, size=(batch, length)).astype(dtype)


def sampler_uniforms(batch, length, num_mixtures=5, seed=999, dtype=np.float32):
    """The two tf.random_uniform draws of ops.py:187,196, U[1e-5, 1-1e-5]."""
    rng = np.random.default_rng(seed)
    u1 = rng.uniform(1e-5, 1 - 1e-5, size=(batch, length, num_mixtures)).astype(dtype)
    u2 = rng.uniform(1e-5, 1 - 1e-5, size=(batch, length)).astype(dtype)
    return u1, u2
