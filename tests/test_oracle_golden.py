"""Oracle vs the committed golden vectors, plus the size-independent properties the domain
offers (causality, receptive field, queue-AR == naive loop, flow composition identity)."""
import os

import numpy as np
import pytest

from conftest import f64, GOLDEN
from oracle import srwn_oracle as orc
import sr_wavenet_b200.synth as synth


def _small_models(g):
    dil = [int(d) for d in g["dilations"]]
    tw = synth.make_teacher_weights(dil, latent_channels=int(g["C"]), num_mixtures=int(g["M"]),
                                    seed=int(g["teacher_seed"]))
    sw = synth.make_student_weights(dil, num_flows=int(g["F"]), latent_channels=int(g["C"]),
                                    seed=int(g["student_seed"]))
    return dil, f64(tw), f64(sw)


def test_golden_small_teacher(golden_small):
    g = golden_small
    dil, tw, _ = _small_models(g)
    x, enc = g["x"].astype(np.float64), g["enc"].astype(np.float64)
    logits = orc.teacher_decoder_logits(tw, x, enc, dil, int(g["P"]))
    np.testing.assert_allclose(logits, g["logits"], rtol=1e-12, atol=1e-12)
    nll = orc.discretized_mix_logistic_loss(x[:, :, None], logits, sum_all=False)
    np.testing.assert_allclose(nll, g["nll"], rtol=1e-12, atol=1e-12)
    # sum_all=True is the sum of sum_all=False (ops.py:172 vs :175)
    np.testing.assert_allclose(orc.discretized_mix_logistic_loss(x[:, :, None], logits, True),
                               nll.sum(), rtol=1e-12)
    s, idx = orc.sample_from_discretized_mix_logistic(
        logits, int(g["M"]), g["u1"].astype(np.float64), g["u2"].astype(np.float64)[:, :, None], True)
    np.testing.assert_allclose(s, g["sample"], rtol=1e-12, atol=1e-12)
    np.testing.assert_array_equal(idx, g["sample_idx"])
    assert s.min() >= -1 and s.max() <= 1


def test_golden_small_student(golden_small):
    g = golden_small
    dil, _, sw = _small_models(g)
    net = orc.student_network(sw, g["z"].astype(np.float64), g["enc"].astype(np.float64), dil,
                              int(g["P"]), int(g["F"]))
    np.testing.assert_allclose(net["out"], g["student_out"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(net["s_tot"], g["s_tot"], rtol=1e-12)
    # z*s_tot + mu_tot is the chained flow output (model.py:517-535 vs :510)
    z = g["z"].astype(np.float64)[:, :, None]
    np.testing.assert_allclose(z * net["s_tot"] + net["mu_tot"], net["x_last"], rtol=1e-9, atol=1e-9)


def test_queue_ar_equals_naive_loop(golden_small):
    """teacher.py:153-170 (one full decoder pass per sample) == per-layer dilation queues."""
    g = golden_small
    dil, tw, _ = _small_models(g)
    T = 24
    enc = g["enc"].astype(np.float64)[:, :2]          # 2 frames * P=16 = 32 >= T; use T=32
    T = 32
    u1, u2 = g["u1"].astype(np.float64)[:, :T], g["u2"].astype(np.float64)[:, :T]
    naive = orc.naive_ar_loop(tw, enc, dil, int(g["P"]), int(g["M"]), u1, u2, T)
    fast = orc.queue_ar(tw, enc, dil, int(g["P"]), int(g["M"]), u1, u2, T)
    np.testing.assert_allclose(fast, naive, rtol=1e-10, atol=1e-12)


def test_golden_small_ar(golden_small):
    g = golden_small
    dil, tw, _ = _small_models(g)
    x, lg = orc.queue_ar(tw, g["enc"].astype(np.float64), dil, int(g["P"]), int(g["M"]),
                         g["u1"].astype(np.float64), g["u2"].astype(np.float64), g["x"].shape[1], return_logits=True)
    np.testing.assert_allclose(x, g["ar_x"], rtol=1e-12, atol=1e-12)
    # teacher-forcing the generated audio reproduces the AR logits (self-consistency)
    tf_logits = orc.teacher_decoder_logits(tw, x, g["enc"].astype(np.float64), dil, int(g["P"]))
    np.testing.assert_allclose(tf_logits, lg, rtol=1e-9, atol=1e-9)


def test_golden_default_cfg(golden_default):
    g = golden_default
    dil = synth.DEFAULT_DILATIONS
    B, T, P = int(g["B"]), int(g["T"]), int(g["P"])
    tw = f64(synth.make_teacher_weights(dil, seed=42))
    x = synth.synthetic_audio(B, T, seed=1234).astype(np.float64)
    enc = synth.synthetic_encoding(B, T // P, 32, seed=4321).astype(np.float64)
    logits = orc.teacher_decoder_logits(tw, x, enc, dil, P)
    np.testing.assert_allclose(logits, g["logits"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(orc.discretized_mix_logistic_loss(x[:, :, None], logits, True),
                               g["nll_sum"], rtol=1e-10)


def test_causality_and_receptive_field():
    """Perturbing x[t0:] leaves logits[:t0+1] unchanged (RightShift + causal convs); logits at t
    depend on x no further back than 1 + 1 + sum(d) samples (3071 at defaults)."""
    dil = [1, 2, 4, 8]
    B, T, P, C = 1, 64, 16, 4
    tw = f64(synth.make_teacher_weights(dil, latent_channels=C, seed=3))
    x = synth.synthetic_audio(B, T, seed=9).astype(np.float64)
    enc = synth.synthetic_encoding(B, T // P, C, seed=10).astype(np.float64)
    base = orc.teacher_decoder_logits(tw, x, enc, dil, P)
    t0 = 40
    x2 = x.copy()
    x2[:, t0:] += 0.3
    pert = orc.teacher_decoder_logits(tw, x2, enc, dil, P)
    np.testing.assert_array_equal(pert[:, :t0 + 1], base[:, :t0 + 1])
    assert np.abs(pert[:, t0 + 1] - base[:, t0 + 1]).max() > 0
    rf = 1 + 1 + sum(dil)          # front conv (K=2) + right shift + dilated taps
    x3 = x.copy()
    x3[:, 10] += 0.5
    pert3 = orc.teacher_decoder_logits(tw, x3, enc, dil, P)
    changed = np.nonzero(np.abs(pert3 - base).max(axis=(0, 2)) > 0)[0]
    assert changed.min() == 11 and changed.max() == 10 + rf


def test_resize_and_shift():
    x = np.arange(12, dtype=np.float64).reshape(1, 3, 4)
    up = orc.resize_embedding_nearest_neighbor(x, 12)
    np.testing.assert_array_equal(up[0, :, 0], np.repeat(x[0, :, 0], 4))   # pure repeat for integer ratios
    y = orc.right_shift(x)
    np.testing.assert_array_equal(y[0, 0], 0)
    np.testing.assert_array_equal(y[0, 1:], x[0, :-1])


def test_gate_is_sigmoid_of_tanh():
    """ops.py:33: gate = sigmoid(tanh(filter conv)); the _gate conv never contributes."""
    rng = np.random.default_rng(0)
    x = rng.normal(size=(1, 16, 32))
    fk, fb = rng.normal(size=(2, 32, 32)) * 0.1, rng.normal(size=(1, 1, 32)) * 0.1
    rk, rb = rng.normal(size=(1, 32, 32)) * 0.1, np.zeros(32)
    sk, sb = rng.normal(size=(1, 32, 128)) * 0.1, np.zeros(128)
    dense, skip = orc.residual_dilation_layer(x, fk, fb, rk, rb, sk, sb, 2)
    f = np.tanh(orc.dilated_causal_conv1d(x, fk, 2) + fb)
    c = f / (1 + np.exp(-f))
    np.testing.assert_allclose(skip, c @ sk[0], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(dense, (x + c @ rk[0]) * 0.7071067811865476, rtol=1e-12, atol=1e-12)


def test_mol_loss_edge_branches():
    """ops.py:167: x<-0.999 / x>0.999 / tiny-probability branches stay finite and positive."""
    M = 5
    rng = np.random.default_rng(1)
    l = rng.normal(size=(1, 6, 4 * M))
    l[0, 3, 2 * M:3 * M] = -9.0            # log-scale clamp at -7 (ops.py:136)
    l[0, 4, M:2 * M] = 5.0                 # mean far away -> cdf_delta < 1e-5 branch
    l[0, 4, 2 * M:3 * M] = -7.0
    x = np.array([[-1.0, 1.0, 0.0, 0.3, 0.0, 0.9995]])[:, :, None]
    nll = orc.discretized_mix_logistic_loss(x, l, sum_all=False)
    assert np.isfinite(nll).all()
    assert nll[0, 4, 0] > 50               # far-tail approximation is a large but finite penalty


@pytest.mark.parametrize("dtype", [np.float32])
def test_oracle_fp32_tracks_fp64(golden_small, dtype):
    g = golden_small
    dil = [int(d) for d in g["dilations"]]
    tw = synth.make_teacher_weights(dil, latent_channels=int(g["C"]), num_mixtures=int(g["M"]),
                                    seed=int(g["teacher_seed"]))
    lg32 = orc.teacher_decoder_logits(tw, g["x"], g["enc"], dil, int(g["P"]))
    assert lg32.dtype == np.float32
    assert np.abs(lg32 - g["logits"]).max() < 1e-4


def test_torch_cpu_port_matches_numpy_oracle(golden_default):
    """The timed CPU baseline (oracle/torch_cpu.py) computes the same thing as the NumPy oracle."""
    from oracle.torch_cpu import TeacherCPU
    g = golden_default
    dil = synth.DEFAULT_DILATIONS
    B, T, P = int(g["B"]), int(g["T"]), int(g["P"])
    tw = synth.make_teacher_weights(dil, seed=42)
    x, enc = synth.synthetic_audio(B, T, seed=1234), synth.synthetic_encoding(B, T // P, 32, seed=4321)
    cpu = TeacherCPU(tw, dil, P, 5)
    assert np.abs(cpu.logits(x, enc).numpy() - g["logits"]).max() < 1e-4
    assert abs(cpu.nll(x, enc) - float(g["nll_sum"])) < 1e-4 * abs(float(g["nll_sum"]))


def test_encoder_golden_and_same_padding():
    """Teacher encoder restatement (model.py:137-155, ops.py:48-57) against the committed vectors, plus a
    hand-checked case of the K=2 SAME-padding tap order: out[t] = relu(x)[t] W0 + relu(x)[t+1] W1, zero past the end."""
    import os
    from conftest import GOLDEN
    with np.load(os.path.join(GOLDEN, "encoder_small.npz")) as z:
        g = {k: z[k] for k in z.files}
    L, E, S, C, P = (int(g[k]) for k in "LESCP")
    w = synth.make_encoder_weights(L, 2, E, S, C, seed=int(g["seed"]))
    enc = orc.teacher_encoder(f64(w), g["x"].astype(np.float64), L, P)
    assert np.array_equal(enc, g["encoding"])
    x = np.array([[[1.0], [-2.0], [3.0], [4.0]]])
    ck = np.array([[[1.0]], [[10.0]]])                    # W0 = 1, W1 = 10
    res, skip = orc.residual_dilation_layer_nc(x, ck, np.zeros(1), np.array([[[2.0]]]), np.array([0.5]),
                                               np.array([[[1.0, -1.0]]]), np.zeros(2))
    # relu(x) = [1,0,3,4]; conv = [1+0, 0+30, 3+40, 4+0] = [1,30,43,4]
    assert np.array_equal(res[0, :, 0], np.array([2.5, 60.5, 86.5, 8.5]))
    assert np.array_equal(skip[0], np.array([[1, -1], [30, -30], [43, -43], [4, -4]], dtype=np.float64))


@pytest.mark.parametrize("tag", ["small", "default"])
def test_oracle_matches_reference_fixtures(tag):
    """The oracle against outputs of the reference's own code (tests/golden/reference_*.npz, produced by running
    /root/reference/ops.py + model.py through tests/tf_shim; see tests/test_reference_shim.py for the live comparison).
    This copy of the pin travels to machines without the reference tree."""
    with np.load(os.path.join(GOLDEN, "reference_%s.npz" % tag)) as z:
        g = {k: z[k] for k in z.files}
    dil, P, C, F, M = [int(d) for d in g["dilations"]], int(g["P"]), int(g["C"]), int(g["F"]), int(g["M"])
    tw = synth.make_teacher_weights(dil, latent_channels=C, num_mixtures=M, seed=int(g["teacher_seed"]))
    tw.update(synth.make_encoder_weights(len(dil), 2, 128, 128, C, seed=int(g["teacher_seed"]) + 2))
    tw, sw = f64(tw), f64(synth.make_student_weights(dil, num_flows=F, latent_channels=C, seed=int(g["student_seed"])))
    x, enc, z_ = (g[k].astype(np.float64) for k in ("x", "enc", "z"))
    u1, u2 = g["u1"].astype(np.float64), g["u2"].astype(np.float64)
    tol = dict(rtol=1e-10, atol=1e-11)
    logits = orc.teacher_decoder_logits(tw, x, enc, dil, P)
    np.testing.assert_allclose(logits, g["logits"], **tol)
    np.testing.assert_allclose(orc.discretized_mix_logistic_loss(x[:, :, None], logits, False), g["nll"], **tol)
    np.testing.assert_allclose(orc.sample_from_discretized_mix_logistic(logits, M, u1, u2[:, :, None])[:, :, 0], g["sample"], **tol)
    np.testing.assert_allclose(orc.teacher_encoder(tw, x, len(dil), P), g["encoding"], **tol)
    net = orc.student_network(sw, z_, enc, dil, P, F)
    for k, ref in (("out", "student_out"), ("s_tot", "s_tot"), ("mu_tot", "mu_tot")):
        np.testing.assert_allclose(net[k], g[ref], **tol)
    loss, power, entropy = orc.distillation_loss(sw, tw, z_, x, enc, dil, P, F, alpha=0.25, beta=1.0, gamma=1.0)
    np.testing.assert_allclose([loss, power, entropy], [g["loss"], g["power_loss"], g["entropy"]], rtol=1e-10)
    if "ar_x" in g:
        n = g["ar_x"].shape[1]
        np.testing.assert_allclose(orc.queue_ar(tw, enc[:, :n // P], dil, P, M, u1[:, :n], u2[:, :n], n), g["ar_x"], rtol=1e-9, atol=1e-10)
