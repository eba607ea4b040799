"""-m gpu: ops.py mirror (through the C ABI) vs the NumPy oracle."""
import numpy as np
import pytest
import torch

from oracle import srwn_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(lib):
    import sr_wavenet_b200 as srwn
    assert torch.cuda.is_available()
    return srwn.ops


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64)


def test_conv_known_answers(ops, conv_kat):
    x = conv_kat["x"].reshape(1, -1, 1)
    for key, f, d in (("l243", [1, 1], 1), ("l244", [1, 0, 1], 1), ("l245", [1, 0, 0, 0, 1], 1),
                      ("l246", [1, 1], 2), ("l247", [1, 1], 3), ("l248", [1, 1], 4), ("l249", [1, 1], 6)):
        w = np.array(f, np.float32).reshape(len(f), 1, 1)
        out = ops._DilatedCausalConv1d(x, w, dilation_rate=d)
        np.testing.assert_array_equal(_np(out).reshape(-1), conv_kat[key])      # ops.py:243-249
    f4 = np.array([[1, 2, 1, 2]], np.float32).reshape(2, 1, 2)
    np.testing.assert_array_equal(_np(ops._DilatedCausalConv1d(x, f4))[0], conv_kat["l252"])   # ops.py:252


@pytest.mark.parametrize("B,T,cin,cout,K,d", [(2, 37, 5, 7, 3, 4), (1, 1, 32, 32, 2, 512), (3, 130, 1, 32, 2, 1)])
def test_conv_random(ops, B, T, cin, cout, K, d):
    rng = np.random.default_rng(0)
    x = rng.normal(size=(B, T, cin)).astype(np.float32)
    w = rng.normal(size=(K, cin, cout)).astype(np.float32)
    ref = orc.dilated_causal_conv1d(x.astype(np.float64), w.astype(np.float64), d)
    np.testing.assert_allclose(_np(ops._DilatedCausalConv1d(x, w, d)), ref, rtol=1e-5, atol=1e-5)


def test_layer_builders_create_reference_variables(ops):
    ops.reset_variables()
    x = torch.ones(1, 8, 1, device="cuda")
    with ops.variable_scope("Decoder"):
        conv1 = ops.DilatedCausalConv1d(x, kernel_size=3, channels=4, dilation_rate=4, name='causal_conv1')   # ops.py:232
        h = torch.ones(1, 8, 8, device="cuda")
        dense, skip = ops.ResidualDilationLayer(h, kernel_size=2, dilation_channels=8, skip_channels=4,
                                                dilation_rate=4, name='dilation_layer1')                     # ops.py:234
    assert conv1.shape == (1, 8, 4) and dense.shape == (1, 8, 8) and skip.shape == (1, 8, 4)
    names = set(ops.global_variables())
    assert {"Decoder/causal_conv1_Kernel", "Decoder/causal_conv1_Bias",
            "Decoder/dilation_layer1_filter/dilation_layer1_Kernel",
            "Decoder/dilation_layer1_gate/dilation_layer1_Kernel",
            "Decoder/conv1d/kernel", "Decoder/conv1d_1/kernel", "Decoder/conv1d_1/bias"} <= names
    v = ops.global_variables()
    ref_d, ref_s = orc.residual_dilation_layer(
        np.ones((1, 8, 8)), _np(v["Decoder/dilation_layer1_filter/dilation_layer1_Kernel"]),
        _np(v["Decoder/dilation_layer1_filter/dilation_layer1_Bias"]), _np(v["Decoder/conv1d/kernel"]),
        _np(v["Decoder/conv1d/bias"]), _np(v["Decoder/conv1d_1/kernel"]), _np(v["Decoder/conv1d_1/bias"]), 4)
    np.testing.assert_allclose(_np(dense), ref_d, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(_np(skip), ref_s, rtol=1e-5, atol=1e-6)
    ops.reset_variables()


def test_shift_and_resize(ops):
    rng = np.random.default_rng(1)
    x = rng.normal(size=(2, 9, 3)).astype(np.float32)
    np.testing.assert_array_equal(_np(ops.RightShift(x)), orc.right_shift(x.astype(np.float64)))
    np.testing.assert_array_equal(_np(ops.RightShift(x, 3)), orc.right_shift(x.astype(np.float64), 3))
    for out in (9, 18, 1152, 20):
        np.testing.assert_array_equal(_np(ops.ResizeEmbeddingNearestNeighbor(x, out)),
                                      orc.resize_embedding_nearest_neighbor(x.astype(np.float64), out))


def test_mol_loss_and_sample(ops, golden_small):
    g = golden_small
    x, l = g["x"][:, :, None], g["logits"].astype(np.float32)
    nll = _np(ops.discretized_mix_logistic_loss(x, l, sum_all=False))
    np.testing.assert_allclose(nll, g["nll"], rtol=1e-4, atol=1e-4)
    tot = float(ops.discretized_mix_logistic_loss(x, l, sum_all=True))
    assert abs(tot - float(g["nll_sum"])) <= 1e-5 * abs(float(g["nll_sum"]))
    s, idx = ops.sample_from_discretized_mix_logistic(l, int(g["M"]), g["u1"], g["u2"], return_index=True)
    np.testing.assert_array_equal(idx.cpu().numpy(), g["sample_idx"])          # identical mixture argmax
    np.testing.assert_allclose(_np(s), g["sample"], rtol=1e-4, atol=1e-5)


def test_mol_loss_edge_cases(ops):
    M = 5
    rng = np.random.default_rng(1)
    l = rng.normal(size=(1, 6, 4 * M)).astype(np.float32)
    l[0, 3, 2 * M:3 * M] = -9.0
    l[0, 4, M:2 * M] = 5.0
    l[0, 4, 2 * M:3 * M] = -7.0
    x = np.array([[-1.0, 1.0, 0.0, 0.3, 0.0, 0.9995]], np.float32)[:, :, None]
    ref = orc.discretized_mix_logistic_loss(x.astype(np.float64), l.astype(np.float64), sum_all=False)
    got = _np(ops.discretized_mix_logistic_loss(x, l, sum_all=False))
    np.testing.assert_allclose(got, ref, rtol=2e-4, atol=2e-4)


def test_sampler_draws_stay_in_range(ops):
    l = torch.randn(2, 64, 20, device="cuda")
    s = ops.sample_from_discretized_mix_logistic(l, 5)
    assert s.shape == (2, 64, 1) and float(s.min()) >= -1 and float(s.max()) <= 1
