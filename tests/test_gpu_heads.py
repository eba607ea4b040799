"""-m gpu: WaveNet (model.py:8-72) and SiameseWaveNet (model.py:660-798) forward passes on the device against the fixture
written by executing the reference's classes (tests/golden/make_reference_heads_golden.py).  fp32 path: <= 1e-4 relative."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from test_heads import _fixture, _build

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _weights(g, tag):
    return {k[len(tag) + 3:]: g[k] for k in g if k.startswith(tag + "_w/")}


def _close(got, ref, name):
    err = np.abs(np.asarray(got, np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30)
    print("%s: max error relative to the tensor's scale %.2e" % (name, err))
    assert err <= TOL, (name, err)


def test_wavenet_classifier_matches_reference(lib):
    g = _fixture()
    wn, _ = _build(g)
    wn.set_weights(_weights(g, "wn"))
    _close(wn.get_logits(g["wn_x"]).cpu().numpy(), g["wn_logits"], "WaveNet logits")
    out = wn.predict(g["wn_x"])
    assert out.shape == g["wn_out"].shape == (3, 1, 7)
    _close(out, g["wn_out"], "WaveNet softmax")
    np.testing.assert_allclose(out.sum(-1), 1.0, rtol=1e-5)
    np.testing.assert_allclose(wn.loss(g["wn_x"], g["wn_targets"]), float(g["wn_loss"]), rtol=TOL)
    _close(wn.get_logits(g["wn_x_long"]).cpu().numpy(), g["wn_logits_long"], "WaveNet logits, sliding window")      # 6 frames


def test_siamese_embedding_distance_loss_match_reference(lib):
    g = _fixture()
    _, si = _build(g)
    si.set_weights(_weights(g, "si"))
    _close(si.get_embedding(None, g["si_xl"]), g["si_embedding"], "Siamese embedding")
    _close(si.get_distance(None, g["si_xl"], g["si_xr"]), g["si_distance"], "Siamese distance")
    loss, d = si.loss(None, g["si_xl"], g["si_xr"], g["si_labels"])
    np.testing.assert_allclose(loss, float(g["si_loss"]), rtol=TOL)


def test_siamese_checkpoint_roundtrip(lib, tmp_path):
    g = _fixture()
    _, si = _build(g)
    si.set_weights(_weights(g, "si"))
    assert si.save(None, str(tmp_path), 3, force=True)
    _, other = _build(g)
    assert other.load(None, str(tmp_path)) is True
    np.testing.assert_array_equal(other.get_embedding(None, g["si_xl"]), si.get_embedding(None, g["si_xl"]))
    assert other.load(None, str(tmp_path / "missing")) is None


@pytest.mark.parametrize("K", [2, 3])
def test_noncausal_block_matches_reference(lib, K):
    """ResidualDilationLayerNC (ops.py:48-57): same variable names as the reference's graph, residual and skip within 1e-4."""
    from sr_wavenet_b200 import ops
    g = _fixture()
    ops.reset_variables()
    x = g["nc%d_x" % K]
    with ops.variable_scope("NCtest"):
        ops.ResidualDilationLayerNC(x, K, dilation_channels=6, skip_channels=4, dilation_rate=7, name="blk")      # creates the variables
    assert sorted(ops.global_variables()) == [str(n) for n in g["nc%d_names" % K]]
    for n in ops.global_variables():
        ops._variables[n] = torch.from_numpy(np.ascontiguousarray(g["nc%d_w/%s" % (K, n)])).cuda()
    ops._layer_counts.clear()
    with ops.variable_scope("NCtest"):
        res, skip = ops.ResidualDilationLayerNC(x, K, dilation_channels=6, skip_channels=4, dilation_rate=7, name="blk")
    _close(res.cpu().numpy(), g["nc%d_residual" % K], "NC block residual, K = %d" % K)
    _close(skip.cpu().numpy(), g["nc%d_skip" % K], "NC block skip, K = %d" % K)
    ops.reset_variables()


def test_log_helpers_match_reference(lib):
    """log_prob_from_logits / log_sum_exp (ops.py:111-122) on logits shifted by +40."""
    from sr_wavenet_b200 import ops
    g = _fixture()
    _close(ops.log_prob_from_logits(g["lse_x"]).cpu().numpy(), g["lse_log_prob"], "log_prob_from_logits")
    lse = ops.log_sum_exp(g["lse_x"])
    assert tuple(lse.shape) == g["lse_log_sum_exp"].shape
    _close(lse.cpu().numpy(), g["lse_log_sum_exp"], "log_sum_exp")
