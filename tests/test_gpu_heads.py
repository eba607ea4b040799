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
