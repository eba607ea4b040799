"""-m gpu: the CUDA path against fixtures produced by EXECUTING the reference's own ops.py / model.py
(tests/golden/reference_*.npz, written by tests/golden/make_reference_golden.py through the NumPy TensorFlow stand-in
tests/tf_shim; the reference tree itself does not travel to the GPU box).  Tolerances: fp32 path <= 1e-4 relative,
fp16 tensor-core path <= 2e-2 max-abs (asserted at 1e-2) -- BASELINE.json north_star."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from sr_wavenet_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def srwn(lib):
    import sr_wavenet_b200
    assert torch.cuda.is_available()
    return sr_wavenet_b200


def _load(tag):
    with np.load("%s/reference_%s.npz" % (GOLDEN, tag)) as z:
        g = {k: z[k] for k in z.files}
    g["dil"] = [int(d) for d in g["dilations"]]
    return g


def _models(srwn, g, alpha=0.25):
    dil, T, P, C, F, M = g["dil"], int(g["T"]), int(g["P"]), int(g["C"]), int(g["F"]), int(g["M"])
    t = srwn.WaveNetAutoEncoder(input_size=T, condition_size=0, num_mixtures=M, dilations=dil, skip_channels=128,
                                latent_channels=C, pool_stride=P)
    tw = synth.make_teacher_weights(dil, latent_channels=C, num_mixtures=M, seed=int(g["teacher_seed"]))
    tw.update(synth.make_encoder_weights(len(dil), 2, 128, 128, C, seed=int(g["teacher_seed"]) + 2))
    t.set_weights(tw)
    s = srwn.ParallelWaveNet(input_size=T, condition_size=0, dilations=dil, teacher=t, num_flows=F, skip_channels=128,
                             latent_channels=C, pool_stride=P, alpha=alpha, beta=1.0, gamma=1.0)
    s.set_weights(synth.make_student_weights(dil, num_flows=F, latent_channels=C, seed=int(g["student_seed"])))
    return t, s


def _rel(a, ref):
    return float(np.abs(np.asarray(a, np.float64) - ref).max() / max(1.0, np.abs(ref).max()))


@pytest.mark.parametrize("tag", ["small", "default"])
def test_teacher_against_reference_fixture(srwn, tag):
    g = _load(tag)
    t, _ = _models(srwn, g)
    x, enc = g["x"], g["enc"]
    for prec in t.available_precisions():
        lg = t.get_logits(x, enc, precision=prec)
        nll = t.nll(x, enc, sum_all=False, precision=prec)
        tot = t.nll(x, enc, precision=prec)
        if prec == "fp32":
            assert _rel(lg, g["logits"]) <= 1e-4 and _rel(nll, g["nll"]) <= 1e-4
            assert abs(tot - float(g["nll_sum"])) <= 1e-4 * abs(float(g["nll_sum"]))
            rec = t.reconstruct_with_encoding(x, enc, u1=g["u1"], u2=g["u2"], precision=prec)
            assert np.abs(rec - g["sample"]).max() <= 1e-4                     # same mixture picks, same samples
        else:
            e = np.abs(lg - g["logits"]).max()
            print("teacher %s %s: max|dlogits| vs reference fixture %.3e" % (tag, prec, e))
            assert e <= 1e-2 and np.abs(nll - g["nll"]).max() <= 2e-2
            assert abs(tot - float(g["nll_sum"])) <= 2e-3 * abs(float(g["nll_sum"]))
    # encoder (model.py:137-155) on the fp32 path
    assert _rel(t.encode(x, precision="fp32"), g["encoding"]) <= 1e-4


def test_naive_ar_loop_against_reference_fixture(srwn):
    """teacher.py:153-170 run on the reference's code (T full decoder evaluations) == the queue kernel."""
    g = _load("small")
    t, _ = _models(srwn, g)
    n = g["ar_x"].shape[1]
    x = t.generate(g["enc"][:, :n // int(g["P"])], u1=g["u1"][:, :n], u2=g["u2"][:, :n], precision="fp32")
    assert np.abs(x - g["ar_x"]).max() <= 1e-4


@pytest.mark.parametrize("tag", ["small", "default"])
def test_student_against_reference_fixture(srwn, tag):
    g = _load(tag)
    _, s = _models(srwn, g)
    z, enc, x = g["z"], g["enc"], g["x"]
    for prec in s.available_precisions():
        r = s.forward_all(z, enc, precision=prec)
        tol = 1e-4 if prec == "fp32" else 2e-2
        assert np.abs(r["out"] - g["student_out"][:, :, 0]).max() <= tol, prec
        assert np.abs(r["s_tot"] / g["s_tot"][:, :, 0] - 1).max() <= 2 * tol, prec
        assert np.all(np.abs(r["mu_tot"] - g["mu_tot"][:, :, 0]) <= tol * (1 + np.abs(g["mu_tot"][:, :, 0]))), prec
    ent = s.getEntropy_fast(None, z, enc)
    assert abs(ent - float(g["entropy"])) <= 1e-4 * abs(float(g["entropy"]))
    # distillation loss graph (model.py:356-379), teacher teacher-forced on the truth (F5), fp32 teacher path
    loss, power, entropy, _ = s.loss_and_grads(z, x, enc, teacher_precision="fp32")
    print("student %s: loss %.6f (reference %.6f)  power %.6e (reference %.6e)" % (tag, float(loss), float(g["loss"]), float(power), float(g["power_loss"])))
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    assert abs(float(power) - float(g["power_loss"])) <= 1e-4 * abs(float(g["power_loss"]))
