"""-m gpu: the generator entry points of teacher.py / student.py run end to end on small clips."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def srwn(lib):
    import sr_wavenet_b200
    assert torch.cuda.is_available()
    return sr_wavenet_b200


def test_teacher_driver_fast_and_slow(srwn, tmp_path):
    import teacher as drv
    tdir = str(tmp_path / "teacher")
    t = srwn.WaveNetAutoEncoder(input_size=1024, condition_size=0, num_mixtures=5, dilations=drv.DILATIONS,
                                latent_channels=32, skip_channels=128, pool_stride=128)
    assert t.save(tdir, 7, force=True)
    out = str(tmp_path / "out")
    res = drv.main(["--teacher", tdir, "--test-fast", "--test-slow", "--num-samples", "1024", "--batch-size", "2",
                    "--clips", "1", "--out-dir", out, "--precision", "fp32", "--check-naive", "48"])
    assert res["test_fast_samples_per_s"] > 0 and res["test_slow_samples_per_s"] > 0
    # the queue kernel and the reference's per-sample loop (teacher.py:153-170) produce the same audio
    assert res["naive_max_abs_diff"] <= 1e-4
    from scipy.io import wavfile
    rate, regen = wavfile.read(os.path.join(out, "regen_wav_0.wav"))
    assert rate == 4000 and regen.shape == (1024,) and np.all(np.abs(regen) <= 1.0) and regen[-1] == 0.0   # teacher.py:170
    with pytest.raises(SystemExit):
        drv.main(["--train"])


def test_student_driver_train_and_test(srwn, tmp_path):
    import teacher as tdrv
    import student as drv
    tdir, sdir, out = str(tmp_path / "teacher"), str(tmp_path / "student"), str(tmp_path / "out")
    t = srwn.WaveNetAutoEncoder(input_size=1024, condition_size=0, num_mixtures=5, dilations=tdrv.DILATIONS,
                                latent_channels=32, skip_channels=128, pool_stride=128)
    t.save(tdir, 1, force=True)
    common = ["--teacher", tdir, "--student", sdir, "--num-samples", "1024", "--batch-size", "2", "--out-dir", out,
              "--seed", "3", "--learning-rate", "1e-3"]
    res = drv.main(common + ["--train", "--steps", "4", "--print-steps", "2"])
    assert len(res["losses"]) == 4 and all(np.isfinite(res["losses"]))
    assert os.path.exists(os.path.join(sdir, "checkpoint"))
    # the checkpoint holds the TRAINED weights, and a second run restores them
    s2 = srwn.ParallelWaveNet(input_size=1024, condition_size=0, dilations=drv.DILATIONS, teacher=tdir, num_flows=4,
                              skip_channels=128, latent_channels=32, pool_stride=128)
    init = s2.get_weights()
    assert s2.load(None, sdir)
    name = "ParallelWaveNet/Flow0/Flow0/conv1d_1/kernel"
    assert np.abs(s2.get_weights()[name] - init[name]).max() > 0
    assert isinstance(s2.teacher, srwn.WaveNetAutoEncoder) and s2.teacher.num_mixtures == 5
    np.testing.assert_array_equal(s2.teacher.get_weights()["WaveNetAutoEncoder/Decoder/causal_conv_Kernel"],
                                  t.get_weights()["WaveNetAutoEncoder/Decoder/causal_conv_Kernel"])
    # --test reads an NSynth-style TFRecord through the TensorFlow-free reader (nsynth.py:5-52)
    from sr_wavenet_b200 import nsynth, synth
    rec = str(tmp_path / "clips.tfrecord")
    clips = synth.synthetic_audio(4, 2048)
    nsynth.write_tfrecord(rec, [{"audio": c, "pitch": np.array([60])} for c in clips])
    res = drv.main(common + ["--test", "--clips", "1", "--data", rec, "--audio-max-length", "2048"])
    assert res["output_shape"] == (2, 1024, 1)
