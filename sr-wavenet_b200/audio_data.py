"""Host-side audio source for the drivers (teacher.py / student.py at the repo root).

The reference reads NSynth TFRecords through a private tf.Session (nsynth.py:5-52: parse, shuffle, repeat, batch;
``next()`` returns ``(audio[:, :num_samples], one_hot(pitch, 128))``).  The same ``next()`` contract is served from a
``*.tfrecord`` file (``nsynth.NsynthDataReader``, a TensorFlow-free reader of the same format), from a directory of
.wav files (int16 / 32767 like data.py:7-123) or from the synthetic waves of simple_audio.py:40-61.  Pure host I/O:
nothing here touches the GPU path."""
import glob
import os

import numpy as np

from . import synth


class AudioReader(object):
    def __init__(self, source, batch_size, num_samples, seed=1234, audio_max_length=16000):
        self.batch_size, self.num_samples = batch_size, num_samples
        self._cursor = 0
        self._seed = seed
        self._files = None
        self._tfrecord = None
        if source not in (None, "synthetic") and os.path.isfile(source):
            from .nsynth import NsynthDataReader              # teacher.py:53: NsynthDataReader(path, batch, num_samples, audio_max_length=16000)
            self._tfrecord = NsynthDataReader(source, batch_size, num_samples, audio_max_length=audio_max_length)
        elif source not in (None, "synthetic"):
            self._files = sorted(glob.glob(os.path.join(source, "*.wav")))
            if not self._files:
                raise ValueError("no .wav files under %s" % source)

    def _read_wav(self, path):
        from scipy.io import wavfile
        _, a = wavfile.read(path)
        a = a[:, 0] if a.ndim == 2 else a
        a = a.astype(np.float32) / (32767.0 if a.dtype.kind == "i" else 1.0)
        if a.shape[0] < self.num_samples:
            a = np.pad(a, (0, self.num_samples - a.shape[0]))
        return np.clip(a[:self.num_samples], -1.0, 1.0)

    def next(self):
        """-> (audio [B, num_samples] float32 in [-1, 1], one-hot pitch [B, 128])."""
        B = self.batch_size
        if self._tfrecord is not None:
            return self._tfrecord.next()
        if self._files is None:
            x = synth.synthetic_audio(B, self.num_samples, seed=self._seed + self._cursor)
        else:
            x = np.stack([self._read_wav(self._files[(self._cursor + i) % len(self._files)]) for i in range(B)])
        self._cursor += B
        y = np.zeros((B, 128), dtype=np.float32)
        y[:, 60] = 1.0                                       # the reference trains on the pitch-60 subset (filter_tfrecord.py)
        return x.astype(np.float32), y


def write_wav(path, sample_rate, x):
    from scipy.io import wavfile
    wavfile.write(path, sample_rate, np.asarray(x, dtype=np.float32))
