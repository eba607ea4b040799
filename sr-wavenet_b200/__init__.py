"""srwn-b200: B200-native SR-WaveNet hot path (dilated causal residual stack for the teacher
decoder, autoregressive generation and the student's IAF flows).

The directory is named ``sr-wavenet_b200`` (not importable as-is); ``import sr_wavenet_b200``
works through the alias module at the repo root.
"""
from . import _lib, synth, shard  # noqa: F401
from . import ops, model   # noqa: F401
from .model import WaveNetAutoEncoder, ParallelWaveNet  # noqa: F401
from .heads import WaveNet, SiameseWaveNet  # noqa: F401

__all__ = ["ops", "model", "synth", "WaveNetAutoEncoder", "ParallelWaveNet", "WaveNet", "SiameseWaveNet"]
