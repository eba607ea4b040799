"""-m gpu: a fused launch whose pipeline wait expires must fail LOUDLY (the 16-bit paths returned rc = OK with
partially written outputs).  The wait limit is a handle option (srwn_set_wait_limit); one clock makes the first wait that
is not already satisfied expire, so the launch aborts, walks its barriers to the end and leaves its abort words in pinned
host memory: the host-boundary call raises, device-resident callers are refused at their next call."""
import numpy as np
import pytest
import torch

from sr_wavenet_b200 import synth, _lib

pytestmark = pytest.mark.gpu
DIL = synth.DEFAULT_DILATIONS


def test_aborted_fused_launch_raises(lib):
    import sr_wavenet_b200 as srwn
    B, T = 2, 4096
    t = srwn.WaveNetAutoEncoder(input_size=T, condition_size=0, num_mixtures=5, dilations=DIL, skip_channels=128,
                                latent_channels=32, pool_stride=128)
    t.set_weights(synth.make_teacher_weights(DIL))
    x, enc = synth.synthetic_audio(B, T), synth.synthetic_encoding(B, T // 128)
    good = t.get_logits(x, enc, precision="fp16")
    _lib.check(lib.srwn_set_wait_limit(t._eng.h, 1))
    with pytest.raises(_lib.SrwnError, match="aborted"):
        t.get_logits(x, enc, precision="fp16")                       # NumPy in -> NumPy out: raises instead of returning garbage
    # device-resident caller: no synchronisation in the call itself; the next call on the handle is refused
    xd, ed = torch.from_numpy(x).cuda(), torch.from_numpy(enc).cuda()
    t.get_logits(xd, ed, precision="fp16")
    torch.cuda.synchronize()
    with pytest.raises(_lib.SrwnError, match="aborted"):
        t.get_logits(xd, ed, precision="fp16")
    with pytest.raises(_lib.SrwnError):
        t._eng.check_async(_lib.OP_TEACHER_LOGITS, B, T, _lib.FP16)   # reports and clears
    _lib.check(lib.srwn_set_wait_limit(t._eng.h, 1000000000))
    again = t.get_logits(x, enc, precision="fp16")
    np.testing.assert_array_equal(again, good)
