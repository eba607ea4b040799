"""One small invocation of every tensor-core path (fused teacher / student incl. a 3-CTA team hand-off and the on-device
noise, the generation kernel, the encoder) for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_small.py`.
Sizes are tiny: the sanitizer serialises and instruments every access."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sr_wavenet_b200 as srwn
from sr_wavenet_b200 import synth, _lib

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dil = synth.DEFAULT_DILATIONS
B, T = 2, 1152          # 3 chunks per utterance
if which in ("all", "teacher"):
    t = srwn.WaveNetAutoEncoder(T, 0, 5, dil, skip_channels=128, latent_channels=32, pool_stride=128)
    w = synth.make_teacher_weights(dil)
    w.update(synth.make_encoder_weights(len(dil)))
    t.set_weights(w)
    x, enc = synth.synthetic_audio(B, T), synth.synthetic_encoding(B, T // 128)
    for G in (1, 3):
        t._eng.set_team_size(G)
        lg = t.get_logits(x, enc, precision="fp16")
        nll = t.nll(x, enc, precision="fp16")
        print("teacher fp16 G=%d partition %s: nll %.3f finite %s" % (G, t._eng.last_partition(), nll, np.isfinite(lg).all()))
    e = t.encode(x[:, :1024], precision="fp16")
    print("encoder fp16:", e.shape, np.isfinite(e).all())
    u1, u2 = synth.sampler_uniforms(B, 256)
    g = t.generate(enc[:, :2], u1=u1, u2=u2, precision="fp16")
    print("generate fp16:", g.shape, np.isfinite(g).all())
if which in ("all", "student"):
    s = srwn.ParallelWaveNet(T, 0, dil, None, num_flows=2, skip_channels=128, latent_channels=32, pool_stride=128)
    s.set_weights(synth.make_student_weights(dil, 2))
    enc = synth.synthetic_encoding(B, T // 128)
    s._eng.set_team_size(3)
    out = s.generate(None, None, enc, precision="fp16")
    print("student fp16 with on-device noise, partition %s:" % (s._eng.last_partition(),), out.shape, np.isfinite(out).all())
torch.cuda.synchronize()
print("sanitize_small done")
