"""Times the teacher encoder: python tools/prof_enc.py 32x64000 fp16"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sr_wavenet_b200 as srwn
from sr_wavenet_b200 import synth, _lib
B, T = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "32x64000").split("x"))
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dil = synth.DEFAULT_DILATIONS
m = srwn.WaveNetAutoEncoder(T, 0, 5, dil, skip_channels=128, latent_channels=32, pool_stride=128)
m.set_weights(synth.make_encoder_weights(len(dil)))
x = torch.from_numpy(synth.synthetic_audio(B, T)).cuda()
m._enc_eng.set_profiling(True)
ms = []
for i in range(iters):
    enc = m.encode(x, precision=prec)
    ms.append(m._enc_eng.last_ms())
best = min(ms)
flop = 2.0 * (2 * 128 + 128 * 128) + 30 * 2.0 * (256 * 128 + 128 * 128) - 2.0 * 128 * 128
print("encode %dx%d %s: ms %s -> %.1f Msamples/s, %.1f TFLOP/s (conv+residual GEMMs), %.0f GB/s of 16-bit activations" % (
    B, T, prec, ["%.2f" % v for v in ms], B * T / best / 1e3, flop * B * T / best / 1e9,
    B * T * 512.0 * 30 / best / 1e6))
