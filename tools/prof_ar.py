"""Times autoregressive generation (CUDA events around the kernel): python tools/prof_ar.py 256x1024 fp16"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sr_wavenet_b200 as srwn
from sr_wavenet_b200 import synth, _lib
B, T = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256x1024").split("x"))
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dil = synth.DEFAULT_DILATIONS
m = srwn.WaveNetAutoEncoder(T, 0, 5, dil, skip_channels=128, latent_channels=32, pool_stride=128)
m.set_weights(synth.make_teacher_weights(dil))
enc = torch.from_numpy(synth.synthetic_encoding(B, T // 128)).cuda()
u1, u2 = (torch.from_numpy(a).cuda() for a in synth.sampler_uniforms(B, T))
m._eng.set_profiling(True)
ms = []
for i in range(iters):
    x = m.generate(enc, u1=u1, u2=u2, precision=prec)
    ms.append(m._eng.last_kernel_ms()[0])
best = min(ms)
print("generate %dx%d %s: kernel ms %s -> %.2f Msamples/s, %.2f us/step" % (
    B, T, prec, ["%.2f" % v for v in ms], B * T / best / 1e3, best * 1e3 / T))
