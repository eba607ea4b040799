"""Runs the fused teacher scoring kernel a few times and prints its CUDA-event time (for ncu / tuning)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sr_wavenet_b200 as srwn
from sr_wavenet_b200 import synth, _lib
B, T = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "32x64000").split("x"))
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
what = sys.argv[4] if len(sys.argv) > 4 else "teacher"
team = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dil = synth.DEFAULT_DILATIONS
enc = torch.from_numpy(synth.synthetic_encoding(B, T // 128)).cuda()
if what == "teacher":
    m = srwn.WaveNetAutoEncoder(T, 0, 5, dil, skip_channels=128, latent_channels=32, pool_stride=128)
    m.set_weights(synth.make_teacher_weights(dil))
    x = torch.from_numpy(synth.synthetic_audio(B, T)).cuda()
    run = lambda: m.nll(x, enc, precision=prec)
    op = _lib.OP_TEACHER_NLL
else:
    m = srwn.ParallelWaveNet(T, 0, dil, None, num_flows=4, skip_channels=128, latent_channels=32, pool_stride=128)
    m.set_weights(synth.make_student_weights(dil, 4))
    x = torch.from_numpy(synth.logistic_noise(B, T)).cuda()
    run = lambda: m.generate(None, x, enc, precision=prec)
    op = _lib.OP_STUDENT_FORWARD
m._eng.set_profiling(True)
m._eng.set_team_size(team)
ms = []
for i in range(iters):
    r = run()
    ms.append(m._eng.last_kernel_ms()[0])
m._eng.check_async(op, B, T, _lib.PRECISIONS[prec])
best = min(ms)
print("%s %dx%d %s team %s: kernel ms %s  best %.3f ms -> %.1f Msamples/s (per launch)" % (
    what, B, T, prec, m._eng.last_partition(), ["%.3f" % v for v in ms], best, B * T / best / 1e3))
