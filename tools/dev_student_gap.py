"""Step time of student synthesis through the public call (CUDA events around generate), e.g. to compare builds of the
conditioning fold: SRWN_LIB=tools/exp/libsrwn_<x>.so python tools/dev_student_gap.py 64x64000"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sr_wavenet_b200 as srwn
from sr_wavenet_b200 import synth
B, T = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "64x64000").split("x"))
dil = synth.DEFAULT_DILATIONS
m = srwn.ParallelWaveNet(T, 0, dil, None, num_flows=4, skip_channels=128, latent_channels=32, pool_stride=128)
m.set_weights(synth.make_student_weights(dil, 4))
enc = torch.from_numpy(synth.synthetic_encoding(B, T // 128)).cuda()
x = torch.from_numpy(synth.logistic_noise(B, T)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(2): m.generate(None, x, enc, precision="fp16")
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(6)]
for a, b in ev:
    flush.fill_(1)
    a.record(); m.generate(None, x, enc, precision="fp16"); b.record()
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in ev]
print("student %dx%d: step ms %s  best %.3f" % (B, T, ["%.3f" % v for v in ms], min(ms)), flush=True)
