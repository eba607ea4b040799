"""Diagnostic: run-to-run repeatability of the distillation gradient per variable class, over sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sr_wavenet_b200 as srwn
from sr_wavenet_b200 import synth
DIL = synth.DEFAULT_DILATIONS
M = 5
for B, T in [(1, 1024), (1, 8192), (2, 8192), (1, 64000), (3, 64000), (4, 64000)]:
    s = srwn.ParallelWaveNet(input_size=T, condition_size=0, dilations=DIL, teacher=None, num_flows=4, skip_channels=128,
                             latent_channels=32, pool_stride=128, alpha=0.25, beta=1.0, gamma=1.0)
    w = synth.make_student_weights(DIL, 4)
    s.set_weights(w)
    rng = np.random.default_rng(5)
    z = synth.logistic_noise(B, T); truth = synth.synthetic_audio(B, T); enc = synth.synthetic_encoding(B, T // 128)
    tl = (rng.normal(0, 1, size=(B, T, 4 * M)) * 0.5).astype(np.float32)
    gs = []
    for _ in range(3):
        _, _, _, g = s.loss_and_grads(z, truth, enc, teacher_logits=tl); gs.append(g.clone())
    cls = {}
    for name in w:
        try: a, b, c = (s.grad_of(x, name) for x in gs)
        except Exception: continue
        if a.numel() == 0: continue
        i = int(name.split("conv1d_")[1].split("/")[0]) if "conv1d_" in name else -1
        k = ("res" if i % 3 == 1 else "skip" if i % 3 == 2 else "cond/head") + "/" + name.split("/")[-1] if i >= 0 else name.split("/")[-1]      # student: conv1d_{3i} cond, _{3i+1} res, _{3i+2} skip
        d = max((a - b).abs().max().item(), (a - c).abs().max().item()) / max(a.abs().max().item(), 1e-30)
        cls[k] = max(cls.get(k, 0.0), d)
    print(B, T, {k: "%.1e" % v for k, v in cls.items() if v > 0} or "bit-identical")
