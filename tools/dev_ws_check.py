"""Diagnostic: reads the training workspace after a backward pass and checks dcond_0 = per-frame sums of dx_0 and
dWc_0 = enc^T dcond_0 (flow 0, the last one the backward pass visits)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sr_wavenet_b200 as srwn
from sr_wavenet_b200 import synth
dil, F, C, P, M, B, T = [1, 2, 4, 3], 2, 8, 64, 3, 3, 832
L, R = len(dil), 32
s = srwn.ParallelWaveNet(input_size=T, condition_size=0, dilations=dil, teacher=None, num_flows=F, skip_channels=128,
                         latent_channels=C, pool_stride=P, alpha=0.25, beta=1.0, gamma=1.0)
w = synth.make_student_weights(dil, F, latent_channels=C)
s.set_weights(w)
rng = np.random.default_rng(5)
z = synth.logistic_noise(B, T); truth = synth.synthetic_audio(B, T)
enc = rng.normal(0, 1, size=(B, T // P, C)).astype(np.float32)
tl = (rng.normal(0, 1, size=(B, T, 4 * M)) * 0.5).astype(np.float32)
_, _, _, g = s.loss_and_grads(z, truth, enc, teacher_logits=tl)
torch.cuda.synchronize()
ws = s._eng._ws
def rnd(nbytes): return (nbytes + 255) & ~255
n, frames = B * T, T // P
sizes = [F * (L + 1) * n * R, B * frames * L * R, F * n, F * n, F * n, n * R, n * R, n * R, L * B * frames * R]
names = ["acts", "cond", "scales", "means", "xs", "g0", "g1", "da", "dcond"]
off, view = 0, {}
for nm, cnt in zip(names, sizes):
    view[nm] = ws[off:off + cnt * 4].view(torch.float32)
    off += rnd(cnt * 4)
dx0 = view["g0"].view(B, T, R).double()                       # L even: the last dx lands in g0
dcond0 = view["dcond"].view(L, B * frames, R)[0].double()
ref = dx0.view(B, frames, P, R).sum(2).view(B * frames, R)
print("dcond_0 vs per-frame sums of dx_0: max diff %.3e (max %.3e)" % ((dcond0 - ref).abs().max().item(), ref.abs().max().item()))
e = torch.from_numpy(enc).cuda().double().view(B * frames, C)
dWc = (e.t() @ dcond0)                                         # [C][R]
name = "ParallelWaveNet/Flow0/Flow0/conv1d_1/kernel"
got = s.grad_of(g, name).view(-1, R).double(); print("library dWc_0 shape", tuple(got.shape)); got = got[:C]
print("dWc_0 (library) vs enc^T dcond_0 (from the workspace): max diff %.3e (max %.3e)" % ((got - dWc).abs().max().item(), dWc.abs().max().item()))
print("dWc_0 (library) vs enc^T (sums of dx_0): max diff %.3e" % ((got - e.t() @ ref).abs().max().item()))
