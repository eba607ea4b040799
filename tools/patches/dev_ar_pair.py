"""Generation kernel, single-CTA form against the CTA-pair form (SRWN_AR_PAIR=0/1): bit-identical samples, kernel time.
python tools/dev_ar_pair.py 256x1024"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sr_wavenet_b200 as srwn
from sr_wavenet_b200 import synth
B, T = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256x1024").split("x"))
dil = synth.DEFAULT_DILATIONS
m = srwn.WaveNetAutoEncoder(T, 0, 5, dil, skip_channels=128, latent_channels=32, pool_stride=128)
m.set_weights(synth.make_teacher_weights(dil))
enc = torch.from_numpy(synth.synthetic_encoding(B, T // 128)).cuda()
u1, u2 = (torch.from_numpy(a).cuda() for a in synth.sampler_uniforms(B, T))
m._eng.set_profiling(True)
out = {}
for mode in ("0", "1", "0", "1"):
    os.environ["SRWN_AR_PAIR"] = mode
    x = m.generate(enc, u1=u1, u2=u2, precision="fp16")
    ms = m._eng.last_kernel_ms()[0]
    x = x.cpu().numpy() if torch.is_tensor(x) else np.asarray(x)
    print("pair=%s %dx%d: kernel %.3f ms -> %.2f Msamples/s, %.3f us/step" % (mode, B, T, ms, B * T / ms / 1e3, ms * 1e3 / T), flush=True)
    if mode in out:
        assert np.array_equal(out[mode], x), "not repeatable"
    out[mode] = x
print("bit-identical:", np.array_equal(out["0"], out["1"]), "max|d| = %.3g" % np.abs(out["0"] - out["1"]).max())
