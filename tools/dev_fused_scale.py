"""Scale-up check of the fused kernel: timing + abort flag + error vs the fp32 GPU path."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sr_wavenet_b200 as srwn
from sr_wavenet_b200 import synth, _lib
dil = synth.DEFAULT_DILATIONS
cases = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(1, 64000), (4, 64000), (32, 64000)]
tw = synth.make_teacher_weights(dil)
for B, T in cases:
    t = srwn.WaveNetAutoEncoder(T, 0, 5, dil, skip_channels=128, latent_channels=32, pool_stride=128)
    t.set_weights(tw)
    x = torch.from_numpy(synth.synthetic_audio(B, T)).cuda(); enc = torch.from_numpy(synth.synthetic_encoding(B, T // 128)).cuda()
    ref = t.get_logits(x, enc, precision="fp32")
    for prec in ("fp16",):
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.time()
            lg = t.get_logits(x, enc, precision=prec)
            try:
                t._eng.check_async(_lib.OP_TEACHER_LOGITS, B, T, _lib.PRECISIONS[prec]); msg = "ok"
            except Exception as e:
                msg = str(e)
            dt = time.time() - t0
        d = (lg - ref).abs()
        print("B=%d T=%d %s: %.2f ms  %.1f Msamples/s  max|d|=%.4f  %s" % (B, T, prec, dt * 1e3, B * T / dt / 1e6, float(d.max()), msg), flush=True)
        if float(d.max()) > 0.05:
            bad = (d.amax(dim=2) > 0.05).nonzero()
            print("   bad count", bad.shape[0], "first", bad[:5].tolist(), "last", bad[-3:].tolist())
