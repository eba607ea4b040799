"""Bring-up check of the fused tcgen05 kernel against the fp32 GPU path and the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sr_wavenet_b200 as srwn
from sr_wavenet_b200 import synth, _lib
from oracle import srwn_oracle as orc

dil = synth.DEFAULT_DILATIONS
cases = [(1, 384), (1, 1152), (2, 4096), (3, 12800)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
tw = synth.make_teacher_weights(dil)
for B, T in cases:
    t = srwn.WaveNetAutoEncoder(T, 0, 5, dil, skip_channels=128, latent_channels=32, pool_stride=128)
    t.set_weights(tw)
    x, enc = synth.synthetic_audio(B, T), synth.synthetic_encoding(B, T // 128)
    ref = t.get_logits(x, enc, precision="fp32")
    for prec in ("fp16", "bf16"):
        t0 = time.time()
        try:
            lg = t.get_logits(x, enc, precision=prec)
            t._eng.check_async(_lib.OP_TEACHER_LOGITS, B, T, _lib.PRECISIONS[prec])
        except Exception as e:
            print("B=%d T=%d %s: ERROR %s" % (B, T, prec, e)); continue
        d = np.abs(lg - ref)
        bad = np.argwhere(d > 0.05)
        print("B=%d T=%d %s: max|d|=%.4f mean=%.5f nan=%d (%.2fs) first bad=%s" % (
            B, T, prec, np.nanmax(d), np.nanmean(d), int(np.isnan(lg).sum()), time.time() - t0,
            bad[:3].tolist()))
        if d.max() > 0.05:
            tt = np.unique(bad[:, 1]); print("   bad t range:", tt[:8], "...", tt[-8:], "count", len(tt))
        nl = t.nll(x, enc, precision=prec); nr = t.nll(x, enc, precision="fp32")
        print("   nll %s=%.3f fp32=%.3f rel=%.2e" % (prec, nl, nr, abs(nl - nr) / abs(nr)))
sw = synth.make_student_weights(dil, 4)
for B, T in cases[:3]:
    s = srwn.ParallelWaveNet(T, 0, dil, None, num_flows=4, skip_channels=128, latent_channels=32, pool_stride=128)
    s.set_weights(sw)
    z, enc = synth.logistic_noise(B, T), synth.synthetic_encoding(B, T // 128)
    ref = s.forward_all(z, enc, precision="fp32")
    for prec in ("fp16", "bf16"):
        try:
            r = s.forward_all(z, enc, precision=prec)
            s._eng.check_async(_lib.OP_STUDENT_FORWARD, B, T, _lib.PRECISIONS[prec])
        except Exception as e:
            print("student B=%d T=%d %s: ERROR %s" % (B, T, prec, e)); continue
        print("student B=%d T=%d %s: max|dout|=%.4f max|ds_tot/s_tot|=%.4f max|dx_last|=%.4f" % (
            B, T, prec, np.abs(r["out"] - ref["out"]).max(), np.abs(r["s_tot"] / ref["s_tot"] - 1).max(),
            np.abs(r["x_last"] - ref["x_last"]).max()))
print("done")
