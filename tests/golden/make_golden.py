"""Generates tests/golden/*.npz from the NumPy oracle (float64) on seeded inputs.

The reference (TF 1.x) cannot run in this environment, so these vectors pin the ORACLE (and,
through the -m gpu tests, the CUDA path) against regressions; they are not outputs of the
reference itself ("parity unpinned", see oracle/__init__.py).  The only reference-held
known answers (ops.py:243-254) are in conv_kat.npz, typed in from the formulas at ops.py:6-10
as listed in SURVEY.md 8(c).

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import sr_wavenet_b200.synth as synth  # noqa: E402
from oracle import srwn_oracle as orc  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def f64(d):
    return {k: v.astype(np.float64) for k, v in d.items()}


def conv_kat():
    """ops.py:243-254 (x = 1..8)."""
    np.savez(os.path.join(OUT, "conv_kat.npz"),
             x=np.arange(1, 9, dtype=np.float32),
             l243=np.array([1, 3, 5, 7, 9, 11, 13, 15], np.float32),      # f=[1,1], d=1
             l244=np.array([1, 2, 4, 6, 8, 10, 12, 14], np.float32),      # f=[1,0,1]
             l245=np.array([1, 2, 3, 4, 6, 8, 10, 12], np.float32),       # f=[1,0,0,0,1]
             l246=np.array([1, 2, 4, 6, 8, 10, 12, 14], np.float32),      # f=[1,1], d=2
             l247=np.array([1, 2, 3, 5, 7, 9, 11, 13], np.float32),       # d=3
             l248=np.array([1, 2, 3, 4, 6, 8, 10, 12], np.float32),       # d=4
             l249=np.array([1, 2, 3, 4, 5, 6, 8, 10], np.float32),        # d=6
             l252=np.array([[1, 2], [3, 6], [5, 10], [7, 14], [9, 18], [11, 22], [13, 26], [15, 30]], np.float32),
             l254=np.array([[3, 6], [5, 10], [7, 14], [9, 18], [11, 22], [13, 26], [15, 30]], np.float32))


def small():
    """6 layers, T=64, P=16, 8 latent channels: teacher logits/NLL/sample/AR + 2-flow student."""
    dil = [1, 2, 4, 1, 2, 4]
    B, T, P, C, M, F = 2, 64, 16, 8, 5, 2
    tw = synth.make_teacher_weights(dil, latent_channels=C, num_mixtures=M, seed=11)
    sw = synth.make_student_weights(dil, num_flows=F, latent_channels=C, seed=12)
    x = synth.synthetic_audio(B, T, seed=5)
    x[0, 10] = 1.0      # exercise the x > 0.999 / x < -0.999 branches of ops.py:167
    x[1, 20] = -1.0
    enc = synth.synthetic_encoding(B, T // P, C, seed=6)
    z = synth.logistic_noise(B, T, seed=7)
    u1, u2 = synth.sampler_uniforms(B, T, M, seed=8)
    x64, e64 = x.astype(np.float64), enc.astype(np.float64)
    logits = orc.teacher_decoder_logits(f64(tw), x64, e64, dil, P)
    nll_t = orc.discretized_mix_logistic_loss(x64[:, :, None], logits, sum_all=False)
    samp, idx = orc.sample_from_discretized_mix_logistic(logits, M, u1.astype(np.float64),
                                                         u2.astype(np.float64)[:, :, None], return_index=True)
    ar_x, ar_logits = orc.queue_ar(f64(tw), e64, dil, P, M, u1.astype(np.float64), u2.astype(np.float64), T,
                                   return_logits=True)
    net = orc.student_network(f64(sw), z.astype(np.float64), e64, dil, P, F)
    np.savez(os.path.join(OUT, "small.npz"), dilations=np.array(dil), P=P, C=C, M=M, F=F,
             teacher_seed=11, student_seed=12, x=x, enc=enc, z=z, u1=u1, u2=u2,
             logits=logits, nll=nll_t, nll_sum=nll_t.sum(), sample=samp, sample_idx=idx,
             ar_x=ar_x, ar_logits=ar_logits, student_out=net["out"], s_tot=net["s_tot"],
             mu_tot=net["mu_tot"], x_last=net["x_last"])


def default_cfg():
    """teacher.py:55-62 hyper-parameters (30 layers, P=128, 32 latent), B=2, T=512."""
    dil = synth.DEFAULT_DILATIONS
    B, T, P = 2, 512, 128
    tw = synth.make_teacher_weights(dil, seed=42)
    sw = synth.make_student_weights(dil, num_flows=4, seed=43)
    x = synth.synthetic_audio(B, T, seed=1234)
    enc = synth.synthetic_encoding(B, T // P, 32, seed=4321)
    z = synth.logistic_noise(B, T, seed=777)
    logits = orc.teacher_decoder_logits(f64(tw), x.astype(np.float64), enc.astype(np.float64), dil, P)
    nll = orc.discretized_mix_logistic_loss(x.astype(np.float64)[:, :, None], logits, True)
    net = orc.student_network(f64(sw), z.astype(np.float64), enc.astype(np.float64), dil, P, 4)
    np.savez(os.path.join(OUT, "default_cfg.npz"), B=B, T=T, P=P, logits=logits.astype(np.float32),
             nll_sum=nll, student_out=net["out"].astype(np.float32),
             s_tot=net["s_tot"].astype(np.float32), mu_tot=net["mu_tot"].astype(np.float32))


def encoder():
    """Teacher encoder (model.py:137-155): a tiny generic shape with a ragged tail (fp32 path only) and
    the teacher.py:55-62 shape (128 channels, P=128, 30 layers), B=2, T=512."""
    L, E, S, C, P, B, T = 3, 16, 8, 4, 8, 2, 43
    w = synth.make_encoder_weights(L, 2, E, S, C, seed=21)
    x = synth.synthetic_audio(B, T, seed=9)
    enc = orc.teacher_encoder(f64(w), x.astype(np.float64), L, P)
    np.savez(os.path.join(OUT, "encoder_small.npz"), L=L, E=E, S=S, C=C, P=P, seed=21, x=x, encoding=enc)
    L, B, T, P = len(synth.DEFAULT_DILATIONS), 2, 512, 128
    w = synth.make_encoder_weights(L, seed=44)
    x = synth.synthetic_audio(B, T, seed=1234)
    enc = orc.teacher_encoder(f64(w), x.astype(np.float64), L, P)
    np.savez(os.path.join(OUT, "encoder_default.npz"), L=L, P=P, seed=44, B=B, T=T, encoding=enc)


if __name__ == "__main__":
    conv_kat()
    small()
    default_cfg()
    encoder()
    print("golden vectors written to", OUT)
