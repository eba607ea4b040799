"""-m gpu: the BASELINE.json configurations at FULL size.

(1) Oracle comparisons on the exact bench.py inputs (same seeds): the float64 NumPy oracle on whole utterances that land
    mid-piece in the 148-way work partition, the fp32 CPU port over all 2.048 M positions for the exact mixture-index
    flip count, one full student utterance, two full generated utterances.  Tolerances are the north star's: fp32 path
    <= 1e-4 relative, fp16 tensor-core path <= 2e-2 max-abs on logits / per-sample log-likelihood.
(2) Size-independent properties: utterance independence (bit-exact under a permutation of the batch, whatever piece of
    the work partition an utterance lands in), causality, sum of per-sample terms == the reduced value, composition
    identities, teacher forcing of generated audio, additivity of the gradient over batch shards."""
import numpy as np
import pytest
import torch

from conftest import f64
from oracle import srwn_oracle as orc
from sr_wavenet_b200 import synth
from test_gpu_models import mixture_flips

pytestmark = pytest.mark.gpu

DIL = synth.DEFAULT_DILATIONS


@pytest.fixture(scope="module")
def srwn(lib):
    import sr_wavenet_b200
    assert torch.cuda.is_available()
    return sr_wavenet_b200


def _teacher(srwn, T):
    t = srwn.WaveNetAutoEncoder(input_size=T, condition_size=0, num_mixtures=5, dilations=DIL, skip_channels=128,
                                latent_channels=32, pool_stride=128)
    t.set_weights(synth.make_teacher_weights(DIL))
    return t


def test_teacher_scoring_32x64000(srwn):
    """configs[1]: per-sample NLL is a function of the utterance alone (bit-exact under batch permutation), causal in the
    input, and its sum is the reduced log-likelihood."""
    B, T = 32, 64000
    t = _teacher(srwn, T)
    x = torch.from_numpy(synth.synthetic_audio(B, T)).cuda()
    enc = torch.from_numpy(synth.synthetic_encoding(B, T // 128)).cuda()
    per = t.nll(x, enc, sum_all=False, precision="fp16")[:, :, 0]
    assert torch.isfinite(per).all()
    total = t.nll(x, enc, sum_all=True, precision="fp16")
    np.testing.assert_allclose(float(total), per.double().sum().item(), rtol=1e-6)
    perm = torch.from_numpy(np.random.default_rng(0).permutation(B)).cuda()
    per_p = t.nll(x[perm].contiguous(), enc[perm].contiguous(), sum_all=False, precision="fp16")[:, :, 0]
    assert torch.equal(per_p, per[perm])
    # causality (RightShift + causal convs): samples from t0 on do not reach the likelihood of samples before t0
    t0 = 40000
    x2 = x.clone()
    x2[:, t0:] = -x2[:, t0:]
    per2 = t.nll(x2, enc, sum_all=False, precision="fp16")[:, :, 0]
    assert torch.equal(per2[:, :t0], per[:, :t0]) and not torch.equal(per2[:, t0:], per[:, t0:])


def test_student_synthesis_8x64000(srwn):
    """configs[2] (per-GPU share): out == clip(z * s_tot + mu_tot) (model.py:535) and utterance independence."""
    B, T = 8, 64000
    s = srwn.ParallelWaveNet(input_size=T, condition_size=0, dilations=DIL, teacher=None, num_flows=4, skip_channels=128,
                             latent_channels=32, pool_stride=128)
    s.set_weights(synth.make_student_weights(DIL, 4))
    z = torch.from_numpy(synth.logistic_noise(B, T)).cuda()
    enc = torch.from_numpy(synth.synthetic_encoding(B, T // 128)).cuda()
    r = s.forward_all(z, enc, precision="fp16")
    assert all(torch.isfinite(v).all() for v in r.values())
    ref = torch.clamp(z * r["s_tot"] + r["mu_tot"], -1.0, 1.0)
    assert (r["out"] - ref).abs().max().item() <= 1e-5
    assert (r["s_tot"] > 0).all()
    perm = torch.from_numpy(np.random.default_rng(1).permutation(B)).cuda()
    rp = s.forward_all(z[perm].contiguous(), enc[perm].contiguous(), precision="fp16")
    assert torch.equal(rp["out"], r["out"][perm]) and torch.equal(rp["s_tot"], r["s_tot"][perm])


def test_generation_256x16000(srwn):
    """configs[3]: teacher-forcing the generated audio reproduces the logits the generator saw (fp16 operand bound) and the
    generated samples are the mixture samples of those logits (ops.py:178-201); utterances are independent."""
    B, T, M = 256, 16000, 5
    t = _teacher(srwn, T)
    enc = torch.from_numpy(synth.synthetic_encoding(B, T // 128)).cuda()
    u1, u2 = (torch.from_numpy(a).cuda() for a in synth.sampler_uniforms(B, T))
    x, lg = t.generate(enc, u1=u1, u2=u2, return_logits=True, precision="fp16")
    assert torch.isfinite(x).all() and x.abs().max().item() <= 1.0
    sub = slice(0, 16)                                            # 16 utterances through the scoring kernel
    tf = t.get_logits(x[sub].contiguous(), enc[sub].contiguous(), precision="fp16")
    assert (tf - lg[sub]).abs().max().item() <= 2e-2
    # the samples are ops.py:178-201 applied to the generator's own logits with the injected noise
    k = torch.argmax(lg[..., :M] - torch.log(-torch.log(u1)), dim=-1, keepdim=True)
    mean = torch.gather(lg[..., M:2 * M], 2, k)[..., 0]
    ls = torch.clamp(torch.gather(lg[..., 2 * M:3 * M], 2, k)[..., 0], min=-7.0)
    xs = torch.clamp(mean + torch.exp(ls) * (torch.log(u2) - torch.log(1.0 - u2)), -1.0, 1.0)
    assert (xs - x).abs().max().item() <= 1e-5
    # one CTA owns 8 utterances: moving an utterance to another CTA / row does not change it
    perm = torch.from_numpy(np.random.default_rng(2).permutation(B)).cuda()
    xp = t.generate(enc[perm].contiguous(), u1=u1[perm].contiguous(), u2=u2[perm].contiguous(), precision="fp16")
    assert torch.equal(xp, x[perm])


def test_encoder_32x64000(srwn):
    """encodings depend on the utterance alone and only look one pooling window plus the layer count ahead."""
    B, T = 32, 64000
    t = _teacher(srwn, T)
    x = torch.from_numpy(synth.synthetic_audio(B, T)).cuda()
    e = t.encode(x, precision="fp16")
    assert e.shape == (B, T // 128, 32) and torch.isfinite(e).all()
    perm = torch.from_numpy(np.random.default_rng(3).permutation(B)).cuda()
    assert torch.equal(t.encode(x[perm].contiguous(), precision="fp16"), e[perm])
    x2 = x.clone()
    x2[:, 32000:] = 0.0                                           # frame f covers samples 128 f .. 128 f + 127 (+31 of look-ahead)
    e2 = t.encode(x2, precision="fp16")
    f0 = (32000 - 32) // 128
    assert torch.equal(e2[:, :f0], e[:, :f0])


def test_distillation_gradient_is_additive_over_shards_4x64000(srwn):
    """configs[4]: with the loss normalised by the global batch (model.py:379), the gradients of two batch shards add up to
    the full-batch gradient -- what the NCCL all-reduce of train_fast relies on."""
    B, T, M = 4, 64000, 5
    s = srwn.ParallelWaveNet(input_size=T, condition_size=0, dilations=DIL, teacher=None, num_flows=4, skip_channels=128,
                             latent_channels=32, pool_stride=128, alpha=0.25, beta=1.0, gamma=1.0)
    s.set_weights(synth.make_student_weights(DIL, 4))
    rng = np.random.default_rng(5)
    z = synth.logistic_noise(B, T)
    truth = synth.synthetic_audio(B, T)
    enc = synth.synthetic_encoding(B, T // 128)
    tl = (rng.normal(0, 1, size=(B, T, 4 * M)) * 0.5).astype(np.float32)
    loss, _, ent, g = s.loss_and_grads(z, truth, enc, teacher_logits=tl)
    g = g.clone()
    # every reduction of the backward pass has a fixed order (the conditioning gradient included: one add per frame and tile)
    _, _, _, g_again = s.loss_and_grads(z, truth, enc, teacher_logits=tl)
    assert torch.equal(g, g_again), "distillation gradient is not reproducible run to run"
    parts, losses = [], []
    for sl in (slice(0, 2), slice(2, 4)):
        l_i, _, _, g_i = s.loss_and_grads(z[sl], truth[sl], enc[sl], teacher_logits=tl[sl], batch_norm=B)
        parts.append(g_i.clone())
        losses.append(float(l_i))
    assert np.isfinite(float(loss)) and torch.isfinite(g).all()
    np.testing.assert_allclose(sum(losses), float(loss), rtol=1e-5)
    scale = g.abs().max().item()
    assert (parts[0] + parts[1] - g).abs().max().item() <= 2e-4 * scale


# ---- (1) oracle comparisons on the bench.py inputs -----------------------------------------------------------------
def _bench_inputs(B, T):
    """bench.py rank 0: synthetic_audio(B, T, seed=1234), synthetic_encoding(B, T // 128, seed=4321)."""
    return synth.synthetic_audio(B, T, seed=1234), synth.synthetic_encoding(B, T // 128, seed=4321)


def test_teacher_bench_inputs_vs_oracle_32x64000(srwn):
    """configs[1] as bench.py runs it (32 x 64000, fp16 operands, 148 pieces with mid-utterance warm-up) and configs[0]
    (1 x 64000): logits and per-sample NLL of whole utterances against the float64 oracle."""
    B, T = 32, 64000
    t = _teacher(srwn, T)
    w = f64(synth.make_teacher_weights(DIL))
    x, enc = _bench_inputs(B, T)
    xd, ed = torch.from_numpy(x).cuda(), torch.from_numpy(enc).cuda()
    lg16 = t.get_logits(xd, ed, precision="fp16")
    nll16 = t.nll(xd, ed, sum_all=False, precision="fp16")[:, :, 0]
    t._eng.check_async(srwn._lib.OP_TEACHER_NLL, B, T, srwn._lib.FP16)
    pick = [0, 5, 22]                  # utterance 0 starts a piece; 5 and 22 start and end inside pieces
    ref = orc.teacher_decoder_logits(w, x[pick].astype(np.float64), enc[pick].astype(np.float64), DIL, 128)
    ref_nll = orc.discretized_mix_logistic_loss(x[pick].astype(np.float64)[:, :, None], ref, False)[:, :, 0]
    e_lg = np.abs(lg16[pick].cpu().numpy() - ref).max()
    e_nll = np.abs(nll16[pick].cpu().numpy() - ref_nll).max()
    print("teacher fp16 32x64000 vs oracle on utterances %s: max|dlogits| %.3e  max|dnll| %.3e" % (pick, e_lg, e_nll))
    assert e_lg <= 1e-2 and e_nll <= 2e-2
    # fp32-grade path on the same utterances, and configs[0] (B = 1) through both paths
    lg32 = t.get_logits(x[pick], enc[pick], precision="fp32")
    rel = np.abs(lg32 - ref).max() / np.abs(ref).max()
    print("teacher fp32 3x64000 vs oracle: max rel %.3e" % rel)
    assert rel <= 1e-4
    one16 = t.get_logits(x[:1], enc[:1], precision="fp16")
    one32 = t.get_logits(x[:1], enc[:1], precision="fp32")
    assert np.abs(one16 - ref[:1]).max() <= 1e-2
    assert np.abs(one32 - ref[:1]).max() <= 1e-4 * np.abs(ref[:1]).max()
    nll_one = t.nll(x[:1], enc[:1], precision="fp32")
    assert abs(nll_one - ref_nll[0].sum()) <= 1e-4 * abs(ref_nll[0].sum())


def test_teacher_mixture_flip_count_2M_positions(srwn):
    """'Identical argmax sample indices under teacher forcing' over all 2.048 M positions of the bench batch: the exact
    number of mixture-index flips of the fp16 path against the CPU restatement (fp32 PyTorch port of the same graph, its
    own error ~1e-5), every flip lying where the reference's top-two perturbed logits are closer than twice the measured
    error (mixture_flips asserts that)."""
    from oracle.torch_cpu import TeacherCPU
    B, T = 32, 64000
    t = _teacher(srwn, T)
    x, enc = _bench_inputs(B, T)
    lg16 = t.get_logits(x, enc, precision="fp16")
    cpu = TeacherCPU(synth.make_teacher_weights(DIL), DIL, 128, 5)
    ref = np.concatenate([np.asarray(cpu.logits(x[i:i + 4], enc[i:i + 4]), np.float64) for i in range(0, B, 4)])
    u1, u2 = synth.sampler_uniforms(B, T)
    flips, reach = mixture_flips(srwn, ref, lg16, u1, u2)
    err = np.abs(lg16 - ref).max()
    print("teacher fp16 32x64000: max|dlogits| vs CPU port %.3e; %d mixture-index flips of %d positions (%.4f %%), "
          "%d positions within reach of the measured error" % (err, flips, B * T, 100.0 * flips / (B * T), reach))
    assert err <= 1e-2
    assert flips <= 2e-3 * B * T


def test_student_bench_inputs_vs_oracle_8x64000(srwn):
    """configs[2] per-GPU share as bench.py runs it: one whole utterance of the 8 x 64000 batch against the oracle."""
    B, T = 8, 64000
    s = srwn.ParallelWaveNet(input_size=T, condition_size=0, dilations=DIL, teacher=None, num_flows=4, skip_channels=128,
                             latent_channels=32, pool_stride=128)
    sw = synth.make_student_weights(DIL, 4)
    s.set_weights(sw)
    z, enc = synth.logistic_noise(B, T, seed=777), synth.synthetic_encoding(B, T // 128, seed=4321)
    r16 = s.forward_all(z, enc, precision="fp16")
    s._eng.check_async(srwn._lib.OP_STUDENT_FORWARD, B, T, srwn._lib.FP16)
    b = 3
    net = orc.student_network(f64(sw), z[b:b + 1].astype(np.float64), enc[b:b + 1].astype(np.float64), DIL, 128, 4)
    e16 = np.abs(r16["out"][b] - net["out"][0, :, 0]).max()
    s16 = np.abs(r16["s_tot"][b] / net["s_tot"][0, :, 0] - 1).max()
    r32 = s.forward_all(z[b:b + 1], enc[b:b + 1], precision="fp32")
    e32 = np.abs(r32["out"][0] - net["out"][0, :, 0]).max()
    s32 = np.abs(r32["s_tot"][0] / net["s_tot"][0, :, 0] - 1).max()
    print("student 8x64000 utterance %d vs oracle: fp16 max|dout| %.3e rel s_tot %.3e; fp32 %.3e / %.3e" % (b, e16, s16, e32, s32))
    assert e16 <= 2e-2 and s16 <= 4e-2
    assert e32 <= 1e-4 and s32 <= 1e-4


def test_generation_vs_oracle_256x16000(srwn):
    """configs[3]: the tensor-core generator at 256 x 16000; for two whole utterances the oracle, teacher-forced on the
    generated audio, reproduces the logits the generator used (<= 2e-2) and the audio is ops.py:178-201 applied to the
    oracle's logits with the same noise, except at counted mixture-index flips within reach of the measured error."""
    B, T, M = 256, 16000, 5
    t = _teacher(srwn, T)
    w = f64(synth.make_teacher_weights(DIL))
    enc = synth.synthetic_encoding(B, T // 128, seed=4321)
    u1, u2 = synth.sampler_uniforms(B, T, seed=999)
    x, lg = t.generate(enc, u1=u1, u2=u2, return_logits=True, precision="fp16")
    pick = [7, 130]
    ref = orc.teacher_decoder_logits(w, x[pick].astype(np.float64), enc[pick].astype(np.float64), DIL, 128)
    err = np.abs(lg[pick] - ref).max()
    flips, reach = mixture_flips(srwn, ref, lg[pick], u1[pick], u2[pick])
    xs, k = orc.sample_from_discretized_mix_logistic(ref, M, u1[pick].astype(np.float64), u2[pick].astype(np.float64)[:, :, None], True)
    _, kg = srwn.ops.sample_from_discretized_mix_logistic(lg[pick], M, u1[pick], u2[pick], return_index=True)
    same = k == (kg.cpu().numpy() if hasattr(kg, "cpu") else kg)
    ex = np.abs(xs[:, :, 0] - x[pick])[same].max()
    print("generation 256x16000 utterances %s vs oracle: max|dlogits| %.3e, max|dx| off-flip %.3e, %d flips of %d (%d within reach)"
          % (pick, err, ex, flips, len(pick) * T, reach))
    assert err <= 2e-2 and ex <= 2e-2
