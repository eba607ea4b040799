"""-m gpu: the BASELINE.json configurations at FULL size, through size-independent properties (the oracle cannot run
at these sizes): utterance independence (bit-exact under a permutation of the batch, whatever piece of the work
partition an utterance lands in), causality, sum of per-sample terms == the reduced value, composition identities,
teacher forcing of generated audio, additivity of the gradient over batch shards."""
import numpy as np
import pytest
import torch

from sr_wavenet_b200 import synth

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
    parts, losses = [], []
    for sl in (slice(0, 2), slice(2, 4)):
        l_i, _, _, g_i = s.loss_and_grads(z[sl], truth[sl], enc[sl], teacher_logits=tl[sl], batch_norm=B)
        parts.append(g_i.clone())
        losses.append(float(l_i))
    assert np.isfinite(float(loss)) and torch.isfinite(g).all()
    np.testing.assert_allclose(sum(losses), float(loss), rtol=1e-5)
    scale = g.abs().max().item()
    assert (parts[0] + parts[1] - g).abs().max().item() <= 2e-4 * scale
