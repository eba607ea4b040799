"""-m gpu: the distillation step (model.py:356-401) through the C ABI vs the torch float64 gradient oracle."""
import numpy as np
import pytest
import torch

from oracle import distill_torch as dt
from sr_wavenet_b200 import synth

pytestmark = pytest.mark.gpu
GRAD_TOL = 1e-4      # fp32-grade backward (3xTF32 GEMMs, ex2-based gate recompute): the north star's bound for the fp32 path


@pytest.fixture(scope="module")
def srwn(lib):
    import sr_wavenet_b200
    assert torch.cuda.is_available()
    return sr_wavenet_b200


def _student(srwn, dil, F, C, P, seed=11, lr=1e-3, **kw):
    s = srwn.ParallelWaveNet(input_size=0, condition_size=0, dilations=dil, teacher=None, num_flows=F,
                             skip_channels=128, latent_channels=C, pool_stride=P, learning_rate=lr, **kw)
    w = synth.make_student_weights(dil, num_flows=F, latent_channels=C, seed=seed)
    s.set_weights(w)
    return s, w


def _inputs(B, T, P, C, M, seed=5):
    rng = np.random.default_rng(seed)
    z = rng.logistic(0, 1, size=(B, T)).astype(np.float32)
    truth = synth.synthetic_audio(B, T)
    enc = rng.normal(0, 1, size=(B, T // P, C)).astype(np.float32)
    tl = (rng.normal(0, 1, size=(B, T, 4 * M)) * 0.5).astype(np.float32)
    return z, truth, enc, tl


def test_mol_loss_grad_matches_autograd(srwn):
    lib = srwn._lib.load()
    rng = np.random.default_rng(3)
    B, T, M = 2, 512, 5
    l = (rng.normal(0, 1.5, size=(B, T, 4 * M))).astype(np.float32)
    l[:, :, 2 * M:3 * M] -= 3.0                 # narrow components: exercises every branch of ops.py:150-167
    x = np.clip(rng.normal(0, 0.7, size=(B, T)), -1, 1).astype(np.float32)
    x[0, :8] = -1.0
    x[1, :8] = 1.0
    xt = torch.tensor(x.astype(np.float64), requires_grad=True)
    nll = dt.mol_nll(xt, torch.tensor(l.astype(np.float64)), M)
    g_ref, = torch.autograd.grad(nll, xt)
    xd, ld = torch.from_numpy(x).cuda(), torch.from_numpy(l).cuda()
    dx, out = torch.empty_like(xd), torch.empty_like(xd)
    srwn._lib.check(lib.srwn_mol_loss_grad(xd.data_ptr(), ld.data_ptr(), dx.data_ptr(), out.data_ptr(), B, T, M,
                                           torch.cuda.current_stream().cuda_stream))
    np.testing.assert_allclose(out.double().sum().item(), float(nll), rtol=1e-5)
    g = dx.cpu().numpy().astype(np.float64)
    scale = np.abs(g_ref.numpy()).max()
    assert np.abs(g - g_ref.numpy()).max() <= 2e-4 * scale


def _stft_ws(srwn, lib, B, T, N, step):
    import ctypes
    n = ctypes.c_size_t()
    srwn._lib.check(lib.srwn_stft_workspace_bytes(B, T, N, step, ctypes.byref(n)))
    return torch.empty(n.value, dtype=torch.uint8, device="cuda")


@pytest.mark.parametrize("B,T,N,step", [(2, 1280, 512, 256), (3, 1000, 512, 256), (1, 64000, 512, 256), (2, 700, 64, 48),
                                         (1, 4096, 2048, 512)])
def test_stft_power_and_loss_match_oracle(srwn, B, T, N, step):
    """model.py:360-371 on the device vs the NumPy oracle (value) and torch float64 autograd (gradient)."""
    from oracle import srwn_oracle as orc
    lib = srwn._lib.load()
    st = torch.cuda.current_stream().cuda_stream
    truth = synth.synthetic_audio(B, T)
    rng = np.random.default_rng(17)
    out = np.clip(truth * 0.8 + rng.normal(0, 0.1, size=truth.shape), -1, 1).astype(np.float32)
    ws = _stft_ws(srwn, lib, B, T, N, step)
    xd, od = torch.from_numpy(truth).cuda(), torch.from_numpy(out).cuda()
    K = N // 2 + 1
    pw = torch.empty(B, K, dtype=torch.float32, device="cuda")
    srwn._lib.check(lib.srwn_stft_power(xd.data_ptr(), pw.data_ptr(), B, T, N, step, ws.data_ptr(), ws.numel(), st))
    ref = orc.stft_power(truth.astype(np.float64), N, step)
    np.testing.assert_allclose(pw.cpu().numpy(), ref, rtol=2e-5, atol=2e-6 * ref.max())
    gamma = 0.7
    loss = torch.empty(1, dtype=torch.float64, device="cuda")
    g = torch.empty_like(od)
    srwn._lib.check(lib.srwn_stft_power_loss(xd.data_ptr(), od.data_ptr(), gamma, loss.data_ptr(), g.data_ptr(), B, T, N, step,
                                             ws.data_ptr(), ws.numel(), st))
    ot = torch.tensor(out.astype(np.float64), requires_grad=True)
    win = torch.hann_window(N, periodic=True, dtype=torch.float64)

    def power(sig):
        spec = torch.stft(sig, n_fft=N, hop_length=step, win_length=N, window=win, center=False, return_complex=True)
        return (spec.real ** 2 + spec.imag ** 2).mean(dim=2)
    ref_loss = gamma * ((power(torch.tensor(truth.astype(np.float64))) - power(ot)) ** 2).sum()
    g_ref, = torch.autograd.grad(ref_loss, ot)
    np.testing.assert_allclose(float(ref_loss), gamma * np.sum((ref - orc.stft_power(out.astype(np.float64), N, step)) ** 2),
                               rtol=1e-9)
    np.testing.assert_allclose(loss.item(), float(ref_loss), rtol=2e-4)
    scale = np.abs(g_ref.numpy()).max()
    assert np.abs(g.cpu().numpy() - g_ref.numpy()).max() <= 2e-4 * scale
    # samples past the last full frame do not enter the loss (pad_end=False)
    F = 1 + (T - N) // step
    assert not g[:, (F - 1) * step + N:].any()
    # value-only call, and determinism
    loss2 = torch.empty(1, dtype=torch.float64, device="cuda")
    srwn._lib.check(lib.srwn_stft_power_loss(xd.data_ptr(), od.data_ptr(), gamma, loss2.data_ptr(), None, B, T, N, step,
                                             ws.data_ptr(), ws.numel(), st))
    assert loss2.item() == loss.item()


def test_distill_loss_glue(srwn):
    lib = srwn._lib.load()
    rng = np.random.default_rng(23)
    B, T = 3, 5000
    z = rng.logistic(0, 1, size=(B, T)).astype(np.float32)
    s = np.exp(rng.normal(-1, 0.5, size=(B, T))).astype(np.float32)
    mu = rng.normal(0, 0.3, size=(B, T)).astype(np.float32)
    nll, dce, dpow = (rng.normal(0, 1, size=(B, T)).astype(np.float32) for _ in range(3))
    dev = [torch.from_numpy(a).cuda() for a in (z, s, mu, nll, dce, dpow)]
    d_pre, d_s = torch.empty_like(dev[0]), torch.empty_like(dev[0])
    sums = torch.empty(srwn._lib.DISTILL_SUMS_LEN, dtype=torch.float64, device="cuda")
    alpha, beta, inv = 0.25, 1.5, 1.0 / 6
    srwn._lib.check(lib.srwn_distill_loss_grad(*[t.data_ptr() for t in dev], alpha, beta, inv, d_pre.data_ptr(), d_s.data_ptr(),
                                               sums.data_ptr(), B, T, torch.cuda.current_stream().cuda_stream))
    pre = z * s + mu
    mask = (pre >= -1) & (pre <= 1)
    assert 0.05 < mask.mean() < 0.999
    np.testing.assert_allclose(d_pre.cpu().numpy(), (beta * dce + dpow) * mask * np.float32(inv), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(d_s.cpu().numpy(), -(alpha * inv) / s, rtol=2e-6)
    np.testing.assert_allclose(sums[0].item(), nll.astype(np.float64).sum(), rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(sums[1].item(), (np.log(s.astype(np.float64)) + 2).sum(), rtol=1e-6)


@pytest.mark.parametrize("cfg", ["small", "default", "many_tiles", "odd_stride"])
def test_gradients_match_oracle(srwn, cfg):
    if cfg == "small":
        dil, F, C, P, M, B, T = [1, 2, 4, 3], 2, 8, 64, 3, 3, 832     # ragged: T not a multiple of the 128-row tile
    elif cfg == "many_tiles":                                         # 304 tiles: every CTA walks two or three (accumulators,
        dil, F, C, P, M, B, T = [1, 2, 16, 512, 4], 1, 4, 512, 3, 2, 19456     # buffers and barriers live across tiles), 4 tiles per frame
    elif cfg == "odd_stride":                                         # frames that straddle row groups: the per-element path of dcond
        dil, F, C, P, M, B, T = [1, 3], 1, 3, 6, 3, 1, 1026
    else:
        dil, F, C, P, M, B, T = synth.DEFAULT_DILATIONS, 4, 32, 128, 5, 1, 1280
    s, w = _student(srwn, dil, F, C, P, alpha=0.25, beta=1.0, gamma=1.0)
    z, truth, enc, tl = _inputs(B, T, P, C, M)
    loss, power, ent, flat = s.loss_and_grads(z, truth, enc, teacher_logits=tl)
    rl, rp, re, rg = dt.loss_and_grads({k: v.astype(np.float64) for k, v in w.items()}, z, truth, enc, tl, dil, P, F,
                                       alpha=0.25, beta=1.0, gamma=1.0)
    print("distill %s: rel err entropy %.2e  power %.2e  loss %.2e" % (cfg, abs(float(ent) / re - 1), abs(float(power) / rp - 1),
                                                                     abs(float(loss) / rl - 1)))
    np.testing.assert_allclose(float(ent), re, rtol=1e-4)
    np.testing.assert_allclose(float(power), rp, rtol=1e-4)
    np.testing.assert_allclose(float(loss), rl, rtol=1e-4)
    gmax = max(np.abs(v).max() for v in rg.values())
    checked, worst, errs = 0, (0.0, ""), []
    for name, ref in rg.items():
        if "_gate/" in name or not np.any(ref):
            continue                                             # dead variables are not stored
        got = s.grad_of(flat, name).cpu().numpy().reshape(ref.shape)
        rel = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-3 * gmax)
        worst = max(worst, (float(rel), name))
        errs.append((float(rel), name))
        checked += 1
    if worst[0] > GRAD_TOL:
        print("distill %s: variables beyond the tolerance:" % cfg, sorted(errs, reverse=True)[:12])
    print("distill %s: worst gradient error relative to the variable's scale %.2e (%s), %d variables" % (cfg, worst[0], worst[1], checked))
    # fp32 kernels (3xTF32 GEMMs, fp32 loss gradients, fp32 sums over B*T positions) against float64 autograd
    assert worst[0] <= GRAD_TOL, worst
    assert checked == F * (2 + 6 * len(dil) + 2)


def test_adam_step_and_training_reduces_loss(srwn):
    dil, F, C, P, M, B, T = [1, 2, 4, 8], 2, 8, 128, 3, 2, 1024
    s, w = _student(srwn, dil, F, C, P, lr=2e-3, alpha=0.25, beta=1.0, gamma=1.0)
    z, truth, enc, tl = _inputs(B, T, P, C, M)
    # one step against the reference formulas (tf.clip_by_global_norm + AdamOptimizer)
    _, _, _, flat = s.loss_and_grads(z, truth, enc, teacher_logits=tl)
    names = [k for k in w if "_gate/" not in k and not k.endswith(("conv1d_2/kernel", "conv1d_2/bias", "conv1d_5/kernel",
             "conv1d_5/bias", "conv1d_8/kernel", "conv1d_8/bias", "conv1d_11/kernel", "conv1d_11/bias"))]
    g = [s.grad_of(flat, k).cpu().numpy().astype(np.float64).reshape(w[k].shape) for k in names]
    gnorm = float(torch.linalg.vector_norm(flat.double()).item())
    ref = dt.adam_reference([w[k].astype(np.float64) for k in names], g, [np.zeros_like(x) for x in g],
                            [np.zeros_like(x) for x in g], step=1, lr=2e-3, gnorm=gnorm)
    s.apply_gradients(flat)
    s.sync_weights()
    new = s.get_weights()
    for k, (wr, _, _) in zip(names, ref):
        np.testing.assert_allclose(new[k], wr, rtol=0, atol=2e-6)
    # a few more steps: the loss goes down, and the re-packed 16-bit path follows the trained weights
    l0, _ = s.train_fast(None, z, truth, enc, teacher_logits=tl)
    for _ in range(5):
        l1, p1 = s.train_fast(None, z, truth, enc, teacher_logits=tl)
    assert np.isfinite(l1) and l1 < l0
    out32 = s.generate(None, z, enc, precision="fp32")
    if "fp16" in s.available_precisions():
        out16 = s.generate(None, z, enc, precision="fp16")
        assert np.abs(out16 - out32).max() <= 2e-2


def test_per_example_train_matches_reference_semantics(srwn):
    """ParallelWaveNet.train (model.py:603-632): per example a batch-of-one noise row against the whole encoding / truth
    batch, loss / 1, gradients clipped per example, then averaged and applied without further clipping."""
    dil, F, C, P, M, B, T = [1, 2, 4, 8], 2, 8, 128, 3, 2, 1024
    s, w = _student(srwn, dil, F, C, P, lr=1e-3, alpha=0.25, beta=1.0, gamma=1.0)
    z, truth, enc, tl = _inputs(B, T, P, C, M)
    w64 = {k: v.astype(np.float64) for k, v in w.items()}
    names = [k for k in w if "_gate/" not in k and not any(k.endswith("conv1d_%d/%s" % (3 * i + 2, t)) for i in range(len(dil)) for t in ("kernel", "bias"))]
    mean_g, losses, powers = None, [], []
    for i in range(B):
        zi = np.repeat(z[i:i + 1], B, axis=0)
        rl, rp, _, rg = dt.loss_and_grads(w64, zi, truth, enc, tl, dil, P, F, alpha=0.25, beta=1.0, gamma=1.0, batch_norm=1)
        g = [rg[k] for k in names]
        gn = np.sqrt(sum((x ** 2).sum() for x in g))
        g = [x * (1.0 / max(gn, 1.0)) for x in g]                                  # tf.clip_by_global_norm(., 1.0)
        mean_g = g if mean_g is None else [a + b for a, b in zip(mean_g, g)]
        losses.append(rl); powers.append(rp)
    mean_g = [x / B for x in mean_g]
    ref = dt.adam_reference([w64[k] for k in names], mean_g, [np.zeros_like(x) for x in mean_g], [np.zeros_like(x) for x in mean_g],
                            step=1, lr=1e-3, clip=None)
    loss, power = s.train(None, z, truth, enc, teacher_logits=tl)
    np.testing.assert_allclose(loss, np.mean(losses), rtol=1e-4)
    np.testing.assert_allclose(power, np.mean(powers), rtol=1e-4)
    s.sync_weights()
    new = s.get_weights()
    # first Adam step: update = lr * g / (|g| + eps), i.e. +-lr for every entry whose gradient is well above eps = 1e-8.  An
    # entry whose per-example gradients cancel in the average (|g| ~ eps) can land anywhere in [-lr, lr] for a relative
    # gradient error of 1e-5, so a fraction of a percent of the entries may differ by up to 2 lr; all others agree to 5e-6.
    n_bad = n_all = 0
    for k, (wr, _, _) in zip(names, ref):
        diff = np.abs(new[k] - wr)
        assert diff.max() <= 2.1e-3, (k, diff.max())
        n_bad += int((diff > 5e-6).sum())
        n_all += diff.size
    print("per-example train: %d of %d weights differ by more than 5e-6 after one Adam step" % (n_bad, n_all))
    assert n_bad <= 0.005 * n_all


def test_create_flow_matches_oracle(srwn):
    """createFlow / createPartialFlow (model.py:415-486) for one flow of the network."""
    from conftest import f64
    from oracle import srwn_oracle as orc
    dil, F, C, P, B, T = synth.DEFAULT_DILATIONS, 3, 32, 128, 2, 1024
    s, w = _student(srwn, dil, F, C, P)
    x = synth.logistic_noise(B, T, seed=3)
    enc = synth.synthetic_encoding(B, T // P, C, seed=4)
    for f in (0, 2):
        scale, mean, out = s.createFlow(x[:, :, None], enc, 'Flow%d' % f)
        rs, rm, ro = orc.student_flow(f64(w), x.astype(np.float64)[:, :, None], enc.astype(np.float64), dil, P, f)
        assert scale.shape == (B, T, 1)
        np.testing.assert_allclose(scale, rs, rtol=1e-4)
        np.testing.assert_allclose(mean, rm, rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(out, ro, rtol=1e-4, atol=1e-5)
        params = s.createPartialFlow(x[:, :, None], enc, 'Flow%d' % f)
        np.testing.assert_allclose(params[..., 0:1], np.log(rs), rtol=1e-4, atol=1e-5)
