"""Differentiable restatement (PyTorch CPU, float64) of the student network and the distillation loss
(model.py:415-535, model.py:356-379) -- TEST INFRASTRUCTURE ONLY.

The reference obtains the student's gradients from TensorFlow's autodiff (model.py:384); TF 1.x cannot be
installed here, so the gradient oracle is torch.autograd over the same forward formulas.  The forward is
pinned against oracle/srwn_oracle.py (NumPy) in tests/test_oracle_golden.py; gradients are additionally
spot-checked by central finite differences there.  Parity unpinned against the reference itself.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import math

import numpy as np
import torch

SQRT_HALF = 0.7071067811865476


def _causal(x, k, d):
    """ops.py:6-10 for K=2: k[0] pairs with x[t-d], k[1] with x[t] (zero padding on the left)."""
    T = x.shape[1]
    tap = torch.nn.functional.pad(x, (0, 0, d, 0))[:, :T]
    return tap @ k[0] + x @ k[1]


def student_forward(W, z, enc, dilations, pool_stride, num_flows):
    """W: dict name -> torch tensor (requires_grad as needed); z [B,T]; enc [B,T/P,C].
    Returns out [B,T] (clipped, model.py:535), s_tot, mu_tot [B,T]."""
    n = len(dilations)
    x = z[:, :, None]
    scales, means = [], []
    for f in range(num_flows):
        p = 'ParallelWaveNet/Flow%d/Flow%d/' % (f, f)
        T = x.shape[1]
        h = torch.nn.functional.pad(x, (0, 0, 1, 0))[:, :T]                           # RightShift, ops.py:78-80
        h = _causal(h, W[p + 'causal_conv_Kernel'], 1) + W[p + 'causal_conv_Bias'].reshape(-1)   # model.py:424
        for i, d in enumerate(dilations):
            cname = 'conv1d' if i == 0 else 'conv1d_%d' % (3 * i)
            name = 'dilated_conv_%d' % i
            cond = enc @ W[p + cname + '/kernel'][0] + W[p + cname + '/bias']         # model.py:431
            h = h + torch.repeat_interleave(cond, pool_stride, dim=1)                 # model.py:432-434
            fl = torch.tanh(_causal(h, W[p + '%s_filter/%s_Kernel' % (name, name)], d) +
                            W[p + '%s_filter/%s_Bias' % (name, name)].reshape(-1))    # ops.py:27-28
            c = fl * torch.sigmoid(fl)                                                # ops.py:33,36 (F1)
            res = c @ W[p + 'conv1d_%d/kernel' % (3 * i + 1)][0] + W[p + 'conv1d_%d/bias' % (3 * i + 1)]
            h = (h + res) * SQRT_HALF                                                 # ops.py:39-40
        prm = torch.relu(h) @ W[p + 'conv1d_%d/kernel' % (3 * n)][0] + W[p + 'conv1d_%d/bias' % (3 * n)]   # model.py:451-452
        s, m = torch.exp(prm[:, :, 0:1]), prm[:, :, 1:2]                              # model.py:479-480
        x = x * s + m                                                                 # model.py:482
        scales.append(s)
        means.append(m)
    s_tot = torch.ones_like(scales[0])
    mu_tot = torch.zeros_like(scales[0])
    for i in range(num_flows):                                                        # model.py:520-533
        s_tot = s_tot * scales[i]
        mu = means[i]
        for j in range(i + 1, num_flows):
            mu = mu * scales[j]
        mu_tot = mu_tot + mu
    out = torch.minimum(torch.maximum(z[:, :, None] * s_tot + mu_tot, torch.tensor(-1.0, dtype=z.dtype)),
                        torch.tensor(1.0, dtype=z.dtype))                            # model.py:535
    return out[:, :, 0], s_tot[:, :, 0], mu_tot[:, :, 0]


def mol_nll(x, l, M):
    """ops.py:124-175, sum_all=True.  x [B,T], l [B,T,4M] -> scalar."""
    xt = x[:, :, None].expand(-1, -1, M)
    logit_probs, means = l[:, :, :M], l[:, :, M:2 * M]
    log_scales = torch.clamp(l[:, :, 2 * M:3 * M], min=-7.0)
    centered = xt - means
    inv = torch.exp(-log_scales)
    plus_in, min_in, mid_in = inv * (centered + 1 / 255.), inv * (centered - 1 / 255.), inv * centered
    sp = torch.nn.functional.softplus
    cdf_delta = torch.sigmoid(plus_in) - torch.sigmoid(min_in)
    log_probs = torch.where(
        xt < -0.999, plus_in - sp(plus_in),
        torch.where(xt > 0.999, -sp(min_in),
                    torch.where(cdf_delta > 1e-5, torch.log(torch.clamp(cdf_delta, min=1e-12)),
                                mid_in - log_scales - 2. * sp(mid_in) - math.log(127.5))))
    log_probs = log_probs + torch.log_softmax(logit_probs, dim=-1)
    return -torch.logsumexp(log_probs, dim=-1).sum()


def stft_power(x):
    """model.py:360-367: tf.contrib.signal.stft(x, 512, 256) (periodic Hann, no padding) -> mean_t |.|^2, [B,257]."""
    win = torch.hann_window(512, periodic=True, dtype=x.dtype)
    spec = torch.stft(x, n_fft=512, hop_length=256, win_length=512, window=win, center=False, return_complex=True)
    return (spec.real ** 2 + spec.imag ** 2).mean(dim=2)


def distillation_loss(W, z, truth, enc, teacher_logits, dilations, pool_stride, num_flows, alpha, beta, gamma,
                      batch_norm=None):
    """model.py:356-379 -> (loss, power_loss, entropy).  ``teacher_logits`` are constants (stop_gradient,
    model.py:333).  ``batch_norm`` overrides the divisor B (global batch under data parallelism)."""
    out, s_tot, _ = student_forward(W, z, enc, dilations, pool_stride, num_flows)
    M = teacher_logits.shape[2] // 4
    entropy = torch.sum(torch.log(s_tot) + 2.0)
    diff = stft_power(truth) - stft_power(out)
    power = torch.sum(diff ** 2) * gamma
    ce = mol_nll(torch.clamp(out, -1, 1), teacher_logits, M) * beta
    loss = (ce - entropy * alpha + power) / float(batch_norm or z.shape[0])
    return loss, power, entropy


def loss_and_grads(weights, z, truth, enc, teacher_logits, dilations, pool_stride, num_flows,
                   alpha=1.0, beta=1.0, gamma=1.0, batch_norm=None):
    """NumPy in / NumPy out: (loss, power_loss, entropy, {name: dLoss/dvar}) in float64."""
    W = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in weights.items()}
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64))
    loss, power, ent = distillation_loss(W, t(z), t(truth), t(enc), t(teacher_logits), dilations, pool_stride,
                                         num_flows, alpha, beta, gamma, batch_norm)
    loss.backward()
    grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros(v.shape)) for k, v in W.items()}
    return float(loss.detach()), float(power.detach()), float(ent.detach()), grads


def adam_reference(w, g, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8, clip=1.0, gnorm=None):
    """tf.clip_by_global_norm + tf.train.AdamOptimizer._apply_dense (model.py:382-401), float64 NumPy."""
    gn = math.sqrt(sum(float((x.astype(np.float64) ** 2).sum()) for x in g)) if gnorm is None else gnorm
    scale = clip / max(gn, clip) if clip else 1.0       # clip=None: gradients arrive clipped (model.py:603-632)
    lr_t = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
    out = []
    for wi, gi, mi, vi in zip(w, g, m, v):
        gi = gi * scale
        mi = b1 * mi + (1 - b1) * gi
        vi = b2 * vi + (1 - b2) * gi * gi
        out.append((wi - lr_t * mi / (np.sqrt(vi) + eps), mi, vi))
    return out
