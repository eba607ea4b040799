"""PyTorch-CPU fp32 port of the teacher decoder + MoL likelihood (TEST / BASELINE infrastructure).

This is the timed stand-in for "the reference's TensorFlow CPU path" (BASELINE.md section 4): the
reference needs TensorFlow <= 1.15, which cannot be installed here, so `bench.py`'s
``cpu_baseline`` leg and ``--impl reference`` time this port instead (kind "port").  It mirrors the
op structure the TF graph executes (one conv per tap, separate 1x1 convs, a stacked skip sum) with
torch's multi-threaded CPU kernels, and is validated against oracle/srwn_oracle.py in
tests/test_oracle_golden.py.  Parity unpinned against the reference itself (see oracle/__init__.py).
"""
import numpy as np
import torch

SQRT_HALF = 0.7071067811865476


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


class TeacherCPU(object):
    def __init__(self, weights, dilations, pool_stride, num_mixtures,
                 prefix='WaveNetAutoEncoder/Decoder/'):
        self.dil, self.P, self.M, self.n = list(dilations), pool_stride, num_mixtures, len(dilations)
        g = lambda n: _t(weights[prefix + n])
        self.front_k, self.front_b = g('causal_conv_Kernel'), g('causal_conv_Bias').reshape(-1)
        self.layers = []
        for i in range(self.n):
            cname = 'conv1d' if i == 0 else 'conv1d_%d' % (3 * i)
            name = 'dilated_conv_%d' % i
            self.layers.append(dict(
                ck=g(cname + '/kernel')[0], cb=g(cname + '/bias'),
                fk=g('%s_filter/%s_Kernel' % (name, name)), fb=g('%s_filter/%s_Bias' % (name, name)).reshape(-1),
                rk=g('conv1d_%d/kernel' % (3 * i + 1))[0], rb=g('conv1d_%d/bias' % (3 * i + 1)),
                sk=g('conv1d_%d/kernel' % (3 * i + 2))[0], sb=g('conv1d_%d/bias' % (3 * i + 2))))
        self.h1k, self.h1b = g('conv1d_%d/kernel' % (3 * self.n))[0], g('conv1d_%d/bias' % (3 * self.n))
        self.h2k, self.h2b = g('conv1d_%d/kernel' % (3 * self.n + 1))[0], g('conv1d_%d/bias' % (3 * self.n + 1))

    @staticmethod
    def _causal(x, k, d):
        """ops.py:6-10 for K=2: k[0] pairs with x[t-d], k[1] with x[t]."""
        tap = torch.nn.functional.pad(x, (0, 0, d, 0))[:, :x.shape[1]]
        return tap @ k[0] + x @ k[1]

    @torch.no_grad()
    def logits(self, x, enc):
        x, enc = _t(x), _t(enc)
        h = torch.nn.functional.pad(x[:, :, None], (0, 0, 1, 0))[:, :x.shape[1]]      # ops.py:78-80
        h = self._causal(h, self.front_k, 1) + self.front_b                           # model.py:173
        skips = []
        for lw, d in zip(self.layers, self.dil):
            cond = enc @ lw['ck'] + lw['cb']                                          # model.py:180
            h = h + torch.repeat_interleave(cond, self.P, dim=1)                      # model.py:181-183
            f = torch.tanh(self._causal(h, lw['fk'], d) + lw['fb'])                   # ops.py:27-28
            c = f * torch.sigmoid(f)                                                  # ops.py:33,36
            skips.append(c @ lw['sk'] + lw['sb'])                                     # ops.py:44
            h = (h + (c @ lw['rk'] + lw['rb'])) * SQRT_HALF                           # ops.py:39-40
        total = torch.relu(torch.stack(skips, 0).sum(0))                              # model.py:190-191
        total = torch.relu(total @ self.h1k + self.h1b)                               # model.py:193-194
        return total @ self.h2k + self.h2b                                            # model.py:196

    @torch.no_grad()
    def nll(self, x, enc):
        """model.py:114-115 with ops.py:124-175 (sum_all=True), same formulas as the TF graph."""
        l = self.logits(x, enc)
        M = self.M
        xt = _t(x)[:, :, None].expand(-1, -1, M)
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
                                    mid_in - log_scales - 2. * sp(mid_in) - float(np.log(127.5)))))
        log_probs = log_probs + torch.log_softmax(logit_probs, dim=-1)
        return float(-torch.logsumexp(log_probs, dim=-1).double().sum())
