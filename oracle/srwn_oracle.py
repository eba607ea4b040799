"""NumPy restatement of the SR-WaveNet hot path (TEST INFRASTRUCTURE, not product).

Each function cites the reference file:line (relative to /root/reference) whose
arithmetic it restates.  Layout everywhere is channels-last ``[B, T, C]`` like the
reference.  Computation runs in the dtype of the inputs (float64 for golden
vectors, float32 to look at rounding drift).

Parity status
-------------
**Pinned by executing the reference's own source.**  TensorFlow 1.x (where the reference's arithmetic
lives: ``tf.nn.convolution``, ``tf.layers.conv1d``, ``tf.image.resize_nearest_neighbor`` ..., not
vendored, no version pinned) cannot be installed here, so ``tests/tf_shim/tensorflow`` provides a
NumPy stand-in for the ~60 ``tf.*`` symbols ``ops.py`` / ``model.py`` touch and
``tests/test_reference_shim.py`` imports ``/root/reference/ops.py`` and ``model.py`` UNMODIFIED through
it.  Every function below is compared at 1e-11 with what the reference's code computes: the block
(``ops.py:23-46``), shift / resize, both mixture-of-logistics functions incl. every branch, the teacher
built by ``WaveNetAutoEncoder.__init__`` (logits, ``loss_encoding``, ``encode``,
``reconstruct_with_encoding``, the training loss), the naive autoregressive loop of
``teacher.py:153-170``, the student built by ``ParallelWaveNet.__init__`` with the teacher imported
through ``import_meta_graph`` + ``input_map`` (``generate``, ``s_tot``, ``mu_tot``, entropy, power loss,
total loss), the variable names and shapes both constructors create, and the set of variables without
gradient (gate convs, student skip convs, the last layer's residual conv).  Graph structure, operation
order, scoping and all quirks F1-F8 therefore come from the reference; what remains restated is the
semantics of single TensorFlow operations (listed in the stand-in's docstring).
* the reference's printed known answers (``ops.py:243-254``) are reproduced both by this file
  (``tests/test_oracle_kat.py``) and by the reference's ``_DilatedCausalConv1d`` run through the stand-in;
* ``tests/golden/reference_*.npz`` are outputs of that reference code path (generator:
  ``tests/golden/make_reference_golden.py``); ``tests/test_oracle_golden.py`` holds this file to them on
  machines without the reference tree and ``tests/test_gpu_reference_golden.py`` the CUDA path.

Weight dictionaries are keyed by the TF1 variable names the reference graph would
create; kernels keep TF's ``[K, Cin, Cout]`` layout.
"""
import numpy as np

SQRT_HALF = 0.7071067811865476  # literal at ops.py:40


# --------------------------------------------------------------------------- ops.py
def dilated_causal_conv1d(inputs, filters, dilation_rate=1):
    """ops.py:6-10  _DilatedCausalConv1d: left-pad d*(K-1), VALID dilated cross-correlation.

    inputs [B,T,Cin], filters [K,Cin,Cout] -> [B,T,Cout];
    out[b,t] = sum_k inputs[b, t - d*(K-1-k)] @ filters[k]   (zero for negative time).
    """
    K = filters.shape[0]
    B, T, _ = inputs.shape
    pad = dilation_rate * (K - 1)
    padded = np.concatenate([np.zeros((B, pad, inputs.shape[2]), inputs.dtype), inputs], axis=1)
    out = np.zeros((B, T, filters.shape[2]), dtype=np.result_type(inputs, filters))
    for k in range(K):
        out += padded[:, k * dilation_rate:k * dilation_rate + T, :] @ filters[k]
    return out


def valid_conv1d(inputs, filters, dilation_rate=1):
    """ops.py:254  tf.nn.convolution(padding='VALID'): no padding, T-d*(K-1) outputs."""
    K = filters.shape[0]
    Tout = inputs.shape[1] - dilation_rate * (K - 1)
    out = 0
    for k in range(K):
        out = out + inputs[:, k * dilation_rate:k * dilation_rate + Tout, :] @ filters[k]
    return out


def dilated_causal_conv1d_layer(inputs, kernel, bias, dilation_rate=1):
    """ops.py:13-20  DilatedCausalConv1d: conv + ``<name>_Bias`` of shape [1,1,C]."""
    conv = dilated_causal_conv1d(inputs, kernel, dilation_rate)
    if bias is not None:
        conv = conv + bias.reshape(1, 1, -1)
    return conv


def sigmoid(x):
    """tf.nn.sigmoid, written so that neither branch overflows."""
    e = np.exp(-np.abs(x))
    return np.where(x >= 0, 1.0 / (1.0 + e), e / (1.0 + e)).astype(x.dtype)


def softplus(x):
    """tf.nn.softplus = log(1 + exp(x))."""
    return np.logaddexp(0, x).astype(x.dtype)


def residual_dilation_layer(inputs, filt_kernel, filt_bias, res_kernel, res_bias,
                            skip_kernel, skip_bias, dilation_rate=1):
    """ops.py:23-46  ResidualDilationLayer -> (dense, skip).

    The reference's gate is ``sigmoid(filter_conv)`` taken AFTER the tanh
    (ops.py:33; the ``_gate`` conv output is discarded), so
    combined = tanh(f) * sigmoid(tanh(f)).  Both 1x1 convs carry a bias
    (tf.layers.conv1d default use_bias=True).  ``skip_kernel`` may be None (student:
    the skip conv is dead code, model.py:438-454).
    res_kernel [1,R,R], skip_kernel [1,R,S].
    """
    f = np.tanh(dilated_causal_conv1d_layer(inputs, filt_kernel, filt_bias, dilation_rate))
    g = sigmoid(f)                                   # ops.py:33
    combined = f * g                                 # ops.py:36
    residual = combined @ res_kernel[0] + res_bias   # ops.py:39
    dense = (inputs + residual) * inputs.dtype.type(SQRT_HALF)  # ops.py:40
    skip = None
    if skip_kernel is not None:
        skip = combined @ skip_kernel[0] + skip_bias  # ops.py:44
    return dense, skip


def resize_embedding_nearest_neighbor(inputs, output_size):
    """ops.py:64-74  resize_nearest_neighbor(align_corners=False):
    src = min(floor(dst * in / out), in - 1)."""
    L = inputs.shape[1]
    src = np.minimum((np.arange(output_size) * L) // output_size, L - 1)
    return inputs[:, src, :]


def right_shift(inputs, shift_size=1):
    """ops.py:78-80  pad ``shift_size`` zeros on the left of time, drop the tail."""
    B, T, C = inputs.shape
    p = np.concatenate([np.zeros((B, shift_size, C), inputs.dtype), inputs], axis=1)
    return p[:, :T, :]


def log_prob_from_logits(x):
    """ops.py:111-115  stable log-softmax over the last axis."""
    m = x.max(axis=-1, keepdims=True)
    return x - m - np.log(np.sum(np.exp(x - m), axis=-1, keepdims=True))


def log_sum_exp(x):
    """ops.py:117-122  stable logsumexp over the last axis."""
    m = x.max(axis=-1)
    m2 = x.max(axis=-1, keepdims=True)
    return m + np.log(np.sum(np.exp(x - m2), axis=-1))


def discretized_mix_logistic_loss(x, l, sum_all=True):
    """ops.py:124-175  PixelCNN++ discretized mixture of logistics, one channel.

    x [B,T,1] in [-1,1]; l [B,T,4*M].  Returns the scalar -sum(log p) or, for
    ``sum_all=False``, ``[B,T,1]`` of -log p.  ``coeffs`` (ops.py:137) is computed by
    the reference but never used.
    """
    dt = l.dtype.type
    nr_mix = l.shape[-1] // 4                                  # ops.py:131
    logit_probs = l[:, :, :nr_mix]
    means = l[:, :, nr_mix:2 * nr_mix]                          # ops.py:135
    log_scales = np.maximum(l[:, :, 2 * nr_mix:3 * nr_mix], dt(-7.0))  # ops.py:136
    xt = np.repeat(x, nr_mix, axis=2)                           # ops.py:139
    centered_x = xt - means
    inv_stdv = np.exp(-log_scales)
    plus_in = inv_stdv * (centered_x + dt(1.0 / 255.0))
    cdf_plus = sigmoid(plus_in)
    min_in = inv_stdv * (centered_x - dt(1.0 / 255.0))
    cdf_min = sigmoid(min_in)
    log_cdf_plus = plus_in - softplus(plus_in)                  # ops.py:152
    log_one_minus_cdf_min = -softplus(min_in)                   # ops.py:153
    cdf_delta = cdf_plus - cdf_min
    mid_in = inv_stdv * centered_x
    log_pdf_mid = mid_in - log_scales - dt(2.0) * softplus(mid_in)  # ops.py:156
    log_probs = np.where(
        xt < dt(-0.999), log_cdf_plus,
        np.where(xt > dt(0.999), log_one_minus_cdf_min,
                 np.where(cdf_delta > dt(1e-5),
                          np.log(np.maximum(cdf_delta, dt(1e-12))),
                          log_pdf_mid - dt(np.log(127.5)))))    # ops.py:167
    log_probs = log_probs + log_prob_from_logits(logit_probs)   # ops.py:169
    lse = log_sum_exp(log_probs)
    if sum_all:
        return -np.sum(lse)                                     # ops.py:172
    return -lse[:, :, None]                                     # ops.py:175


def sample_from_discretized_mix_logistic(l, nr_mix, u1, u2, return_index=False):
    """ops.py:178-201 with the two tf.random_uniform draws injected:
    u1 [B,T,M] and u2 [B,T,1], both in [1e-5, 1-1e-5].

    k = argmax(logit_probs - log(-log(u1))); x = clip(mu_k + exp(max(ls_k,-7)) *
    (log u2 - log(1-u2)), -1, 1).  Returns [B,T,1].
    """
    dt = l.dtype.type
    logit_probs = l[:, :, :nr_mix]
    sel_idx = np.argmax(logit_probs - np.log(-np.log(u1)), axis=2)       # ops.py:187
    means = np.take_along_axis(l[:, :, nr_mix:2 * nr_mix], sel_idx[..., None], axis=2)
    log_scales = np.maximum(
        np.take_along_axis(l[:, :, 2 * nr_mix:3 * nr_mix], sel_idx[..., None], axis=2), dt(-7.0))
    x = means + np.exp(log_scales) * (np.log(u2) - np.log(dt(1.0) - u2))  # ops.py:197
    x = np.minimum(np.maximum(x, dt(-1.0)), dt(1.0))                      # ops.py:199
    if return_index:
        return x, sel_idx
    return x


# --------------------------------------------------------------------------- model.py
def _stack(weights, prefix, inputs, encoding, dilations, pool_stride, with_skip):
    """Shared body of createDecoder (model.py:172-187) and createPartialFlow
    (model.py:423-440): RightShift -> causal conv -> per layer (1x1 conditioning at
    latent rate -> nearest-neighbour upsample -> add to the residual stream -> block).
    Returns (h, sum_of_skips or None)."""
    W = lambda n: weights[prefix + n]
    h = right_shift(inputs)                                               # :172 / :423
    h = dilated_causal_conv1d_layer(h, W('causal_conv_Kernel'), W('causal_conv_Bias'), 1)
    T = inputs.shape[1]
    total = None
    for i, d in enumerate(dilations):
        cname = 'conv1d' if i == 0 else 'conv1d_%d' % (3 * i)
        cond = encoding @ W(cname + '/kernel')[0] + W(cname + '/bias')    # :180 / :431
        up = resize_embedding_nearest_neighbor(cond, pool_stride * cond.shape[1])  # :181
        assert up.shape[1] == T, "T must equal pool_stride * latent frames (model.py:183)"
        h = h + up                                                        # :183 / :435
        name = 'dilated_conv_%d' % i
        fk = W('%s_filter/%s_Kernel' % (name, name))
        fb = W('%s_filter/%s_Bias' % (name, name))
        rk, rb = W('conv1d_%d/kernel' % (3 * i + 1)), W('conv1d_%d/bias' % (3 * i + 1))
        sk = sb = None
        if with_skip:
            sk, sb = W('conv1d_%d/kernel' % (3 * i + 2)), W('conv1d_%d/bias' % (3 * i + 2))
        h, skip = residual_dilation_layer(h, fk, fb, rk, rb, sk, sb, d)   # :185 / :438
        if with_skip:
            total = skip if total is None else total + skip               # :190 reduce_sum
    return h, total


def teacher_decoder_logits(weights, truth, encoding, dilations, pool_stride,
                           prefix='WaveNetAutoEncoder/Decoder/'):
    """model.py:158-196  createDecoder up to the logits.  truth [B,T] (teacher-forcing
    audio), encoding [B,T/P,C] -> logits [B,T,4M]."""
    n = len(dilations)
    _, total = _stack(weights, prefix, truth[:, :, None], encoding, dilations, pool_stride, True)
    total = np.maximum(total, 0)                                          # :191
    total = total @ weights[prefix + 'conv1d_%d/kernel' % (3 * n)][0] + \
        weights[prefix + 'conv1d_%d/bias' % (3 * n)]                      # :193
    total = np.maximum(total, 0)                                          # :194
    return total @ weights[prefix + 'conv1d_%d/kernel' % (3 * n + 1)][0] + \
        weights[prefix + 'conv1d_%d/bias' % (3 * n + 1)]                  # :196


def residual_dilation_layer_nc(inputs, conv_kernel, conv_bias, res_kernel, res_bias,
                               skip_kernel=None, skip_bias=None):
    """ops.py:48-57  ResidualDilationLayerNC -> (residual, skip).

    x = relu(inputs); x = relu(conv1d(x, k=K, padding='SAME')) -- NOT causal and NOT dilated (the
    ``dilation_rate`` argument is never passed on, ops.py:51): for K=2 TF's SAME pads one zero on the
    RIGHT, so out[t] = x[t] @ W[0] + x[t+1] @ W[1].  ``residual`` is the 1x1 conv alone (no skip
    connection to ``inputs``, ops.py:54,57); ``skip`` the second 1x1 conv."""
    x = np.maximum(inputs, 0)                                             # :49
    K = conv_kernel.shape[0]
    B, T, _ = x.shape
    left = (K - 1) // 2                                                   # SAME: extra padding goes to the end
    padded = np.concatenate([np.zeros((B, left, x.shape[2]), x.dtype), x,
                             np.zeros((B, K - 1 - left, x.shape[2]), x.dtype)], axis=1)
    conv = 0
    for k in range(K):
        conv = conv + padded[:, k:k + T, :] @ conv_kernel[k]
    x = np.maximum(conv + conv_bias, 0)                                   # :51-52
    residual = x @ res_kernel[0] + res_bias                               # :54
    skip = None if skip_kernel is None else x @ skip_kernel[0] + skip_bias  # :55
    return residual, skip


def teacher_encoder(weights, inputs, n_layers, pool_stride, prefix='WaveNetAutoEncoder/Encoder/'):
    """model.py:137-155  createEncoder: inputs [B,T] -> encoding [B, T/P, latent].

    ``nc_conv`` (1 -> E channels, its skip discarded, :141-142) then ``n_layers`` =
    len(dilations) layers of ResidualDilationLayerNC (:144-149), sum of skips (:151), 1x1 to the
    latent channels (:152), average pooling window = stride = pool_stride, VALID (:154).
    Variable names follow tf.layers' default numbering inside the ``Encoder`` scope: layer j
    (0 = nc_conv, i+1 = dilated_conv_i) owns ``<name>_NC/conv1d`` (the K=2 conv), ``conv1d_{2j}``
    (residual) and ``conv1d_{2j+1}`` (skip; ``conv1d`` without suffix for index 0); the latent conv is
    ``conv1d_{2(n_layers+1)}``."""
    W = lambda n: weights[prefix + n]
    cname = lambda idx: 'conv1d' if idx == 0 else 'conv1d_%d' % idx
    h = inputs[:, :, None]
    h, _ = residual_dilation_layer_nc(h, W('nc_conv_NC/conv1d/kernel'), W('nc_conv_NC/conv1d/bias'),
                                      W(cname(0) + '/kernel'), W(cname(0) + '/bias'))
    total = None
    for i in range(n_layers):
        name = 'dilated_conv_%d_NC/conv1d' % i
        j = i + 1
        h, skip = residual_dilation_layer_nc(h, W(name + '/kernel'), W(name + '/bias'),
                                             W(cname(2 * j) + '/kernel'), W(cname(2 * j) + '/bias'),
                                             W(cname(2 * j + 1) + '/kernel'), W(cname(2 * j + 1) + '/bias'))
        total = skip if total is None else total + skip                   # :151
    lat = cname(2 * (n_layers + 1))
    reduced = total @ W(lat + '/kernel')[0] + W(lat + '/bias')            # :152
    B, T, C = reduced.shape
    frames = T // pool_stride                                             # VALID pooling drops a ragged tail
    return reduced[:, :frames * pool_stride].reshape(B, frames, pool_stride, C).mean(axis=2)   # :154


def teacher_nll(weights, truth, encoding, dilations, pool_stride, sum_all=True):
    """model.py:114-115  loss_encoding = discretized_mix_logistic_loss(labels_truth, logits)."""
    logits = teacher_decoder_logits(weights, truth, encoding, dilations, pool_stride)
    return discretized_mix_logistic_loss(truth[:, :, None], logits, sum_all)


def naive_ar_loop(weights, encoding, dilations, pool_stride, num_mixtures, u1, u2, T):
    """teacher.py:153-170  the reference's only autoregressive path: one full decoder
    evaluation per generated sample, keep column i.  u1 [B,T,M], u2 [B,T] are the
    injected uniforms (the reference redraws for all T every iteration and keeps
    column i; with injected noise that is the same thing).

    Returns x [B,T] with ALL T samples.  (teacher.py:170 then zeroes the last sample,
    ``x_so_far[:, i:] = 0`` with i = T-1; callers wanting that quirk apply it.)
    """
    B = encoding.shape[0]
    x = np.zeros((B, T), dtype=encoding.dtype)
    for i in range(T):
        x[:, i:] = 0                                                      # teacher.py:164
        logits = teacher_decoder_logits(weights, x, encoding, dilations, pool_stride)
        s = sample_from_discretized_mix_logistic(logits, num_mixtures, u1, u2[:, :, None])
        x[:, i] = s[:, i, 0]                                              # teacher.py:167
    return x


def queue_ar(weights, encoding, dilations, pool_stride, num_mixtures, u1, u2, T,
             prefix='WaveNetAutoEncoder/Decoder/', return_logits=False):
    """Per-layer dilation-queue restatement of ``naive_ar_loop`` (same arithmetic per
    sample, O(T) instead of O(T^2)); exists so tests can check long horizons.  It is
    validated against ``naive_ar_loop`` in tests/test_oracle_golden.py."""
    W = lambda n: weights[prefix + n]
    B = encoding.shape[0]
    dt = encoding.dtype
    n = len(dilations)
    R = W('causal_conv_Kernel').shape[2]
    x = np.zeros((B, T), dtype=dt)
    all_logits = np.zeros((B, T, 4 * num_mixtures), dtype=dt)
    conds = []
    for i in range(n):
        cname = 'conv1d' if i == 0 else 'conv1d_%d' % (3 * i)
        conds.append(encoding @ W(cname + '/kernel')[0] + W(cname + '/bias'))
    hist = [np.zeros((B, T, R), dtype=dt) for _ in range(n)]   # inputs of each layer
    ck, cb = W('causal_conv_Kernel'), W('causal_conv_Bias').reshape(-1)
    sq = dt.type(SQRT_HALF)
    for t in range(T):
        xm1 = x[:, t - 1] if t >= 1 else np.zeros(B, dt)
        xm2 = x[:, t - 2] if t >= 2 else np.zeros(B, dt)
        h = xm2[:, None] * ck[0, 0][None, :] + xm1[:, None] * ck[1, 0][None, :] + cb
        total = 0
        for i, d in enumerate(dilations):
            h = h + conds[i][:, min(t // pool_stride, conds[i].shape[1] - 1), :]
            hist[i][:, t] = h
            tap = hist[i][:, t - d] if t >= d else np.zeros_like(h)
            name = 'dilated_conv_%d' % i
            fk = W('%s_filter/%s_Kernel' % (name, name))
            fb = W('%s_filter/%s_Bias' % (name, name)).reshape(-1)
            f = np.tanh(tap @ fk[0] + h @ fk[1] + fb)
            c = f * sigmoid(f)
            res = c @ W('conv1d_%d/kernel' % (3 * i + 1))[0] + W('conv1d_%d/bias' % (3 * i + 1))
            total = total + (c @ W('conv1d_%d/kernel' % (3 * i + 2))[0] + W('conv1d_%d/bias' % (3 * i + 2)))
            h = (h + res) * sq
        total = np.maximum(total, 0)
        total = np.maximum(total @ W('conv1d_%d/kernel' % (3 * n))[0] + W('conv1d_%d/bias' % (3 * n)), 0)
        logits = total @ W('conv1d_%d/kernel' % (3 * n + 1))[0] + W('conv1d_%d/bias' % (3 * n + 1))
        all_logits[:, t] = logits
        s = sample_from_discretized_mix_logistic(
            logits[:, None, :], num_mixtures, u1[:, t:t + 1], u2[:, t:t + 1, None])
        x[:, t] = s[:, 0, 0]
    if return_logits:
        return x, all_logits
    return x


def student_partial_flow(weights, inputs, encoding, dilations, pool_stride, flow):
    """model.py:415-454  createPartialFlow: the stack with the skip path discarded
    (model.py:438-454: skip_layers is never consumed) -> relu -> 1x1 conv to 2 ch.
    Variables live under ``ParallelWaveNet/Flow{f}/Flow{f}/`` (scope entered twice,
    model.py:469 + :416)."""
    prefix = 'ParallelWaveNet/Flow%d/Flow%d/' % (flow, flow)
    n = len(dilations)
    h, _ = _stack(weights, prefix, inputs, encoding, dilations, pool_stride, False)
    h = np.maximum(h, 0)                                                  # :451
    return h @ weights[prefix + 'conv1d_%d/kernel' % (3 * n)][0] + \
        weights[prefix + 'conv1d_%d/bias' % (3 * n)]                      # :452


def student_flow(weights, inputs, encoding, dilations, pool_stride, flow):
    """model.py:457-486  createFlow -> (scale, mean, out); no clamp on the log-scale."""
    params = student_partial_flow(weights, inputs, encoding, dilations, pool_stride, flow)
    scale = np.exp(params[:, :, 0:1])                                     # :479
    mean = params[:, :, 1:2]                                              # :480
    return scale, mean, inputs * scale + mean                             # :482


def student_network(weights, z, encoding, dilations, pool_stride, num_flows):
    """model.py:489-535  createNetwork: chain the flows, compose s_tot / mu_tot in the
    reference's loop order, out = clip(z*s_tot + mu_tot, -1, 1).
    z [B,T] -> dict(out [B,T,1], s_tot, mu_tot, x_last, scales, means)."""
    x = z[:, :, None]
    scales, means = [], []
    for f in range(num_flows):
        s, m, x = student_flow(weights, x, encoding, dilations, pool_stride, f)  # :510
        scales.append(s)
        means.append(m)
    s_tot = np.ones_like(scales[0])                                       # :517
    mu_tot = np.zeros_like(scales[0])                                     # :518
    for i in range(num_flows):
        s_tot = s_tot * scales[i]                                         # :521
        mu = means[i]
        for j in range(i + 1, num_flows):
            mu = mu * scales[j]                                           # :529
        mu_tot = mu_tot + mu                                              # :533
    out = np.minimum(np.maximum(z[:, :, None] * s_tot + mu_tot, -1), 1)   # :535
    return dict(out=out, s_tot=s_tot, mu_tot=mu_tot, x_last=x, scales=scales, means=means)


def stft_power(x, frame_length=512, frame_step=256):
    """model.py:360-367  tf.contrib.signal.stft(x,512,256) (periodic Hann, fft_length=512,
    pad_end=False) -> mean over frames of |.|^2.  x [B,T] -> [B,257]."""
    B, T = x.shape
    n_frames = 1 + (T - frame_length) // frame_step
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(frame_length) / frame_length)
    idx = np.arange(frame_length)[None, :] + frame_step * np.arange(n_frames)[:, None]
    frames = x[:, idx] * win.astype(x.dtype)
    spec = np.fft.rfft(frames, n=frame_length, axis=-1)
    return np.mean(np.abs(spec) ** 2, axis=1)


def distillation_loss(student_w, teacher_w, z, truth, encoding, dilations, pool_stride,
                      num_flows, alpha=1.0, beta=1.0, gamma=1.0):
    """model.py:356-379.  The teacher is teacher-forced on the REAL audio ``truth``
    (model.py:323-334); the student's clipped output enters only as the ``x`` argument
    of the mixture likelihood (model.py:374).  Returns (loss, power_loss, entropy)."""
    net = student_network(student_w, z, encoding, dilations, pool_stride, num_flows)
    out = net['out']
    teacher_logits = teacher_decoder_logits(teacher_w, truth, encoding, dilations, pool_stride)
    entropy = np.sum(np.log(net['s_tot']) + 2.0)                          # :356
    s1 = stft_power(truth)
    s2 = stft_power(out[:, :, 0])
    power_loss = np.sum((s1 - s2) ** 2) * gamma                           # :371
    h_pt_ps = discretized_mix_logistic_loss(np.clip(out, -1, 1), teacher_logits, True) * beta
    loss = (h_pt_ps - entropy * alpha + power_loss) / z.shape[0]          # :378-379
    return loss, power_loss, entropy
