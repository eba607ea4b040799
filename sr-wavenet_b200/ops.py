"""Python-facing mirror of the reference's ``ops.py`` (same names, argument order and
meaning), backed by the CUDA kernels in libsrwn.so.  Tensors are fp32 CUDA torch tensors,
channels-last ``[B, T, C]`` like the reference's TF tensors; torch is only the memory /
stream host.  TF1's variable store is mirrored by a small name -> tensor registry
(``variable_scope`` / ``get_variable``) so that ``DilatedCausalConv1d`` and
``ResidualDilationLayer`` create the same variable names as the reference graph.
"""
import contextlib
import math

import numpy as np
import torch

from . import _lib

SQRT_HALF = 0.7071067811865476
SEED = int.from_bytes(__import__("os").urandom(7), "little")     # key of the sampler's Philox draws (the reference is unseeded)
_draws = 0

# ---- TF1-style variable store -----------------------------------------------------------
_scope_stack = []
_variables = {}
_layer_counts = {}


@contextlib.contextmanager
def variable_scope(name):
    """tf.variable_scope(name): prefixes variable names created inside."""
    _scope_stack.append(name)
    try:
        yield
    finally:
        _scope_stack.pop()


def current_scope():
    return "/".join(s for s in _scope_stack if s)


def reset_variables():
    _variables.clear()
    _layer_counts.clear()


def global_variables():
    return dict(_variables)


def get_variable(name, shape, initializer="xavier", dtype=torch.float32, seed=None):
    """tf.get_variable: returns the existing variable of that (scoped) name or creates it.
    'xavier' = tf.contrib.layers.xavier_initializer (uniform, ops.py:15); 'zeros' = ops.py:18."""
    full = (current_scope() + "/" if current_scope() else "") + name
    if full in _variables:
        v = _variables[full]
        if tuple(v.shape) != tuple(shape):
            raise ValueError("variable %s exists with shape %s, wanted %s" % (full, tuple(v.shape), tuple(shape)))
        return v
    if initializer == "zeros":
        v = torch.zeros(*shape, dtype=dtype, device="cuda")
    else:
        K, cin, cout = shape
        limit = math.sqrt(6.0 / (K * cin + K * cout))
        rng = np.random.default_rng(seed)
        v = torch.from_numpy(rng.uniform(-limit, limit, size=shape).astype(np.float32)).cuda()
    _variables[full] = v
    return v


def _unique_layer_name(base):
    """tf.layers default naming inside the current scope: conv1d, conv1d_1, conv1d_2, ..."""
    key = (current_scope(), base)
    n = _layer_counts.get(key, 0)
    _layer_counts[key] = n + 1
    return base if n == 0 else "%s_%d" % (base, n)


# ---- plumbing ---------------------------------------------------------------------------
def _prep(t, name):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(t, dtype=np.float32))
    if not t.is_cuda:
        t = t.cuda()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


# ---- ops.py ------------------------------------------------------------------------------
def _DilatedCausalConv1d(inputs, filters, dilation_rate=1):
    """ops.py:6-10.  inputs [B,T,Cin], filters [K,Cin,Cout] -> [B,T,Cout]."""
    x, w = _prep(inputs, "inputs"), _prep(filters, "filters")
    B, T, cin = x.shape
    K, cin2, cout = w.shape
    if cin != cin2:
        raise ValueError("filters expect %d input channels, inputs have %d" % (cin2, cin))
    y = torch.empty(B, T, cout, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().srwn_dilated_causal_conv1d(
        _ptr(x), _ptr(w), None, _ptr(y), B, T, cin, cout, K, int(dilation_rate), _stream()))
    return y


def DilatedCausalConv1d(inputs, kernel_size, channels, dilation_rate=1, name='', dtype=torch.float32,
                        use_bias=True):
    """ops.py:13-20: creates ``<name>_Kernel [K,Cin,Cout]`` / ``<name>_Bias [1,1,C]`` and applies them."""
    x = _prep(inputs, "inputs")
    filters = get_variable(name + '_Kernel', [kernel_size, x.shape[-1], channels], "xavier", dtype)
    bias = get_variable(name + '_Bias', [1, 1, channels], "zeros", dtype) if use_bias else None
    B, T, cin = x.shape
    y = torch.empty(B, T, channels, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().srwn_dilated_causal_conv1d(
        _ptr(x), _ptr(filters), _ptr(bias), _ptr(y), B, T, cin, channels, kernel_size,
        int(dilation_rate), _stream()))
    return y


def ResidualDilationLayer(inputs, kernel_size, dilation_channels, skip_channels, dilation_rate=1,
                          name='', dtype=torch.float32, use_bias=True):
    """ops.py:23-46 -> (dense, skip).  Creates the same variables as the reference, including the
    dead ``<name>_gate`` conv (ops.py:31-33: its output is discarded; the gate is the sigmoid of
    the already tanh'd filter conv)."""
    x = _prep(inputs, "inputs")
    B, T, R = x.shape
    if R != dilation_channels:
        raise ValueError("inputs + residual (ops.py:40) needs inputs.shape[-1] == dilation_channels")
    with variable_scope(name + '_filter'):
        fk = get_variable(name + '_Kernel', [kernel_size, R, dilation_channels], "xavier", dtype)
        fb = get_variable(name + '_Bias', [1, 1, dilation_channels], "zeros", dtype) if use_bias else None
    with variable_scope(name + '_gate'):       # dead variables, kept for checkpoint compatibility
        get_variable(name + '_Kernel', [kernel_size, R, dilation_channels], "xavier", dtype)
        if use_bias:
            get_variable(name + '_Bias', [1, 1, dilation_channels], "zeros", dtype)
    with variable_scope(_unique_layer_name('conv1d')):
        rk = get_variable('kernel', [1, dilation_channels, dilation_channels], "xavier", dtype)
        rb = get_variable('bias', [dilation_channels], "zeros", dtype)
    with variable_scope(_unique_layer_name('conv1d')):
        sk = get_variable('kernel', [1, dilation_channels, skip_channels], "xavier", dtype)
        sb = get_variable('bias', [skip_channels], "zeros", dtype)
    dense = torch.empty(B, T, dilation_channels, dtype=torch.float32, device=x.device)
    skip = torch.empty(B, T, skip_channels, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().srwn_residual_dilation_layer(
        _ptr(x), _ptr(fk), _ptr(fb), _ptr(rk), _ptr(rb), _ptr(sk), _ptr(sb), _ptr(dense), _ptr(skip),
        B, T, dilation_channels, skip_channels, kernel_size, int(dilation_rate), _stream()))
    return dense, skip


def ResidualDilationLayerNC(inputs, kernel_size, dilation_channels, skip_channels, dilation_rate=1,
                            name='', dtype=torch.float32, use_bias=True):
    """ops.py:48-57 -> (residual, skip): relu -> conv1d(kernel_size, SAME) -> relu -> two 1x1 convs.  Like the reference,
    ``dilation_rate`` and ``use_bias`` are accepted and ignored (``tf.layers.conv1d`` is called without a dilation and with
    its default bias), and the block returns the residual itself, not ``inputs + residual``.  The teacher's encoder runs
    this block in its own tcgen05 kernel (csrc/encoder.cu); this is the shape-generic building block."""
    x = _prep(inputs, "inputs")
    B, T, cin = x.shape
    lib = _lib.load()
    with variable_scope(name + '_NC'):
        with variable_scope(_unique_layer_name('conv1d')):
            k0 = get_variable('kernel', [kernel_size, cin, dilation_channels], "xavier", dtype)
            b0 = get_variable('bias', [dilation_channels], "zeros", dtype)
    with variable_scope(_unique_layer_name('conv1d')):
        rk = get_variable('kernel', [1, dilation_channels, dilation_channels], "xavier", dtype)
        rb = get_variable('bias', [dilation_channels], "zeros", dtype)
    with variable_scope(_unique_layer_name('conv1d')):
        sk = get_variable('kernel', [1, dilation_channels, skip_channels], "xavier", dtype)
        sb = get_variable('bias', [skip_channels], "zeros", dtype)
    h = torch.empty(B, T, dilation_channels, dtype=torch.float32, device=x.device)
    _lib.check(lib.srwn_conv1d_same(_ptr(x), _ptr(k0), _ptr(b0), _ptr(h), B, T, cin, dilation_channels, kernel_size, 3, _stream()))
    residual = torch.empty(B, T, dilation_channels, dtype=torch.float32, device=x.device)
    skip = torch.empty(B, T, skip_channels, dtype=torch.float32, device=x.device)
    _lib.check(lib.srwn_conv1d_same(_ptr(h), _ptr(rk), _ptr(rb), _ptr(residual), B, T, dilation_channels, dilation_channels, 1, 0, _stream()))
    _lib.check(lib.srwn_conv1d_same(_ptr(h), _ptr(sk), _ptr(sb), _ptr(skip), B, T, dilation_channels, skip_channels, 1, 0, _stream()))
    return residual, skip


def log_prob_from_logits(x):
    """ops.py:111-115: numerically stable log-softmax over the last axis."""
    t = _prep(x, "x")
    y = torch.empty_like(t)
    _lib.check(_lib.load().srwn_log_softmax(_ptr(t), _ptr(y), t.numel() // t.shape[-1], t.shape[-1], 0, _stream()))
    return y


def log_sum_exp(x):
    """ops.py:117-122: numerically stable log-sum-exp over the last axis (which it removes)."""
    t = _prep(x, "x")
    y = torch.empty(t.shape[:-1], dtype=torch.float32, device=t.device)
    _lib.check(_lib.load().srwn_log_softmax(_ptr(t), _ptr(y), t.numel() // t.shape[-1], t.shape[-1], 1, _stream()))
    return y


def ResizeEmbeddingNearestNeighbor(inputs, output_size):
    """ops.py:64-74: [B,L,C] -> [B,output_size,C], nearest neighbour, align_corners=False."""
    x = _prep(inputs, "inputs")
    B, L, C = x.shape
    y = torch.empty(B, int(output_size), C, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().srwn_resize_nearest(_ptr(x), _ptr(y), B, L, C, int(output_size), _stream()))
    return y


def RightShift(inputs, shift_size=1):
    """ops.py:78-80."""
    x = _prep(inputs, "inputs")
    B, T, C = x.shape
    y = torch.empty_like(x)
    _lib.check(_lib.load().srwn_right_shift(_ptr(x), _ptr(y), B, T, C, int(shift_size), _stream()))
    return y


def discretized_mix_logistic_loss(x, l, sum_all=True):
    """ops.py:124-175.  x [B,T,1], l [B,T,4M] -> scalar tensor, or [B,T,1] for sum_all=False."""
    xt, lt = _prep(x, "x"), _prep(l, "l")
    B, T, C = lt.shape
    if C % 4 or xt.numel() != B * T:
        raise ValueError("l must be [B,T,4*M] and x [B,T,1]")
    if sum_all:
        out = torch.empty(1, dtype=torch.float32, device=lt.device)
        _lib.check(_lib.load().srwn_mol_loss(_ptr(xt), _ptr(lt), None, _ptr(out), B, T, C // 4, _stream()))
        return out[0]
    out = torch.empty(B, T, 1, dtype=torch.float32, device=lt.device)
    _lib.check(_lib.load().srwn_mol_loss(_ptr(xt), _ptr(lt), _ptr(out), None, B, T, C // 4, _stream()))
    return out


def sample_from_discretized_mix_logistic(l, nr_mix, u1=None, u2=None, return_index=False):
    """ops.py:178-201 -> [B,T,1].  ``u1`` [B,T,M] / ``u2`` [B,T] are the two uniform draws of
    ops.py:187,196; when omitted they are drawn on the device in [1e-5, 1-1e-5]."""
    lt = _prep(l, "l")
    B, T, C = lt.shape
    if C != 4 * nr_mix:
        raise ValueError("l must have 4*nr_mix channels")
    lo, hi = 1e-5, 1.0 - 1e-5
    if u1 is None or u2 is None:       # tf.random_uniform (ops.py:187, 196): Philox draws on the device
        global _draws
        _draws += 1

        def _uniform(shape, stream_id):
            t = torch.empty(shape, dtype=torch.float32, device=lt.device)
            _lib.check(_lib.load().srwn_random_uniform(_ptr(t), t.numel(), SEED, 2 * _draws + stream_id, lo, hi, _stream()))
            return t
    u1 = _uniform((B, T, nr_mix), 0) if u1 is None else _prep(u1, "u1")
    u2 = _uniform((B, T), 1) if u2 is None else _prep(u2, "u2")
    out = torch.empty(B, T, 1, dtype=torch.float32, device=lt.device)
    idx = torch.empty(B, T, dtype=torch.int32, device=lt.device) if return_index else None
    _lib.check(_lib.load().srwn_mol_sample(_ptr(lt), _ptr(u1), _ptr(u2), _ptr(out), _ptr(idx), B, T,
                                           nr_mix, _stream()))
    return (out, idx) if return_index else out
