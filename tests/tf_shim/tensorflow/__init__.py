"""Eager-enough NumPy stand-in for the TensorFlow 1.x symbols the reference touches.

TEST INFRASTRUCTURE ONLY.  TensorFlow <= 1.15 cannot be installed in this image (no CPython 3.12
wheel, no network), so the reference's own ``ops.py`` / ``model.py`` could never be executed.  This
package lets them be imported UNMODIFIED (``sys.path = [tests/tf_shim, /root/reference]``) and run:
the graph structure, the order of operations, the order in which variables are created and every
quirk of the reference then come from the reference's source, and only the
semantics of the individual ``tf.*`` operations below are restated.  ``tests/test_reference_shim.py``
compares the result with ``oracle/srwn_oracle.py``; ``tests/golden/make_reference_golden.py`` writes
fixtures from it.

What is restated here (TF 1.x behaviour, from its documentation / source):
  * ``tf.nn.convolution`` = cross-correlation over NWC input with a ``[K, Cin, Cout]`` filter,
    VALID padding, dilation; ``tf.layers.conv1d`` = the same with ``kernel``/``bias`` variables,
    SAME padding (K=2, stride 1: 0 left / 1 right), glorot-uniform kernel, zero bias;
  * variable scoping: ``variable_scope`` nesting, ``reuse``, and the per-scope ``conv1d``,
    ``conv1d_1``, ... numbering of un-named layers (``_get_unique_variable_scope``: a counter per full
    scope name, sub-scope counters reset when the enclosing scope is left);
  * ``tf.image.resize_nearest_neighbor(align_corners=False)``: ``src = min(floor(dst*in/out), in-1)``;
  * ``tf.nn.pool`` AVG/VALID, ``tf.contrib.signal.stft`` (periodic Hann, ``fft_length`` = next power
    of two, ``pad_end=False``), ``tf.norm`` (Frobenius), ``tf.where``/``one_hot``/``argmax`` etc.;
  * ``compute_gradients`` returns ``None`` for variables the loss does not depend on (reachability
    through the graph, stopping at ``tf.stop_gradient``) -- no derivative is computed here.
Arithmetic runs in ``tf._FLOAT`` (float64 by default, so comparisons at 1e-12 are meaningful).
Random draws come from ``tf._random_hook`` so tests can inject the uniforms.
"""
import builtins
import contextlib
import os
import pickle
import types

import numpy as np

_FLOAT = np.float64          # the arithmetic type standing in for tf.float32
float32 = 'float32'
int32 = 'int32'
int64 = 'int64'


def _np_dtype(dt):
    if dt in (float32, None):
        return _FLOAT
    return {'int32': np.int32, 'int64': np.int64}.get(dt, dt)


# ---------------------------------------------------------------------------------------------
# graph, tensors
# ---------------------------------------------------------------------------------------------
class Tensor(object):
    """A node of the lazy graph: ``fn(*input values) -> ndarray``; ``shape`` is the static shape."""
    _is_variable = False
    __array_priority__ = 1000
    __array_ufunc__ = None      # numpy scalars on the left defer to the reflected operators

    def __init__(self, fn, inputs, shape, name=None, stops_gradient=False):
        self._fn, self._inputs, self.name = fn, list(inputs), name
        self._shape = None if shape is None else tuple(shape)
        self._stops_gradient = stops_gradient
        self.graph = get_default_graph()

    @property
    def shape(self):
        return TensorShape(self._shape)

    def get_shape(self):
        return self.shape

    __hash__ = object.__hash__

    def __eq__(self, other):
        return self is other

    def __add__(self, o): return _binary(np.add, self, o)
    def __radd__(self, o): return _binary(np.add, o, self)
    def __sub__(self, o): return _binary(np.subtract, self, o)
    def __rsub__(self, o): return _binary(np.subtract, o, self)
    def __mul__(self, o): return _binary(np.multiply, self, o)
    def __rmul__(self, o): return _binary(np.multiply, o, self)
    def __truediv__(self, o): return _binary(np.true_divide, self, o)
    def __rtruediv__(self, o): return _binary(np.true_divide, o, self)
    __div__, __rdiv__ = __truediv__, __rtruediv__
    def __pow__(self, o): return _binary(np.power, self, o)
    def __rpow__(self, o): return _binary(np.power, o, self)
    def __neg__(self): return _unary(np.negative, self)
    def __abs__(self): return _unary(np.abs, self)
    def __lt__(self, o): return _binary(np.less, self, o)
    def __gt__(self, o): return _binary(np.greater, self, o)
    def __le__(self, o): return _binary(np.less_equal, self, o)
    def __ge__(self, o): return _binary(np.greater_equal, self, o)

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        shp = None
        if self._shape is not None:
            try:
                shp = np.empty([0 if d is None else d for d in self._shape])[idx].shape
                # dims that were unknown and kept whole stay unknown
                shp = list(shp)
                out_axis = 0
                for ax, it in enumerate(idx):
                    if isinstance(it, builtins.slice):
                        if self._shape[ax] is None:
                            shp[out_axis] = None
                        out_axis += 1
                for ax in range(len(idx), len(self._shape)):
                    if self._shape[ax] is None:
                        shp[out_axis] = None
                    out_axis += 1
            except Exception:
                shp = None
        return Tensor(lambda v: v[idx], [self], shp)


class TensorShape(object):
    def __init__(self, dims):
        self.dims = dims

    def __getitem__(self, i):
        return self.dims[i]

    def __len__(self):
        return len(self.dims)

    def __iter__(self):
        return iter(self.dims)

    def as_list(self):
        return list(self.dims)

    def __repr__(self):
        return 'TensorShape(%r)' % (self.dims,)


class Variable(Tensor):
    _is_variable = True

    def __init__(self, name, shape, value):
        Tensor.__init__(self, None, [], shape, name=name + ':0')
        self.var_name = name
        self.value = value

    @property
    def op(self):
        return types.SimpleNamespace(name=self.var_name)


class Operation(object):
    """Stand-in for training ops: building them is allowed, running them is not."""

    def __init__(self, what):
        self.what = what


class Graph(object):
    def __init__(self):
        self.collections = {}
        self.variables = {}            # full name -> Variable, in creation order
        self.scope_counts = {}
        self.scope_stack = [_VarScope('', False)]
        self.placeholder_names = {}
        self.placeholders = {}         # tensor name -> placeholder

    @contextlib.contextmanager
    def as_default(self):
        _graph_stack.append(self)
        try:
            yield self
        finally:
            _graph_stack.pop()

    def get_collection(self, name, scope=None):
        items = list(self.collections.get(name, []))
        if scope:
            items = [v for v in items if getattr(v, 'var_name', v.name or '').startswith(scope)]
        return items

    def add_to_collection(self, name, value):
        self.collections.setdefault(name, []).append(value)


_graph_stack = [None]
_session_stack = [None]


def get_default_graph():
    if _graph_stack[-1] is None:
        _graph_stack[-1] = Graph()
    return _graph_stack[-1]


def reset_default_graph():
    _graph_stack[-1] = Graph()


def as_tensor(x):
    if isinstance(x, Tensor):
        return x
    if isinstance(x, (list, tuple)) and any(isinstance(e, Tensor) for e in x):
        parts = [as_tensor(e) for e in x]
        return Tensor(lambda *v: np.stack([np.asarray(e) for e in v]), parts,
                      (len(parts),) + tuple(parts[0]._shape or ()) if parts[0]._shape is not None else None)
    arr = np.asarray(x)
    if arr.dtype.kind == 'f':
        arr = arr.astype(_FLOAT)
    return Tensor(lambda: arr, [], arr.shape)


constant = lambda value, dtype=None, shape=None, name=None: as_tensor(
    np.full(shape, value) if shape is not None else value)


def _bshape(a, b):
    if a is None or b is None:
        return None
    out = []
    for x, y in zip(((1,) * (len(b) - len(a)) + tuple(a)), ((1,) * (len(a) - len(b)) + tuple(b))):
        out.append(y if x == 1 else (x if y == 1 or y == x else (x if y is None else (y if x is None else x))))
    return tuple(out)


def _binary(f, a, b):
    a, b = as_tensor(a), as_tensor(b)
    return Tensor(lambda x, y: f(x, y), [a, b], _bshape(a._shape, b._shape))


def _unary(f, a):
    a = as_tensor(a)
    return Tensor(lambda x: f(x), [a], a._shape)


def evaluate(fetch, feeds):
    """Iterative post-order evaluation with memoisation (graphs here are thousands of nodes deep)."""
    cache = dict(feeds)
    stack = [fetch]
    while stack:
        t = stack[-1]
        if t in cache:
            stack.pop()
            continue
        if t._is_variable:
            cache[t] = t.value
            stack.pop()
            continue
        if isinstance(t, _Remapped):
            cache[t] = t.evaluate(cache)
            stack.pop()
            continue
        if t._fn is None:
            raise ValueError('placeholder %r was not fed' % (t.name,))
        todo = [i for i in t._inputs if i not in cache]
        if todo:
            stack.extend(todo)
            continue
        cache[t] = t._fn(*[cache[i] for i in t._inputs])
        stack.pop()
    return cache[fetch]


# ---------------------------------------------------------------------------------------------
# variable scopes (tensorflow/python/ops/variable_scope.py, TF 1.x)
# ---------------------------------------------------------------------------------------------
class _VarScope(object):
    def __init__(self, name, reuse):
        self.name, self.reuse = name, reuse


def get_variable_scope():
    return get_default_graph().scope_stack[-1]


@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, reuse=None, **_):
    g = get_default_graph()
    cur = g.scope_stack[-1]
    if isinstance(name_or_scope, _VarScope):
        new_name, by_object = name_or_scope.name, True
        inherit = name_or_scope.reuse
    else:
        name = name_or_scope
        if name is None:            # _get_unique_variable_scope(default_name)
            full = cur.name + '/' + default_name if cur.name else default_name
            name = default_name
            if g.scope_counts.get(full, 0) > 0:
                idx = 1
                while g.scope_counts.get(full + '_%d' % idx, 0) > 0:
                    idx += 1
                name = default_name + '_%d' % idx
        new_name, by_object = (cur.name + '/' + name if cur.name else name), False
        inherit = cur.reuse
    scope = _VarScope(new_name, inherit if not reuse else True)     # reuse=None/False inherits
    g.scope_counts[new_name] = g.scope_counts.get(new_name, 0) + 1   # open_variable_scope
    saved = dict(g.scope_counts) if by_object else None
    g.scope_stack.append(scope)
    try:
        yield scope
    finally:
        g.scope_stack.pop()
        if by_object:
            g.scope_counts = saved
        else:                                                        # close_variable_subscopes
            for k in list(g.scope_counts):
                if k.startswith(new_name + '/'):
                    g.scope_counts[k] = 0


@contextlib.contextmanager
def name_scope(name, *a, **k):
    yield name


class GraphKeys(object):
    TRAINABLE_VARIABLES = 'trainable_variables'
    VARIABLES = 'variables'
    GLOBAL_VARIABLES = 'variables'


_init_rng = np.random.RandomState(0)


def constant_initializer(value=0.0):
    return lambda shape: np.full(shape, value, dtype=_FLOAT)


def _glorot_uniform(shape):
    rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    limit = np.sqrt(6.0 / (rf * shape[-2] + rf * shape[-1]))
    return _init_rng.uniform(-limit, limit, size=shape).astype(_FLOAT)


def get_variable(name, shape=None, initializer=None, dtype=float32, **_):
    g = get_default_graph()
    sc = g.scope_stack[-1]
    full = sc.name + '/' + name if sc.name else name
    shape = tuple(int(d) for d in shape)
    if sc.reuse:
        if full not in g.variables:
            raise ValueError('Variable %s does not exist, or was not created with tf.get_variable()' % full)
        return g.variables[full]
    if full in g.variables:
        raise ValueError('Variable %s already exists, disallowed. Did you mean to set reuse=True?' % full)
    v = Variable(full, shape, (initializer or _glorot_uniform)(shape))
    g.variables[full] = v
    g.add_to_collection(GraphKeys.TRAINABLE_VARIABLES, v)
    g.add_to_collection(GraphKeys.VARIABLES, v)
    return v


def get_collection(name, scope=None):
    return get_default_graph().get_collection(name, scope)


def add_to_collection(name, value):
    get_default_graph().add_to_collection(name, value)


def global_variables_initializer():
    return Operation('init')


def placeholder(dtype, shape=None, name=None):
    g = get_default_graph()
    base = name or 'Placeholder'
    sc = g.scope_stack[-1].name
    full = sc + '/' + base if sc else base
    n = g.placeholder_names.get(full, 0)
    g.placeholder_names[full] = n + 1
    if n:
        full = '%s_%d' % (full, n)
    t = Tensor(None, [], shape, name=full + ':0')
    g.placeholders[t.name] = t
    return t


# ---------------------------------------------------------------------------------------------
# array / math ops
# ---------------------------------------------------------------------------------------------
def shape(x):
    x = as_tensor(x)
    return Tensor(lambda v: np.asarray(np.shape(v), dtype=np.int64), [x], (None if x._shape is None else len(x._shape),))


def expand_dims(x, axis):
    x = as_tensor(x)
    shp = None
    if x._shape is not None:
        shp = list(x._shape)
        shp.insert(axis if axis >= 0 else len(shp) + 1 + axis, 1)
    return Tensor(lambda v: np.expand_dims(v, axis), [x], shp)


def squeeze(x, axis=None):
    x = as_tensor(x)
    ax = tuple(axis) if isinstance(axis, (list, tuple)) else axis
    shp = None
    if x._shape is not None and ax is not None:
        axs = [a % len(x._shape) for a in (ax if isinstance(ax, tuple) else (ax,))]
        shp = [d for i, d in enumerate(x._shape) if i not in axs]
    return Tensor(lambda v: np.squeeze(v, axis=ax), [x], shp)


def pad(x, paddings):
    x = as_tensor(x)
    shp = None if x._shape is None else [None if d is None else d + p[0] + p[1] for d, p in zip(x._shape, paddings)]
    return Tensor(lambda v: np.pad(v, paddings), [x], shp)


def _resolve(items):
    """list whose entries may be tensors -> (tensor inputs, builder(values) -> python list)"""
    tens = [e for e in items if isinstance(e, Tensor)]

    def build(vals):
        it = iter(vals)
        return [int(next(it)) if isinstance(e, Tensor) else e for e in items]
    return tens, build


def slice(x, begin, size):
    x = as_tensor(x)
    tens, build = _resolve(list(size))

    def f(v, *vals):
        sz = build(vals)
        return v[tuple(np.s_[b:(None if s == -1 else b + s)] for b, s in zip(begin, sz))]
    shp = None
    if x._shape is not None:
        shp = [(x._shape[i] if s == -1 and begin[i] == 0 else (None if isinstance(s, Tensor) or s == -1 else s))
               for i, s in enumerate(size)]
    return Tensor(f, [x] + tens, shp)


def tile(x, multiples):
    x = as_tensor(x)
    tens, build = _resolve(list(multiples))
    shp = None
    if x._shape is not None:
        shp = [None if (isinstance(m, Tensor) or d is None) else d * m for d, m in zip(x._shape, multiples)]
    return Tensor(lambda v, *vals: np.tile(v, build(vals)), [x] + tens, shp)


def concat(values, axis):
    values = [as_tensor(v) for v in values]
    shp = None
    if all(v._shape is not None for v in values):
        shp = list(values[0]._shape)
        dims = [v._shape[axis] for v in values]
        shp[axis] = None if any(d is None for d in dims) else sum(dims)
    return Tensor(lambda *v: np.concatenate(v, axis=axis), values, shp)


def reshape(x, shp):
    x = as_tensor(x)
    tens, build = _resolve(list(shp))
    return Tensor(lambda v, *vals: np.reshape(v, build(vals)), [x] + tens,
                  [None if (isinstance(s, Tensor) or s == -1) else s for s in shp])


def _fill(value):
    def f(shp, dtype=float32):
        tens, build = _resolve(list(shp))
        return Tensor(lambda *vals: np.full(build(vals), value, dtype=_np_dtype(dtype)), tens,
                      [None if isinstance(s, Tensor) else s for s in shp])
    return f


ones, zeros = _fill(1.0), _fill(0.0)


def _reduce(npf):
    def f(x, axis=None, keepdims=False, keep_dims=None):
        if keep_dims is not None:
            keepdims = keep_dims
        x = as_tensor(x)
        ax = tuple(axis) if isinstance(axis, (list, tuple)) else axis
        shp = None
        if x._shape is not None:
            if ax is None:
                shp = [1] * len(x._shape) if keepdims else []
            else:
                axs = [a % len(x._shape) for a in (ax if isinstance(ax, tuple) else (ax,))]
                shp = [(1 if i in axs else d) for i, d in enumerate(x._shape) if keepdims or i not in axs]
        return Tensor(lambda v: npf(v, axis=ax, keepdims=keepdims), [x], shp)
    return f


reduce_sum, reduce_max, reduce_mean = _reduce(np.sum), _reduce(np.max), _reduce(np.mean)

exp = lambda x: _unary(np.exp, x)
log = lambda x: _unary(np.log, x)
sqrt = lambda x: _unary(np.sqrt, x)
sign = lambda x: _unary(np.sign, x)
abs = lambda x: _unary(np.abs, x)          # complex input -> magnitude, as tf.abs
log1p = lambda x: _unary(np.log1p, x)
maximum = lambda a, b: _binary(np.maximum, a, b)
minimum = lambda a, b: _binary(np.minimum, a, b)
pow = lambda a, b: _binary(np.power, a, b)
to_float = lambda x: _unary(lambda v: np.asarray(v, dtype=_FLOAT), x)
to_int32 = lambda x: _unary(lambda v: np.asarray(v).astype(np.int32), x)
cast = lambda x, dtype: _unary(lambda v: np.asarray(v).astype(_np_dtype(dtype)), x)
clip_by_value = lambda x, lo, hi: _unary(lambda v: np.clip(v, lo, hi), x)


def stop_gradient(x):
    x = as_tensor(x)
    return Tensor(lambda v: v, [x], x._shape, stops_gradient=True)


def where(c, a, b):
    c, a, b = as_tensor(c), as_tensor(a), as_tensor(b)
    return Tensor(lambda cv, av, bv: np.where(cv, av, bv), [c, a, b], _bshape(a._shape, b._shape))


select = where


def argmax(x, axis=None):
    x = as_tensor(x)
    shp = None if x._shape is None else [d for i, d in enumerate(x._shape) if i != axis % len(x._shape)]
    return Tensor(lambda v: np.argmax(v, axis=axis), [x], shp)


def one_hot(idx, depth, dtype=float32):
    idx = as_tensor(idx)
    return Tensor(lambda v: (np.asarray(v)[..., None] == np.arange(depth)).astype(_np_dtype(dtype)), [idx],
                  None if idx._shape is None else list(idx._shape) + [depth])


def norm(x, ord='euclidean', axis=None):
    assert ord == 'euclidean' and axis is None
    return _unary(lambda v: np.sqrt(np.sum(np.abs(v) ** 2)), x) if True else None


def _default_uniform(shp, minval, maxval):
    return _init_rng.uniform(minval, maxval, size=shp)


_random_hook = _default_uniform     # tests replace this to inject the reference's uniforms


def random_uniform(shp, minval=0.0, maxval=1.0, dtype=float32):
    shp = as_tensor(shp)
    return Tensor(lambda s: np.asarray(_random_hook(tuple(int(d) for d in s), minval, maxval), dtype=_FLOAT), [shp], None)


def multinomial(*a, **k):
    raise NotImplementedError('tf.multinomial is not on the hot path')


# ---------------------------------------------------------------------------------------------
# tf.nn / tf.layers / tf.image / tf.contrib
# ---------------------------------------------------------------------------------------------
def _sigmoid(v):
    return np.where(v >= 0, 1.0 / (1.0 + np.exp(-np.abs(v))), np.exp(-np.abs(v)) / (1.0 + np.exp(-np.abs(v))))


def _softplus(v):
    return np.maximum(v, 0.0) + np.log1p(np.exp(-np.abs(v)))


def _conv1d_valid(x, w, dilation=1):
    """cross-correlation: out[b, t, :] = sum_k x[b, t + k*d, :] @ w[k]   (tf.nn.convolution, NWC, VALID)"""
    K = w.shape[0]
    Tout = x.shape[1] - dilation * (K - 1)
    out = 0.0
    for k in range(K):
        out = out + x[:, k * dilation:k * dilation + Tout, :] @ w[k]
    return out


def _convolution(input, filter, padding, strides=None, dilation_rate=None, name=None, data_format=None):
    assert padding == 'VALID'
    d = 1 if dilation_rate is None else int(dilation_rate[0])
    x, w = as_tensor(input), as_tensor(filter)
    shp = None
    if x._shape is not None and w._shape is not None:
        K = w._shape[0]
        shp = [x._shape[0], None if x._shape[1] is None else x._shape[1] - d * (K - 1), w._shape[2]]
    return Tensor(lambda xv, wv: _conv1d_valid(xv, wv, d), [x, w], shp)


def _pool(input, window_shape, pooling_type, padding, strides=None, **_):
    assert pooling_type == 'AVG' and padding == 'VALID'
    w, s = int(window_shape[0]), int(strides[0])
    x = as_tensor(input)

    def f(v):
        n = (v.shape[1] - w) // s + 1
        return np.stack([v[:, i * s:i * s + w, :].mean(axis=1) for i in range(n)], axis=1)
    return Tensor(f, [x], None if x._shape is None else [x._shape[0], None, x._shape[2]])


def _softmax(v, axis=-1):
    e = np.exp(v - v.max(axis=axis, keepdims=True))
    return e / e.sum(axis=axis, keepdims=True)


nn = types.SimpleNamespace(
    convolution=_convolution,
    pool=_pool,
    tanh=lambda x: _unary(np.tanh, x),
    sigmoid=lambda x: _unary(_sigmoid, x),
    relu=lambda x: _unary(lambda v: np.maximum(v, 0.0), x),
    softplus=lambda x: _unary(_softplus, x),
    softmax=lambda x, axis=-1: _unary(lambda v: _softmax(v, axis), x),
    log_softmax=lambda x, axis=-1: _unary(lambda v: np.log(_softmax(v, axis)), x),
    softmax_cross_entropy_with_logits_v2=lambda logits, labels: Tensor(
        lambda l, y: -(y * np.log(_softmax(l))).sum(-1), [as_tensor(logits), as_tensor(labels)], None),
)


def _layers_conv1d(inputs, filters, kernel_size, strides=1, padding='valid', use_bias=True, name=None, reuse=None, **_):
    """tf.layers.conv1d: Conv1D layer, default name 'conv1d' made unique per enclosing variable scope."""
    x = as_tensor(inputs)
    K = int(kernel_size[0] if isinstance(kernel_size, (list, tuple)) else kernel_size)
    assert strides == 1
    cin = int(x._shape[-1])
    with variable_scope(name, default_name='conv1d', reuse=reuse):
        kernel = get_variable('kernel', [K, cin, filters], initializer=_glorot_uniform)
        bias = get_variable('bias', [filters], initializer=constant_initializer(0.0)) if use_bias else None
    pad_total = K - 1 if padding.upper() == 'SAME' else 0
    pl = pad_total // 2
    pr = pad_total - pl

    def f(xv, kv, *bv):
        if pad_total:
            xv = np.pad(xv, [[0, 0], [pl, pr], [0, 0]])
        out = _conv1d_valid(xv, kv, 1)
        return out + bv[0] if bv else out
    shp = [x._shape[0], x._shape[1] if padding.upper() == 'SAME' else None, filters]
    return Tensor(f, [x, kernel] + ([bias] if use_bias else []), shp)


layers = types.SimpleNamespace(conv1d=_layers_conv1d)


def _resize_nearest_neighbor(images, size, align_corners=False):
    assert not align_corners
    x = as_tensor(images)
    tens, build = _resolve(list(size))

    def f(v, *vals):
        oh, ow = build(vals)
        ih, iw = v.shape[1], v.shape[2]
        # ResizeNearestNeighbor kernel: in = min(floor(out * (in_size / out_size)), in_size - 1), float scale
        hs = np.minimum(np.floor(np.arange(oh) * np.float32(ih / np.float32(oh))).astype(np.int64), ih - 1)
        ws = np.minimum(np.floor(np.arange(ow) * np.float32(iw / np.float32(ow))).astype(np.int64), iw - 1)
        return v[:, hs][:, :, ws]
    shp = None if x._shape is None else [x._shape[0], None if isinstance(size[0], Tensor) else size[0],
                                         None if isinstance(size[1], Tensor) else size[1], x._shape[3]]
    return Tensor(f, [x] + tens, shp)


image = types.SimpleNamespace(resize_nearest_neighbor=_resize_nearest_neighbor)


def _stft(signals, frame_length, frame_step, fft_length=None, pad_end=False):
    """tf.contrib.signal.stft: frames of frame_length every frame_step (no end padding), periodic Hann window,
    rfft of length fft_length (default: the next power of two >= frame_length)."""
    assert not pad_end
    if fft_length is None:
        fft_length = 1 << int(np.ceil(np.log2(frame_length)))
    x = as_tensor(signals)
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(frame_length) / frame_length)

    def f(v):
        n = 1 + (v.shape[-1] - frame_length) // frame_step if v.shape[-1] >= frame_length else 0
        frames = np.stack([v[..., i * frame_step:i * frame_step + frame_length] for i in range(n)], axis=-2) \
            if n else np.zeros(v.shape[:-1] + (0, frame_length))
        return np.fft.rfft(frames * win, n=fft_length, axis=-1)
    return Tensor(f, [x], None)


contrib = types.SimpleNamespace(
    layers=types.SimpleNamespace(xavier_initializer=lambda: _glorot_uniform),
    signal=types.SimpleNamespace(stft=_stft),
)


# ---------------------------------------------------------------------------------------------
# sessions, savers, optimizers, meta graphs
# ---------------------------------------------------------------------------------------------
class Session(object):
    def __init__(self, graph=None, **_):
        self.graph = graph or get_default_graph()

    def run(self, fetches, feed_dict=None):
        feeds = {}
        for k, v in (feed_dict or {}).items():
            arr = np.asarray(v)
            feeds[k] = arr.astype(_FLOAT) if arr.dtype.kind == 'f' else arr

        def one(f):
            if isinstance(f, Operation):
                if f.what == 'init':
                    return None
                raise NotImplementedError('the TensorFlow stand-in builds %s but cannot run it' % f.what)
            if isinstance(f, (list, tuple)):
                return [one(e) for e in f]
            return evaluate(as_tensor(f), feeds)
        return one(fetches)

    @contextlib.contextmanager
    def as_default(self):
        _session_stack.append(self)
        try:
            yield self
        finally:
            _session_stack.pop()

    def __enter__(self):
        _session_stack.append(self)
        return self

    def __exit__(self, *a):
        _session_stack.pop()

    def close(self):
        pass


def get_default_session():
    return _session_stack[-1]


class errors(object):
    class NotFoundError(Exception):
        pass


_META_REGISTRY = {}      # '<prefix>.meta' -> Graph (import_meta_graph works inside one process)


class _Saver(object):
    def __init__(self, var_list=None, graph=None):
        self.graph = graph or get_default_graph()
        self.var_list = list(var_list) if var_list is not None else list(self.graph.variables.values())

    def save(self, sess, path, global_step=None):
        prefix = path if global_step is None else '%s-%d' % (path, global_step)
        with open(prefix + '.data', 'wb') as f:
            pickle.dump({v.var_name: v.value for v in self.var_list}, f)
        open(prefix + '.meta', 'wb').close()
        _META_REGISTRY[os.path.abspath(prefix + '.meta')] = self.graph
        with open(os.path.join(os.path.dirname(prefix), 'checkpoint'), 'w') as f:
            f.write('model_checkpoint_path: "%s"\n' % os.path.basename(prefix))
        return prefix

    def restore(self, sess, prefix):
        if not os.path.exists(prefix + '.data'):
            raise errors.NotFoundError(prefix)
        with open(prefix + '.data', 'rb') as f:
            vals = pickle.load(f)
        for v in self.var_list:
            if v.var_name not in vals:
                raise errors.NotFoundError('Key %s not found in checkpoint' % v.var_name)
            v.value = np.asarray(vals[v.var_name], dtype=_FLOAT)


class _CheckpointState(object):
    def __init__(self, path):
        self.model_checkpoint_path = path


def _get_checkpoint_state(logdir):
    f = os.path.join(logdir, 'checkpoint')
    if not os.path.exists(f):
        return None
    line = open(f).readline()
    return _CheckpointState(os.path.join(logdir, line.split('"')[1]))


class _Remapped(Tensor):
    """A tensor of an imported graph evaluated with some of that graph's placeholders replaced by tensors of the
    importing graph (``tf.train.import_meta_graph(..., input_map=...)``)."""

    def __init__(self, inner, mapping):
        Tensor.__init__(self, None, list(mapping.values()), inner._shape, name=inner.name)
        self.inner, self.mapping = inner, mapping
        self._stops_gradient = False

    def evaluate(self, cache):
        feeds = {}
        for src, dst in self.mapping.items():
            try:
                feeds[src] = evaluate(dst, cache)
            except ValueError:
                pass                      # mapped input not fed: only an error if the fetch needs it
        for k, v in cache.items():        # direct feeds of imported placeholders (wrappers of them)
            if isinstance(k, _Remapped) and k.inner._fn is None and not k.inner._is_variable:
                feeds[k.inner] = v
        return evaluate(self.inner, feeds)


def _import_meta_graph(meta_path, input_map=None, **_):
    src = _META_REGISTRY.get(os.path.abspath(meta_path))
    if src is None:
        raise errors.NotFoundError('no graph was saved as %s in this process' % meta_path)
    g = get_default_graph()
    mapping = {}
    for name, dst in (input_map or {}).items():
        if name not in src.placeholders:
            raise ValueError('input_map key %r is not a tensor of the imported graph' % name)
        mapping[src.placeholders[name]] = as_tensor(dst)
    for cname, items in src.collections.items():
        for it in items:
            if isinstance(it, Variable):
                continue
            g.add_to_collection(cname, _Remapped(it, mapping) if isinstance(it, Tensor) else it)
    return _Saver(list(src.variables.values()), graph=src)


class _Grad(Tensor):
    def __init__(self, var):
        Tensor.__init__(self, self._no, [], var._shape)

    @staticmethod
    def _no():
        raise NotImplementedError('the TensorFlow stand-in does not differentiate')


def _reachable_variables(t):
    seen, stack, out = set(), [as_tensor(t)], set()
    while stack:
        n = stack.pop()
        if id(n) in seen:
            continue
        seen.add(id(n))
        if n._is_variable:
            out.add(n)
        if n._stops_gradient:
            continue
        if isinstance(n, _Remapped):
            stack.append(n.inner)
            # (placeholders of the imported graph are leaves there; the mapped tensors feed them)
        stack.extend(n._inputs)
    return out


class _AdamOptimizer(object):
    def __init__(self, learning_rate=0.001, **_):
        self.learning_rate = learning_rate

    def compute_gradients(self, loss, var_list=None):
        var_list = list(var_list) if var_list is not None else get_collection(GraphKeys.TRAINABLE_VARIABLES)
        live = _reachable_variables(loss)
        return [((_Grad(v) if v in live else None), v) for v in var_list]

    def apply_gradients(self, grads_and_vars, **_):
        return Operation('apply_gradients')

    def minimize(self, loss, var_list=None, **_):
        return Operation('minimize')


def clip_by_global_norm(t_list, clip_norm):
    return [None if t is None else t for t in t_list], Operation('global_norm')


train = types.SimpleNamespace(
    Saver=_Saver,
    AdamOptimizer=_AdamOptimizer,
    get_checkpoint_state=_get_checkpoint_state,
    import_meta_graph=_import_meta_graph,
)
