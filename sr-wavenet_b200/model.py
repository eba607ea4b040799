"""Python-facing mirror of the reference's ``model.py`` for the hot path: same class names,
constructor arguments and method names/argument meaning as ``WaveNetAutoEncoder``
(model.py:75-285) and ``ParallelWaveNet`` (model.py:290-656).  Each former ``sess.run`` is one
call into libsrwn.so (include/srwn.h).  Like the reference's ``feed_dict`` boundary, methods take
NumPy arrays (or torch tensors) and return NumPy arrays; CUDA tensors in -> CUDA tensors out,
with no host copies (used for device-resident benchmarking).

The teacher encoder (``encode`` / ``reconstruct``) runs through its own handle
(``srwn_encoder_*``); teacher training (``train``) raises NotImplementedError.
"""
import ctypes
import os
import time

import numpy as np
import torch

from . import _lib, synth


def _as_i32_array(values):
    arr = (ctypes.c_int32 * len(values))(*[int(v) for v in values])
    return arr


class _Engine(object):
    """Owns one libsrwn handle plus the caller-side workspace and pinned staging buffers."""

    def __init__(self, kind, dilations, filter_width, dilation_channels, skip_channels, cond_channels,
                 pool_stride, num_mixtures=0, num_flows=0):
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("the SR-WaveNet hot path needs a CUDA device (no CPU fallback)")
        self._dil = _as_i32_array(dilations)
        cfg = _lib.Config(kind, len(dilations), self._dil, filter_width, dilation_channels, skip_channels,
                          cond_channels, pool_stride, num_mixtures, num_flows)
        h = ctypes.c_void_p()
        _lib.check(lib.srwn_create(ctypes.byref(cfg), ctypes.byref(h)))
        self.lib, self.h, self.kind = lib, h, kind
        self.cond_channels = cond_channels
        self.pool_stride = pool_stride
        self._ws = None
        self._pinned = {}

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.srwn_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_weights(self, weights, prefix_filter=None):
        for name, arr in weights.items():
            if prefix_filter and not name.startswith(prefix_filter):
                continue
            a = np.ascontiguousarray(arr, dtype=np.float32)
            shape = (ctypes.c_int64 * a.ndim)(*a.shape)
            _lib.check(self.lib.srwn_set_weight(self.h, name.encode(), a.ctypes.data_as(ctypes.c_void_p),
                                                shape, a.ndim))
        _lib.check(self.lib.srwn_commit_weights(self.h, torch.cuda.current_stream().cuda_stream))

    def get_weight(self, name, shape):
        out = np.empty(shape, dtype=np.float32)
        _lib.check(self.lib.srwn_get_weight(self.h, name.encode(), out.ctypes.data_as(ctypes.c_void_p), out.size))
        return out

    def set_profiling(self, enable):
        _lib.check(self.lib.srwn_set_profiling(self.h, int(bool(enable))))

    def last_kernel_ms(self):
        """(elapsed ms, launches, kernel name) of the dominant kernel(s) of the last call."""
        ms, n, name = ctypes.c_float(), ctypes.c_int32(), ctypes.c_char_p()
        _lib.check(self.lib.srwn_last_kernel_ms(self.h, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(name)))
        return ms.value, n.value, name.value.decode()

    def random_uniform(self, shape, seed, stream_id, lo, hi):
        """U(lo, hi) draws on the device (Philox, csrc/philox.cuh) for the sampler's noise (ops.py:187, 196)."""
        out = torch.empty(shape, dtype=torch.float32, device="cuda")
        _lib.check(self.lib.srwn_random_uniform(out.data_ptr(), out.numel(), seed, stream_id, lo, hi, _stream()))
        return out

    def set_team_size(self, ctas_per_team):
        """CTAs per team of the fused kernel (0 = chosen per (B, T)); results do not depend on it."""
        _lib.check(self.lib.srwn_set_team_size(self.h, int(ctas_per_team)))

    def last_partition(self):
        """(teams, CTAs per team) of the last fused call."""
        t, g = ctypes.c_int32(), ctypes.c_int32()
        _lib.check(self.lib.srwn_last_partition(self.h, ctypes.byref(t), ctypes.byref(g)))
        return t.value, g.value

    def check_async(self, op, B, T, precision):
        """Synchronises and raises if the last fused launch aborted on the device."""
        ws, wsn = self.workspace(op, B, T, precision)
        _lib.check(self.lib.srwn_check_async_error(self.h, op, B, T, precision, ws, wsn,
                                                   torch.cuda.current_stream().cuda_stream))

    def workspace(self, op, B, T, precision):
        n = ctypes.c_size_t()
        _lib.check(self.lib.srwn_workspace_bytes(self.h, op, B, T, precision, ctypes.byref(n)))
        if self._ws is None or self._ws.numel() < n.value:
            self._ws = None
            self._ws = torch.empty(max(n.value, 256), dtype=torch.uint8, device="cuda")
        return self._ws.data_ptr(), self._ws.numel()

    def to_device(self, a, key):
        """NumPy / CPU tensor -> CUDA fp32 tensor through a persistent pinned staging buffer."""
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return a.float().contiguous(), True
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        t = t.float().contiguous()
        if not t.is_pinned():
            buf, ev = self._pinned.get(key, (None, None))
            if buf is None or buf.shape != t.shape:
                buf, ev = torch.empty(t.shape, dtype=torch.float32, pin_memory=True), torch.cuda.Event()
                self._pinned[key] = (buf, ev)
            else:
                ev.synchronize()        # the previous H2D copy out of this staging buffer may still be queued
            buf.copy_(t)
            d = buf.to("cuda", non_blocking=True)
            ev.record()
            return d, False
        return t.to("cuda", non_blocking=True), False

    def to_host(self, t, on_device):
        if on_device:
            return t
        key = ("out", tuple(t.shape), t.dtype)
        buf = self._pinned.get(key)
        if buf is None:
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            self._pinned[key] = buf
        buf.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return buf.numpy().copy()


class _EncoderEngine(object):
    """Owns one libsrwn encoder handle (include/srwn.h, srwn_encoder_*) and its workspace."""

    def __init__(self, n_layers, filter_width, encoder_channels, skip_channels, latent_channels, pool_stride):
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("the SR-WaveNet encoder needs a CUDA device (no CPU fallback)")
        cfg = _lib.EncoderConfig(n_layers, filter_width, encoder_channels, skip_channels, latent_channels,
                                 pool_stride)
        h = ctypes.c_void_p()
        _lib.check(lib.srwn_encoder_create(ctypes.byref(cfg), ctypes.byref(h)))
        self.lib, self.h = lib, h
        self.latent_channels, self.pool_stride = latent_channels, pool_stride
        self._ws = None
        self.committed = False

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.srwn_encoder_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_weights(self, weights):
        for name, arr in weights.items():
            a = np.ascontiguousarray(arr, dtype=np.float32)
            shape = (ctypes.c_int64 * a.ndim)(*a.shape)
            _lib.check(self.lib.srwn_encoder_set_weight(self.h, name.encode(), a.ctypes.data_as(ctypes.c_void_p),
                                                        shape, a.ndim))
        _lib.check(self.lib.srwn_encoder_commit(self.h, torch.cuda.current_stream().cuda_stream))
        self.committed = True

    def supports(self, prec):
        return bool(self.lib.srwn_encoder_supports(self.h, prec))

    def set_profiling(self, enable):
        _lib.check(self.lib.srwn_encoder_set_profiling(self.h, int(bool(enable))))

    def last_ms(self):
        ms = ctypes.c_float()
        _lib.check(self.lib.srwn_encoder_last_ms(self.h, ctypes.byref(ms)))
        return ms.value

    def workspace(self, B, T, prec):
        n = ctypes.c_size_t()
        _lib.check(self.lib.srwn_encoder_workspace_bytes(self.h, B, T, prec, ctypes.byref(n)))
        if self._ws is None or self._ws.numel() < n.value:
            self._ws = None
            self._ws = torch.empty(max(n.value, 256), dtype=torch.uint8, device="cuda")
        return self._ws.data_ptr(), self._ws.numel()

    def encode(self, x, prec, check=True):
        """x: CUDA fp32 [B,T] -> CUDA fp32 [B, T // P, latent]."""
        B, T = x.shape
        ws, wsn = self.workspace(B, T, prec)
        out = torch.empty(B, T // self.pool_stride, self.latent_channels, dtype=torch.float32, device="cuda")
        _lib.check(self.lib.srwn_teacher_encode(self.h, x.data_ptr(), out.data_ptr(), B, T, prec, ws, wsn, _stream()))
        if check and prec != _lib.FP32:
            _lib.check(self.lib.srwn_encoder_check_async_error(self.h, B, T, prec, ws, wsn, _stream()))
        return out


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _with_conditions(encoding, conditions, condition_size):
    """model.py:161-167 / 495-500: tile the global condition over frames and concatenate."""
    if condition_size > 0:
        if conditions is None:
            raise ValueError("condition_size > 0 needs `conditions`")
        if isinstance(encoding, torch.Tensor):
            c = torch.as_tensor(conditions, dtype=torch.float32, device=encoding.device)
            return torch.cat([encoding, c[:, None, :].expand(-1, encoding.shape[1], -1)], dim=2)
        c = np.asarray(conditions, dtype=np.float32)
        return np.concatenate([encoding, np.tile(c[:, None, :], [1, encoding.shape[1], 1])], axis=2)
    return encoding


class _CheckpointMixin(object):
    """Stand-in for tf.train.Saver (model.py:217-239).  ``save`` writes ``<logdir>/model.ckpt-<step>.npz`` (name-keyed,
    TF variable names) plus the ``checkpoint`` state file, throttled to once per 60 s unless ``force``; with
    ``checkpoint_format = "tf"`` it writes a TensorFlow tensor bundle (``.index`` + ``.data-00000-of-00001``) instead.
    ``load`` restores whichever of the two the state file points to, so a checkpoint directory written by the
    reference's Saver loads as it is (``tf_checkpoint.py``; optimizer slots and other variables the graph does not have
    are ignored)."""

    checkpoint_format = "npz"

    def _save(self, logdir, global_step, force):
        if force or time.time() - self.last_checkpoint_time > 60:
            if not os.path.isdir(logdir):
                os.makedirs(logdir)
            path = os.path.join(logdir, 'model.ckpt-%d' % global_step)
            if self.checkpoint_format == "tf":
                from . import tf_checkpoint
                tf_checkpoint.write_checkpoint(path, self.get_weights())
            else:
                np.savez(path + '.npz', **self.get_weights())
            if isinstance(self, WaveNetAutoEncoder):      # tf.train.Saver.save also exports the meta-graph the student imports
                from . import tf_meta
                tf_meta.write_meta(path + '.meta', *tf_meta.teacher_meta_skeleton(self.name))
            with open(os.path.join(logdir, 'checkpoint'), 'w') as f:
                f.write('model_checkpoint_path: "%s"\n' % os.path.basename(path))
            self.last_checkpoint_time = time.time()
            return True
        return False

    def _load(self, logdir):
        if logdir is not None and os.path.exists(logdir):
            state = os.path.join(logdir, 'checkpoint')
            if os.path.exists(state):
                with open(state) as f:
                    name = f.readline().split('"')[1]
                path = name if os.path.isabs(name) else os.path.join(logdir, name)
                if os.path.exists(path + '.npz'):
                    with np.load(path + '.npz') as z:
                        self.set_weights({k: z[k] for k in z.files})
                elif os.path.exists(path + '.index'):
                    from . import tf_checkpoint
                    known = self.get_weights()
                    found = {k: v for k, v in tf_checkpoint.read_checkpoint(path).items()
                             if k in known and tuple(v.shape) == tuple(known[k].shape)}
                    if not found:
                        print('No variable of this graph in %s' % path)
                        return False
                    self.set_weights(found)
                else:
                    print('Could not find checkpoint at %s' % path)
                    return False
                print('Restoring previous session')
                return True
        return None


class WaveNetAutoEncoder(_CheckpointMixin):
    """model.py:75-285: decoder (the hot path) and encoder; constructor signature kept verbatim."""

    def __init__(self, input_size, condition_size, num_mixtures, dilations, filter_width=2,
                 encoder_channels=128, dilation_channels=32, skip_channels=256, latent_channels=16,
                 pool_stride=512, name='WaveNetAutoEncoder', learning_rate=0.001):
        self.input_size = input_size
        self.condition_size = condition_size
        self.num_mixtures = num_mixtures
        self.dilations = list(dilations)
        self.filter_width = filter_width
        self.encoder_channels = encoder_channels
        self.dilation_channels = dilation_channels
        self.skip_channels = skip_channels
        self.latent_channels = latent_channels
        self.pool_stride = pool_stride
        self.name = name
        self.learning_rate = learning_rate
        self.precision = "fp32"
        self.seed = int.from_bytes(os.urandom(7), "little")      # the reference's RNG is unseeded; set for reproducible draws
        self.last_checkpoint_time = time.time()
        self.createNetwork()

    # -- graph construction stand-ins ------------------------------------------------------
    def createNetwork(self):
        """model.py:202-215: the reference builds placeholders + encoder + two decoder instances;
        here it creates the decoder engine and Glorot-initialised variables."""
        self._eng = _Engine(_lib.TEACHER, self.dilations, self.filter_width, self.dilation_channels,
                            self.skip_channels, self.latent_channels + self.condition_size,
                            self.pool_stride, num_mixtures=self.num_mixtures)
        w = synth.make_teacher_weights(self.dilations, self.filter_width, self.dilation_channels,
                                       self.skip_channels, self.latent_channels + self.condition_size,
                                       self.num_mixtures, seed=None, dead_vars=True)
        self._enc_eng = _EncoderEngine(len(self.dilations), self.filter_width, self.encoder_channels,
                                       self.skip_channels, self.latent_channels, self.pool_stride)
        we = synth.make_encoder_weights(len(self.dilations), self.filter_width, self.encoder_channels,
                                        self.skip_channels, self.latent_channels, seed=None, gain=1.0)
        w.update(we)
        for k in w:                      # TF initialises biases to zero (ops.py:18, tf.layers default)
            if k.endswith('Bias') or k.endswith('/bias'):
                w[k][...] = 0
        self._weights = {}
        self.set_weights(w)

    def createEncoder(self, h, reuse=False, precision=None):
        """model.py:137-155 -> encoding [B, T/pool_stride, latent].  h [B,T,1] or [B,T]."""
        h2 = h[..., 0] if getattr(h, "ndim", 2) == 3 else h
        return self.encode(h2, precision=precision)

    def createDecoder(self, truth, encoding, conditions, reuse=False, u1=None, u2=None, precision=None):
        """model.py:158-200 -> (logits [B,T,4M], out [B,T]).  truth [B,T,1] or [B,T]."""
        from . import ops
        truth2 = truth[..., 0] if getattr(truth, "ndim", 2) == 3 else truth
        logits = self.get_logits(truth2, encoding, conditions, precision=precision)
        out = ops.sample_from_discretized_mix_logistic(logits, self.num_mixtures, u1, u2)
        out = out[..., 0]
        if not isinstance(logits, torch.Tensor):
            out = out.cpu().numpy()
        return logits, out

    # -- weights -----------------------------------------------------------------------------
    def set_weights(self, weights):
        """name -> ndarray with TF variable names (``WaveNetAutoEncoder/Decoder/...`` and
        ``WaveNetAutoEncoder/Encoder/...``); either half may be given alone."""
        dec = {k: v for k, v in weights.items() if '/Encoder/' not in k}
        enc = {k: v for k, v in weights.items() if '/Encoder/' in k}
        if dec:
            self._eng.set_weights(dec)
        if enc:
            self._enc_eng.set_weights(enc)
        self._weights.update({k: np.array(v, dtype=np.float32) for k, v in weights.items()})

    def get_weights(self):
        return dict(self._weights)

    def available_precisions(self):
        return ["fp32", "fp16"] if _fused_available(self._eng) else ["fp32"]

    def load(self, logdir):
        return self._load(logdir)

    def save(self, logdir, global_step, force=False):
        return self._save(logdir, global_step, force)

    @classmethod
    def from_checkpoint(cls, logdir, dilations, latent_channels, pool_stride, condition_size=0, filter_width=2,
                        dilation_channels=32, input_size=0):
        """A teacher built from a checkpoint directory alone (what ParallelWaveNet(teacher=<dir>) needs,
        model.py:323-334): mixtures, skip and encoder widths come from the stored variable shapes."""
        state = os.path.join(logdir, 'checkpoint')
        if not os.path.exists(state):
            raise IOError("no teacher checkpoint under %s (expected the files WaveNetAutoEncoder.save writes)" % logdir)
        with open(state) as f:
            name = f.readline().split('"')[1]
        path = name if os.path.isabs(name) else os.path.join(logdir, name)
        L = len(dilations)
        k_head2 = '%sconv1d_%d/kernel' % (synth.TEACHER_PREFIX, 3 * L + 1)
        k_enc0 = '%snc_conv_NC/conv1d/kernel' % synth.ENCODER_PREFIX
        if os.path.exists(path + '.npz'):
            with np.load(path + '.npz') as z:
                head2, enc0 = z[k_head2].shape, z[k_enc0].shape
        else:                                              # a tf.train.Saver checkpoint (tensor bundle)
            from . import tf_checkpoint
            shapes = {n: sh for n, sh, _ in tf_checkpoint.list_variables(path)}
            head2, enc0 = shapes[k_head2], shapes[k_enc0]
        t = cls(input_size=input_size, condition_size=condition_size, num_mixtures=head2[2] // 4, dilations=dilations,
                filter_width=filter_width, encoder_channels=enc0[2], dilation_channels=dilation_channels,
                skip_channels=head2[1], latent_channels=latent_channels, pool_stride=pool_stride)
        if not t.load(logdir):
            raise IOError("could not restore the teacher from %s" % logdir)
        if os.path.exists(path + '.meta'):
            # the part of tf.train.import_meta_graph(..., input_map=...) + get_collection(...)[0] (model.py:326-341) that can
            # fail: the placeholders the student re-wires and the collections it reads must exist in the teacher's graph
            from . import tf_meta
            t.meta_collections = tf_meta.check_teacher_contract(tf_meta.read_meta(path + '.meta'), t.name)
        return t

    # -- sess.run wrappers -------------------------------------------------------------------
    def train(self, inputs, conditions=None):
        raise NotImplementedError("teacher training (model.py:242-248) is outside the hot path")

    def encode(self, inputs, conditions=None, precision=None):
        """model.py:250-255 -> encoding [B, T/pool_stride, latent] (``conditions`` does not enter the
        encoder graph, model.py:212).  A ragged tail T % pool_stride is dropped like the VALID pooling
        of model.py:154 does.  The 16-bit tensor-core path needs T % 128 == 0; other lengths run fp32."""
        eng = self._enc_eng
        x, on_dev = self._eng.to_device(inputs, "x")
        prec = self._prec(precision)
        if prec != _lib.FP32 and (not eng.supports(prec) or x.shape[1] % 128 != 0):
            prec = _lib.FP32
        return self._eng.to_host(eng.encode(x, prec), on_dev)

    def reconstruct(self, inputs, conditions=None, u1=None, u2=None, precision=None):
        """model.py:257-262 -> [B,T]: encoder, then the decoder teacher-forced on the same audio, then
        parallel sampling from the logits (``self.out``)."""
        x, on_dev = self._eng.to_device(inputs, "x")
        enc = self.encode(x, precision=precision)
        out = self.reconstruct_with_encoding(x, enc, conditions, u1=u1, u2=u2, precision=precision)
        return out if on_dev else out.cpu().numpy()

    def _prec(self, precision):
        return _lib.PRECISIONS[precision or self.precision]

    def _finish(self, t, on_dev, op, B, T, prec):
        """Result hand-over.  Host boundary (NumPy in -> NumPy out): the copy synchronises anyway, so a fused launch that
        aborted on the device raises here instead of returning partial data.  Device-resident results are not
        synchronised; an abort then refuses the NEXT call on the handle (pinned abort words, include/srwn.h)."""
        out = self._eng.to_host(t, on_dev)
        if not on_dev and prec != _lib.FP32:
            self._eng.check_async(op, B, T, prec)
        return out

    def get_logits(self, inputs, encoding, conditions=None, precision=None):
        """model.py:279-285 -> logits [B,T,4M]."""
        eng = self._eng
        enc = _with_conditions(encoding, conditions, self.condition_size)
        x, on_dev = eng.to_device(inputs, "x")
        e, _ = eng.to_device(enc, "enc")
        B, T = x.shape
        self._check(e, B, T)
        prec = self._prec(precision)
        ws, wsn = eng.workspace(_lib.OP_TEACHER_LOGITS, B, T, prec)
        logits = torch.empty(B, T, 4 * self.num_mixtures, dtype=torch.float32, device="cuda")
        _lib.check(eng.lib.srwn_teacher_logits(eng.h, x.data_ptr(), e.data_ptr(), logits.data_ptr(), B, T,
                                               prec, ws, wsn, _stream()))
        return self._finish(logits, on_dev, _lib.OP_TEACHER_LOGITS, B, T, prec)

    def nll(self, inputs, encoding, conditions=None, scored=None, sum_all=True, precision=None):
        """``loss_encoding`` of model.py:114-115: teacher-forced mixture-of-logistics negative
        log-likelihood.  ``scored`` (default: ``inputs``) is the audio whose likelihood is taken
        (model.py:374 scores the student's output against logits of the real audio).
        Returns a Python float (sum_all=True, ops.py:172) or [B,T,1] (ops.py:175)."""
        eng = self._eng
        enc = _with_conditions(encoding, conditions, self.condition_size)
        x, on_dev = eng.to_device(inputs, "x")
        e, _ = eng.to_device(enc, "enc")
        xs = x if scored is None else eng.to_device(scored, "xs")[0]
        B, T = x.shape
        self._check(e, B, T)
        prec = self._prec(precision)
        ws, wsn = eng.workspace(_lib.OP_TEACHER_NLL, B, T, prec)
        if sum_all:
            out = torch.empty(1, dtype=torch.float32, device="cuda")
            _lib.check(eng.lib.srwn_teacher_nll(eng.h, x.data_ptr(), e.data_ptr(), xs.data_ptr(), None,
                                                out.data_ptr(), None, B, T, prec, ws, wsn, _stream()))
            return out if on_dev else float(self._finish(out, False, _lib.OP_TEACHER_NLL, B, T, prec)[0])
        out = torch.empty(B, T, 1, dtype=torch.float32, device="cuda")
        _lib.check(eng.lib.srwn_teacher_nll(eng.h, x.data_ptr(), e.data_ptr(), xs.data_ptr(), out.data_ptr(),
                                            None, None, B, T, prec, ws, wsn, _stream()))
        return self._finish(out, on_dev, _lib.OP_TEACHER_NLL, B, T, prec)

    def nll_stream(self, batches, precision=None, depth=2):
        """Scores a sequence of host batches: ``batches`` yields ``(inputs [B,T], encoding [B,T/P,C])`` (NumPy arrays or CPU
        tensors, pinned or not); yields one float per batch, in order -- the value ``nll(inputs, encoding)`` returns.
        Throughput path of a scoring service: the upload of batch i+1 runs on a copy stream while batch i is scored, and the
        4-byte result of batch i is read back while batch i+1 runs, so a step costs max(copy, kernel) instead of their sum.
        ``depth`` device-side input slots (>= 2)."""
        eng = self._eng
        prec = self._prec(precision)
        depth = max(2, int(depth))
        compute = torch.cuda.current_stream()
        copy = getattr(self, "_copy_stream", None) or torch.cuda.Stream()
        self._copy_stream = copy
        slots = [dict(x=None, e=None, px=None, pe=None, up=torch.cuda.Event(), used=torch.cuda.Event(), done=torch.cuda.Event(),
                      out=torch.empty(1, dtype=torch.float32, device="cuda"), host=torch.empty(1, dtype=torch.float32, pin_memory=True))
                 for _ in range(depth)]

        def host32(a):
            t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
            return t.float().contiguous()

        def upload(slot, batch):
            x, e = host32(batch[0]), host32(_with_conditions(batch[1], None, self.condition_size))
            if slot["x"] is None or slot["x"].shape != x.shape or slot["e"].shape != e.shape:
                slot["x"] = torch.empty(x.shape, dtype=torch.float32, device="cuda")
                slot["e"] = torch.empty(e.shape, dtype=torch.float32, device="cuda")
                slot["px"] = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
                slot["pe"] = torch.empty(e.shape, dtype=torch.float32, pin_memory=True)
            slot["used"].synchronize() if slot.get("busy") else None      # the kernel that read this slot last has finished
            src_x, src_e = x, e
            if not x.is_pinned():
                slot["px"].copy_(x); src_x = slot["px"]
            if not e.is_pinned():
                slot["pe"].copy_(e); src_e = slot["pe"]
            with torch.cuda.stream(copy):
                slot["x"].copy_(src_x, non_blocking=True)
                slot["e"].copy_(src_e, non_blocking=True)
                slot["up"].record(copy)

        def launch(slot):
            B, T = slot["x"].shape
            self._check(slot["e"], B, T)
            ws, wsn = eng.workspace(_lib.OP_TEACHER_NLL, B, T, prec)
            compute.wait_event(slot["up"])
            _lib.check(eng.lib.srwn_teacher_nll(eng.h, slot["x"].data_ptr(), slot["e"].data_ptr(), slot["x"].data_ptr(), None,
                                                slot["out"].data_ptr(), None, B, T, prec, ws, wsn, compute.cuda_stream))
            slot["used"].record(compute)
            slot["host"].copy_(slot["out"], non_blocking=True)
            slot["done"].record(compute)
            slot["busy"], slot["shape"] = True, (B, T)

        def collect(slot):
            slot["done"].synchronize()
            if prec != _lib.FP32:                      # abort words of that launch, without draining the launches behind it
                _lib.check(eng.lib.srwn_peek_async_error(eng.h))
            return float(slot["host"][0])

        it = iter(batches)
        pending = []                                  # slots launched, oldest first
        i = 0
        try:
            first = next(it)
        except StopIteration:
            return
        try:
            upload(slots[0], first)
            nxt = 0
            while nxt is not None:
                cur = slots[nxt]
                launch(cur)
                pending.append(cur)
                i += 1
                try:
                    batch = next(it)
                    if len(pending) == depth:         # every slot is in flight: hand the oldest result out before reusing its slot
                        yield collect(pending.pop(0))
                    nxt = i % depth
                    upload(slots[nxt], batch)
                except StopIteration:
                    nxt = None
            while pending:
                yield collect(pending.pop(0))
        finally:
            # the slots' device buffers go back to the allocator when this frame dies: nothing on the copy stream (or, when
            # the consumer stopped early or an error came up, on the compute stream) may still be using them
            copy.synchronize()
            if pending:
                compute.synchronize()

    def reconstruct_with_encoding(self, inputs, encoding, conditions=None, u1=None, u2=None,
                                  precision=None):
        """model.py:264-270 -> [B,T]: one teacher-forced pass + parallel sampling from the logits
        (``out_from_encoding``).  ``u1``/``u2`` inject the sampler's uniforms (ops.py:187,196)."""
        _, out = self.createDecoder(inputs, encoding, conditions, reuse=True, u1=u1, u2=u2,
                                    precision=precision)
        return out

    def generate(self, encoding, conditions=None, u1=None, u2=None, num_samples=None,
                 return_logits=False, zero_last=False, precision="fp32"):
        """Autoregressive generation, the queue-based restatement of teacher.py:153-170:
        x[t] = clip(MoL_sample(logits_t)) with logits_t from x[<t] and encoding[t // pool_stride].
        ``zero_last=True`` reproduces teacher.py:170, which zeroes the final sample.
        ``precision``: "fp32" (parity-grade FFMA kernel) or "fp16" (tensor-core kernel, fp16 operands
        and queue state, fp32 accumulate / residual stream)."""
        eng = self._eng
        enc = _with_conditions(encoding, conditions, self.condition_size)
        e, on_dev = eng.to_device(enc, "enc")
        B = e.shape[0]
        T = num_samples if num_samples is not None else e.shape[1] * self.pool_stride
        self._check(e, B, T)
        lo, hi = 1e-5, 1.0 - 1e-5            # ops.py:187, 196
        if u1 is None or u2 is None:
            self._noise_calls = getattr(self, "_noise_calls", 0) + 1
        u1 = eng.random_uniform((B, T, self.num_mixtures), self.seed, 2 * self._noise_calls, lo, hi) if u1 is None \
            else eng.to_device(u1, "u1")[0]
        u2 = eng.random_uniform((B, T), self.seed, 2 * self._noise_calls + 1, lo, hi) if u2 is None \
            else eng.to_device(u2, "u2")[0]
        prec = _lib.PRECISIONS[precision]
        if not eng.lib.srwn_supports(eng.h, _lib.OP_TEACHER_GENERATE, prec):
            raise RuntimeError("generate(precision=%r) is not supported for this configuration" % precision)
        ws, wsn = eng.workspace(_lib.OP_TEACHER_GENERATE, B, T, prec)
        x = torch.empty(B, T, dtype=torch.float32, device="cuda")
        logits = torch.empty(B, T, 4 * self.num_mixtures, dtype=torch.float32, device="cuda") \
            if return_logits else None
        _lib.check(eng.lib.srwn_teacher_generate(eng.h, e.data_ptr(), u1.data_ptr(), u2.data_ptr(),
                                                 x.data_ptr(), logits.data_ptr() if return_logits else None,
                                                 B, T, prec, ws, wsn, _stream()))
        if prec != _lib.FP32:
            eng.check_async(_lib.OP_TEACHER_GENERATE, B, T, prec)
        if zero_last:
            x[:, T - 1:] = 0            # teacher.py:170
        if return_logits:
            return eng.to_host(x, on_dev), eng.to_host(logits, on_dev)
        return eng.to_host(x, on_dev)

    def _check(self, e, B, T):
        if e.ndim != 3 or e.shape[0] != B or e.shape[2] != self.latent_channels + self.condition_size:
            raise ValueError("encoding must be [B, T/pool_stride, %d]" % (self.latent_channels + self.condition_size))
        if e.shape[1] * self.pool_stride != T:
            raise ValueError("T=%d must equal pool_stride * frames = %d (model.py:183)"
                             % (T, e.shape[1] * self.pool_stride))


def _fused_available(eng):
    op = _lib.OP_TEACHER_LOGITS if eng.kind == _lib.TEACHER else _lib.OP_STUDENT_FORWARD
    return bool(eng.lib.srwn_supports(eng.h, op, _lib.FP16))


class ParallelWaveNet(_CheckpointMixin):
    """model.py:290-656.  ``teacher`` is a checkpoint directory in the reference; here it may also
    be a ``WaveNetAutoEncoder`` instance or None.  Methods keep the explicit ``sess`` first argument
    (ignored)."""

    def __init__(self, input_size, condition_size, dilations, teacher, num_flows=2, filter_width=2,
                 dilation_channels=32, skip_channels=256, latent_channels=16, pool_stride=512,
                 name='ParallelWaveNet', alpha=1.0, beta=1.0, gamma=1.0, learning_rate=0.001):
        self.input_size = input_size
        self.condition_size = condition_size
        self.dilations = list(dilations)
        self.teacher = teacher
        if isinstance(teacher, str):
            # model.py:326-331 imports the teacher's meta-graph from this directory; here the checkpoint written by
            # WaveNetAutoEncoder.save names every variable, and the constructor arguments the graph would carry
            # (mixtures, skip / encoder channels) are read off the variable shapes
            self.teacher = WaveNetAutoEncoder.from_checkpoint(teacher, dilations, latent_channels=latent_channels,
                                                              pool_stride=pool_stride, condition_size=condition_size,
                                                              filter_width=filter_width,
                                                              dilation_channels=dilation_channels, input_size=input_size)
        self.num_flows = num_flows
        self.filter_width = filter_width
        self.dilation_channels = dilation_channels
        self.skip_channels = skip_channels
        self.latent_channels = latent_channels
        self.pool_stride = pool_stride
        self.name = name
        self.alpha, self.beta, self.gamma = alpha, beta, gamma
        self.learning_rate = learning_rate
        self.precision = "fp32"
        self.seed = int.from_bytes(os.urandom(7), "little")      # key of the on-device logistic noise (generate(sess, None, ...))
        self._noise_calls = 0
        self.last_checkpoint_time = time.time()
        self.createNetwork()

    def createNetwork(self):
        """model.py:489-535."""
        self._eng = _Engine(_lib.STUDENT, self.dilations, self.filter_width, self.dilation_channels,
                            self.skip_channels, self.latent_channels + self.condition_size,
                            self.pool_stride, num_flows=self.num_flows)
        w = synth.make_student_weights(self.dilations, self.num_flows, self.filter_width,
                                       self.dilation_channels, self.skip_channels,
                                       self.latent_channels + self.condition_size, seed=None)
        for k in w:
            if k.endswith('Bias') or k.endswith('/bias'):
                w[k][...] = 0
        self._weights = {}
        self.set_weights(w)

    def _flow_model(self, scope):
        """A one-flow network holding the variables of ``scope`` ('Flow<i>'): model.py:457-486 builds each flow as its own
        sub-graph; here it is the same kernels run for a single stack."""
        f = int(str(scope).replace('Flow', ''))
        if not 0 <= f < self.num_flows:
            raise ValueError("scope must be 'Flow0' .. 'Flow%d'" % (self.num_flows - 1))
        cache = self.__dict__.setdefault("_flow_models", {})
        if f not in cache:
            m = ParallelWaveNet(self.input_size, self.condition_size, self.dilations, None, num_flows=1,
                                filter_width=self.filter_width, dilation_channels=self.dilation_channels,
                                skip_channels=self.skip_channels, latent_channels=self.latent_channels,
                                pool_stride=self.pool_stride)
            cache[f] = m
        src, dst = synth.student_prefix(f), synth.student_prefix(0)
        cache[f].set_weights({dst + k[len(src):]: v for k, v in self.get_weights().items() if k.startswith(src)})
        return cache[f]

    def createFlow(self, inputs, encoding, scope, precision=None):
        """model.py:457-486 -> (scale, mean, out), each [B,T,1]: out = inputs * scale + mean with
        scale = exp(params[..., 0]), mean = params[..., 1] (no clamp on the log-scale)."""
        x = inputs[..., 0] if getattr(inputs, "ndim", 2) == 3 else inputs
        r = self._flow_model(scope).forward_all(x, encoding, precision=precision)
        return r["s_tot"][..., None], r["mu_tot"][..., None], r["x_last"][..., None]

    def createPartialFlow(self, inputs, encoding, scope, precision=None):
        """model.py:415-454 -> params [B,T,2] = (log scale, mean) of one flow."""
        scale, mean, _ = self.createFlow(inputs, encoding, scope, precision=precision)
        log = torch.log if isinstance(scale, torch.Tensor) else np.log
        cat = torch.cat if isinstance(scale, torch.Tensor) else np.concatenate
        return cat([log(scale), mean], -1)

    def set_weights(self, weights):
        """name -> ndarray with TF variable names (``ParallelWaveNet/Flow{f}/Flow{f}/...``)."""
        self._eng.set_weights(weights)
        self._weights.update({k: np.array(v, dtype=np.float32) for k, v in weights.items()})

    def get_weights(self):
        if getattr(self, "_packed_stale", False):        # trained on the device since the last read-back
            self.sync_weights()
        return dict(self._weights)

    def available_precisions(self):
        return ["fp32", "fp16"] if _fused_available(self._eng) else ["fp32"]

    def load(self, sess, logdir):
        if isinstance(self.teacher, str):
            pass   # the reference restores the imported teacher meta-graph here (model.py:543-544)
        return self._load(logdir)

    def save(self, sess, logdir, global_step, force=False):
        return self._save(logdir, global_step, force)

    def _forward(self, inputs, encoding, conditions, precision=None, want=("out",)):
        eng = self._eng
        enc = _with_conditions(encoding, conditions, self.condition_size)
        e, on_dev = eng.to_device(enc, "enc")
        if inputs is None:          # student.py:104 / :172 draws the logistic noise on the host; here it is drawn on the device
            B, T = e.shape[0], e.shape[1] * self.pool_stride
            z = None
        else:
            z, on_dev = eng.to_device(inputs, "z")
            B, T = z.shape
        if e.ndim != 3 or e.shape[0] != B or e.shape[1] * self.pool_stride != T:
            raise ValueError("encoding must be [B, T/pool_stride, C] with T == pool_stride * frames")
        prec = _lib.PRECISIONS[precision or self.precision]
        if prec != _lib.FP32 and getattr(self, "_packed_stale", False):
            self.sync_weights()
        ws, wsn = eng.workspace(_lib.OP_STUDENT_FORWARD, B, T, prec)
        bufs = {k: torch.empty(B, T, dtype=torch.float32, device="cuda") for k in set(want) | {"out"}}
        g = lambda k: bufs[k].data_ptr() if k in bufs else None
        if z is None:
            self._noise_calls += 1
            _lib.check(eng.lib.srwn_student_sample(eng.h, self.seed, self._noise_calls, e.data_ptr(), g("out"), g("s_tot"),
                                                   g("mu_tot"), g("x_last"), g("z"), B, T, prec, ws, wsn, _stream()))
        else:
            bufs.pop("z", None)
            _lib.check(eng.lib.srwn_student_forward(eng.h, z.data_ptr(), e.data_ptr(), g("out"), g("s_tot"),
                                                    g("mu_tot"), g("x_last"), B, T, prec, ws, wsn, _stream()))
        self._last_call = (B, T, prec)
        return bufs, on_dev

    def _to_host(self, t, on_dev):
        """see WaveNetAutoEncoder._finish"""
        out = self._eng.to_host(t, on_dev)
        B, T, prec = self._last_call
        if not on_dev and prec != _lib.FP32:
            self._eng.check_async(_lib.OP_STUDENT_FORWARD, B, T, prec)
        return out

    def generate(self, sess, inputs, encoding, conditions=None, precision=None):
        """model.py:570-576 -> out [B,T,1] = clip(z*s_tot + mu_tot, -1, 1).  ``inputs=None`` draws the logistic noise of
        student.py:172 on the device (inside the fused flow kernel on the fp16 path): nothing but the encoding is uploaded."""
        bufs, on_dev = self._forward(inputs, encoding, conditions, precision)
        return self._to_host(bufs["out"][:, :, None], on_dev)

    def forward_all(self, inputs, encoding, conditions=None, precision=None):
        """out, s_tot, mu_tot (model.py:517-535) and the chained flow output, each [B,T]; with ``inputs=None`` also the
        noise ``z`` that was drawn."""
        want = ("out", "s_tot", "mu_tot", "x_last") + (("z",) if inputs is None else ())
        bufs, on_dev = self._forward(inputs, encoding, conditions, precision, want=want)
        return {k: self._to_host(v, on_dev) for k, v in bufs.items()}

    def _entropies(self, inputs, encoding, conditions):
        bufs, _ = self._forward(inputs, encoding, conditions, want=("s_tot",))
        B, T = bufs["s_tot"].shape
        per = torch.empty(B, dtype=torch.float64, device="cuda")
        _lib.check(self._eng.lib.srwn_entropy(bufs["s_tot"].data_ptr(), per.data_ptr(), B, T, _stream()))
        return per.cpu().numpy()

    def getEntropy_fast(self, sess, inputs, encoding, conditions=None):
        """model.py:595-600: reduce_sum(log(s_tot) + 2) over the whole batch (model.py:356)."""
        return float(self._entropies(inputs, encoding, conditions).sum())

    def getEntropy(self, sess, inputs, encoding, conditions=None):
        """model.py:578-593: per-example entropies."""
        return self._entropies(inputs, encoding, conditions)

    # ---- distillation (model.py:316-401, 603-642) -------------------------------------------------
    def _teacher_logits(self, truth, enc, teacher_logits, teacher_precision):
        """stop_gradient(teacher decoder teacher-forced on the REAL audio) (model.py:326-334, F5)."""
        if teacher_logits is not None:
            return self._eng.to_device(teacher_logits, "tlogits")[0]
        if not isinstance(self.teacher, WaveNetAutoEncoder):
            raise RuntimeError("train needs `teacher` to be a WaveNetAutoEncoder (or pass teacher_logits=)")
        prec = teacher_precision or ("fp16" if "fp16" in self.teacher.available_precisions() else "fp32")
        return self.teacher.get_logits(truth, enc, precision=prec)

    def loss_and_grads(self, inputs, truth, encoding, conditions=None, teacher_logits=None,
                       teacher_precision=None, batch_norm=None):
        """One forward + backward of the distillation graph (model.py:356-384) on the device.
        Returns (loss, power_loss, entropy, flat_grads) with flat_grads a CUDA fp32 tensor laid out like the
        library's weight arena (see ``grad_of``).  ``batch_norm`` overrides the divisor B of model.py:379
        (the global batch under data parallelism)."""
        eng = self._eng
        enc = _with_conditions(encoding, conditions, self.condition_size)
        z, _ = eng.to_device(inputs, "z")
        x, _ = eng.to_device(truth, "truth")
        e, _ = eng.to_device(enc, "enc")
        B, T = z.shape
        if e.ndim != 3 or e.shape[0] != B or e.shape[1] * self.pool_stride != T or x.shape != z.shape:
            raise ValueError("inputs/truth must be [B,T], encoding [B, T/pool_stride, C]")
        tl = self._teacher_logits(x, e, teacher_logits, teacher_precision).contiguous()
        M = tl.shape[2] // 4
        ws, wsn = eng.workspace(_lib.OP_STUDENT_TRAIN, B, T, _lib.FP32)
        out = torch.empty(B, T, dtype=torch.float32, device="cuda")
        s_tot, mu_tot = torch.empty_like(out), torch.empty_like(out)
        _lib.check(eng.lib.srwn_student_forward_train(eng.h, z.data_ptr(), e.data_ptr(), out.data_ptr(),
                                                      s_tot.data_ptr(), mu_tot.data_ptr(), B, T, ws, wsn, _stream()))
        norm = float(batch_norm or B)
        # cross-entropy against the teacher's mixture (model.py:374-375): value + d/d out
        d_ce, nll = torch.empty_like(out), torch.empty_like(out)
        _lib.check(eng.lib.srwn_mol_loss_grad(out.data_ptr(), tl.data_ptr(), d_ce.data_ptr(), nll.data_ptr(), B, T, M, _stream()))
        # spectral power loss (model.py:360-371) and its gradient: FFT kernels of csrc/stft_loss.cu
        power, d_pow = self._power_loss(x, out)
        # entropy (model.py:356), clip mask (model.py:535) and the loss gradients the backward pass starts from
        d_pre, d_s = torch.empty_like(out), torch.empty_like(out)
        sums = torch.empty(_lib.DISTILL_SUMS_LEN, dtype=torch.float64, device="cuda")
        _lib.check(eng.lib.srwn_distill_loss_grad(z.data_ptr(), s_tot.data_ptr(), mu_tot.data_ptr(), nll.data_ptr(),
                                                  d_ce.data_ptr(), d_pow.data_ptr(), float(self.alpha), float(self.beta),
                                                  1.0 / norm, d_pre.data_ptr(), d_s.data_ptr(), sums.data_ptr(), B, T,
                                                  _stream()))
        entropy = sums[1]
        # one bucket [flat gradient | loss | power_loss]: what a data-parallel step all-reduces in a single call
        n = ctypes.c_int64()
        _lib.check(eng.lib.srwn_param_count(eng.h, ctypes.byref(n)))
        bucket = torch.empty(n.value + 2, dtype=torch.float32, device="cuda")
        grads, tail = bucket[:n.value], bucket[n.value:]
        _lib.check(eng.lib.srwn_distill_finish(sums.data_ptr(), power.data_ptr(), float(self.alpha), float(self.beta),
                                               1.0 / norm, tail.data_ptr(), _stream()))        # model.py:374-379
        _lib.check(eng.lib.srwn_student_backward(eng.h, z.data_ptr(), e.data_ptr(), d_pre.data_ptr(), d_s.data_ptr(),
                                                 grads.data_ptr(), B, T, ws, wsn, _stream()))
        self._bucket = bucket
        return tail[0], tail[1], entropy, grads

    def _power_loss(self, truth, out, frame_length=512, frame_step=256):
        """gamma * || mean_t |STFT(truth)|^2 - mean_t |STFT(out)|^2 ||^2 (model.py:360-371) and d/d out, on the device.
        Returns (power_loss: float64 CUDA tensor [1], d_out [B,T] fp32)."""
        eng = self._eng
        B, T = out.shape
        n = ctypes.c_size_t()
        _lib.check(eng.lib.srwn_stft_workspace_bytes(B, T, frame_length, frame_step, ctypes.byref(n)))
        if getattr(self, "_stft_ws", None) is None or self._stft_ws.numel() < n.value:
            self._stft_ws = torch.empty(max(n.value, 256), dtype=torch.uint8, device="cuda")
        p = torch.empty(1, dtype=torch.float64, device="cuda")
        g = torch.empty_like(out)
        _lib.check(eng.lib.srwn_stft_power_loss(truth.data_ptr(), out.data_ptr(), float(self.gamma), p.data_ptr(),
                                                g.data_ptr(), B, T, frame_length, frame_step, self._stft_ws.data_ptr(),
                                                self._stft_ws.numel(), _stream()))
        return p, g

    def grad_of(self, flat_grads, name):
        """View of one variable's gradient inside ``flat_grads`` (TF variable name)."""
        off, cnt = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self._eng.lib.srwn_weight_offset(self._eng.h, name.encode(), ctypes.byref(off), ctypes.byref(cnt)))
        return flat_grads[off.value:off.value + cnt.value]

    def apply_gradients(self, flat_grads, clip_norm=1.0, beta1=0.9, beta2=0.999, epsilon=1e-8):
        """tf.clip_by_global_norm(grads, 1.0) + AdamOptimizer(learning_rate).apply_gradients (model.py:382-401).
        Under torch.distributed the caller all-reduces ``flat_grads`` first (``train_fast`` does)."""
        eng = self._eng
        if not hasattr(self, "_adam"):
            self._adam = dict(m=torch.zeros_like(flat_grads), v=torch.zeros_like(flat_grads),
                              scratch=torch.zeros(4, dtype=torch.float32, device="cuda"), step=0)
        a = self._adam
        a["step"] += 1
        _lib.check(eng.lib.srwn_adam_step(eng.h, flat_grads.data_ptr(), a["m"].data_ptr(), a["v"].data_ptr(),
                                          a["scratch"].data_ptr(), float(clip_norm), float(self.learning_rate),
                                          float(beta1), float(beta2), float(epsilon), a["step"], _stream()))
        self._packed_stale = True

    def sync_weights(self):
        """Re-packs the 16-bit operand images from the (trained) device weights and refreshes ``get_weights``."""
        eng = self._eng
        _lib.check(eng.lib.srwn_commit_weights(eng.h, _stream()))
        for k, v in list(self._weights.items()):
            try:
                self._weights[k] = eng.get_weight(k, v.shape)
            except _lib.SrwnError:
                pass      # dead variables (gate conv, student skip conv) are not stored
        self._packed_stale = False

    def _sync_replicas(self, local_B):
        """Data parallelism: every rank must hold the same weights and optimizer state, or the shared gradient moves
        different parameters and the replicas drift apart.  On the first distributed step rank 0's weight arena and Adam
        state are broadcast; the global batch is the sum of the ranks' local batches (recomputed when the local one
        changes).  Returns (world, global batch)."""
        from . import shard
        if not shard.is_distributed():
            return 1, local_B
        import torch.distributed as dist
        st = self.__dict__.setdefault("_dp", {})
        if st.get("local_B") != local_B:
            st["local_B"], st["global_B"] = local_B, shard.global_batch(local_B)
        if not st.get("synced"):
            eng = self._eng
            n = ctypes.c_int64()
            _lib.check(eng.lib.srwn_param_count(eng.h, ctypes.byref(n)))
            w = torch.empty(n.value, dtype=torch.float32, device="cuda")
            _lib.check(eng.lib.srwn_weights_flat(eng.h, w.data_ptr(), n.value, 0, _stream()))      # device arena -> w
            state = [w]
            if hasattr(self, "_adam"):
                step = torch.tensor([float(self._adam["step"])], device="cuda")
                state += [self._adam["m"], self._adam["v"], step]
            shard.broadcast_from_rank0(state)
            _lib.check(eng.lib.srwn_weights_flat(eng.h, w.data_ptr(), n.value, 1, _stream()))      # w -> device arena
            if hasattr(self, "_adam"):
                self._adam["step"] = int(step.item())
            self._packed_stale = True
            st["synced"] = True
        return dist.get_world_size(), st["global_B"]

    def train_fast(self, sess, inputs, truth, encoding, conditions=None, teacher_logits=None, teacher_precision=None):
        """model.py:634-642: one optimisation step, returns (loss, power_loss).  With torch.distributed
        initialised, ranks hold batch shards: the loss is normalised by the global batch, the bucket
        [flat gradient | loss | power_loss] is summed with ONE all-reduce before the global-norm clip,
        and every rank applies the same Adam update to the same weights (``_sync_replicas``).  CUDA tensors in ->
        0-d CUDA tensors out with no host synchronisation inside the step; NumPy in -> Python floats."""
        B = int(inputs.shape[0])
        world, global_B = self._sync_replicas(B)
        on_dev = isinstance(inputs, torch.Tensor) and inputs.is_cuda
        loss, power, _, grads = self.loss_and_grads(inputs, truth, encoding, conditions, teacher_logits,
                                                    teacher_precision, batch_norm=global_B)
        if world > 1:
            from . import shard
            ev = self.__dict__.get("_coll_events")       # bench.py: CUDA events around the collective
            if ev:
                ev[0].record()
            shard.all_reduce_sum(self._bucket)
            if ev:
                ev[1].record()
        self.apply_gradients(grads)
        if on_dev:
            return loss, power
        lp = self._eng.to_host(self._bucket[-2:], False)
        return float(lp[0]), float(lp[1])

    def train(self, sess, inputs, truth, encoding, conditions=None, teacher_logits=None, teacher_precision=None):
        """model.py:603-632, literally: one evaluation of the graph per EXAMPLE of ``inputs`` -- its noise row fed as a
        batch of one, broadcast against the whole ``encoding`` / ``truth`` batch, the loss divided by shape(inputs)[0] = 1
        (model.py:379), gradients clipped per example (``self.grads`` are the clipped ones, model.py:385-392) -- then the
        clipped gradients, losses and power losses are averaged and one Adam update is applied without further clipping.
        Returns (mean loss, mean power loss)."""
        eng = self._eng
        z_all, _ = eng.to_device(inputs, "z_all")
        x, _ = eng.to_device(truth, "truth")
        e, _ = eng.to_device(_with_conditions(encoding, conditions, self.condition_size), "enc")
        N, Be = z_all.shape[0], e.shape[0]
        mean = None
        scratch = torch.zeros(1, dtype=torch.float32, device="cuda")
        for i in range(N):
            zi = z_all[i:i + 1].expand(Be, -1).contiguous()          # [1,T] against [Be, ...]: TF broadcasts the batch axis
            self.loss_and_grads(zi, x, e, None, teacher_logits, teacher_precision, batch_norm=1)
            b = self._bucket
            _lib.check(eng.lib.srwn_clip_by_global_norm(b.data_ptr(), b.numel() - 2, 1.0, scratch.data_ptr(), _stream()))
            if mean is None:
                mean = torch.zeros_like(b)
            _lib.check(eng.lib.srwn_axpy(mean.data_ptr(), b.data_ptr(), 1.0 / N, b.numel(), _stream()))
        self.apply_gradients(mean[:-2], clip_norm=0.0)
        lp = eng.to_host(mean[-2:], False)
        return float(lp[0]), float(lp[1])

    def _need_teacher(self):
        if not isinstance(self.teacher, WaveNetAutoEncoder):
            raise RuntimeError("this call evaluates the imported teacher graph (model.py:326-334): construct "
                               "ParallelWaveNet with teacher=<WaveNetAutoEncoder instance>")
        return self.teacher

    def encode(self, sess, inputs, conditions=None, precision=None):
        """model.py:644-649: the imported teacher's ``Encoding_output`` on ``inputs``."""
        return self._need_teacher().encode(inputs, conditions, precision=precision)

    def reconstruct(self, sess, inputs, conditions=None, **kw):
        """model.py:651-656: the imported teacher's ``Out_e`` (encoder + teacher-forced decoder + sampling)."""
        return self._need_teacher().reconstruct(inputs, conditions, **kw)
