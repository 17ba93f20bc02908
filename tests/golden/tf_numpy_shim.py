"""A numpy (float64) stand-in for the handful of TensorFlow 2.2 / Keras symbols that the reference's
hot path calls, so that the reference's OWN graph-wiring code -- ``Network/SR4DFlowNet.py``
(``build_network``, ``upsample3d``, ``conv3d``, ``resnet_block``), ``Network/loss_utils.py`` and the loss /
metric methods of ``Network/TrainerController.py`` -- can be imported UNMODIFIED from /root/reference and
executed eagerly here (TensorFlow itself is not installable in this image).

TEST INFRASTRUCTURE, used only by ``tests/golden/make_graph_golden.py`` / ``make_predictor_golden.py`` in the
build container.  Tensors are plain float64 numpy arrays; ``tf.keras.layers.Input`` returns a symbolic ``Sym`` node
and every op below records itself instead of executing when it meets one, so ``tf.keras.Model(inputs, outputs)``
(predictor.py:11-29) can replay the recorded graph on real batches later.

What this pins and what it does not: the layer order, channel concat order, skip connections, activation
placement, padding calls, the reshape/transpose choreography of ``upsample3d``, the loss / metric formulas
and the Keras variable creation order all come from the reference's code.  Only the primitive op semantics
below are restated, each from TensorFlow's published definition:

* ``tf.pad(x, paddings, 'SYMMETRIC')``            == ``numpy.pad(mode='symmetric')`` (edge included).
* ``tf.keras.layers.Conv3D`` (defaults)           channels-last, stride 1, padding 'valid', cross-correlation
                                                  with kernel (kd,kh,kw,Cin,Cout), ``+ bias``, then activation.
* ``tf.compat.v1.image.resize_bilinear(align_corners=True)``  NHWC; scale=(in-1)/(out-1) (0 when out==1);
                                                  src=dst*scale; lower=floor(src); upper=min(ceil(src),in-1);
                                                  lerp=src-lower, all three in C `float`;
                                                  ``top=tl+(tr-tl)*xl; bot=bl+(br-bl)*xl;
                                                  out=top+(bot-top)*yl`` (resize_bilinear_op.cc).
* ``tf.keras.layers.LeakyReLU(alpha)``            x if x > 0 else alpha*x.
* ``tf.keras.regularizers.l2(l)``                 l * sum(w**2).
* ``tf.round``                                    round half to even (== ``numpy.round``).
* ``tf.keras.metrics.Mean``                       running sum(values)/count(values) over all elements.
"""
import sys
import types

import numpy as np

F64 = np.float64
_state = {"weights": None, "layers": []}


def set_weight_source(fn):
    """fn(kernel_shape, use_bias) -> (kernel, bias|None); called once per Conv3D in creation order."""
    _state["weights"] = fn
    _state["layers"] = []


def created_layers():
    return list(_state["layers"])


# ---------------------------------------------------------------- symbolic tensors (Input / Model)
class Sym:
    """A recorded op (fn, args, kwargs); ``probe`` is its value on a batch of one zero patch (static shapes)."""
    __array_priority__ = 1000
    __array_ufunc__ = None

    def __init__(self, fn, args, kwargs, probe, name=None):
        self.fn, self.args, self.kwargs, self.probe, self.name = fn, args, kwargs, probe, name

    @property
    def shape(self):
        return (None,) + tuple(self.probe.shape[1:])

    def __pow__(self, e): return _lazy(np.power)(self, e)
    def __add__(self, o): return _lazy(np.add)(self, o)
    def __radd__(self, o): return _lazy(np.add)(o, self)
    def __sub__(self, o): return _lazy(np.subtract)(self, o)
    def __rsub__(self, o): return _lazy(np.subtract)(o, self)
    def __mul__(self, o): return _lazy(np.multiply)(self, o)
    def __rmul__(self, o): return _lazy(np.multiply)(o, self)
    def __truediv__(self, o): return _lazy(np.divide)(self, o)
    def __getitem__(self, idx): return _lazy(lambda a: a[idx])(self)


def _any_sym(obj):
    if isinstance(obj, Sym):
        return True
    if isinstance(obj, (list, tuple)):
        return any(_any_sym(o) for o in obj)
    return False


def _subst(obj, get):
    if isinstance(obj, Sym):
        return get(obj)
    if isinstance(obj, (list, tuple)):
        return type(obj)(_subst(o, get) for o in obj)
    return obj


def _lazy(fn):
    def wrapped(*args, **kwargs):
        if not (_any_sym(args) or _any_sym(list(kwargs.values()))):
            return fn(*args, **kwargs)
        probe = fn(*_subst(args, lambda n: n.probe), **{k: _subst(v, lambda n: n.probe) for k, v in kwargs.items()})
        return Sym(fn, args, kwargs, probe)
    return wrapped


def _evaluate(node, env, memo):
    if not isinstance(node, Sym):
        return node
    key = id(node)
    if key not in memo:
        if node.fn is None:
            memo[key] = env[key]
        else:
            get = lambda n: _evaluate(n, env, memo)  # noqa: E731
            memo[key] = node.fn(*_subst(node.args, get), **{k: _subst(v, get) for k, v in node.kwargs.items()})
    return memo[key]


def _Input(shape, name=None, **kw):
    return Sym(None, (), {}, np.zeros((1,) + tuple(shape), F64), name=name)


class _Model:
    """tf.keras.Model(inputs, outputs) over a recorded graph: predict / __call__ / load_weights / layers."""

    def __init__(self, inputs, outputs):
        self.inputs, self.outputs = list(inputs), outputs
        self.layers = created_layers()

    def __call__(self, xs, training=False):
        env = {id(i): np.asarray(x, F64) for i, x in zip(self.inputs, xs)}
        return _evaluate(self.outputs, env, {})

    def predict(self, xs, batch_size=None):
        return self(xs)

    @property
    def variable_names(self):
        return [f"{l.name}/{w}" for l in self.layers for w in (("kernel", "bias") if l.use_bias else ("kernel",))]

    def load_weights(self, path):
        import importlib
        h5io = importlib.import_module("4dflownet_b200.utils.h5io")     # the repo's HDF5 reader (no h5py here)
        w = h5io.load_keras_weights(path, self.variable_names)
        for l in self.layers:
            assert w[f"{l.name}/kernel"].shape == l.kernel.shape
            l.kernel = w[f"{l.name}/kernel"]
            if l.use_bias:
                l.bias = w[f"{l.name}/bias"]


# ---------------------------------------------------------------- primitive ops
def _pad(x, paddings, mode="CONSTANT"):
    assert mode == "SYMMETRIC", mode
    return np.pad(np.asarray(x, F64), [tuple(p) for p in paddings], mode="symmetric")


def _conv3d_valid(x, kernel):
    kd, kh, kw, ci, co = kernel.shape
    B, X, Y, Z, _ = x.shape
    ox, oy, oz = X - kd + 1, Y - kh + 1, Z - kw + 1
    if B * ox * oy * oz * ci * kd * kh * kw * 8 <= (1 << 30):
        win = np.lib.stride_tricks.sliding_window_view(x, (kd, kh, kw), axis=(1, 2, 3))  # (B,X',Y',Z',Ci,kd,kh,kw)
        return np.einsum("bxyzcijk,ijkco->bxyzo", win, kernel, optimize=True)
    out = np.zeros((B * ox * oy * oz, co), F64)          # big tensors: one GEMM per tap instead of an im2col copy
    for i in range(kd):
        for j in range(kh):
            for k in range(kw):
                out += np.ascontiguousarray(x[:, i:i + ox, j:j + oy, k:k + oz, :]).reshape(-1, ci) @ kernel[i, j, k]
    return out.reshape(B, ox, oy, oz, co)


class _Conv3D:
    def __init__(self, filters, kernel_size, strides=(1, 1, 1), padding="valid", activation=None, use_bias=True,
                 kernel_initializer=None, kernel_regularizer=None, bias_regularizer=None, **kw):
        assert padding == "valid" and not kw, (padding, kw)
        self.filters, self.k, self.activation, self.use_bias = filters, kernel_size, activation, use_bias
        self.kernel_regularizer, self.bias_regularizer = kernel_regularizer, bias_regularizer
        self.kernel = self.bias = None

    def __call__(self, x):
        shape = (self.k, self.k, self.k, x.shape[-1], self.filters)
        if _state["weights"] is None:       # no source: zeros until Model.load_weights fills them in
            self.kernel, self.bias = np.zeros(shape, np.float32), (np.zeros(self.filters, np.float32) if self.use_bias else None)
        else:
            self.kernel, self.bias = _state["weights"](shape, self.use_bias)
        self.name = "conv3d" if not _state["layers"] else f"conv3d_{len(_state['layers'])}"
        _state["layers"].append(self)
        return _lazy(self._apply)(x)

    def _apply(self, x):
        x = np.asarray(x, F64)
        y = _conv3d_valid(x, np.asarray(self.kernel, F64))
        if self.use_bias:
            y = y + np.asarray(self.bias, F64)
        if self.activation == "relu":
            y = np.maximum(y, 0.0)
        else:
            assert self.activation is None, self.activation
        return y


class _LeakyReLU:
    def __init__(self, alpha=0.3):
        self.alpha = alpha

    def __call__(self, x):
        return _lazy(lambda a: np.where(a > 0, a, self.alpha * a))(x)


def _resize_bilinear(images, size, align_corners=False, name=None):
    assert align_corners is True
    x = np.asarray(images, F64)
    _, h, w, _ = x.shape
    oh, ow = int(size[0]), int(size[1])

    def weights(n_in, n_out):
        # TF computes the scale, the source coordinate and the lerp weight in `float` (CalculateResizeScale,
        # LegacyScaler, compute_interpolation_weights in image_resizer_state.h / resize_bilinear_op.cc)
        f32 = np.float32
        scale = f32(n_in - 1) / f32(n_out - 1) if n_out > 1 else f32(0)
        src = np.arange(n_out, dtype=f32) * scale
        lo = np.floor(src).astype(np.int64)
        hi = np.minimum(np.ceil(src).astype(np.int64), n_in - 1)
        return lo, hi, (src - np.floor(src)).astype(F64)

    ylo, yhi, yl = weights(h, oh)
    xlo, xhi, xl = weights(w, ow)
    xl = xl[None, None, :, None]
    yl = yl[None, :, None, None]
    tl, tr = x[:, ylo][:, :, xlo], x[:, ylo][:, :, xhi]
    bl, br = x[:, yhi][:, :, xlo], x[:, yhi][:, :, xhi]
    top = tl + (tr - tl) * xl
    bot = bl + (br - bl) * xl
    return top + (bot - top) * yl


class _L2:
    def __init__(self, l=0.01):
        self.l2 = l

    def __call__(self, w):
        return self.l2 * np.sum(np.square(np.asarray(w, F64)))


class _Mean:
    def __init__(self, name=None):
        self.name = name
        self.reset_states()

    def update_state(self, values):
        v = np.asarray(values, F64)
        self.total += v.sum()
        self.count += v.size

    def result(self):
        return self.total / self.count if self.count else 0.0

    def reset_states(self):
        self.total, self.count = 0.0, 0


def install():
    """Register the stand-in as ``tensorflow`` in sys.modules (callers that import reference modules which
    ``import h5py`` register a stand-in for it themselves)."""
    tf = types.ModuleType("tensorflow")
    tf.float32 = np.float32          # tf.cast(x, tf.float32): kept in float64 below on purpose (golden = fp64 truth)
    tf.function = lambda f: f
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))      # replaying a ~40-layer recorded graph recurses
    tf.pad = _lazy(_pad)
    tf.reshape = _lazy(lambda x, shape, name=None: np.reshape(x, shape))
    tf.transpose = _lazy(lambda x, perm: np.transpose(x, perm))
    tf.concat = _lazy(lambda xs, axis: np.concatenate(xs, axis=axis))
    tf.constant = lambda v, dtype=None: np.asarray(v, F64)
    tf.less = np.less
    tf.equal = np.equal
    tf.not_equal = np.not_equal
    tf.cast = lambda x, dtype: np.asarray(x, F64)
    tf.reduce_sum = lambda x, axis=None: np.sum(x, axis=None if axis is None else tuple(np.atleast_1d(axis)))
    tf.square = np.square
    tf.sqrt = np.sqrt
    tf.clip_by_value = np.clip
    tf.where = np.where
    tf.round = np.round
    tf.zeros_like = np.zeros_like
    keras = types.ModuleType("tensorflow.keras")
    layers = types.ModuleType("tensorflow.keras.layers")
    layers.Conv3D = _Conv3D
    layers.LeakyReLU = _LeakyReLU
    layers.concatenate = _lazy(lambda xs, axis=-1: np.concatenate([np.asarray(a, F64) for a in xs], axis=axis))
    layers.Input = _Input
    keras.Model = _Model
    regularizers = types.ModuleType("tensorflow.keras.regularizers")
    regularizers.l2 = _L2
    metrics = types.ModuleType("tensorflow.keras.metrics")
    metrics.Mean = _Mean
    keras.layers, keras.regularizers, keras.metrics = layers, regularizers, metrics
    tf.keras = keras
    compat = types.ModuleType("tensorflow.compat")
    v1 = types.ModuleType("tensorflow.compat.v1")
    image = types.ModuleType("tensorflow.compat.v1.image")
    image.resize_bilinear = _lazy(_resize_bilinear)
    v1.image, compat.v1, tf.compat = image, v1, compat
    for m in (tf, keras, layers, regularizers, metrics, compat, v1, image):
        sys.modules[m.__name__] = m
    return tf
