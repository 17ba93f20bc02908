"""A numpy (float64) stand-in for the handful of TensorFlow 2.2 / Keras symbols that the reference's
hot path calls, so that the reference's OWN graph-wiring code -- ``Network/SR4DFlowNet.py``
(``build_network``, ``upsample3d``, ``conv3d``, ``resnet_block``), ``Network/loss_utils.py`` and the loss /
metric methods of ``Network/TrainerController.py`` -- can be imported UNMODIFIED from /root/reference and
executed eagerly here (TensorFlow itself is not installable in this image).

TEST INFRASTRUCTURE, used only by ``tests/golden/make_graph_golden.py`` in the build container.

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


# ---------------------------------------------------------------- primitive ops
def _pad(x, paddings, mode="CONSTANT"):
    assert mode == "SYMMETRIC", mode
    return np.pad(np.asarray(x, F64), [tuple(p) for p in paddings], mode="symmetric")


def _conv3d_valid(x, kernel):
    kd, kh, kw, ci, co = kernel.shape
    win = np.lib.stride_tricks.sliding_window_view(x, (kd, kh, kw), axis=(1, 2, 3))  # (B,X',Y',Z',Ci,kd,kh,kw)
    return np.einsum("bxyzcijk,ijkco->bxyzo", win, kernel, optimize=True)


class _Conv3D:
    def __init__(self, filters, kernel_size, strides=(1, 1, 1), padding="valid", activation=None, use_bias=True,
                 kernel_initializer=None, kernel_regularizer=None, bias_regularizer=None, **kw):
        assert padding == "valid" and not kw, (padding, kw)
        self.filters, self.k, self.activation, self.use_bias = filters, kernel_size, activation, use_bias
        self.kernel_regularizer, self.bias_regularizer = kernel_regularizer, bias_regularizer
        self.kernel = self.bias = None

    def __call__(self, x):
        x = np.asarray(x, F64)
        shape = (self.k, self.k, self.k, x.shape[-1], self.filters)
        self.kernel, self.bias = _state["weights"](shape, self.use_bias)
        self.name = "conv3d" if not _state["layers"] else f"conv3d_{len(_state['layers'])}"
        _state["layers"].append(self)
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
        return np.where(x > 0, x, self.alpha * x)


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
    """Register the stand-in as ``tensorflow`` (and an empty ``h5py``) in sys.modules."""
    tf = types.ModuleType("tensorflow")
    tf.float32 = np.float32          # tf.cast(x, tf.float32): kept in float64 below on purpose (golden = fp64 truth)
    tf.function = lambda f: f
    tf.pad = _pad
    tf.reshape = lambda x, shape, name=None: np.reshape(x, shape)
    tf.transpose = lambda x, perm: np.transpose(x, perm)
    tf.concat = lambda xs, axis: np.concatenate(xs, axis=axis)
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
    layers.concatenate = lambda xs, axis=-1: np.concatenate([np.asarray(a, F64) for a in xs], axis=axis)
    regularizers = types.ModuleType("tensorflow.keras.regularizers")
    regularizers.l2 = _L2
    metrics = types.ModuleType("tensorflow.keras.metrics")
    metrics.Mean = _Mean
    keras.layers, keras.regularizers, keras.metrics = layers, regularizers, metrics
    tf.keras = keras
    compat = types.ModuleType("tensorflow.compat")
    v1 = types.ModuleType("tensorflow.compat.v1")
    image = types.ModuleType("tensorflow.compat.v1.image")
    image.resize_bilinear = _resize_bilinear
    v1.image, compat.v1, tf.compat = image, v1, compat
    for m in (tf, keras, layers, regularizers, metrics, compat, v1, image):
        sys.modules[m.__name__] = m
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    return tf
