"""Known-answer checks of the numpy TensorFlow stand-in (tests/golden/tf_numpy_shim.py) that produced the float-graph
golden vectors.  TensorFlow cannot be executed here, so each primitive is held to (a) the worked examples of
TensorFlow's own API documentation (tf.pad, tf.round, tf.clip_by_value, LeakyReLU, regularizers.l2) and (b)
independent third-party implementations of the same published definition (scipy.ndimage.correlate for the VALID
cross-correlation, torch's align_corners interpolation for resize_bilinear)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import tf_numpy_shim as shim  # noqa: E402


@pytest.fixture(scope="module")
def tf():
    saved = {k: v for k, v in sys.modules.items() if k == "tensorflow" or k.startswith("tensorflow.")}
    mod = shim.install()
    yield mod
    for k in [k for k in sys.modules if k == "tensorflow" or k.startswith("tensorflow.")]:
        del sys.modules[k]
    sys.modules.update(saved)


def test_pad_symmetric_documentation_example(tf):
    # tf.pad API docs: t = [[1, 2, 3], [4, 5, 6]], paddings = [[1, 1], [2, 2]], "SYMMETRIC"
    got = tf.pad(np.array([[1, 2, 3], [4, 5, 6]]), [[1, 1], [2, 2]], "SYMMETRIC")
    want = [[2, 1, 1, 2, 3, 3, 2], [2, 1, 1, 2, 3, 3, 2], [5, 4, 4, 5, 6, 6, 5], [5, 4, 4, 5, 6, 6, 5]]
    assert np.array_equal(got, want)


def test_round_and_clip_documentation_examples(tf):
    # tf.round API docs ("rounds half to even, also known as bankers rounding")
    assert np.array_equal(tf.round(np.array([0.9, 2.5, 2.3, 1.5, -4.5])), [1.0, 2.0, 2.0, 2.0, -4.0])
    # tf.clip_by_value API docs
    t = np.array([[-10., -1., 0.], [0., 2., 10.]])
    assert np.array_equal(tf.clip_by_value(t, -1, 1), [[-1., -1., 0.], [0., 1., 1.]])


def test_leaky_relu_and_l2_definitions(tf):
    x = np.array([-3.0, -1.0, 0.0, 2.0])
    # tf.keras.layers.LeakyReLU docs: f(x) = alpha * x if x < 0, f(x) = x if x >= 0; layer example with alpha=0.1
    assert np.allclose(tf.keras.layers.LeakyReLU(alpha=0.1)(x), [-0.3, -0.1, 0.0, 2.0])
    # tf.keras.regularizers.l2 docs: loss = l2 * reduce_sum(square(x))
    assert np.isclose(tf.keras.regularizers.l2(0.01)(np.array([[1.0, -2.0], [3.0, 0.5]])), 0.01 * (1 + 4 + 9 + 0.25))


def test_conv3d_is_valid_cross_correlation(tf):
    from scipy import ndimage
    g = np.random.default_rng(0)
    x = g.standard_normal((2, 6, 5, 7, 3))
    k = g.standard_normal((3, 3, 3, 3, 4))
    b = g.standard_normal(4)
    shim.set_weight_source(lambda shape, use_bias: (k, b))
    got = tf.keras.layers.Conv3D(4, 3, activation=None, use_bias=True)(x)
    assert got.shape == (2, 4, 3, 5, 4)
    want = np.zeros_like(got)
    for n in range(2):
        for co in range(4):
            acc = sum(ndimage.correlate(x[n, ..., ci], k[..., ci, co], mode="constant") for ci in range(3))
            want[n, ..., co] = acc[1:-1, 1:-1, 1:-1] + b[co]          # 'valid' = the fully covered interior
    assert np.abs(got - want).max() < 1e-12
    # 'relu' activation string
    shim.set_weight_source(lambda shape, use_bias: (k, b))
    assert np.array_equal(tf.keras.layers.Conv3D(4, 3, activation="relu")(x), np.maximum(got, 0))


def test_resize_bilinear_align_corners(tf):
    # corners map to corners and the interior is linear: [[1,2],[3,4]] -> 4x4
    img = np.array([[1.0, 2.0], [3.0, 4.0]]).reshape(1, 2, 2, 1)
    got = tf.compat.v1.image.resize_bilinear(img, [4, 4], align_corners=True)[0, :, :, 0]
    i, j = np.meshgrid(np.arange(4), np.arange(4), indexing="ij")
    assert np.allclose(got, 1 + j / 3 + 2 * i / 3, atol=1e-6)
    # against torch's align_corners=True bilinear interpolation on random data, several non-integer scale factors
    g = np.random.default_rng(1)
    for (h, w, oh, ow) in [(5, 7, 10, 14), (6, 6, 18, 18), (4, 9, 16, 36), (24, 24, 48, 48)]:
        x = g.standard_normal((2, h, w, 3))
        got = tf.compat.v1.image.resize_bilinear(x, [oh, ow], align_corners=True)
        want = torch.nn.functional.interpolate(torch.tensor(x).permute(0, 3, 1, 2), size=(oh, ow), mode="bilinear",
                                               align_corners=True).permute(0, 2, 3, 1).numpy()
        assert np.abs(got - want).max() < 2e-6 * np.abs(want).max()      # TF keeps the weights in C float


def test_metrics_mean_accumulates_over_all_values(tf):
    m = tf.keras.metrics.Mean(name="m")
    m.update_state(np.array([1.0, 3.0]))
    m.update_state(5.0)
    assert m.result() == 3.0
    m.reset_states()
    assert m.result() == 0.0
