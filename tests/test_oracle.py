"""Self-checks of the CPU oracle (parity is unpinned by the reference: see
oracle/sr4d_oracle.py header) and its pinning on the reference-generated integer
golden vectors."""
import hashlib
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "patchgen_golden.npz")


def test_param_table_matches_survey(oracle):
    tab = oracle.param_table(8, 4)
    assert len(tab) == 48
    assert sum(int(np.prod(s)) for _, s in tab) == 3342083
    assert tab[0] == ("conv3d/kernel", (3, 3, 3, 3, 64))
    assert tab[8] == ("conv3d_4/kernel", (1, 1, 1, 128, 64))
    assert tab[-1] == ("conv3d_35/bias", (1,))
    # resblock kernels have no bias
    names = [n for n, _ in tab]
    assert "conv3d_6/bias" not in names and "conv3d_29/bias" not in names


def test_conv_vs_naive_numpy(oracle):
    g = np.random.default_rng(0)
    x = g.standard_normal((2, 5, 4, 6, 3))
    k = g.standard_normal((3, 3, 3, 3, 7))
    b = g.standard_normal(7)
    ref = oracle.conv3d_naive(x, k, b)
    got = oracle.conv3d(torch.tensor(x), torch.tensor(k), torch.tensor(b)).numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)
    # symmetric p=1 == edge replicate, reflect differs
    a = g.standard_normal((4, 5))
    assert np.array_equal(np.pad(a, 1, mode="symmetric"), np.pad(a, 1, mode="edge"))
    assert not np.array_equal(np.pad(a, 1, mode="reflect"), np.pad(a, 1, mode="edge"))


def test_upsample_vs_literal_and_torch(oracle):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 6, 5, 4, 3, generator=g, dtype=torch.float64)
    for r in (1, 2, 3, 4):
        a = oracle.upsample3d(x, r)
        b = oracle.upsample3d_literal(x, r)
        assert a.shape == (2, 6 * r, 5 * r, 4 * r, 3)
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=0, atol=1e-12)
        if r > 1:
            t = torch.nn.functional.interpolate(x.permute(0, 4, 1, 2, 3), scale_factor=r, mode="trilinear",
                                                align_corners=True).permute(0, 2, 3, 4, 1)
            np.testing.assert_allclose(a.numpy(), t.numpy(), rtol=0, atol=5e-6)
    lo, hi, lerp = oracle._resize_weights(24, 48)
    assert lo[-1] == 23 and hi[-1] == 23 and lo[1] == 0 and abs(lerp[1] - 23 / 47) < 1e-6


def test_forward_shapes_and_dtype_agreement(oracle):
    P, r = 6, 2
    params = oracle.glorot_params(1, 1, seed=3, bias_scale=0.05)
    batch = oracle.synthetic_batch(2, P, r, seed=1)
    p64 = {k: torch.tensor(v, dtype=torch.float64) for k, v in params.items()}
    p32 = {k: torch.tensor(v) for k, v in params.items()}
    y64 = oracle.forward(p64, [torch.tensor(b, dtype=torch.float64) for b in batch[:6]], r, 1, 1)
    y32 = oracle.forward(p32, [torch.tensor(b) for b in batch[:6]], r, 1, 1)
    assert y64.shape == (2, P * r, P * r, P * r, 3)
    err = (y32.double() - y64).abs().max() / y64.abs().max()
    assert err < 1e-5


def test_gradients_vs_finite_differences(oracle):
    P, r = 4, 2
    params = oracle.glorot_params(1, 1, seed=5, bias_scale=0.05, dtype=np.float64)
    batch = oracle.synthetic_batch(2, P, r, seed=2)
    grads, metrics = oracle.gradients(params, batch, r, 1, 1)
    assert metrics["loss"].shape == (2,)

    def obj(pd):
        pt = {k: torch.tensor(v, dtype=torch.float64) for k, v in pd.items()}
        bt = [torch.tensor(np.asarray(b), dtype=torch.float64) for b in batch]
        return float(oracle.train_objective(pt, bt, r, 1, 1)[0])

    g = np.random.default_rng(0)
    for name in ["conv3d/kernel", "conv3d_4/kernel", "conv3d_6/kernel", "conv3d_9/kernel", "conv3d_11/kernel",
                 "conv3d_11/bias", "conv3d_5/bias"]:
        for _ in range(2):
            idx = tuple(int(g.integers(0, s)) for s in params[name].shape)
            h = 1e-5
            pp = {k: v.copy() for k, v in params.items()}
            pp[name][idx] += h
            pm = {k: v.copy() for k, v in params.items()}
            pm[name][idx] -= h
            fd = (obj(pp) - obj(pm)) / (2 * h)
            assert abs(fd - grads[name][idx]) <= 1e-6 + 1e-4 * abs(fd), (name, idx, fd, grads[name][idx])


def test_l2_gradient_is_B_times_1e6_w(oracle):
    """tape.gradient of the (B,) loss vector sums it, and l2 is added to every
    entry (TrainerController.py:249,223) => dJ/dw has B * 2 * 5e-7 * w."""
    P, r, B = 4, 1, 3
    params = oracle.glorot_params(0, 0, seed=6, dtype=np.float64)
    batch = list(oracle.synthetic_batch(B, P, r, seed=3))
    g_full, _ = oracle.gradients(params, batch, r, 0, 0)
    # remove regulariser analytically and compare against a zero-L2 run
    name = "conv3d_1/kernel"
    old = oracle.L2_COEFF
    try:
        oracle.L2_COEFF = 0.0
        g_nol2, _ = oracle.gradients(params, batch, r, 0, 0)
    finally:
        oracle.L2_COEFF = old
    np.testing.assert_allclose(g_full[name] - g_nol2[name], B * 1e-6 * params[name], rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(g_full["conv3d_1/bias"], g_nol2["conv3d_1/bias"], rtol=0, atol=0)


def test_adam_hand_example(oracle):
    p, m, v = np.array([1.0]), np.array([0.0]), np.array([0.0])
    p, m, v = oracle.adam_step(p, np.array([0.5]), m, v, t=1, lr=0.1)
    # step 1: m=0.05, v=0.00025, alpha=0.1*sqrt(0.001)/0.1
    assert abs(m[0] - 0.05) < 1e-15 and abs(v[0] - 0.00025) < 1e-15
    alpha = 0.1 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert abs(p[0] - (1 - alpha * 0.05 / (np.sqrt(0.00025) + 1e-7))) < 1e-15
    p2, m2, v2 = oracle.adam_step(p, np.array([-0.25]), m, v, t=2, lr=0.1)
    assert abs(m2[0] - (0.05 + (-0.25 - 0.05) * 0.1)) < 1e-15
    assert abs(v2[0] - (0.00025 + (0.0625 - 0.00025) * 0.001)) < 1e-12


def test_adam_tensorflow_documentation_example(oracle):
    """tf.keras.optimizers.Adam API docs: `opt = Adam(learning_rate=0.1); var1 = tf.Variable(10.0);
    loss = lambda: (var1 ** 2) / 2.0  # d(loss)/d(var1) == var1; opt.minimize(loss, [var1]);
    # The first step is -learning_rate * sign(grad)` and `var1.numpy()` prints 9.9."""
    p, m, v = oracle.adam_step(np.array([10.0]), np.array([10.0]), np.array([0.0]), np.array([0.0]), t=1, lr=0.1)
    assert abs(p[0] - 9.9) < 1e-6
    # a gradient of the other sign and a different magnitude moves by the same 0.1 in the other direction (up to the
    # epsilon = 1e-7 in the denominator: sqrt(v_hat) = 2e-3 here)
    p, m, v = oracle.adam_step(np.array([-3.0]), np.array([-0.002]), np.array([0.0]), np.array([0.0]), t=1, lr=0.1)
    assert abs(p[0] - (-2.9)) < 3e-4


def test_loss_and_metric_small_example(oracle):
    yt = torch.zeros(1, 2, 1, 1, 3)
    yp = torch.zeros(1, 2, 1, 1, 3)
    yt[0, 0, 0, 0] = torch.tensor([3.0, 4.0, 0.0])
    yp[0, 0, 0, 0] = torch.tensor([3.0, 4.0, 1.0])       # fluid voxel, se = 1
    yp[0, 1, 0, 0] = torch.tensor([0.0, 2.0, 0.0])       # non-fluid voxel, se = 4
    mask = torch.tensor([[[[1.0]], [[0.0]]]])
    tot, mse, div = oracle.loss_function(yt, yp, mask)
    assert div == 0 and abs(float(tot) - (1 / 2 + 4 / 2)) < 1e-7
    rel = oracle.calculate_relative_error(yt, yp, mask)
    # diff=1, actual=5 -> 0.2 (rounded to 1e-4) ; masked voxel only ; /(1+1)*100
    assert abs(float(rel) - 0.2 / 2 * 100) < 1e-4


@pytest.mark.skipif(not os.path.exists(GOLD), reason="golden file missing")
def test_tiling_restatement_vs_reference_golden(oracle):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_patchgen_golden import CASES, volume
    z = np.load(GOLD)
    for ci, (shape, P, r, full) in enumerate(CASES):
        meta = z[f"case{ci}_meta"]
        assert tuple(meta[:3]) == shape and meta[3] == P and meta[4] == r
        vol = volume(shape, ci)
        patches, nr = oracle.patchify(vol, P)
        plan = oracle.tiling_plan(shape, P, r)
        assert tuple(meta[5:8]) == tuple(nr) == plan["nr"]
        assert tuple(meta[8:11]) == plan["hr_padding"]
        assert hashlib.sha256(np.ascontiguousarray(patches).tobytes()).digest() == z[f"case{ci}_patch_sha"].tobytes()
        hr = patches.repeat(r, axis=1).repeat(r, axis=2).repeat(r, axis=3)
        st = oracle.patchup(hr, shape, P, r)
        assert tuple(meta[11:14]) == st.shape
        assert hashlib.sha256(np.ascontiguousarray(st).tobytes()).digest() == z[f"case{ci}_stitch_sha"].tobytes()
        if full:
            assert np.array_equal(patches, z[f"case{ci}_patches"])
            assert np.array_equal(st, z[f"case{ci}_stitched"])


def test_synthetic_batch_matches_product_generator(oracle):
    import importlib
    synth = importlib.import_module("4dflownet_b200.utils.synthetic")
    a, b = oracle.synthetic_batch(2, 6, 2, seed=3), synth.synthetic_batch(2, 6, 2, seed=3)
    assert len(a) == len(b) == 11 and all(np.array_equal(x, y) for x, y in zip(a, b))
