"""Parity against golden vectors produced by the reference's OWN graph code.

``tests/golden/graph_golden.npz`` was written by ``tests/golden/make_graph_golden.py``: the reference's
``SR4DFlowNet.build_network`` / ``upsample3d`` / ``conv3d`` / ``resnet_block``, ``loss_utils.calculate_relative_error``
and the loss / metric methods of ``TrainerController`` imported unmodified and executed in float64 on a numpy stand-in
for the TensorFlow primitives (``tests/golden/tf_numpy_shim.py``).  The CPU tests hold the oracle restatement to it
(1e-12: same arithmetic, different summation order), the GPU tests hold the CUDA engine to it through the C ABI
(north-star tolerance 1e-4, max|d|/max|ref|)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import synth  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "graph_golden.npz"))
CASES = [str(c) for c in GOLD["cases"]]


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)


def case(oracle, tag):
    """Rebuild the weights in the ORACLE's parameter order from the generator's sequential stream (a layer-order or
    bias-presence mismatch against the reference's creation order shows up as a shape / value mismatch)."""
    P, r, low, hi, B, seed = (int(x) for x in GOLD[f"{tag}_cfg"])
    draw = synth.graph_weight_source(seed)
    table = oracle.param_table(low, hi)
    params, i = {}, 0
    while i < len(table):
        name, shape = table[i]
        assert name.endswith("/kernel")
        has_bias = i + 1 < len(table) and table[i + 1][0] == name.replace("/kernel", "/bias")
        k, b = draw(shape, has_bias)
        params[name] = k
        if has_bias:
            params[table[i + 1][0]] = b
        i += 2 if has_bias else 1
    lr, hr, mask = synth.graph_inputs(B, P, r, seed)
    assert np.float64(sum(a.sum() for a in lr + hr) + mask.sum()) == GOLD[f"{tag}_input_sum"]
    return (P, r, low, hi, B), params, lr, hr, mask


@pytest.mark.parametrize("tag", CASES)
def test_oracle_layer_order_matches_reference_creation_order(oracle, tag):
    (P, r, low, hi, B), *_ = case(oracle, tag)
    ref_layers = [str(s).split("|") for s in GOLD[f"{tag}_layer_table"]]
    table = oracle.param_table(low, hi)
    kernels = [(n, s) for n, s in table if n.endswith("/kernel")]
    biases = {n for n, _ in table if n.endswith("/bias")}
    assert len(kernels) == len(ref_layers)
    for (n, s), (rname, rshape, rbias) in zip(kernels, ref_layers):
        assert n == rname + "/kernel"
        assert "x".join(map(str, s)) == rshape
        assert ((rname + "/bias") in biases) == bool(int(rbias))


@pytest.mark.parametrize("tag", CASES)
def test_oracle_forward_loss_metric_vs_reference_graph(oracle, tag):
    (P, r, low, hi, B), params, lr, hr, mask = case(oracle, tag)
    p64 = {k: torch.tensor(v, dtype=torch.float64) for k, v in params.items()}
    pred = oracle.forward(p64, [torch.tensor(a) for a in lr], r, low, hi)
    assert pred.shape == GOLD[f"{tag}_pred"].shape
    assert relerr(pred.numpy(), GOLD[f"{tag}_pred"]) < 1e-12
    hires = torch.tensor(np.concatenate(hr, axis=-1))
    m = torch.tensor(mask)
    gp = torch.tensor(GOLD[f"{tag}_pred"])          # feed the golden prediction: isolates the loss / metric formulas
    tot, mse, div = oracle.loss_function(hires, gp, m)
    np.testing.assert_allclose(tot.numpy(), GOLD[f"{tag}_loss"], rtol=1e-12)
    np.testing.assert_allclose(mse.numpy(), GOLD[f"{tag}_mse"], rtol=1e-12)
    assert div == 0
    rel = oracle.calculate_relative_error(hires, gp, m)
    np.testing.assert_allclose(rel.numpy(), GOLD[f"{tag}_rel"], rtol=1e-12)
    l2 = float(oracle.regularizer_loss(p64))
    assert abs(l2 - float(GOLD[f"{tag}_l2"])) < 1e-12 * l2
    np.testing.assert_allclose(tot.numpy() + l2, GOLD[f"{tag}_loss_train"], rtol=1e-12)
    # running means as TrainerController.calculate_and_update_metrics accumulates them (one train + one val call)
    names = [str(n) for n in GOLD["metric_names"]]
    got = dict(zip(names, GOLD[f"{tag}_metrics"]))
    assert abs(got["train_loss"] - float((tot + l2).mean())) < 1e-12
    assert abs(got["val_loss"] - float(tot.mean())) < 1e-12
    assert abs(got["train_accuracy"] - float(rel.mean())) < 1e-10
    assert abs(got["l2_reg_loss"] - l2) < 1e-15
    assert got["train_div"] == 0 and got["val_div"] == 0


@pytest.mark.parametrize("tag", [t for t in CASES if f"{t}_fd_grad" in GOLD.files])
def test_oracle_gradient_vs_finite_differences_of_the_reference_objective(oracle, tag):
    """d/dw of sum_b(loss_b + l2) -- the objective TrainerController.train_step gives to tape.gradient, assembled by the
    reference's own calculate_and_update_metrics(..., 'train') -- by central differences through the reference code
    (stored with the golden) against the oracle's autograd gradient, incl. its B * 2 * 5e-7 * w regulariser share."""
    (P, r, low, hi, B), params, lr, hr, mask = case(oracle, tag)
    batch = [*lr, *hr, np.float64(1.0), mask]
    grads, _ = oracle.gradients({k: v.astype(np.float64) for k, v in params.items()}, batch, r, low, hi)
    for pick, want in zip(GOLD[f"{tag}_fd_picks"], GOLD[f"{tag}_fd_grad"]):
        name, fi = str(pick).split("|")
        got = grads[name].reshape(-1)[int(fi)]
        assert abs(got - want) <= 2e-6 * abs(want) + 1e-9, (name, fi, got, want)


def test_oracle_upsample_vs_reference_resize(oracle):
    """The reference's upsample3d (two resize_bilinear passes + transposes) on a random 5-D tensor, via the shim, vs
    the oracle's separable lerp -- an odd, anisotropy-revealing shape is impossible here (the reference assumes the
    same factor on all axes) so the tensor is non-cubic instead."""
    import tf_numpy_shim
    mods = {k: sys.modules.get(k) for k in list(sys.modules) if k == "tensorflow" or k.startswith("tensorflow.")}
    ref_src = "/root/reference/src"
    if not os.path.isdir(ref_src):
        pytest.skip("reference tree not present (GPU box)")
    tf_numpy_shim.install()
    sys.path.insert(0, ref_src)
    try:
        import importlib
        net = importlib.import_module("Network.SR4DFlowNet")
        g = np.random.default_rng(9)
        for shape, r in [((2, 3, 5, 4, 2), 2), ((1, 4, 2, 3, 3), 3), ((1, 2, 2, 2, 1), 4)]:
            x = g.standard_normal(shape)
            want = net.upsample3d(x, r)
            got = oracle.upsample3d(torch.tensor(x), r).numpy()
            assert got.shape == want.shape
            assert relerr(got, want) < 1e-13
    finally:
        sys.path.remove(ref_src)
        for k in [k for k in sys.modules if k == "tensorflow" or k.startswith("tensorflow.") or k.startswith("Network")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in mods.items() if v is not None})


def test_committed_graph_golden_is_reproducible(tmp_path):
    """Where the reference tree is present (the build container), re-run the generator and compare every array with the
    committed file: guards the fixture against drifting away from the reference's code."""
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference tree not present (GPU box)")
    import subprocess
    script = os.path.join(HERE, "golden", "make_graph_golden.py")
    code = ("import runpy, sys, numpy as np, os; sys.argv=['x']; m = runpy.run_path(%r); "
            "m['HERE'] = %r; m['main'].__globals__['HERE'] = %r; m['main']()") % (script, str(tmp_path), str(tmp_path))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, GRAPH_GOLDEN_NO_FD="1"))      # the finite-difference entries take a minute
    assert out.returncode == 0, out.stderr[-2000:]
    fresh = np.load(os.path.join(str(tmp_path), "graph_golden.npz"))
    keys = [k for k in GOLD.files if "_fd_" not in k]
    assert sorted(fresh.files) == sorted(keys)
    for k in keys:
        a, b = GOLD[k], fresh[k]
        if a.dtype.kind in "fc":
            np.testing.assert_allclose(b, a, rtol=1e-13, atol=0, err_msg=k)
        else:
            assert np.array_equal(a, b), k


def test_oracle_predictor_flow_vs_reference_script_golden(oracle, tmp_path):
    """BASELINE configs[0] end to end: the oracle's patchify -> forward (fp32) -> patchup -> x venc -> zero small values
    against the result file the reference's own src/predictor.py wrote (tests/golden/make_predictor_golden.py)."""
    import importlib
    gold = np.load(os.path.join(HERE, "golden", "predictor_golden.npz"))
    h5io = importlib.import_module("4dflownet_b200.utils.h5io")
    path = os.path.join(str(tmp_path), "example_data.h5")
    synth.make_example_lr(path, int(gold["data_seed"]))
    params = {k: torch.tensor(v) for k, v in synth.keras_weight_dict(8, 4, int(gold["weight_seed"])).items()}
    with h5io.open_file(path, "r") as lr:
        venc = np.float32(max(float(lr[k][0]) for k in ("venc_u", "venc_v", "venc_w")))
        vel = [(np.asarray(lr[c][0]) / venc).astype(np.float32) for c in "uvw"]        # ImageDataset.py:11-35
        mag = [(np.asarray(lr["mag_" + c][0]) / 4095.).astype(np.float32) for c in "uvw"]
    stacks = [oracle.patchify(a, 24)[0] for a in vel + mag]
    assert stacks[0].shape[0] == 12
    with torch.no_grad():
        y = torch.cat([oracle.forward(params, [torch.tensor(s[i:i + 4, ..., None]) for s in stacks], 2, 8, 4)
                       for i in range(0, 12, 4)]).numpy()
    for ci, c in enumerate("uvw"):
        v = oracle.patchup(y[..., ci], vel[0].shape, 24, 2) * venc
        v[np.abs(v) < venc / 2048] = 0                                                 # predictor.py:99-104
        sub, want = v[::2, ::2, ::2].astype(np.float64), gold[c].astype(np.float64)
        assert sub.shape == want.shape
        tol = 1e-4 * float(gold[c + "_moments"][4])
        bad = np.abs(sub - want) > tol
        flip = bad & ((sub == 0) | (want == 0)) & (np.maximum(np.abs(sub), np.abs(want)) < venc / 2048 + 2 * tol)
        assert not np.any(bad & ~flip), float(np.abs(sub - want)[bad & ~flip].max())
        assert flip.sum() <= 4


# ------------------------------------------------------------------ GPU: the CUDA engine against the same golden
@pytest.mark.gpu
@pytest.mark.parametrize("impl", ["simt", "auto"])
@pytest.mark.parametrize("tag", CASES)
def test_engine_forward_vs_reference_graph(pkg, oracle, tag, impl):
    (P, r, low, hi, B), params, lr, hr, mask = case(oracle, tag)
    eng = pkg.Engine(P, r, low, hi, max_batch=B, training=False, device=0)
    eng.set_option(pkg._lib.OPT_CONV_IMPL, pkg._lib.CONV_SIMT if impl == "simt" else pkg._lib.CONV_AUTO)
    eng.set_weights(params)
    y = eng.forward([a.astype(np.float32) for a in lr]).cpu().numpy()
    assert y.shape == GOLD[f"{tag}_pred"].shape
    assert relerr(y, GOLD[f"{tag}_pred"]) < 1e-4
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_engine_train_metrics_vs_reference_graph(pkg, oracle, tag):
    (P, r, low, hi, B), params, lr, hr, mask = case(oracle, tag)
    eng = pkg.Engine(P, r, low, hi, max_batch=B, training=True, device=0)
    eng.set_weights(params)
    per, l2, pred = eng.train_fwd_bwd([a.astype(np.float32) for a in lr], [a[..., 0].astype(np.float32) for a in hr],
                                      mask.astype(np.float32), want_pred=True)
    per = per.cpu().numpy()
    assert relerr(pred.cpu().numpy(), GOLD[f"{tag}_pred"]) < 1e-4
    np.testing.assert_allclose(per[:, 0], GOLD[f"{tag}_loss"], rtol=1e-4)
    np.testing.assert_allclose(per[:, 1], GOLD[f"{tag}_mse"], rtol=1e-4)
    # the metric rounds every voxel to 1e-4 steps: a prediction 1e-6 away can flip single voxels
    np.testing.assert_allclose(per[:, 2], GOLD[f"{tag}_rel"], rtol=1e-3)
    np.testing.assert_allclose(per[:, 3], mask.sum(axis=(1, 2, 3)), rtol=0)
    assert abs(float(l2) - float(GOLD[f"{tag}_l2"])) < 1e-5 * float(GOLD[f"{tag}_l2"])
    np.testing.assert_allclose(per[:, 0] + float(l2), GOLD[f"{tag}_loss_train"], rtol=1e-4)
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", [t for t in CASES if f"{t}_fd_grad" in GOLD.files])
def test_engine_gradient_vs_finite_differences_of_the_reference_objective(pkg, oracle, tag):
    """The engine's flat gradient (sum over the batch, the L2 share is added by the Adam kernel) plus B*2*5e-7*w against
    the finite differences of the reference code's training objective.  Tolerance: 5e-3 of the entry plus 5e-3 of the
    tensor's rms gradient (the bar tests/test_gpu_backward.py applies per tensor)."""
    (P, r, low, hi, B), params, lr, hr, mask = case(oracle, tag)
    eng = pkg.Engine(P, r, low, hi, max_batch=B, training=True, device=0)
    eng.set_weights(params)
    eng.train_fwd_bwd([a.astype(np.float32) for a in lr], [a[..., 0].astype(np.float32) for a in hr],
                      mask.astype(np.float32))
    grads = {name: view.cpu().numpy().astype(np.float64) for name, view in eng.tensor_views(eng.grads)}
    for pick, want in zip(GOLD[f"{tag}_fd_picks"], GOLD[f"{tag}_fd_grad"]):
        name, fi = str(pick).split("|")
        g = grads[name]
        l2_share = B * 2 * oracle.L2_COEFF * float(params[name].reshape(-1)[int(fi)]) if name.endswith("kernel") else 0.0
        got = g.reshape(-1)[int(fi)] + l2_share
        rms = float(np.sqrt(np.mean(g ** 2)))
        assert abs(got - want) <= 5e-3 * abs(want) + 5e-3 * rms, (name, fi, got, want, rms)
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "c"])
def test_trainer_controller_running_means_vs_reference_graph(pkg, oracle, tag):
    """The product TrainerController's calculate_and_update_metrics / loss_metrics (the reference's names) after one
    'train' and one 'val' call on the golden prediction problem, against the running means the reference's own
    calculate_and_update_metrics accumulated in tf.keras.metrics.Mean (stand-in) objects."""
    import importlib
    tcm = importlib.import_module("4dflownet_b200.Network.TrainerController")
    (P, r, low, hi, B), params, lr, hr, mask = case(oracle, tag)
    tc = tcm.TrainerController(P, r, 1e-4, False, "golden", low, hi, max_batch=B)
    tc.model.set_weights([params[n] for n in tc.model.variable_names])
    pred = tc.model([a.astype(np.float32) for a in lr], training=False)
    hires = np.concatenate(hr, axis=-1).astype(np.float32)
    m32 = mask.astype(np.float32)
    loss_train = tc.calculate_and_update_metrics(hires, pred, m32, "train")
    tc.calculate_and_update_metrics(hires, pred, m32, "val")
    np.testing.assert_allclose(np.asarray(loss_train, dtype=np.float64), GOLD[f"{tag}_loss_train"], rtol=1e-4)
    want = dict(zip([str(n) for n in GOLD["metric_names"]], GOLD[f"{tag}_metrics"]))
    for name, ref in want.items():
        got = float(tc.loss_metrics[name].result())
        tol = 1e-3 if "accuracy" in name else 1e-4        # the metric rounds every voxel to 1e-4 steps
        assert abs(got - ref) <= tol * abs(ref) + 1e-12, (name, got, ref)
