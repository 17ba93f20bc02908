"""Edge cases and size-independent properties at BASELINE.json's full sizes, through the C ABI on the GPU:
ragged / partial batches, error behaviour, determinism, batch-permutation equivariance, tiling round trips at the
config-5 volume size, and full-geometry parity of the r=4 forward and of one config-2 train step."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def test_partial_and_ragged_batches(pkg, oracle):
    P, r, low, hi = 8, 2, 1, 1
    params = oracle.glorot_params(low, hi, seed=3, bias_scale=0.05)
    batch = oracle.synthetic_batch(5, P, r, seed=2)
    big = pkg.Engine(P, r, low, hi, max_batch=8, training=True, device=0)
    big.set_weights(params)
    y5 = big.forward(batch[:6]).cpu().numpy()
    for B in (1, 3):                                     # B < max_batch reuses the same buffers with a smaller batch
        yB = big.forward([b[:B] for b in batch[:6]]).cpu().numpy()
        assert np.array_equal(yB, y5[:B])                # per-sample results do not depend on the batch they ride in
    small = pkg.Engine(P, r, low, hi, max_batch=3, training=True, device=0)
    small.set_weights(params)
    sub = [b[:3] for b in batch]
    per_a, l2_a, _ = big.train_fwd_bwd(sub[:6], [b[..., 0] for b in sub[6:9]], sub[10])
    per_b, l2_b, _ = small.train_fwd_bwd(sub[:6], [b[..., 0] for b in sub[6:9]], sub[10])
    assert torch.equal(per_a, per_b) and torch.equal(l2_a, l2_b)
    assert torch.equal(big.grads, small.grads)           # bit-identical gradients: deterministic reductions
    # Keras-style predict with a ragged tail (5 samples, batch 2)
    model = pkg.prepare_network(P, r, low, hi, max_batch=2)
    model.set_weights([params[n] for n in model.variable_names])
    yp = model.predict(list(batch[:6]), batch_size=2)
    assert yp.shape == (5, 16, 16, 16, 3) and np.array_equal(yp, y5)
    big.close(); small.close()


def test_error_behaviour(pkg, oracle):
    L = pkg._lib
    with pytest.raises(pkg.Sr4dError):
        pkg.Engine(2, 2, 1, 1, max_batch=1)              # patch_size below the supported minimum -> SR4D_EINVAL
    with pytest.raises(pkg.Sr4dError):
        pkg.Engine(8, 0, 1, 1, max_batch=1)
    eng = pkg.Engine(8, 2, 1, 1, max_batch=2, training=False, device=0)
    batch = oracle.synthetic_batch(3, 8, 2, seed=0)
    with pytest.raises(pkg.Sr4dError, match="SR4D_EINVAL"):
        eng.forward(batch[:6])                           # B = 3 > max_batch = 2
    assert not hasattr(eng, "grads")                     # inference handle: no gradient / optimizer state
    with pytest.raises(pkg.Sr4dError, match="SR4D_ESTATE"):
        eng._check(eng.lib.sr4d_adam_step(eng._h, 1e-3, 0.9, 0.999, 1e-7, 1, 0.0, None), "sr4d_adam_step")
    with pytest.raises(ValueError):
        eng.set_weights([np.zeros(1, np.float32)])       # wrong number of tensors
    with pytest.raises(pkg.Sr4dError):
        eng.set_option(99, 1)
    eng.close()
    eng.close()                                          # idempotent


def test_determinism_and_batch_permutation(pkg, oracle):
    P, r = 24, 2
    eng = pkg.Engine(P, r, 8, 4, max_batch=4, training=True, device=0)
    eng.set_weights(oracle.glorot_params(8, 4, seed=1234, bias_scale=0.02))
    batch = oracle.synthetic_batch(4, P, r, seed=5)
    hr = [b[..., 0] for b in batch[6:9]]
    y1 = eng.forward(batch[:6]).clone()
    y2 = eng.forward(batch[:6])
    assert torch.equal(y1, y2)
    perm = [2, 0, 3, 1]
    yp = eng.forward([b[perm] for b in batch[:6]])
    assert torch.equal(yp, y1[perm])
    per1, _, _ = eng.train_fwd_bwd(batch[:6], hr, batch[10])
    g1 = eng.grads.clone()
    per2, _, _ = eng.train_fwd_bwd(batch[:6], hr, batch[10])
    assert torch.equal(per1, per2) and torch.equal(g1, eng.grads)      # run-to-run bit-identical gradients
    eng.close()


def test_tiling_round_trip_at_config5_size(pkg):
    """patchify -> (nearest-neighbour x2 'prediction') -> GPU stitch reproduces the repeated volume bit-exactly for
    the 160x160x64 volume of config 5 (256 patches, 320x320x128 output) and a ragged volume."""
    eng = pkg.Engine(24, 2, 0, 0, max_batch=1, training=False, device=0)
    for shape in ((160, 160, 64), (42, 38, 36), (25, 47, 23)):
        g = np.random.default_rng(1)
        vol = g.integers(-1000, 1000, size=shape).astype(np.float32)
        pg = pkg.PatchGenerator(24, 2)
        patches, nx, ny, nz = pg._generate_overlapping_patches(vol)
        pg.nr_x, pg.nr_y, pg.nr_z = nx, ny, nz
        assert len(patches) == pg.count_patches(shape)
        hr = patches.repeat(2, 1).repeat(2, 2).repeat(2, 3)
        pred = torch.from_numpy(np.stack([hr, -hr, 2 * hr], -1)).cuda()
        out = eng.stitch(pred, (nx, ny, nz), pg.stitched_shape(), 4, 1.0, round_small=False).cpu().numpy()
        want = vol.repeat(2, 0).repeat(2, 1).repeat(2, 2)
        assert out.shape == (3,) + want.shape
        assert np.array_equal(out[0], want) and np.array_equal(out[1], -want) and np.array_equal(out[2], 2 * want)
        assert np.array_equal(pg._patchup_with_overlap(hr, nx, ny, nz), want)
    eng.close()


def test_forward_r4_full_geometry_vs_oracle(pkg, oracle):
    """BASELINE config 3 geometry (P=24, r=4 -> 96^3), one patch, against the fp32 oracle."""
    params = oracle.glorot_params(8, 4, seed=77, bias_scale=0.02)
    batch = oracle.synthetic_batch(1, 24, 4, seed=3)
    eng = pkg.Engine(24, 4, 8, 4, max_batch=1, training=False, device=0)
    eng.set_weights(params)
    y = eng.forward(batch[:6]).cpu().numpy()
    p32 = {k: torch.tensor(v) for k, v in params.items()}
    with torch.no_grad():
        ref = oracle.forward(p32, [torch.tensor(b) for b in batch[:6]], 4, 8, 4).numpy()
    assert y.shape == (1, 96, 96, 96, 3)
    assert relerr(y, ref) < 1e-4
    eng.close()


def _rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def _flat_vs_fp64(eng, oracle, params, batch, r, low, hi):
    """(flat rel-L2, worst per-tensor rel-L2, metrics) of the engine's gradient buffer against float64 autograd."""
    B = len(batch[0])
    g64, met = oracle.gradients({k: v.astype(np.float64) for k, v in params.items()}, batch, r, low, hi)
    l2c = oracle.L2_COEFF
    want = {n: g64[n] - (B * 2 * l2c * params[n].astype(np.float64) if n.endswith("kernel") else 0.0) for n, *_ in eng.table}
    got = {n: v.cpu().numpy().astype(np.float64) for n, v in eng.tensor_views(eng.grads)}
    flat = _rel_l2(np.concatenate([got[n].ravel() for n in want]), np.concatenate([want[n].ravel() for n in want]))
    worst = max(_rel_l2(got[n], want[n]) for n in want)
    return flat, worst, met


def test_train_step_full_geometry_vs_oracle(pkg, oracle, bars):
    """BASELINE configs[1] geometry (P=24, r=2, 8/4 blocks), two samples: loss, metric and the gradient of the
    default (tensor-core) path against FLOAT64 autograd of the oracle; flat gradient within the north-star 1e-4.
    (tools/grad_parity.py prints the same comparison at B=1 and B=8 with the SIMT anchor, fp32 autograd and the
    identical-gates variants next to it: profiles/r02_grad_parity.txt.)"""
    params = oracle.glorot_params(8, 4, seed=1234, bias_scale=0.02)
    batch = oracle.synthetic_batch(2, 24, 2, seed=9)
    eng = pkg.Engine(24, 2, 8, 4, max_batch=2, training=True, device=0)
    eng.set_weights(params)
    per, l2, pred = eng.train_fwd_bwd(batch[:6], [b[..., 0] for b in batch[6:9]], batch[10], want_pred=True)
    flat, worst, met = _flat_vs_fp64(eng, oracle, params, batch, 2, 8, 4)
    assert relerr(pred.cpu().numpy(), met["pred"]) < 1e-4
    np.testing.assert_allclose(per[:, 0].cpu().numpy() + float(l2), met["loss"], rtol=1e-4)
    np.testing.assert_allclose(per[:, 2].cpu().numpy(), met["rel_err"], rtol=1e-3, atol=1e-3)
    bars("full_geometry/P24r2l8h4B2/flat", flat, 1e-4)
    bars("full_geometry/P24r2l8h4B2/worst_tensor", worst, 2e-3)
    eng.close()


def test_trainer_default_geometry_vs_oracle(pkg, oracle, bars):
    """trainer.py's own defaults (patch_size=16, res_increase=2, 8/4 blocks; trainer.py:28-39), two samples: exercises the
    TY=16 forward tiles and the 18^3 / 34^3 fused-dgrad grids against FLOAT64 autograd of the oracle."""
    params = oracle.glorot_params(8, 4, seed=21, bias_scale=0.02)
    batch = oracle.synthetic_batch(2, 16, 2, seed=4)
    eng = pkg.Engine(16, 2, 8, 4, max_batch=2, training=True, device=0)
    eng.set_weights(params)
    per, l2, pred = eng.train_fwd_bwd(batch[:6], [b[..., 0] for b in batch[6:9]], batch[10], want_pred=True)
    flat, worst, met = _flat_vs_fp64(eng, oracle, params, batch, 2, 8, 4)
    assert relerr(pred.cpu().numpy(), met["pred"]) < 1e-4
    np.testing.assert_allclose(per[:, 0].cpu().numpy() + float(l2), met["loss"], rtol=1e-4)
    bars("trainer_default/P16r2l8h4B2/flat", flat, 2e-4)
    bars("trainer_default/P16r2l8h4B2/worst_tensor", worst, 3e-3)
    got = eng.grads.clone()
    # the unfused path gives the same gradient up to rounding of the split copies' exponents
    eng.set_option(pkg._lib.OPT_FUSED_DGRAD, 0)
    eng.train_fwd_bwd(batch[:6], [b[..., 0] for b in batch[6:9]], batch[10])
    bars("trainer_default/P16r2l8h4B2/unfused_vs_fused", float((eng.grads - got).norm() / got.norm()), 1e-4)
    eng.close()


def test_three_train_steps_follow_the_oracle_loop(pkg, oracle, bars):
    """Three consecutive TrainerController.train_step calls (forward, backward, Adam, refreshed tensor-core weight
    images) against the oracle's fp64 loop: per-step loss trajectory and the accumulated weight update."""
    import contextlib
    import io
    tcm = importlib.import_module("4dflownet_b200.Network.TrainerController")
    P, r, low, hi, B, lr = 8, 2, 2, 1, 3, 1e-3
    params = oracle.glorot_params(low, hi, seed=17, bias_scale=0.05)
    batches = [oracle.synthetic_batch(B, P, r, seed=30 + i) for i in range(3)]
    with contextlib.redirect_stdout(io.StringIO()):
        ctl = tcm.TrainerController(P, r, lr, False, "t", low, hi, max_batch=B)
    ctl.model.set_weights(params)
    got_losses = []
    for bt in batches:
        ctl.reset_metrics()
        ctl.train_step(bt)
        got_losses.append(ctl.loss_metrics["train_loss"].result())
    w_got = dict(zip(ctl.model.variable_names, ctl.model.get_weights()))
    # oracle loop (fp64 autograd + numpy Adam)
    w = {k: v.astype(np.float64) for k, v in params.items()}
    m = {k: 0.0 for k in w}
    v = {k: 0.0 for k in w}
    want_losses = []
    for t, bt in enumerate(batches, 1):
        g, met = oracle.gradients(w, bt, r, low, hi)
        want_losses.append(float(np.mean(met["loss"])))
        for k in w:
            w[k], m[k], v[k] = oracle.adam_step(w[k], g[k], m[k], v[k], t, lr)
    np.testing.assert_allclose(got_losses, want_losses, rtol=2e-4)
    assert want_losses[2] != want_losses[0]
    num = sum(float(np.sum((w_got[k] - w[k]) ** 2)) for k in w)
    den = sum(float(np.sum((w[k] - params[k]) ** 2)) for k in w)
    # accumulated update (Adam's sign-like first steps amplify gradient noise near zero: the update of a weight whose
    # gradient is ~1e-6 flips with the gradient's last bits; the loss trajectory above is the tight check)
    bars("three_steps/P8r2l2h1B3/update_rel_l2", (num / den) ** 0.5, 2e-2)
    assert ctl.optimizer.iterations == 3


def test_training_reduces_the_loss_on_a_fixed_batch(pkg, oracle):
    """Overfit check through the public controller: 25 Adam steps on one fixed batch must cut the masked-MSE loss
    (the regulariser is ~1e-3 of it) by a large factor -- catches sign / scale errors no single-step parity test sees."""
    import contextlib
    import io
    tcm = importlib.import_module("4dflownet_b200.Network.TrainerController")
    P, r = 12, 2
    with contextlib.redirect_stdout(io.StringIO()):
        ctl = tcm.TrainerController(P, r, 2e-3, False, "t", 2, 1, max_batch=4, seed=3)
    batch = oracle.synthetic_batch(4, P, r, seed=8)
    losses = []
    for _ in range(25):
        ctl.reset_metrics()
        ctl.train_step(batch)
        losses.append(ctl.loss_metrics["train_mse"].result())
    assert np.isfinite(losses).all()
    assert losses[-1] < 0.5 * losses[0], losses
    assert losses[-1] == min(losses[-5:]) or losses[-1] < 0.6 * losses[0]


def test_train_step_from_host_buffers_equals_device_buffers(pkg, oracle):
    """engine.train_fwd_bwd uploads host-resident HR targets / mask on a side stream underneath the forward
    (sr4d_train_forward + sr4d_train_backward); the result must be bit-identical to the single sr4d_train_fwd_bwd call on
    device-resident tensors -- three steps in a row (the staging buffers are reused), pinned and pageable sources."""
    import contextlib
    import io
    import torch
    tcm = importlib.import_module("4dflownet_b200.Network.TrainerController")
    P, r, B = 12, 2, 3
    ctls = []
    for _ in range(3):
        with contextlib.redirect_stdout(io.StringIO()):
            ctls.append(tcm.TrainerController(P, r, 1e-3, False, "t", 2, 1, max_batch=4, seed=5))
    for step in range(3):
        batch = [np.ascontiguousarray(a) for a in oracle.synthetic_batch(B, P, r, seed=20 + step)]
        dev = [torch.from_numpy(a).cuda() for a in batch]
        pinned = [torch.from_numpy(a).pin_memory() for a in batch]
        ctls[0].train_step(dev)
        ctls[1].train_step(pinned)
        ctls[2].train_step(batch)            # numpy (pageable)
        torch.cuda.synchronize()
        for c in ctls[1:]:
            assert torch.equal(c.engine.grads, ctls[0].engine.grads), f"gradients differ at step {step}"
            assert torch.equal(c.engine.params, ctls[0].engine.params), f"weights differ at step {step}"
    for k in ("train_loss", "train_mse", "train_accuracy"):
        assert ctls[1].loss_metrics[k].result() == ctls[0].loss_metrics[k].result()
        assert ctls[2].loss_metrics[k].result() == ctls[0].loss_metrics[k].result()
