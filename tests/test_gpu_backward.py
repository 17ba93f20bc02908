"""GPU parity tests, training path: loss/metric, gradients (vs fp64 autograd of the oracle) and Adam."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


@pytest.mark.parametrize("impl", ["simt", "tcgen05_full", "tcgen05", "tcgen05_hi_only_old_kernel"])
@pytest.mark.parametrize("D,B,dy_scale", [(6, 2, 1.0), (9, 1, 1.0), (24, 1, 3e-7), (26, 1, 1e3), (24, 3, 1.0), (48, 1, 1.0)])
def test_conv64_layer_bwd(pkg, bars, D, B, dy_scale, impl):
    """Backward kernels of one 64->64 layer given identical inputs, against float64 autograd.
    dy_scale exercises the power-of-two rescaling of the split-fp16 gradient operand: loss gradients of this network
    are ~1e-6, far below the fp16 normal range.
    tcgen05_full: both gradient planes in the dgrad, the two-plane weight-gradient kernel (1e-5, like the fp32 SIMT
    anchor).  tcgen05 (the default since round 2): single scaled-fp16 gradient plane in the dgrad and the stacked
    hi-plane weight-gradient kernel -- their rounding errors are independent per element (2^-12 relative) and average
    out over the 1728-term (dgrad) / B*D^3-term (wgrad) sums, so the bar scales with 1/sqrt(#terms)."""
    L = pkg._lib
    eng = pkg.Engine(8, 2, 0, 0, max_batch=2, training=False, device=0)
    single = {"tcgen05_full": (0, 0), "tcgen05": (1, 1), "tcgen05_hi_only_old_kernel": (1, 2)}.get(impl, (0, 0))
    eng.set_option(L.OPT_DGRAD_SINGLE, single[0])
    eng.set_option(L.OPT_WGRAD_SINGLE, single[1])
    g = np.random.default_rng(D)
    x = g.standard_normal((B, D, D, D, 64)).astype(np.float32)
    k = (g.standard_normal((3, 3, 3, 64, 64)) * 0.05).astype(np.float32)
    dy = (g.standard_normal((B, D, D, D, 64)) * dy_scale).astype(np.float32)
    dx, dk, db = eng.conv64_layer_bwd(x, k, dy, impl=L.CONV_SIMT if impl == "simt" else L.CONV_TCGEN05)
    import importlib
    oracle = importlib.import_module("oracle.sr4d_oracle")
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    kt = torch.tensor(k, dtype=torch.float64, requires_grad=True)
    bt = torch.zeros(64, dtype=torch.float64, requires_grad=True)
    y = oracle.conv3d(xt, kt, bt)
    (y * torch.tensor(dy, dtype=torch.float64)).sum().backward()
    tag = f"layer_bwd/D{D}B{B}s{dy_scale:g}/{impl}"
    e_dx, e_dk = rel_l2(dx.cpu().numpy(), xt.grad.numpy()), rel_l2(dk.cpu().numpy(), kt.grad.numpy())
    if single == (0, 0):
        assert e_dx < 1e-5 and e_dk < 1e-5, (e_dx, e_dk)
    else:
        # Single-plane operands: every element carries an independent fp16 rounding error (rms 2^-11/sqrt(3) = 2.8e-4
        # relative).  x, k and dy are INDEPENDENT random tensors here, so every output is a random-sign sum, and a
        # random-sign sum inherits the relative error of its terms: ~2e-4 per element of dx, ~3e-4 per element of dk
        # (both operands rounded), independent from element to element.  This is the worst case by construction: in the
        # network the weight-gradient sums are not pure noise and the per-voxel errors average out over them, which is
        # what test_backward_kernels_given_identical_gates and tools/grad_parity.py measure (1e-5 .. 5e-5 on the flat
        # gradient, the same as the two-plane kernels).
        bars(tag + "/dx", e_dx, 6e-4)
        bars(tag + "/dk", e_dk, 1e-3)
    assert rel_l2(db.cpu().numpy(), bt.grad.numpy()) < 1e-5
    eng.close()


@pytest.mark.parametrize("impl", ["simt", "tcgen05_full", "tcgen05"])
@pytest.mark.parametrize("D,B,g_scale,c", [(6, 2, 1.0, 0), (8, 2, 1.0, 1), (12, 1, 1e-6, 1), (16, 2, 1.0, 2), (24, 3, 3e-7, 0), (36, 1, 1e3, 1),
                                           (48, 2, 1.0, 2), (48, 8, 1e-6, 0)])
def test_head_layer_bwd(pkg, bars, D, B, g_scale, c, impl):
    """Whole backward of one 64->1 head (relu -> clamp-padded conv3d with one filter, SR4DFlowNet.py:40-49) given
    identical inputs, against float64 autograd: input gradient incl. ReluGrad and MirrorPadGrad, kernel gradient, both
    bias gradients.  tcgen05_full (two-plane backward options): the G table, the weights and the saved activations enter
    the tensor cores as hi + lo fp16 pairs (three products), so every output is fp32-accurate; dx is read back from the
    split-fp16 copy the consumers use.  tcgen05 (the training default, single-plane backward): G, the saved activation
    and the output gradient are single fp16 planes -- independent 2^-12 roundings per element, as in the 64->64 layers'
    single-plane dgrad / wgrad (test_conv64_layer_bwd explains why random inputs are the worst case for them)."""
    L = pkg._lib
    if impl != "simt" and D % 4:
        pytest.skip("tensor-core head backward needs D % 4 == 0 (16-byte strides of the planar gradient); the engine falls back to the SIMT kernel")
    eng = pkg.Engine(8, 2, 0, 0, max_batch=2, training=False, device=0)
    single = 0 if impl == "tcgen05_full" else 1
    eng.set_option(L.OPT_DGRAD_SINGLE, single)
    eng.set_option(L.OPT_WGRAD_SINGLE, single)
    rng = np.random.default_rng(100 + D + B)
    x = np.maximum(rng.standard_normal((B, D, D, D, 64)), 0).astype(np.float32)      # a ReLU output: about half zeros
    k = (rng.standard_normal((3, 3, 3, 64, 1)) * 0.05).astype(np.float32)
    g = (rng.standard_normal((B, D, D, D, 3)) * g_scale).astype(np.float32)
    dx, dk, db, db1 = eng.head_layer_bwd(x, k, g, c, impl=L.CONV_SIMT if impl == "simt" else L.CONV_TCGEN05)
    # float64 reference: the head sees a = relu(pre) with pre > 0 exactly where x > 0; d pre = relu'(x) * d a
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    kt = torch.tensor(k, dtype=torch.float64, requires_grad=True)
    xp = torch.nn.functional.pad(xt.permute(0, 4, 1, 2, 3), (1, 1, 1, 1, 1, 1), mode="replicate")
    y = torch.nn.functional.conv3d(xp, kt.permute(4, 3, 0, 1, 2))[:, 0]
    gt = torch.tensor(g[..., c], dtype=torch.float64)
    (y * gt).sum().backward()
    dx_ref = (xt.grad * (xt > 0)).numpy()
    ref_db1 = dx_ref.sum(axis=(0, 1, 2, 3))
    e_dx, e_dk = rel_l2(dx.cpu().numpy(), dx_ref), rel_l2(dk.cpu().numpy().reshape(-1), kt.grad.numpy().reshape(-1))
    e_db1 = rel_l2(db1.cpu().numpy(), ref_db1)
    np.testing.assert_allclose(db.cpu().numpy()[0], gt.sum().item(), rtol=2e-5, atol=1e-5 * g_scale * np.sqrt(gt.numel()))
    if impl == "tcgen05":
        tag = f"head_bwd/D{D}B{B}s{g_scale:g}"
        bars(tag + "/dx", e_dx, 6e-4)
        bars(tag + "/dk", e_dk, 1e-3)
        bars(tag + "/db_prev", e_db1, 1e-3)
    else:
        assert e_dx < 2e-6 and e_dk < 1e-5 and e_db1 < 3e-5, (e_dx, e_dk, e_db1)
    eng.close()


def test_loss_metrics(pkg, oracle):
    P, r, B = 8, 2, 3
    H = P * r
    eng = pkg.Engine(P, r, 0, 0, max_batch=B, training=False, device=0)
    g = np.random.default_rng(0)
    pred = (g.standard_normal((B, H, H, H, 3)) * 0.1).astype(np.float32)
    mask = (g.uniform(size=(B, H, H, H)) < 0.2).astype(np.float32)
    hr = (g.standard_normal((B, H, H, H, 3)) * 0.08).astype(np.float32) * mask[..., None]
    per = eng.loss_metrics(pred, hr[..., 0], hr[..., 1], hr[..., 2], mask).cpu().numpy()
    tot, mse, _ = oracle.loss_function(torch.tensor(hr), torch.tensor(pred), torch.tensor(mask))
    rel = oracle.calculate_relative_error(torch.tensor(hr), torch.tensor(pred), torch.tensor(mask))
    np.testing.assert_allclose(per[:, 0], tot.numpy(), rtol=2e-6)
    np.testing.assert_allclose(per[:, 1], mse.numpy(), rtol=2e-6)
    np.testing.assert_allclose(per[:, 2], rel.numpy(), rtol=1e-4)   # rounding at 1e-4 steps can flip single voxels
    np.testing.assert_allclose(per[:, 3], mask.sum(axis=(1, 2, 3)), rtol=0)
    eng.close()


def _flat(grads, names):
    return np.concatenate([np.asarray(grads[n], np.float64).ravel() for n in names])


def _engine_grads(pkg, P, r, low, hi, B, params, batch, fwd_impl, bwd_impl, dgrad_single=1, wgrad_single=1, fused=1):
    """Gradient of one train step with the forward and the backward convolutions on possibly different kernels
    (sr4d_train_forward / sr4d_train_backward)."""
    L = pkg._lib
    eng = pkg.Engine(P, r, low, hi, max_batch=B, training=True, device=0)
    eng.set_option(L.OPT_FUSED_DGRAD, fused)
    eng.set_option(L.OPT_DGRAD_SINGLE, dgrad_single)
    eng.set_option(L.OPT_WGRAD_SINGLE, wgrad_single)
    eng.set_weights(params)
    eng.set_option(L.OPT_CONV_IMPL, fwd_impl)
    eng.train_forward(batch[:6])
    eng.set_option(L.OPT_CONV_IMPL, bwd_impl)
    eng.train_backward([b[..., 0] for b in batch[6:9]], batch[10])
    g = {n: v.cpu().numpy().astype(np.float64) for n, v in eng.tensor_views(eng.grads)}
    eng.close()
    return g


GEOMS = [(8, 2, 1, 1, 2), (6, 1, 1, 1, 2), (6, 2, 0, 1, 1), (6, 2, 2, 0, 2), (6, 1, 0, 0, 1), (12, 2, 2, 2, 1)]


@pytest.mark.parametrize("P,r,low,hi,B", GEOMS + [(24, 2, 2, 1, 2)])
def test_backward_kernels_given_identical_gates(pkg, oracle, bars, P, r, low, hi, B):
    """The decisive separation: the tensor-core BACKWARD fed the fp32 SIMT forward's saved
    activations -- identical ReLU / LeakyReLU gates -- against the SIMT backward on the same activations.  What is
    left is the backward kernels' own arithmetic: split-fp16 weights, one scaled fp16 gradient plane, hi-plane weight
    gradient, truncating tcgen05 accumulators with chains cut at 384 accumulations."""
    L = pkg._lib
    params = oracle.glorot_params(low, hi, seed=P + r, bias_scale=0.05)
    batch = oracle.synthetic_batch(B, P, r, seed=4)
    names = [n for n, _ in oracle.param_table(low, hi)]
    ref = _flat(_engine_grads(pkg, P, r, low, hi, B, params, batch, L.CONV_SIMT, L.CONV_SIMT), names)
    tag = f"P{P}r{r}l{low}h{hi}B{B}"
    for label, kw in (("default", {}), ("full", dict(dgrad_single=0, wgrad_single=0)), ("unfused", dict(fused=0))):
        got = _flat(_engine_grads(pkg, P, r, low, hi, B, params, batch, L.CONV_SIMT, L.CONV_AUTO, **kw), names)
        # ceiling: the north-star 1e-4 at real patch sizes; the single-plane operands' averaging is weaker on toy grids
        ceiling = 1e-4 if B * (P * r) ** 3 >= 48 ** 3 else 3e-4
        bars(f"identical_gates/{tag}/{label}", rel_l2(got, ref), ceiling)


@pytest.mark.parametrize("P,r,low,hi,B", [(8, 3, 1, 1, 3), (6, 4, 1, 1, 2), (9, 2, 1, 1, 1), (12, 1, 1, 2, 2), (10, 3, 2, 1, 2)])
def test_backward_kernels_on_unusual_geometries(pkg, oracle, bars, P, r, low, hi, B):
    """res_increase 3 / 4 / 1, odd patch edges and batches: grids where the tensor-core head backward is not available
    (edge % 4 != 0: SIMT head kernel), where tile counts are odd (no CTA pairs) and where the forward is chained -- the same
    identical-gates comparison as above, every fallback combination against the SIMT backward."""
    L = pkg._lib
    params = oracle.glorot_params(low, hi, seed=P * 7 + r, bias_scale=0.05)
    batch = oracle.synthetic_batch(B, P, r, seed=5)
    names = [n for n, _ in oracle.param_table(low, hi)]
    ref = _flat(_engine_grads(pkg, P, r, low, hi, B, params, batch, L.CONV_SIMT, L.CONV_SIMT), names)
    got = _flat(_engine_grads(pkg, P, r, low, hi, B, params, batch, L.CONV_SIMT, L.CONV_AUTO), names)
    bars(f"identical_gates_unusual/P{P}r{r}l{low}h{hi}B{B}/default", rel_l2(got, ref), 3e-4)
    # and the whole tensor-core step (forward included) against the fp64 oracle's gradient
    full = _flat(_engine_grads(pkg, P, r, low, hi, B, params, batch, L.CONV_AUTO, L.CONV_AUTO), names)
    grads_ref, _ = oracle.gradients({k: v.astype(np.float64) for k, v in params.items()}, batch, r, low, hi)
    # (the engine's gradient buffer excludes the L2 term, which its Adam kernel folds in)
    want = {n: grads_ref[n] - (B * 2 * oracle.L2_COEFF * params[n] if n.endswith("kernel") else 0.0) for n in names}
    bars(f"train_step_unusual/P{P}r{r}l{low}h{hi}B{B}/flat", rel_l2(full, _flat(want, names)), 1e-3)


@pytest.mark.parametrize("impl", ["simt", "auto", "auto_unfused", "auto_full"])
@pytest.mark.parametrize("P,r,low,hi,B", GEOMS)
def test_train_step_gradients_vs_oracle(pkg, oracle, bars, P, r, low, hi, B, impl):
    params = oracle.glorot_params(low, hi, seed=P + r, bias_scale=0.05)
    batch = oracle.synthetic_batch(B, P, r, seed=4)
    eng = pkg.Engine(P, r, low, hi, max_batch=B, training=True, device=0)
    eng.set_option(pkg._lib.OPT_CONV_IMPL, pkg._lib.CONV_SIMT if impl == "simt" else pkg._lib.CONV_AUTO)
    # "auto": the default tensor-core path (fused dgrad epilogue, single gradient plane, stacked hi-plane wgrad);
    # "auto_unfused": separate raw-dgrad + fold kernel; "auto_full": both gradient planes / two-plane wgrad kernel
    eng.set_option(pkg._lib.OPT_FUSED_DGRAD, 0 if impl == "auto_unfused" else 1)
    if impl == "auto_full":
        eng.set_option(pkg._lib.OPT_DGRAD_SINGLE, 0)
        eng.set_option(pkg._lib.OPT_WGRAD_SINGLE, 0)
    eng.set_weights(params)
    per, l2, pred = eng.train_fwd_bwd(batch[:6], [b[..., 0] for b in batch[6:9]], batch[10], want_pred=True)
    # everything is compared with the FLOAT64 oracle.  fp32 autograd of the same graph (what the reference's TF fp32
    # path amounts to) is computed next to it: on these toy grids fp32 itself is 1e-4..5e-4 off on the ill-conditioned
    # stem kernels, because a forward perturbation of relative size eps flips ~eps of the ReLU / LeakyReLU gates and
    # moves a random-sign gradient sum by ~sqrt(eps).
    gref, met = oracle.gradients({k: v.astype(np.float64) for k, v in params.items()}, batch, r, low, hi)
    g32, _ = oracle.gradients(params, batch, r, low, hi, dtype=torch.float32)
    l2c = oracle.L2_COEFF
    assert abs(float(l2) - float(met["l2"])) <= 1e-5 * float(met["l2"])
    np.testing.assert_allclose(per[:, 0].cpu().numpy() + float(l2), met["loss"], rtol=1e-4)
    np.testing.assert_allclose(per[:, 2].cpu().numpy(), met["rel_err"], rtol=1e-3, atol=1e-3)
    assert np.abs(pred.cpu().numpy() - met["pred"]).max() / np.abs(met["pred"]).max() < 1e-4
    tag = f"P{P}r{r}l{low}h{hi}B{B}/{impl}"
    worst, worst32 = 0.0, 0.0
    for name, view in eng.tensor_views(eng.grads):
        got = view.cpu().numpy()
        want = gref[name] - (B * 2 * l2c * params[name] if name.endswith("kernel") else 0.0)
        worst = max(worst, rel_l2(got, want))
        worst32 = max(worst32, rel_l2(g32[name], gref[name]))
    # the flat gradient the optimizer consumes, and the worst single tensor (small bias tensors of these toy grids feel
    # individual gate flips: fp32 autograd's own worst tensor is printed next to it when the check fails)
    flat_got = np.concatenate([v.cpu().numpy().ravel() for _, v in eng.tensor_views(eng.grads)])
    flat_want = np.concatenate([(gref[n] - (B * 2 * l2c * params[n] if n.endswith("kernel") else 0.0)).ravel()
                                for n, *_ in eng.table])
    flat = rel_l2(flat_got, flat_want)
    bars(f"train_step/{tag}/flat", flat, 1e-4 if impl == "simt" else 3e-4)
    bars(f"train_step/{tag}/worst_tensor", worst, 1e-3 if impl == "simt" else 5e-3)
    assert worst32 < 1e-3, worst32      # sanity of the calibration: fp32 autograd itself on the worst tensor
    # one Adam step (Keras semantics, L2 gradient folded in) vs the oracle's numpy Adam, fed the gradient the
    # engine itself produced: the first step moves a weight by lr*g/(|g|+3.2e-6), so near g ~ 1e-6 the update
    # amplifies gradient noise that the checks above already bound; this isolates the Adam kernel.
    lr = 1e-3
    names = [n for n, *_ in eng.table]
    before = dict(zip(names, eng.get_weights()))
    g_eng = {n: v.cpu().numpy().astype(np.float64) for n, v in eng.tensor_views(eng.grads)}
    eng.adam_step(lr, 1, B * 2 * l2c)
    after = dict(zip(names, eng.get_weights()))
    m_after = {n: v.cpu().numpy() for n, v in eng.tensor_views(eng.adam_m)}
    for name in before:
        w0 = before[name].astype(np.float64)
        g_tot = g_eng[name] + (B * 2 * l2c * w0 if name.endswith("kernel") else 0.0)
        p1, m1, _ = oracle.adam_step(w0, g_tot, 0.0, 0.0, 1, lr)
        upd_ref = p1 - w0
        upd = after[name].astype(np.float64) - w0
        assert np.abs(upd - upd_ref).max() < 2e-3 * lr, name     # fp32 rounding of w - update
        np.testing.assert_allclose(m_after[name], m1, rtol=1e-5, atol=1e-12)
        # and the update implied by the fp64 reference gradient agrees wherever the gradient is well above eps
        p_ref, _, _ = oracle.adam_step(w0, gref[name], 0.0, 0.0, 1, lr)
        big = np.abs(gref[name]) > 1e-2 * np.abs(gref[name]).max()
        if np.abs(gref[name]).max() > 1e-3:
            assert np.percentile(np.abs(upd[big] - (p_ref - w0)[big]), 99) < 0.2 * lr, name
    eng.close()


def test_reference_named_metric_entry_points(pkg, oracle):
    """loss_utils.calculate_relative_error and TrainerController.calculate_and_update_metrics / calculate_mse with the
    reference's argument conventions, against the oracle."""
    import contextlib
    import importlib
    import io
    lu = importlib.import_module("4dflownet_b200.Network.loss_utils")
    tcm = importlib.import_module("4dflownet_b200.Network.TrainerController")
    g = np.random.default_rng(3)
    B, H = 3, 12
    true = (g.standard_normal((B, H, H, H, 3)) * 0.1).astype(np.float32)
    pred = (true + g.standard_normal(true.shape) * 0.02).astype(np.float32)
    mask = (g.uniform(size=(B, H, H, H)) < 0.3).astype(np.float32)
    true[mask == 0] = 0
    want = oracle.calculate_relative_error(torch.tensor(true, dtype=torch.float64), torch.tensor(pred, dtype=torch.float64),
                                           torch.tensor(mask, dtype=torch.float64)).numpy()
    got = lu.calculate_relative_error(pred[..., 0:1], pred[..., 1:2], pred[..., 2:3], true[..., 0:1], true[..., 1:2],
                                      true[..., 2:3], mask).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-3, atol=1e-3)
    with contextlib.redirect_stdout(io.StringIO()):
        ctl = tcm.TrainerController(6, 2, 1e-4, False, "t", 0, 0, max_batch=B)
    loss = ctl.calculate_and_update_metrics(true, pred, mask, 'val')
    lw, _, _ = oracle.loss_function(torch.tensor(true, dtype=torch.float64), torch.tensor(pred, dtype=torch.float64),
                                    torch.tensor(mask, dtype=torch.float64))
    np.testing.assert_allclose(loss, lw.numpy(), rtol=1e-4)
    assert abs(ctl.loss_metrics['val_loss'].result() - float(lw.mean())) < 1e-4 * float(lw.mean())
    mse = ctl.calculate_mse(true[..., 0], true[..., 1], true[..., 2], pred[..., 0], pred[..., 1], pred[..., 2]).cpu().numpy()
    np.testing.assert_allclose(mse, ((pred - true) ** 2).sum(-1), rtol=1e-5, atol=1e-9)
