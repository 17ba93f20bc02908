"""GPU parity tests, forward path: CUDA kernels (through the C ABI) vs the CPU oracle.
Tolerance (north_star): 1e-4 relative, measured as max|d| / max|ref| (element-wise relative
error is meaningless at zero crossings)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-4


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)


@pytest.fixture(scope="module")
def eng_small(pkg):
    return pkg.Engine(8, 2, 1, 1, max_batch=3, training=False, device=0)


@pytest.mark.parametrize("impl", ["simt", "tcgen05"])
@pytest.mark.parametrize("D,B", [(8, 2), (5, 1), (24, 1), (26, 1), (12, 2), (48, 1)])
@pytest.mark.parametrize("variant", ["linear", "bias_relu", "res_lrelu"])
def test_conv64_layer(pkg, oracle, eng_small, D, B, variant, impl):
    if impl == "simt" and D == 48:
        pytest.skip("covered by the tensor-core case")
    g = np.random.default_rng(D * 10 + B)
    x = g.standard_normal((B, D, D, D, 64)).astype(np.float32)
    k = (g.standard_normal((3, 3, 3, 64, 64)) * 0.04).astype(np.float32)
    bias = res = None
    slope = 1.0
    if variant == "bias_relu":
        bias = g.standard_normal(64).astype(np.float32)
        slope = 0.0
    if variant == "res_lrelu":
        res = g.standard_normal((B, D, D, D, 64)).astype(np.float32)
        slope = 0.2
    impl_id = pkg._lib.CONV_SIMT if impl == "simt" else pkg._lib.CONV_TCGEN05
    y = eng_small.conv64_layer(x, k, bias, res, slope, impl=impl_id).cpu().numpy()
    t = oracle.conv3d(torch.tensor(x, dtype=torch.float64), torch.tensor(k, dtype=torch.float64),
                      None if bias is None else torch.tensor(bias, dtype=torch.float64))
    if res is not None:
        t = t + torch.tensor(res, dtype=torch.float64)
    if slope == 0.0:
        t = torch.relu(t)
    elif slope == 0.2:
        t = oracle.leaky_relu(t)
    assert relerr(y, t.numpy()) < 1e-5


@pytest.mark.parametrize("P,r", [(8, 2), (6, 4), (24, 2), (5, 3)])
def test_upsample_layer(pkg, oracle, P, r):
    eng = pkg.Engine(P, r, 0, 0, max_batch=2, training=False, device=0)
    g = np.random.default_rng(P + r)
    x = g.standard_normal((2, P, P, P, 64)).astype(np.float32)
    y = eng.upsample_layer(x).cpu().numpy()
    t = oracle.upsample3d(torch.tensor(x), r).numpy()          # fp32 oracle, same lerp order
    assert relerr(y, t) < 2e-6
    eng.close()


@pytest.mark.parametrize("P,r,low,hi,B", [(8, 2, 1, 1, 3), (6, 1, 2, 1, 2), (10, 3, 0, 2, 1), (12, 2, 2, 0, 2)])
def test_forward_small_vs_oracle_fp64(pkg, oracle, P, r, low, hi, B):
    params = oracle.glorot_params(low, hi, seed=P, bias_scale=0.05)
    batch = oracle.synthetic_batch(B, P, r, seed=3)
    eng = pkg.Engine(P, r, low, hi, max_batch=B, training=False, device=0)
    eng.set_option(pkg._lib.OPT_CONV_IMPL, pkg._lib.CONV_SIMT)
    eng.set_weights(params)
    y = eng.forward(batch[:6]).cpu().numpy()
    p64 = {k: torch.tensor(v, dtype=torch.float64) for k, v in params.items()}
    ref = oracle.forward(p64, [torch.tensor(b, dtype=torch.float64) for b in batch[:6]], r, low, hi).numpy()
    assert y.shape == ref.shape
    assert relerr(y, ref) < TOL
    # batch-size independence / chunked predict() path
    model_y = eng.forward([b[:1] for b in batch[:6]]).cpu().numpy()
    assert relerr(model_y, ref[:1]) < TOL
    eng.close()


def test_forward_full_config_vs_oracle(pkg, oracle):
    """BASELINE config 1/2 geometry: P=24, r=2, 8 low / 4 hi resblocks."""
    params = oracle.glorot_params(8, 4, seed=1234, bias_scale=0.02)
    batch = oracle.synthetic_batch(1, 24, 2, seed=0)
    eng = pkg.Engine(24, 2, 8, 4, max_batch=2, training=False, device=0)
    assert eng.param_count == 3342083 and len(eng.table) == 48
    eng.set_weights(params)
    y = eng.forward(batch[:6]).cpu().numpy()
    p32 = {k: torch.tensor(v) for k, v in params.items()}
    ref = oracle.forward(p32, [torch.tensor(b) for b in batch[:6]], 2, 8, 4).numpy()
    assert relerr(y, ref) < TOL
    eng.close()


@pytest.mark.parametrize("P,r,low,hi,B", [(8, 2, 1, 1, 3), (12, 2, 2, 1, 2), (24, 2, 8, 4, 1), (24, 2, 2, 1, 4)])
def test_chained_forward_launches_are_bit_identical(pkg, oracle, P, r, low, hi, B):
    """SR4D_OPT_FWD_CHAIN: a run of 64->64 layers as one cooperative launch (grid-wide barrier between layers, odd and even
    tile counts per CTA, 1..4 tiles per SM) must produce exactly the bytes of the per-layer launches -- same kernels, same
    arithmetic, only the launch boundaries differ.  Three forwards in a row also exercise the barrier re-arming."""
    L = pkg._lib
    params = oracle.glorot_params(low, hi, seed=P + B, bias_scale=0.02)
    batch = oracle.synthetic_batch(B, P, r, seed=3)
    eng = pkg.Engine(P, r, low, hi, max_batch=B, training=False, device=0)
    eng.set_weights(params)
    eng.set_option(L.OPT_FWD_CHAIN, 0)
    ref = eng.forward(batch[:6]).clone()
    for tiles_per_sm in (4, 64):
        eng.set_option(L.OPT_FWD_CHAIN, tiles_per_sm)
        for _ in range(3):
            got = eng.forward(batch[:6])
            assert torch.equal(got, ref)
    eng.close()


def test_model_protocol_and_stitch(pkg, oracle):
    P, r = 8, 2
    model = pkg.prepare_network(P, r, 1, 1, max_batch=4)
    params = oracle.glorot_params(1, 1, seed=2)
    model.set_weights([params[n] for n in model.variable_names])
    got = model.get_weights()
    assert all(np.array_equal(a, params[n]) for a, n in zip(got, model.variable_names))

    class DS:
        pass
    g = np.random.default_rng(5)
    ds = DS()
    for n in ("u", "v", "w"):
        setattr(ds, n, g.uniform(-1, 1, (10, 9, 11)).astype(np.float32))
    for n in ("mag_u", "mag_v", "mag_w"):
        setattr(ds, n, g.uniform(0, 0.016, (10, 9, 11)).astype(np.float32))
    ds.venc = np.float32(1.5)
    ds.velocity_per_px = ds.venc / 2048
    pg = pkg.PatchGenerator(P, r)
    pred_mod = importlib_predictor(pkg)
    vol_gpu = pred_mod.predict_volume(model, pg, ds, batch_size=4, gpu_stitch=True)
    vol_cpu = pred_mod.predict_volume(model, pg, ds, batch_size=4, gpu_stitch=False)
    assert vol_gpu.shape == (3, 20, 18, 22)
    assert np.array_equal(vol_gpu, vol_cpu)          # integer indexing + identical fp32 ops: bit-exact
    # predict() (numpy in / numpy out, chunked) equals the tensor path
    vel, mag = pg.patchify(ds)
    y = model.predict([*vel, *mag], batch_size=3)
    y2 = model([*vel, *mag][0:6] if len(vel[0]) <= 4 else [a[:4] for a in (*vel, *mag)]).cpu().numpy()
    assert np.array_equal(y[:len(y2)], y2)


def importlib_predictor(pkg):
    import importlib
    return importlib.import_module(pkg.__name__ + ".predictor")
