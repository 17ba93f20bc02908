"""End-to-end drop-in checks on the GPU: the predictor.py flow (HDF5 volume -> patches -> engine -> stitched,
de-normalised result file) against the oracle, and the trainer.py flow (CSV index -> PatchHandler3D batches ->
TrainerController epochs -> loss.csv / best model / optimizer.pkl / quicksave -> restore)."""
import importlib
import os
import pickle
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import synth  # noqa: E402

pytestmark = pytest.mark.gpu


def test_predictor_main_on_hdf5_volume(pkg, oracle, tmp_path):
    h5io = importlib.import_module("4dflownet_b200.utils.h5io")
    predictor = importlib.import_module("4dflownet_b200.predictor")
    d = str(tmp_path)
    synth.make_synthetic_h5(d)                                     # LR volume 10x9x8, two rows, venc differs per row
    P, r, low, hi = 8, 2, 1, 1
    params = oracle.glorot_params(low, hi, seed=8, bias_scale=0.02)
    wpath = os.path.join(d, "weights.h5")
    h5io.save_keras_weights(wpath, params)                          # Keras layout, read back by load_weights
    predictor.main(data_dir=d, filename="synth_LR.h5", output_dir=os.path.join(d, "result"),
                   output_filename="out.h5", model_path=wpath, patch_size=P, res_increase=r, batch_size=3,
                   round_small_values=True, low_resblock=low, hi_resblock=hi)
    p64 = {k: torch.tensor(v, dtype=torch.float64) for k, v in params.items()}
    with h5io.open_file(os.path.join(d, "result", "out.h5"), "r") as res, h5io.open_file(os.path.join(d, "synth_LR.h5"), "r") as lr:
        assert res["u"].shape == (2, 20, 18, 16) and res["dx"].shape == (2, 3)
        np.testing.assert_allclose(res["dx"][0], lr["dx"][0] / r)
        for row in range(2):
            venc = max(float(lr[k][row]) for k in ("venc_u", "venc_v", "venc_w"))
            vel = [np.asarray(lr[c][row]) / np.float32(venc) for c in "uvw"]
            mag = [np.asarray(lr["mag_" + c][row]) / 4095. for c in "uvw"]
            stacks = [oracle.patchify(a.astype(np.float32), P)[0] for a in vel + mag]
            y = oracle.forward(p64, [torch.tensor(s[..., None], dtype=torch.float64) for s in stacks], r, low, hi).numpy()
            for ci, c in enumerate("uvw"):
                want = oracle.patchup(y[..., ci], vel[0].shape, P, r) * venc
                got = np.asarray(res[c][row])
                live = np.abs(want) >= 1.5 * venc / 2048              # away from the zeroing threshold
                assert np.abs(got - want)[live].max() < 1e-4 * np.abs(want).max()
                assert np.all(got[np.abs(want) < 0.5 * venc / 2048] == 0)


def test_predictor_config0_vs_reference_script_golden(pkg, tmp_path):
    """BASELINE configs[0] (patch 24, r=2, 8/4 blocks, 42x38x36 volume): the product's predictor.main against the
    output file of the reference's own src/predictor.py run unmodified on the numpy TF stand-in
    (tests/golden/make_predictor_golden.py).  Inputs are regenerated from the seeds stored with the golden."""
    gold = np.load(os.path.join(HERE, "golden", "predictor_golden.npz"))
    h5io = importlib.import_module("4dflownet_b200.utils.h5io")
    predictor = importlib.import_module("4dflownet_b200.predictor")
    d = str(tmp_path)
    synth.make_example_lr(os.path.join(d, "example_data.h5"), int(gold["data_seed"]))
    wpath = os.path.join(d, "4DFlowNet.h5")
    h5io.save_keras_weights(wpath, synth.keras_weight_dict(8, 4, int(gold["weight_seed"])))
    predictor.main(data_dir=d, filename="example_data.h5", output_dir=os.path.join(d, "result"),
                   output_filename="example_result.h5", model_path=wpath)          # the reference's defaults otherwise
    venc, flips = 1.5, 0
    with h5io.open_file(os.path.join(d, "result", "example_result.h5"), "r") as res:
        np.testing.assert_allclose(res["dx"][...], gold["dx"], rtol=1e-6)
        for c in "uvw":
            got = np.asarray(res[c][...])
            assert got.shape == (1, 84, 76, 72) and got.dtype == np.float32
            want = gold[c]
            sub = got[0, ::2, ::2, ::2].astype(np.float64)
            tol = 1e-4 * float(gold[c + "_moments"][4])                     # 1e-4 of max|ref| (north-star bar)
            bad = np.abs(sub - want) > tol
            # a value within tol of the venc/2048 zeroing threshold may be zeroed on one side only
            flip = bad & ((sub == 0) | (want == 0)) & (np.maximum(np.abs(sub), np.abs(want)) < venc / 2048 + 2 * tol)
            assert not np.any(bad & ~flip), float(np.abs(sub - want)[bad & ~flip].max())
            flips += int(flip.sum())
            g64 = got.astype(np.float64)
            m = gold[c + "_moments"]
            assert abs(g64.sum() - m[0]) < 1e-4 * m[1]
            assert abs(np.abs(g64).sum() - m[1]) < 1e-4 * m[1]
            assert abs((g64 ** 2).sum() - m[2]) < 2e-4 * m[2]
            assert abs(float((got == 0).sum()) - m[3]) <= max(8.0, 0.02 * m[3])
    assert flips <= 8


def test_trainer_flow_and_restore(pkg, oracle, tmp_path, capsys):
    h5io = importlib.import_module("4dflownet_b200.utils.h5io")
    trainer = importlib.import_module("4dflownet_b200.trainer")
    tcm = importlib.import_module("4dflownet_b200.Network.TrainerController")
    d = str(tmp_path)
    synth.make_synthetic_h5(d)
    header = "source,target,index,start_x,start_y,start_z,rotate,rotation_plane,rotation_degree_idx,coverage\n"
    for name, rows in (("train.csv", synth.ROWS), ("validate.csv", synth.ROWS[:4]), ("benchmark.csv", synth.ROWS[2:6])):
        with open(os.path.join(d, name), "w") as f:
            f.write(header + "".join(",".join(r) + "\n" for r in rows))
    net = trainer.main(data_dir=d, QUICKSAVE=True, initial_learning_rate=1e-3, epochs=2, batch_size=4,
                       mask_threshold=0.6, network_name="t4d", patch_size=synth.PATCH, res_increase=synth.R,
                       low_resblock=1, hi_resblock=1, models_root=os.path.join(d, "models"))
    md = net.model_dir
    log = open(os.path.join(md, "loss.csv")).read().splitlines()
    rows = [ln for ln in log if ln[:1].isdigit()]
    assert len(rows) == 2 and rows[0].startswith("1,") and "**" in rows[0]            # epoch 1 is always a best epoch
    cols = [c.strip() for c in [ln for ln in log if ln.startswith("epoch")][0].split(",")]
    assert cols[:10] == ["epoch", "train_loss", "val_loss", "train_accuracy", "val_accuracy", "train_mse", "val_mse",
                         "train_div", "val_div", "l2_reg_loss"]
    vals = [float(x) for x in rows[1].split(",")[1:10]]
    assert np.isfinite(vals).all() and vals[0] > 0 and vals[8] > 0
    # 11 train samples -> 3 steps per epoch, 2 epochs
    assert net.optimizer.iterations == 6
    with open(os.path.join(md, "optimizer.pkl"), "rb") as f:
        ow = pickle.load(f)
    assert len(ow) == 1 + 2 * len(net.engine.table) and int(ow[0]) in (3, 6)
    with h5io.open_file(os.path.join(md, "quicksave_t4d.h5"), "r") as q:
        H = synth.PATCH * synth.R
        assert q["u"].shape[1:] == (4, H, H, H) and q["epoch"][0] == 1 and "lr_u" in q and "mask" in q
    # restore into a fresh controller: same weights, same optimizer state
    best = os.path.join(md, "t4d-best.h5")
    assert os.path.exists(best)
    net2 = tcm.TrainerController(synth.PATCH, synth.R, 1e-3, False, "t4d", 1, 1, max_batch=4)
    net2.restore_model(md, "t4d-best.h5")
    saved = h5io.load_keras_weights(best, net.model.variable_names)
    for n, w2 in zip(net2.model.variable_names, net2.model.get_weights()):
        np.testing.assert_array_equal(w2, saved[n])
    assert net2.optimizer.iterations == int(ow[0])
    np.testing.assert_array_equal(net2.optimizer.weights[1], ow[1])
    # and training continues from there
    batch = oracle.synthetic_batch(4, synth.PATCH, synth.R, seed=1)
    net2.train_step(batch)
    assert net2.optimizer.iterations == int(ow[0]) + 1 and np.isfinite(net2.loss_metrics["train_loss"].result())


@pytest.mark.parametrize("name_offset", [0, 12])
def test_load_weights_from_tf22_model_save_layout(pkg, oracle, tmp_path, name_offset):
    """predictor.py:61 `network.load_weights(model_path)` on a file laid out like TF 2.2's `model.save` (root / group
    attributes, weightless layers, nested <layer>/<layer>/kernel:0; tests/golden/keras_layout.py), incl. a writer whose
    Conv3D names are offset: the loaded network predicts exactly what the same weights set directly predict, and the
    activation-range flag stays clear."""
    import keras_layout
    h5io = importlib.import_module("4dflownet_b200.utils.h5io")
    P, r, low, hi = 8, 2, 1, 1
    params = oracle.glorot_params(low, hi, seed=40, bias_scale=0.05)
    path = str(tmp_path / "4DFlowNet.h5")
    keras_layout.write_tf22_model_file(h5io, path, params, name_offset)
    net = pkg.prepare_network(P, r, low, hi, max_batch=2)
    net.load_weights(path)
    ref = pkg.prepare_network(P, r, low, hi, max_batch=2)
    ref.set_weights([params[n] for n in ref.variable_names])
    batch = oracle.synthetic_batch(2, P, r, seed=1)
    assert np.array_equal(net.predict(list(batch[:6])), ref.predict(list(batch[:6])))
    assert net.engine.activation_overflow() is False
    # a diverged network (huge weights) trips the flag instead of saturating silently
    big = {k: (v * 1e4 if k.endswith("kernel") else v) for k, v in params.items()}
    ref.set_weights([big[n] for n in ref.variable_names])
    ref.predict(list(batch[:6]))
    assert ref.engine.activation_overflow() is True and ref.engine.activation_overflow() is False      # reset on read
