"""CPU tests of the host-side data path around the hot loop: the pure-Python HDF5
shim, the Keras weight file layout, ImageDataset / prediction_utils mirrors and the PatchHandler3D loader, the
latter bit-exact against golden vectors produced by the reference's own PatchHandler3D
(tests/golden/make_patchhandler_golden.py)."""
import importlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import synth  # noqa: E402

GOLD = os.path.join(HERE, "golden", "patchhandler_golden.npz")


@pytest.fixture(scope="module")
def h5io():
    return importlib.import_module("4dflownet_b200.utils.h5io")


@pytest.fixture(scope="module")
def data_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("synth")
    synth.make_synthetic_h5(str(d))
    return str(d)


def test_h5_shim_round_trip(h5io, tmp_path):
    p = str(tmp_path / "t.h5")
    a = np.arange(2 * 3 * 4 * 5, dtype=np.float32).reshape(2, 3, 4, 5)
    with h5io.File(p, "w") as f:
        f.create_dataset("u", data=a[:1], maxshape=(None, 3, 4, 5), compression="gzip")
        f.create_dataset("n", data=np.asarray([7], np.int64))
        g = f.create_group("model_weights/conv3d_1/conv3d_1")
        g.create_dataset("kernel:0", data=np.ones((3, 3, 3, 2, 2), np.float32))
        f["model_weights"].attrs["layer_names"] = np.asarray([b"conv3d_1"])
    with h5io.File(p, "a") as f:                       # the reference's append idiom (prediction_utils.py:24-28)
        f["u"].resize(f["u"].shape[0] + 1, axis=0)
        f["u"][-1:] = a[1:]
    with h5io.File(p, "r") as f:
        assert set(f.keys()) == {"u", "n", "model_weights"}
        np.testing.assert_array_equal(f["u"][...], a)
        assert f["u"].shape == (2, 3, 4, 5) and f["n"][0] == 7
        assert f["model_weights/conv3d_1/conv3d_1/kernel:0"].shape == (3, 3, 3, 2, 2)
        assert f["model_weights"].attrs["layer_names"][0] == b"conv3d_1"
        assert "nope" not in f and f.get("nope") is None
        with pytest.raises(OSError):
            f.create_dataset("x", data=a)
    with pytest.raises(FileNotFoundError):
        h5io.File(str(tmp_path / "missing.h5"), "r")


def test_h5_shim_reads_chunked_deflate_shuffle(h5io, tmp_path):
    """Hand-assembled chunked + shuffle + deflate dataset (what h5py writes for compression='gzip', shuffle=True):
    exercises the v1 chunk B-tree and the filter pipeline of the reader."""
    import struct
    import zlib
    data = (np.arange(6 * 5, dtype=np.float32).reshape(6, 5) * 1.5 - 7).astype("<f4")
    chunk = (4, 5)
    blobs = []
    for r0 in range(0, 6, 4):
        blk = np.zeros(chunk, "<f4")
        blk[:min(4, 6 - r0)] = data[r0:r0 + 4]
        raw = np.frombuffer(blk.tobytes(), np.uint8).reshape(-1, 4).T.tobytes()        # shuffle
        blobs.append(((r0, 0, 0), zlib.compress(raw)))
    # file layout: [superblock 96][root group pieces via the writer] is simpler: write a file with a placeholder
    # contiguous dataset, then patch its layout / filter messages is fragile -- instead build the pieces directly.
    pad8 = h5io._pad8
    buf = bytearray(b"\x00" * 96)
    addrs = []
    for _, z in blobs:
        addrs.append(len(buf))
        buf += pad8(z)
    tree_addr = len(buf)
    tree = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(blobs), h5io.UNDEF, h5io.UNDEF)
    for (offs, z), a in zip(blobs, addrs):
        tree += struct.pack("<II", len(z), 0) + struct.pack("<3Q", *offs) + struct.pack("<Q", a)
    tree += struct.pack("<II", 0, 0) + struct.pack("<3Q", 8, 0, 0)
    buf += pad8(tree)
    msgs = [h5io._msg(0x01, h5io._space_msg(data.shape)), h5io._msg(0x03, h5io._dtype_msg(data.dtype), flags=1),
            h5io._msg(0x0B, struct.pack("<BB6x", 1, 2) + struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<II", 4, 0)
                      + struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<II", 4, 0)),
            h5io._msg(0x08, struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", tree_addr) + struct.pack("<3I", 4, 5, 4))]
    ds_addr = len(buf)
    buf += pad8(h5io._object_header(msgs))
    # root group with one entry
    heap_data_addr = len(buf)
    buf += pad8(b"\x00" * 8 + b"d\x00")
    heap_addr = len(buf)
    buf += b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, 16, h5io.UNDEF, heap_data_addr)
    snod_addr = len(buf)
    buf += pad8(b"SNOD" + struct.pack("<BBH", 1, 0, 1) + struct.pack("<QQII", 8, ds_addr, 0, 0) + b"\x00" * 16)
    gtree_addr = len(buf)
    buf += b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, h5io.UNDEF, h5io.UNDEF) + struct.pack("<QQQ", 0, snod_addr, 8)
    root_addr = len(buf)
    buf += pad8(h5io._object_header([h5io._msg(0x11, struct.pack("<QQ", gtree_addr, heap_addr))]))
    sb = h5io.SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, h5io.UNDEF, len(buf), h5io.UNDEF)
    sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", gtree_addr, heap_addr)
    buf[:len(sb)] = sb
    p = str(tmp_path / "chunked.h5")
    with open(p, "wb") as f:
        f.write(bytes(buf))
    with h5io.File(p, "r") as f:
        np.testing.assert_array_equal(f["d"][...], data)


def test_keras_weight_file_layout(h5io, tmp_path):
    oracle = importlib.import_module("oracle.sr4d_oracle")
    params = oracle.glorot_params(1, 1, seed=2, bias_scale=0.1)
    names = [n for n, _ in oracle.param_table(1, 1)]
    p = str(tmp_path / "w.h5")
    h5io.save_keras_weights(p, params)
    back = h5io.load_keras_weights(p, names)
    for n in names:
        np.testing.assert_array_equal(back[n], params[n])
    with h5io.File(p, "r") as f:                       # Keras layout: model_weights/<layer>/<layer>/kernel:0
        assert f["model_weights/conv3d_4/conv3d_4/kernel:0"].shape == (1, 1, 1, 128, 64)
        assert f["model_weights/conv3d_1"].attrs["weight_names"].tolist() == [b"conv3d_1/kernel:0", b"conv3d_1/bias:0"]
    with pytest.raises(KeyError):
        h5io.load_keras_weights(p, names + ["conv3d_99/kernel"])


@pytest.mark.parametrize("name_offset", [0, 37])
def test_tf22_model_save_layout_loads(h5io, tmp_path, name_offset):
    """A file laid out like TF 2.2's `model.save(path)` (root attributes, every layer of model.layers in layer_names,
    empty weight_names for the weightless ones, nested <layer>/<layer>/kernel:0) is read by name; with a non-zero
    Conv3D uid offset in the writer process (conv3d_37...) the creation-order fallback maps it -- Keras' own
    topological `load_weights` loads such files too."""
    import keras_layout
    oracle = importlib.import_module("oracle.sr4d_oracle")
    params = oracle.glorot_params(1, 1, seed=5, bias_scale=0.1)
    names = [n for n, _ in oracle.param_table(1, 1)]
    p = str(tmp_path / "4DFlowNet-best.h5")
    on_disk = keras_layout.write_tf22_model_file(h5io, p, params, name_offset)
    with h5io.File(p, "r") as f:
        assert bytes(f.attrs["keras_version"]) == b"2.3.0-tf" and bytes(f.attrs["backend"]) == b"tensorflow"
        ln = [bytes(x).decode() for x in f["model_weights"].attrs["layer_names"]]
        assert "u_mag" in ln and "tf_op_layer_MirrorPad" in ln and on_disk["conv3d_4"] in ln
        assert len(f["model_weights/concatenate"].attrs["weight_names"]) == 0
        assert f[f"model_weights/{on_disk['conv3d_4']}/{on_disk['conv3d_4']}/kernel:0"].shape == (1, 1, 1, 128, 64)
    rep = {}
    back = h5io.load_keras_weights(p, names, report=rep)
    assert rep["scheme"].startswith("by name" if name_offset == 0 else "by creation order")
    for n in names:
        np.testing.assert_array_equal(back[n], params[n])
    # a file with a different number of conv layers is refused, not mis-mapped
    few = {k: v for k, v in params.items() if not k.startswith("conv3d_15")}
    keras_layout.write_tf22_model_file(h5io, p, few, 3)
    with pytest.raises(KeyError):
        h5io.load_keras_weights(p, names)


def test_rotation_table_matches_reference_golden():
    ph = importlib.import_module("4dflownet_b200.Network.PatchHandler3D")
    z = np.load(GOLD)
    u, v, w = z["rot_in"]
    for plane in (1, 2, 3):
        for k in (1, 2, 3):
            for phase in (True, False):
                got = np.stack(ph.apply_rotation(u.copy(), v.copy(), w.copy(), k, plane, phase))
                np.testing.assert_array_equal(got, z[f"rot_p{plane}_k{k}_{int(phase)}"], err_msg=f"{plane},{k},{phase}")
    np.testing.assert_array_equal(np.stack(ph.apply_rotation(u, v, w, 1, 7, True)), z["rot_in"])


def test_patchhandler_rows_bit_exact_vs_reference_golden(data_dir):
    ph = importlib.import_module("4dflownet_b200.Network.PatchHandler3D")
    z = np.load(GOLD)
    h = ph.PatchHandler3D(data_dir, synth.PATCH, synth.R, 2, 0.6)
    for i, row in enumerate(synth.ROWS):
        items = h.load_patches_from_index_file(row)
        assert len(items) == 11
        for k, a in enumerate(items):
            want = z[f"row{i}_{k}"]
            assert np.asarray(a).dtype == want.dtype and np.asarray(a).shape == want.shape, (i, k)
            np.testing.assert_array_equal(np.asarray(a), want, err_msg=f"row {i} item {k}")


def test_patchhandler_batches_and_shuffle(data_dir):
    ph = importlib.import_module("4dflownet_b200.Network.PatchHandler3D")
    rows = np.asarray(synth.ROWS)
    h = ph.PatchHandler3D(data_dir, synth.PATCH, synth.R, 4, 0.6)
    ds = h.initialize_dataset(rows, shuffle=False)
    batches = list(ds)
    assert len(ds) == 3 and [len(b[0]) for b in batches] == [4, 4, 3]          # ragged tail like tf.data.batch
    P, H = synth.PATCH, synth.PATCH * synth.R
    assert batches[0][0].shape == (4, P, P, P, 1) and batches[0][6].shape == (4, H, H, H, 1)
    assert batches[0][9].shape == (4,) and batches[0][10].shape == (4, H, H, H)
    assert set(np.unique(batches[0][10])) <= {0.0, 1.0}                          # mask binarised at the threshold
    again = list(ds)                                                            # re-iterable, same order
    np.testing.assert_array_equal(again[2][0], batches[2][0])
    a = [b[0] for b in h.initialize_dataset(rows, shuffle=True, seed=5)]
    b = [b[0] for b in h.initialize_dataset(rows, shuffle=True, seed=5)]
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert not all(np.array_equal(x, y) for x, y in zip(a, [bb[0] for bb in batches]))


def test_patchhandler_row_level_sharding_equals_shard_batch(data_dir):
    """Data parallel feeding: with shard=(rank, world) a rank loads only its rows of every global batch; the result is
    exactly parallel.shard_batch() of the batch a single process would have loaded (same seed, same order)."""
    ph = importlib.import_module("4dflownet_b200.Network.PatchHandler3D")
    parallel = importlib.import_module("4dflownet_b200.parallel")
    rows = np.asarray(synth.ROWS)                                                # 11 rows
    full = list(ph.PatchHandler3D(data_dir, synth.PATCH, synth.R, 4, 0.6).initialize_dataset(rows, shuffle=True, seed=3))
    assert [len(b[0]) for b in full] == [4, 4, 3]
    for world in (2, 3):
        for rank in range(world):
            ds = ph.PatchHandler3D(data_dir, synth.PATCH, synth.R, 4, 0.6).initialize_dataset(
                rows, shuffle=True, seed=3, shard=(rank, world))
            got = list(ds)
            assert len(got) == len(ds) == 3
            for g, f in zip(got, full):
                want = parallel.shard_batch(f, rank, world)
                assert all(np.array_equal(a, b) for a, b in zip(g, want))
    # a tail batch with fewer rows than ranks is dropped on every rank alike
    ds = ph.PatchHandler3D(data_dir, synth.PATCH, synth.R, 5, 0.6).initialize_dataset(rows, shuffle=False, shard=(1, 2))
    assert len(ds) == 2 and [len(b[0]) for b in ds] == [2, 2]                    # 11 = 5 + 5 + 1: the 1-row tail goes


def test_patchify_device_equals_patchify():
    """The torch (device) tiling used by predict_volume against the numpy tiling that is pinned to the reference's
    golden vectors: bit-identical stacks, padding and patch counts, whole list and per-rank ranges."""
    import torch
    pgm = importlib.import_module("4dflownet_b200.Network.PatchGenerator")

    class DS:
        pass
    g = np.random.default_rng(4)
    for shape, P, r in [((42, 38, 36), 24, 2), ((10, 9, 11), 8, 2), ((21, 22, 23), 12, 3), ((7, 13, 5), 8, 1)]:
        ds = DS()
        for n in ("u", "v", "w", "mag_u", "mag_v", "mag_w"):
            setattr(ds, n, g.standard_normal(shape).astype(np.float32))
        a, b = pgm.PatchGenerator(P, r), pgm.PatchGenerator(P, r)
        vel, mag = a.patchify(ds)
        dev = b.patchify_device(ds, "cpu")
        assert (a.nr_x, a.nr_y, a.nr_z, a.padding) == (b.nr_x, b.nr_y, b.nr_z, b.padding)
        for want, got in zip((*vel, *mag), dev):
            assert got.dtype == torch.float32 and got.is_contiguous()
            assert np.array_equal(got.numpy(), want[..., 0])
        n = a.count_patches(shape)
        lo, hi = n // 3, n - 1
        vel, mag = a.patchify(ds, lo, hi)
        for want, got in zip((*vel, *mag), b.patchify_device(ds, "cpu", lo, hi)):
            assert np.array_equal(got.numpy(), want[..., 0])


def test_image_dataset_and_result_writer(h5io, data_dir, tmp_path):
    ids = importlib.import_module("4dflownet_b200.utils.ImageDataset")
    pu = importlib.import_module("4dflownet_b200.utils.prediction_utils")
    lr = os.path.join(data_dir, "synth_LR.h5")
    ds = ids.ImageDataset()
    assert ds.get_dataset_len(lr) == 2
    ds.load_vectorfield(lr, 1)
    with h5io.File(lr, "r") as f:
        venc = max(float(f[k][1]) for k in ("venc_u", "venc_v", "venc_w"))
        np.testing.assert_array_equal(ds.u, (f["u"][1] / np.float32(venc)).astype(np.float32))
        np.testing.assert_array_equal(ds.mag_w, (f["mag_w"][1] / 4095.).astype(np.float32))
    assert ds.venc == np.float32(2.0) and ds.velocity_per_px == np.float32(2.0) / 2048 and ds.dx.shape == (3,)
    out = str(tmp_path / "res.h5")
    vol = np.random.default_rng(0).standard_normal((1, 4, 5, 6))
    pu.save_to_h5(out, "u", vol, compression="gzip")
    pu.save_to_h5(out, "u", vol * 2, compression="gzip")
    with h5io.open_file(out, "r") as f:
        assert f["u"].shape == (2, 4, 5, 6) and f["u"].dtype == np.float32
        np.testing.assert_array_equal(f["u"][1], (vol * 2).astype(np.float32)[0])
