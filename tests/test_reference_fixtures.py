"""The reference's own data fixtures (data/example_data.h5, data/example_data_HR.h5, written by real h5py) through the
repo's pure-Python HDF5 reader and data-path mirrors.  Runs only where the reference tree is present (the build
container); the GPU box has no /root/reference."""
import importlib
import os
import sys
import types

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "data")), reason="reference tree not present")


@pytest.fixture(scope="module")
def h5io():
    return importlib.import_module("4dflownet_b200.utils.h5io")


def test_shim_reads_the_reference_lr_and_hr_files(h5io):
    """Shapes and value ranges recorded in SURVEY 8c for the two example files."""
    with h5io.File(os.path.join(REF, "data", "example_data.h5"), "r") as f:
        assert {"u", "v", "w", "mag_u", "mag_v", "mag_w", "venc_u", "venc_v", "venc_w", "dx"} <= set(f.keys())
        u = np.asarray(f["u"][0])
        assert u.shape == (42, 38, 36) and u.dtype == np.float32
        for c in "uvw":
            a = np.asarray(f[c][0])
            assert -1.5 <= a.min() < -1.4 and 1.4 < a.max() <= 1.5          # noise-filled to +-venc
            m = np.asarray(f["mag_" + c][0])
            assert 0 < m.min() < 0.1 and 60 < m.max() < 70
            assert abs(float(np.asarray(f["venc_" + c][0])) - 1.5) < 1e-6
        np.testing.assert_allclose(np.asarray(f["dx"][0]), 1.1875)
    with h5io.File(os.path.join(REF, "data", "example_data_HR.h5"), "r") as f:
        hu = np.asarray(f["u"][0])
        mask = np.asarray(f["mask"][0])
        assert hu.shape == (84, 76, 72) and mask.shape == (84, 76, 72)
        assert np.abs(hu).max() <= 0.66 and 0.10 < mask.mean() < 0.14
        assert np.all(hu[mask == 0] == 0)                                     # zero outside the fluid region


def test_product_data_path_equals_reference_modules_on_the_real_file(h5io):
    """The reference's ImageDataset + PatchGenerator (their own code, h5py replaced by the shim) and the product mirrors
    on data/example_data.h5: identical normalised volumes, identical 12 x 24^3 patch stacks and stitched shape."""
    saved = {k: sys.modules.get(k) for k in ("h5py", "utils", "utils.ImageDataset", "Network", "Network.PatchGenerator")}
    shim = types.ModuleType("h5py")
    shim.File, shim.Group, shim.Dataset, shim.__shim__ = h5io.File, h5io.Group, h5io.Dataset, True
    sys.modules["h5py"] = shim
    for k in ("utils", "utils.ImageDataset", "Network", "Network.PatchGenerator"):
        sys.modules.pop(k, None)
    sys.path.insert(0, os.path.join(REF, "src"))
    try:
        ref_ds_mod = importlib.import_module("utils.ImageDataset")             # reference code, unmodified
        ref_pg_mod = importlib.import_module("Network.PatchGenerator")         # reference code, unmodified
        path = os.path.join(REF, "data", "example_data.h5")
        rds = ref_ds_mod.ImageDataset()
        assert rds.get_dataset_len(path) == 1
        rds.load_vectorfield(path, 0)
        rpg = ref_pg_mod.PatchGenerator(24, 2)
        rvel, rmag = rpg.patchify(rds)
    finally:
        sys.path.remove(os.path.join(REF, "src"))
        for k in ("utils", "utils.ImageDataset", "Network", "Network.PatchGenerator"):
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)
    pds = importlib.import_module("4dflownet_b200.utils.ImageDataset").ImageDataset()
    assert pds.get_dataset_len(path) == 1
    pds.load_vectorfield(path, 0)
    for n in ("u", "v", "w", "mag_u", "mag_v", "mag_w"):
        a, b = getattr(rds, n), getattr(pds, n)
        assert a.dtype == b.dtype == np.float32 and np.array_equal(a, b), n
    assert np.float32(rds.venc) == pds.venc and np.float32(rds.velocity_per_px) == pds.velocity_per_px
    np.testing.assert_array_equal(rds.dx, pds.dx)
    ppg = importlib.import_module("4dflownet_b200.Network.PatchGenerator").PatchGenerator(24, 2)
    pvel, pmag = ppg.patchify(pds)
    assert (rpg.nr_x, rpg.nr_y, rpg.nr_z) == (ppg.nr_x, ppg.nr_y, ppg.nr_z) == (3, 2, 2)
    assert tuple(rpg.padding) == tuple(ppg.padding)
    for a, b in zip((*rvel, *rmag), (*pvel, *pmag)):
        assert a.shape == (12, 24, 24, 24, 1) and np.array_equal(a, b)
    dev = ppg.patchify_device(pds, "cpu")
    for a, b in zip((*rvel, *rmag), dev):
        assert np.array_equal(a[..., 0], b.numpy())


def test_product_patch_loader_equals_reference_loader_on_the_real_training_rows(h5io):
    """Every row of the reference's data/train.csv, validate.csv and benchmark.csv (real LR / HR files, rotations
    included) through the reference's own PatchHandler3D.load_patches_from_index_file (tf stubbed: only tf.newaxis is
    touched) and through the product loader: the 11 arrays of a sample are bit-identical."""

    class FakeTensor:                       # what tf.py_function hands to the reference loader
        def __init__(self, s):
            self.s = s

        def numpy(self):
            return self.s.encode()

        def __int__(self):
            return int(self.s)

        def __float__(self):
            return float(self.s)

    saved = {k: sys.modules.get(k) for k in ("h5py", "tensorflow", "Network", "Network.PatchHandler3D")}
    shim = types.ModuleType("h5py")
    shim.File, shim.Group, shim.Dataset, shim.__shim__ = h5io.File, h5io.Group, h5io.Dataset, True
    tf = types.ModuleType("tensorflow")
    tf.newaxis = None
    sys.modules["h5py"], sys.modules["tensorflow"] = shim, tf
    for k in ("Network", "Network.PatchHandler3D"):
        sys.modules.pop(k, None)
    sys.path.insert(0, os.path.join(REF, "src"))
    data_dir = os.path.join(REF, "data")
    rows = []
    for name in ("train.csv", "validate.csv", "benchmark.csv"):
        rows += [ln.strip().split(",") for ln in open(os.path.join(data_dir, name)).read().splitlines()[1:] if ln.strip()]
    assert len(rows) == 70
    try:
        ref = importlib.import_module("Network.PatchHandler3D")                 # reference code, unmodified
        rh = ref.PatchHandler3D(data_dir, 16, 2, 4, 0.6)                       # trainer.py defaults: patch 16, r 2
        want = [[np.asarray(a) for a in rh.load_patches_from_index_file([FakeTensor(x) for x in row])] for row in rows]
    finally:
        sys.path.remove(os.path.join(REF, "src"))
        for k in ("Network", "Network.PatchHandler3D"):
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)
    ph = importlib.import_module("4dflownet_b200.Network.PatchHandler3D").PatchHandler3D(data_dir, 16, 2, 4, 0.6)
    rotated = 0
    for row, w in zip(rows, want):
        got = ph.load_patches_from_index_file(np.asarray(row))
        assert len(got) == len(w) == 11
        for k, (a, b) in enumerate(zip(got, w)):
            a = np.asarray(a)
            assert a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b), (row, k)
        rotated += int(row[6])
    assert rotated > 10                     # the CSVs do exercise the rotation augmentation
