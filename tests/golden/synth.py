"""Deterministic synthetic LR / HR 4D-flow HDF5 pair (same column layout as the reference's data/example_data*.h5),
written with the repo's pure-Python HDF5 shim.  Shared by the golden generator and the tests."""
import importlib
import os

import numpy as np

LR_SHAPE, R = (10, 9, 8), 2


def make_synthetic_h5(directory, rows=2, seed=123):
    h5io = importlib.import_module("4dflownet_b200.utils.h5io")
    g = np.random.default_rng(seed)
    hr_shape = tuple(d * R for d in LR_SHAPE)
    lr = os.path.join(directory, "synth_LR.h5")
    hr = os.path.join(directory, "synth_HR.h5")
    for p in (lr, hr):
        if os.path.exists(p):
            os.remove(p)
    with h5io.File(lr, "w") as f:
        for c in "uvw":
            f.create_dataset(c, data=g.uniform(-1.5, 1.5, (rows,) + LR_SHAPE).astype(np.float32))
            f.create_dataset("mag_" + c, data=g.uniform(0, 65, (rows,) + LR_SHAPE).astype(np.float32))
        f.create_dataset("venc_u", data=np.asarray([1.5, 1.0][:rows], np.float32))
        f.create_dataset("venc_v", data=np.asarray([1.2, 2.0][:rows], np.float32))
        f.create_dataset("venc_w", data=np.asarray([0.9, 1.0][:rows], np.float32))
        f.create_dataset("dx", data=np.full((rows, 3), 1.1875, np.float32))
    with h5io.File(hr, "w") as f:
        for c in "uvw":
            f.create_dataset(c, data=g.uniform(-0.7, 0.7, (rows,) + hr_shape).astype(np.float32))
        f.create_dataset("mask", data=g.uniform(0, 1, (1,) + hr_shape).astype(np.float32))
    return "synth_LR.h5", "synth_HR.h5"


# CSV rows (source,target,index,start_x,start_y,start_z,rotate,rotation_plane,rotation_degree_idx,coverage), P = 4
ROWS = [
    ["synth_LR.h5", "synth_HR.h5", "0", "0", "0", "0", "0", "0", "0", "0.5"],
    ["synth_LR.h5", "synth_HR.h5", "1", "6", "5", "4", "0", "0", "0", "0.5"],
] + [["synth_LR.h5", "synth_HR.h5", str(i % 2), str(1 + p), str(2 + k), str(k), "1", str(p), str(k), "0.3"]
     for i, (p, k) in enumerate((p, k) for p in (1, 2, 3) for k in (1, 2, 3))]
PATCH = 4
