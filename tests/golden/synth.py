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


# ---- floating-point graph golden (make_graph_golden.py and the tests that replay it) ----
def graph_weight_source(seed):
    """fn(kernel_shape, use_bias) -> (kernel, bias|None), drawn sequentially in layer creation order: glorot-uniform
    kernels (Keras default) and small normal biases (so the bias path matters)."""
    g = np.random.default_rng(seed)

    def draw(shape, use_bias):
        k3 = shape[0] * shape[1] * shape[2]
        lim = np.sqrt(6.0 / (k3 * shape[3] + k3 * shape[4]))
        kernel = g.uniform(-lim, lim, size=shape).astype(np.float32)
        bias = (0.05 * g.standard_normal(shape[4])).astype(np.float32) if use_bias else None
        return kernel, bias
    return draw


def graph_inputs(B, P, r, seed):
    """6 LR inputs (B,P,P,P,1), 3 HR targets (B,rP,rP,rP,1) with exact zeros, binary mask (B,rP,rP,rP); float32
    values held in float64 arrays."""
    g = np.random.default_rng(seed + 1000)
    lr = [g.uniform(-1, 1, (B, P, P, P, 1)).astype(np.float32).astype(np.float64) for _ in range(3)] + \
         [g.uniform(0, 0.016, (B, P, P, P, 1)).astype(np.float32).astype(np.float64) for _ in range(3)]
    H = P * r
    keep = g.uniform(size=(B, H, H, H, 1)) < 0.7
    hr = [(0.3 * g.standard_normal((B, H, H, H, 1)) * keep).astype(np.float32).astype(np.float64) for _ in range(3)]
    mask = (g.uniform(size=(B, H, H, H)) < 0.3).astype(np.float64)
    return lr, hr, mask


# ---- predictor.py golden (make_predictor_golden.py and tests/test_gpu_integration.py) ----
EXAMPLE_LR_SHAPE = (42, 38, 36)      # the shape of the reference's data/example_data.h5


def make_example_lr(path, seed=7):
    """A one-row LR file with the column layout, shape and value ranges of the reference's data/example_data.h5
    (u,v,w in [-1.5,1.5], magnitudes in [0,65], venc 1.5, dx 1.1875), written with the repo's HDF5 shim."""
    h5io = importlib.import_module("4dflownet_b200.utils.h5io")
    g = np.random.default_rng(seed)
    if os.path.exists(path):
        os.remove(path)
    with h5io.File(path, "w") as f:
        for c in "uvw":
            f.create_dataset(c, data=g.uniform(-1.5, 1.5, (1,) + EXAMPLE_LR_SHAPE).astype(np.float32))
            f.create_dataset("mag_" + c, data=g.uniform(0, 65, (1,) + EXAMPLE_LR_SHAPE).astype(np.float32))
            f.create_dataset("venc_" + c, data=np.asarray([1.5], np.float32))
        f.create_dataset("dx", data=np.full((1, 3), 1.1875, np.float32))


def keras_weight_dict(low, hi, seed):
    """{'conv3d_k/kernel'|'bias': array} for the 8/4 (or any) network in Keras creation order, drawn from
    graph_weight_source; layer shapes follow Network/SR4DFlowNet.py:17-46."""
    draw = graph_weight_source(seed)
    C = 64
    spec = [(3, 3, C, True), (3, C, C, True), (3, 3, C, True), (3, C, C, True), (1, 2 * C, C, True), (3, C, C, True)]
    spec += [(3, C, C, False)] * (2 * low + 2 * hi)
    spec += [(3, C, C, True), (3, C, 1, True)] * 3
    out = {}
    for i, (k, ci, co, bias) in enumerate(spec):
        name = "conv3d" if i == 0 else f"conv3d_{i}"
        kern, b = draw((k, k, k, ci, co), bias)
        out[f"{name}/kernel"] = kern
        if bias:
            out[f"{name}/bias"] = b
    return out
