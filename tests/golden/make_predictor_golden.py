"""Golden output of the reference's own ``src/predictor.py`` (BASELINE configs[0]: patch_size=24, res_increase=2,
8 low / 4 hi resblocks, a 42x38x36 volume -> 12 patches -> 84x76x72 x 3).

Run in the build container only (needs /root/reference; ~10 minutes of float64 numpy):
    python tests/golden/make_predictor_golden.py
The reference script is executed UNMODIFIED with ``runpy`` as ``__main__`` from a scratch directory laid out like the
reference tree (``../data/example_data.h5``, ``../models/4DFlowNet/4DFlowNet.h5``, ``../result``).  Its imports resolve
to the reference's own ``Network/*`` and ``utils/*`` modules; ``tensorflow`` is the float64 numpy stand-in of
``tf_numpy_shim.py`` (symbolic ``Input`` / ``Model`` replaying the recorded graph), ``h5py`` is the repo's pure-Python
HDF5 shim.  Inputs that do not ship to the GPU box are synthetic and regenerable: the LR file
(``synth.make_example_lr``: same layout, shape and value ranges as data/example_data.h5) and the Keras-layout weight
file (``synth.keras_weight_dict``).  Output: tests/golden/predictor_golden.npz -- every second voxel of the stitched
u, v, w volumes plus float64 moments of the full volumes.
"""
import importlib
import os
import runpy
import shutil
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import synth  # noqa: E402
import tf_numpy_shim  # noqa: E402

REF = "/root/reference/src"
WEIGHT_SEED, DATA_SEED = 2024, 7


def main():
    h5io = importlib.import_module("4dflownet_b200.utils.h5io")
    work = tempfile.mkdtemp(prefix="sr4d_predgold_")
    for d in ("src", "data", "models/4DFlowNet"):
        os.makedirs(os.path.join(work, d))
    synth.make_example_lr(os.path.join(work, "data", "example_data.h5"), DATA_SEED)
    h5io.save_keras_weights(os.path.join(work, "models", "4DFlowNet", "4DFlowNet.h5"),
                            synth.keras_weight_dict(8, 4, WEIGHT_SEED))
    assert h5io.install_as_h5py()
    tf_numpy_shim.install()
    tf_numpy_shim.set_weight_source(None)
    sys.path.insert(0, REF)
    os.chdir(os.path.join(work, "src"))
    t0 = time.time()
    runpy.run_path(os.path.join(REF, "predictor.py"), run_name="__main__")     # reference script, unmodified
    print(f"\nreference predictor.py finished in {time.time() - t0:.0f} s")
    out = {"weight_seed": WEIGHT_SEED, "data_seed": DATA_SEED}
    with h5io.File(os.path.join(work, "result", "example_result.h5"), "r") as f:
        for c in "uvw":
            v = np.asarray(f[c][...])
            assert v.shape == (1, 84, 76, 72) and v.dtype == np.float32, (v.shape, v.dtype)
            v64 = v.astype(np.float64)
            out[c] = v[0, ::2, ::2, ::2]
            out[c + "_moments"] = np.array([v64.sum(), np.abs(v64).sum(), (v64 ** 2).sum(), float((v == 0).sum()),
                                            np.abs(v64).max()])
            print(c, out[c + "_moments"])
        out["dx"] = np.asarray(f["dx"][...])
    np.savez_compressed(os.path.join(HERE, "predictor_golden.npz"), **out)
    print("wrote predictor_golden.npz", os.path.getsize(os.path.join(HERE, "predictor_golden.npz")))
    os.chdir(ROOT)
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
