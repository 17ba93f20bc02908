"""Golden vectors for the training-side patch loader, produced by the reference's own code.

Run in the build container only (needs /root/reference):
    python tests/golden/make_patchhandler_golden.py
Imports /root/reference/src/Network/PatchHandler3D.py unmodified with (a) a stub `tensorflow` module (only
`tf.newaxis` is touched by the per-row loader) and (b) this repo's pure-Python HDF5 shim registered as `h5py`,
builds the synthetic LR/HR pair of tests/golden/synth.py, and records the 11 arrays the reference's
`load_patches_from_index_file` returns for every row of synth.ROWS, plus the rotation helpers on fresh copies
(the reference flips signs IN PLACE, so each call gets its own copy).  Output: tests/golden/patchhandler_golden.npz
"""
import importlib
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"


class FakeTensor:
    def __init__(self, s):
        self.s = s

    def numpy(self):
        return self.s.encode()

    def __int__(self):
        return int(self.s)


def main():
    import synth
    h5io = importlib.import_module("4dflownet_b200.utils.h5io")
    assert h5io.install_as_h5py(), "real h5py present: regenerate with the shim for reproducibility"
    tf = types.ModuleType("tensorflow")
    tf.newaxis = None
    sys.modules["tensorflow"] = tf
    sys.path.insert(0, REF)
    from Network import PatchHandler3D as ref       # reference code, unmodified

    out = {}
    with tempfile.TemporaryDirectory() as d:
        synth.make_synthetic_h5(d)
        h = ref.PatchHandler3D(d, synth.PATCH, synth.R, 2, 0.6)
        for i, row in enumerate(synth.ROWS):
            items = h.load_patches_from_index_file([FakeTensor(x) for x in row])
            for k, a in enumerate(items):
                out[f"row{i}_{k}"] = np.asarray(a)
    g = np.random.default_rng(99)
    u, v, w = (g.standard_normal((4, 4, 4)).astype(np.float32) for _ in range(3))
    out["rot_in"] = np.stack([u, v, w])
    for plane in (1, 2, 3):
        for k in (1, 2, 3):
            for phase in (True, False):
                a, b, c = u.copy(), v.copy(), w.copy()
                r3 = ref.rotate180_3d(a, b, c, plane, phase) if k == 2 else ref.rotate90(a, b, c, plane, k, phase)
                out[f"rot_p{plane}_k{k}_{int(phase)}"] = np.stack(r3)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "patchhandler_golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
