"""Generate golden vectors for the INTEGER tiling logic from the reference's own code.

Run in the build container only (needs /root/reference):
    python tests/golden/make_patchgen_golden.py
It imports /root/reference/src/Network/PatchGenerator.py unmodified (with empty stub
modules standing in for ``h5py`` and ``utils.ImageDataset`` which PatchGenerator only
imports, never calls) and records, for several (volume, patch_size, res_increase):
nr_x/y/z, HR padding, sha256 digests of the patch stack and of the stitched output
and, for the tiny cases, the full arrays.  Output: tests/golden/patchgen_golden.npz
(+ the rotation golden from PatchHandler3D helpers).
"""
import hashlib
import os
import sys
import types

import numpy as np

REF = "/root/reference/src"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def volume(shape, seed):
    # deterministic, cheap to regenerate in the test: small integers in float32
    g = np.random.default_rng(seed)
    return g.integers(-1000, 1000, size=shape).astype(np.float32)


CASES = [
    # (shape, patch_size, res_increase, store_full)
    ((42, 38, 36), 24, 2, False),
    ((42, 38, 36), 24, 4, False),
    ((42, 38, 36), 16, 2, False),
    ((42, 38, 36), 12, 2, False),
    ((160, 160, 64), 24, 2, False),
    ((10, 9, 11), 8, 2, True),
    ((7, 13, 5), 8, 1, True),
    ((20, 20, 20), 24, 2, False),
    ((21, 22, 23), 12, 3, False),
]


def main():
    _stub("h5py")
    _stub("tensorflow")
    utils = _stub("utils")
    utils.ImageDataset = _stub("utils.ImageDataset", ImageDataset=object)
    sys.path.insert(0, REF)
    from Network.PatchGenerator import PatchGenerator  # reference code, unmodified

    out = {}
    for ci, (shape, P, r, full) in enumerate(CASES):
        vol = volume(shape, ci)
        pg = PatchGenerator(P, r)
        patches, nx, ny, nz = pg._generate_overlapping_patches(vol)
        # HR "prediction" stand-in: nearest-neighbour repeat of each LR patch
        hr = patches.repeat(r, axis=1).repeat(r, axis=2).repeat(r, axis=3)
        stitched = pg._patchup_with_overlap(hr, nx, ny, nz)
        key = f"case{ci}"
        out[key + "_meta"] = np.array([*shape, P, r, nx, ny, nz, *pg.padding, *stitched.shape], dtype=np.int64)
        out[key + "_patch_sha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(patches).tobytes()).digest(), dtype=np.uint8)
        out[key + "_stitch_sha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(stitched).tobytes()).digest(), dtype=np.uint8)
        if full:
            out[key + "_patches"] = patches
            out[key + "_stitched"] = stitched
        print(key, shape, P, r, (nx, ny, nz), pg.padding, stitched.shape)

    # rotation helpers of the training iterator (PatchHandler3D.py:166-274), tiny case
    from Network import PatchHandler3D as ph
    g = np.random.default_rng(99)
    u, v, w = (g.standard_normal((4, 4, 4)).astype(np.float32) for _ in range(3))
    out["rot_in"] = np.stack([u, v, w])
    for plane in (1, 2, 3):
        for k in (1, 2, 3):
            for phase in (True, False):
                if k == 2:
                    r3 = ph.rotate180_3d(u, v, w, plane, phase)
                else:
                    r3 = ph.rotate90(u, v, w, plane, k, phase)
                out[f"rot_p{plane}_k{k}_{int(phase)}"] = np.stack(r3)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "patchgen_golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
