"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/sr4d.h declares; host-side tiling logic is bit-exact with the reference's golden
vectors.  No compute calls (no GPU here)."""
import hashlib
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    return g.build()


def test_library_exports_every_header_symbol(built_lib, pkg):
    header = open(os.path.join(ROOT, "include", "sr4d.h")).read()
    declared = set(re.findall(r"\b(sr4d_[a-z0-9_]+)\s*\(", header))
    declared -= {"sr4d_t"}
    assert len(declared) >= 25
    lib = pkg._lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in sr4d.h but not exported"
    # and the Python binding table covers exactly the header
    assert declared == set(pkg._lib.SYMBOLS), declared ^ set(pkg._lib.SYMBOLS)
    assert lib.sr4d_version().decode().startswith("sr4d")


def test_option_constants_match_the_header(pkg):
    """The Python binding's option / implementation constants are the header's #defines (SR4D_OPT_*, SR4D_CONV_*)."""
    header = open(os.path.join(ROOT, "include", "sr4d.h")).read()
    defines = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+SR4D_((?:OPT|CONV)_[A-Z0-9_]+)\s+(\d+)", header)}
    assert len([k for k in defines if k.startswith("OPT_")]) >= 8
    for name, value in defines.items():
        assert getattr(pkg._lib, name) == value, name


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.Sr4dError):
        pkg.Engine(24, 2)


def test_product_never_imports_oracle():
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "4dflownet_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                txt = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle", txt, re.M) or "sr4d_oracle" in txt:
                    bad.append(f)
    assert not bad, bad


def test_patchgenerator_bit_exact_vs_reference_golden(pkg):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_patchgen_golden import CASES, volume
    z = np.load(os.path.join(ROOT, "tests", "golden", "patchgen_golden.npz"))
    for ci, (shape, P, r, full) in enumerate(CASES):
        pg = pkg.PatchGenerator(P, r)
        vol = volume(shape, ci)
        patches, nx, ny, nz = pg._generate_overlapping_patches(vol)
        meta = z[f"case{ci}_meta"]
        assert (nx, ny, nz) == tuple(meta[5:8]) and tuple(pg.padding) == tuple(meta[8:11])
        assert hashlib.sha256(patches.tobytes()).digest() == z[f"case{ci}_patch_sha"].tobytes()
        hr = patches.repeat(r, 1).repeat(r, 2).repeat(r, 3)
        st = pg._patchup_with_overlap(hr, nx, ny, nz)
        assert st.shape == tuple(meta[11:14])
        assert hashlib.sha256(np.ascontiguousarray(st).tobytes()).digest() == z[f"case{ci}_stitch_sha"].tobytes()
        if full:
            assert np.array_equal(patches, z[f"case{ci}_patches"]) and np.array_equal(st, z[f"case{ci}_stitched"])
        pg.nr_x, pg.nr_y, pg.nr_z = nx, ny, nz
        assert pg.stitched_shape() == st.shape


def test_patchify_dataset_object(pkg):
    class DS:
        pass
    g = np.random.default_rng(0)
    ds = DS()
    for n in ("u", "v", "w", "mag_u", "mag_v", "mag_w"):
        setattr(ds, n, g.standard_normal((10, 9, 11)).astype(np.float32))
    pg = pkg.PatchGenerator(8, 2)
    vel, mag = pg.patchify(ds)
    assert vel[0].shape == (27, 8, 8, 8, 1) and mag[2].shape == (27, 8, 8, 8, 1)
    assert (pg.nr_x, pg.nr_y, pg.nr_z) == (3, 3, 3)
