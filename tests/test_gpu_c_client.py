"""The C ABI used from plain C (examples/c_abi_forward.c: dlopen + cudart, no Python / torch in the client) gives the
same prediction as the Python mirror for the same weights and inputs, and reports errors through return codes."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def pattern(i, scale):
    i = np.asarray(i, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = i * np.uint64(6364136223846793005) + np.uint64(1442695040888963407)
    x = x ^ (x >> np.uint64(33))
    return (np.float32(scale) * ((x % np.uint64(20001)).astype(np.float32) / np.float32(10000.0) - np.float32(1.0))).astype(np.float32)


def test_c_client_matches_python_mirror(pkg):
    exe = os.path.join(ROOT, "examples", "c_abi_forward")
    src = exe + ".c"
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", src, "-o", exe,
                               "-L", "/usr/local/cuda/lib64", "-lcudart", "-ldl", "-lm"])
    env = dict(os.environ, LD_LIBRARY_PATH="/usr/local/cuda/lib64:" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe, os.path.join(ROOT, "4dflownet_b200", "libsr4d.so")], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"tensors (\d+) flat (\d+) out (\d+) sum (\S+) abssum (\S+) first (\S+) last (\S+)", r.stdout)
    assert m, r.stdout
    assert "oversized batch -> rc -1" in r.stdout
    nt, flat, nout = int(m.group(1)), int(m.group(2)), int(m.group(3))
    c_sum, c_abs, c_first, c_last = (float(m.group(k)) for k in (4, 5, 6, 7))

    P, R, LOW, HI, B = 8, 2, 1, 1, 2
    eng = pkg.Engine(P, R, LOW, HI, max_batch=B, training=False, device=0)
    assert len(eng.table) == nt and eng.flat_size == flat
    ws = []
    for t, (name, off, cnt, shape, is_kernel) in enumerate(eng.table):
        scale = 0.02
        if is_kernel:
            k3 = shape[0] * shape[1] * shape[2]
            scale = np.sqrt(np.float32(6.0) / np.float32(k3 * shape[3] + k3 * shape[4])).astype(np.float32)
        ws.append(pattern(np.uint64(1000003 * t) + np.arange(cnt, dtype=np.uint64), scale).reshape(shape))
    eng.set_weights(ws)
    nin = B * P ** 3
    idx = np.arange(nin, dtype=np.uint64)
    xs = []
    for c in range(6):
        if c < 3:
            xs.append(pattern(np.uint64(7 + 31 * c) + np.uint64(6) * idx, 1.0))
        else:
            xs.append((np.float32(0.008) * (np.float32(1.0) + pattern(np.uint64(11 + 17 * c) + np.uint64(6) * idx, 1.0))).astype(np.float32))
    y = eng.forward([x.reshape(B, P, P, P) for x in xs]).cpu().numpy().ravel()
    assert y.size == nout
    scale = np.abs(y).max()
    assert abs(float(y[0]) - c_first) < 1e-5 * scale and abs(float(y[-1]) - c_last) < 1e-5 * scale
    assert abs(float(np.abs(y.astype(np.float64)).sum()) - c_abs) < 1e-5 * c_abs
    assert abs(float(y.astype(np.float64).sum()) - c_sum) < 1e-5 * c_abs
    eng.close()
