import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (its directory name starts with a digit, so importlib)."""
    return importlib.import_module("4dflownet_b200")


@pytest.fixture(scope="session")
def oracle():
    return importlib.import_module("oracle.sr4d_oracle")


# ---- gradient-parity bars ----------------------------------------------------------------------------------------
# tests/gradient_bars.json holds, per named check, the MEASURED relative error on a B200 (written by running the GPU
# suite with SR4D_RECORD_BARS=<file>, see tools/gpu_r02_*.sh) -- the tests assert err <= BAR_FACTOR * measured, and
# never more than the check's absolute ceiling.  A check without a recorded value is held to its ceiling only.
# tools/make_gradient_bars.py turns the recorded file into tests/gradient_bars.json.
BAR_FACTOR = 2.0
_BARS_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gradient_bars.json")


def _load_bars():
    import json
    if os.path.exists(_BARS_PATH):
        with open(_BARS_PATH) as f:
            return json.load(f)
    return {}


@pytest.fixture(scope="session")
def bars():
    """check(name, err, ceiling): assert err against min(ceiling, BAR_FACTOR * measured) and record it when asked."""
    import json
    table = _load_bars()
    rec_path = os.environ.get("SR4D_RECORD_BARS")

    def check(name, err, ceiling, floor=1e-6):
        """floor: the bar never drops below it (measured values of ~1e-7 are fp32 summation-order noise)."""
        err = float(err)
        if rec_path:
            with open(rec_path, "a") as f:
                f.write(json.dumps({"name": name, "err": err, "ceiling": ceiling}) + "\n")
        bar = ceiling
        if name in table:
            bar = min(ceiling, max(floor, BAR_FACTOR * float(table[name])))
        assert err <= bar, f"{name}: {err:.3e} > bar {bar:.3e} (measured {table.get(name)}, ceiling {ceiling:.1e})"
        return err
    return check
