"""pytest configuration: `gpu` marker (tests that need a B200) and shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def iniconds():
    import astrophysical_sph_b200.iniconds as ic

    return ic


def make_case(eos, ic_type, N, **kw):
    """(pos, vel, K, constants, sph-kwargs) of an initial condition built by the package's generators."""
    import astrophysical_sph_b200.iniconds as ic

    d = ic.make_ic(eos, ic_type, N=N, **kw)
    c = d["constants"]
    args = dict(m=c["m"], G=c["G"], theta=c["theta"], alpha=c["alpha"], beta=c["beta"])
    if eos == "isothermal":
        args.update(cs=c["cs"], U_iso=c["U"])
    else:
        args.update(gamma=c["gamma"])
    return d["pos"], d["vel"], d["K"], c, args


def oracle_kwargs(O, eos, c, K):
    return dict(eos=O.ISOTHERMAL if eos == "isothermal" else O.POLYTROPIC, cs=c.get("cs", 0.0), Kent=K,
                gamma=c.get("gamma", 5 / 3), alpha=c["alpha"], beta=c["beta"])


def vec_rel(a, b, floor=0.0):
    """max_i |a_i - b_i| / max(|b_i|, floor) over rows of N x 3 arrays."""
    return float((np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), max(floor, 1e-300))).max())


GOLDEN = os.path.join(ROOT, "tests", "golden")
