"""Generates tests/golden/*.npz.

PARITY UNPINNED: the reference (pure Julia, no tests, no fixtures) cannot be executed in
the build container, so these vectors are outputs of the CPU oracle (oracle/sph_oracle.cpp), cross-checked
against the independent numpy/scipy twin (oracle/sph_numpy.py) by tests/test_oracle.py.  They pin the oracle
against silent regressions and give the GPU tests a committed, machine-independent target.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import astrophysical_sph_b200.iniconds as ic  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def one(eos, N, name, Kh=50, steps=2):
    d = ic.make_ic(eos, "gaussian_sphere", N=N, R=ic.R0, Kh=Kh, seed=7)
    c = d["constants"]
    pos, vel, K = d["pos"], d["vel"], d["K"]
    # give the particles some motion so that the artificial viscosity and dK/dt are exercised
    rng = np.random.default_rng(11)
    vel = np.asfortranarray(vel + 0.3 * c.get("cs", 4.5e7) * rng.standard_normal(vel.shape) - 2e-10 * pos)
    kw = dict(eos=O.ISOTHERMAL if eos == "isothermal" else O.POLYTROPIC, cs=c.get("cs", 0.0), Kent=K,
              gamma=c.get("gamma", 5 / 3), alpha=c["alpha"], beta=c["beta"])
    hy = O.hydro(pos, vel, c["m"], Kh, **kw)
    g, phi, st = O.gravity(np.abs(pos).max(), c["m"], pos, c["theta"], hy["h"])
    stp = O.step(pos, vel, c["m"], Kh, c["G"], c["theta"], 0.0, steps, U_iso=c.get("U", 0.0), **kw)
    np.savez_compressed(
        os.path.join(HERE, name), pos=pos, vel=vel, K=np.zeros(0) if K is None else K,
        consts=np.array([c["m"], c.get("cs", 0.0), c.get("gamma", 5 / 3), c["G"], c["theta"], c["alpha"], c["beta"],
                         c.get("U", 0.0), Kh]),
        idx=hy["idx"], rK=hy["r"][:, -1].copy(), rho=hy["rho"], h=hy["h"], ahyd=hy["ahyd"], sum_vdw=hy["sum_vdw"],
        mumax=hy["mumax"], cs_i=hy["cs_i"], dkdt=hy["dkdt"], g=g, phi=phi, tree_stats=st, dts=stp["dts"],
        stats=stp["stats"], pos_end=stp["pos"], vel_end=stp["vel"], K_end=np.zeros(0) if stp["K"] is None else stp["K"],
        t_end=stp["t"])
    print(name, "written")


if __name__ == "__main__":
    one("isothermal", 1024, "gauss_iso_1024.npz")
    one("polytropic", 1024, "gauss_poly_1024.npz")
