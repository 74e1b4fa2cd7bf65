"""Prints the error levels of one evaluation against the oracle (GPU box): python tools/err_levels.py [N]
Used to see how far inside the tolerances (1e-9 hydro, 1e-6 gravity) the fast reciprocal / rsqrt sequences keep the results."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import astrophysical_sph_b200.iniconds as ic  # noqa: E402
from astrophysical_sph_b200.libsph import SphB200  # noqa: E402
from oracle import oracle as O  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
for eos in ("isothermal", "polytropic"):
    d = ic.make_ic(eos, "gaussian_sphere", N=N, R=ic.R0, seed=11)
    c = d["constants"]
    rng = np.random.default_rng(5)
    vel = np.asfortranarray(d["vel"] + 2e7 * rng.standard_normal(d["vel"].shape))
    kw = dict(m=c["m"], G=c["G"], theta=c["theta"], alpha=c["alpha"], beta=c["beta"])
    okw = dict(alpha=c["alpha"], beta=c["beta"])
    if eos == "isothermal":
        kw.update(cs=c["cs"], U_iso=c["U"]); okw.update(eos=O.ISOTHERMAL, cs=c["cs"])
    else:
        kw.update(gamma=c["gamma"]); okw.update(eos=O.POLYTROPIC, Kent=d["K"], gamma=c["gamma"])
    with SphB200(N, c["Kh"], eos, **kw) as s:
        s.eval_acc(d["pos"], vel, d["K"])
        hy = s.hydro()
        g, phi = s.grav()
    nt = O.max_threads()
    oh = O.hydro(d["pos"], vel, c["m"], c["Kh"], nthreads=nt, **okw)
    og, ophi, _ = O.gravity(np.abs(d["pos"]).max(), c["m"], d["pos"], c["theta"], oh["h"], nthreads=nt)

    def vrel(a, b):
        return float((np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)).max())

    print(eos, "rho %.2e ahyd %.2e sum_vdw %.2e dkdt %.2e g %.2e phi %.2e" % (
        np.abs(hy["rho"] / oh["rho"] - 1).max(), vrel(hy["ahyd"], oh["ahyd"]),
        np.abs(hy["sum_vdw"] - oh["sum_vdw"]).max() / np.abs(oh["sum_vdw"]).max(),
        np.abs(hy["dkdt"] - oh["dkdt"]).max() / max(np.abs(oh["dkdt"]).max(), 1e-300), vrel(g, og), np.abs(phi / ophi - 1).max()), flush=True)
