"""Runs one parity check of a 20 000-particle polytropic case against the oracle; used by the tests to exercise the
alternative code paths selected by environment variables (classic sort passes, level-wise COM sweep, batched walk,
serial force/walk, warp-per-target search).  Exit code 0 = parity within the usual tolerances."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import astrophysical_sph_b200.iniconds as ic  # noqa: E402
from astrophysical_sph_b200.libsph import SphB200  # noqa: E402
from oracle import oracle as O  # noqa: E402

N = 20000
d = ic.make_ic("polytropic", "gaussian_sphere", N=N, R=ic.R0, seed=11)
c = d["constants"]
rng = np.random.default_rng(5)
vel = np.asfortranarray(d["vel"] + 2e7 * rng.standard_normal(d["vel"].shape))
s = SphB200(N, c["Kh"], "polytropic", m=c["m"], gamma=c["gamma"], G=c["G"], theta=c["theta"], alpha=c["alpha"], beta=c["beta"])
s.upload(d["pos"], vel, d["K"], 0.0)
info = s.step(2)                     # second step runs the hinted search
idx, r = s.neighbors()
p, v, K, t = s.download()
s.close()
nt = O.max_threads()
kw = dict(eos=O.POLYTROPIC, Kent=d["K"], gamma=c["gamma"], alpha=c["alpha"], beta=c["beta"])
oo = O.step(d["pos"], vel, c["m"], c["Kh"], c["G"], c["theta"], 0.0, 2, nthreads=nt, **kw)
# neighbour lists of the last evaluation = second force evaluation of step 2: recompute its inputs with the oracle
errs = dict(dt=float(np.abs(info["dts"] / oo["dts"] - 1).max()), pos=float(np.abs(p - oo["pos"]).max() / np.abs(oo["pos"]).max()),
            K=float(np.abs(K / oo["K"] - 1).max()), E=float(np.abs(info["stats"][:, 4] / oo["stats"][:, 4] - 1).max()))
ok = errs["dt"] < 1e-9 and errs["pos"] < 1e-9 and errs["K"] < 1e-9 and errs["E"] < 1e-9
ok = ok and (idx[:, 0] == np.arange(1, N + 1)).all() and (np.diff(r, axis=1) >= 0).all()
print("alt_paths_check", {k: os.environ[k] for k in os.environ if k.startswith("SPH_B200_")}, errs, "OK" if ok else "FAILED", flush=True)
sys.exit(0 if ok else 1)
