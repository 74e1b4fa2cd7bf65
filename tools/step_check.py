"""Stepping run with sanity checks after every step: python tools/step_check.py [N] [steps] [--oracle]"""
import os
import sys

import numpy as np

os.environ.setdefault("SPH_B200_GRAPH_N", "0")   # plain launches: the phase timers printed below are not recorded in a graph replay

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import astrophysical_sph_b200.iniconds as ic  # noqa: E402
from astrophysical_sph_b200.libsph import SphB200  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d = ic.make_ic("isothermal", "boss_bodenheimer", N=N, T=10)
c = d["constants"]
s = SphB200(N, c["Kh"], "isothermal", m=c["m"], cs=c["cs"], G=c["G"], theta=c["theta"], alpha=c["alpha"], beta=c["beta"],
            U_iso=c["U"])
s.upload(d["pos"], d["vel"], None, 0.0)
opos, ovel = d["pos"], d["vel"]
for k in range(steps):
    info = s.step(1)
    p, v, _, t = s.download()
    hy = s.hydro()
    print("step", k, "dt", info["dts"], "knn_ms", round(s.timings()["knn_ms"], 3), "retries", s.timings()["knn_retries"],
          "| nan pos", int(np.isnan(p).sum()), "nan vel", int(np.isnan(v).sum()), "h min/max", hy["h"].min(), hy["h"].max(),
          "rho min", hy["rho"].min(), "nan ahyd", int(np.isnan(hy["ahyd"]).sum()), flush=True)
    if "--oracle" in sys.argv:
        from oracle import oracle as O
        oo = O.step(opos, ovel, c["m"], c["Kh"], c["G"], c["theta"], 0.0, 1, cs=c["cs"], alpha=c["alpha"], beta=c["beta"],
                    U_iso=c["U"], nthreads=O.max_threads())
        opos, ovel = oo["pos"], oo["vel"]
        dp = np.abs(p - opos).max() / np.abs(opos).max()
        dv = np.linalg.norm(v - ovel, axis=1) / np.maximum(np.linalg.norm(ovel, axis=1), 1e-3 * np.abs(ovel).max())
        print("   vs oracle: dt", info["dts"][0], oo["dts"][0], "pos maxdiff/scale", dp, "vel rel max", dv.max(),
              "n(vel rel > 1e-6)", int((dv > 1e-6).sum()), flush=True)
s.close()
print("ok")
