"""Times the five BASELINE.json configs on one GPU (device-resident stepping) and checks the state stays finite.
usage: python tools/config_sweep.py [names...]   names: c0 c1 c2 c3 c4"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import astrophysical_sph_b200.iniconds as ic  # noqa: E402
from astrophysical_sph_b200.libsph import SphB200  # noqa: E402

CASES = {
    "c0": ("polytropic", "gaussian_sphere", 5000, dict(R=ic.R0), 100),
    "c1": ("polytropic", "sample_plummer_sphere", 100_000, {}, 10),
    "c2": ("isothermal", "boss_bodenheimer", 1_000_000, dict(T=10), 5),
    "c3": ("isothermal", "turbulent_molecular_cloud", 4_000_000, dict(T=10), 3),
    "c4": ("isothermal", "bonnor_ebert_sphere", 16_000_000, dict(T=10), 2),
}

for name in (sys.argv[1:] or list(CASES)):
    eos, ict, N, kw, steps = CASES[name]
    t0 = time.time()
    d = ic.make_ic(eos, ict, N=N, **kw)
    c = d["constants"]
    args = dict(m=c["m"], G=c["G"], theta=c["theta"], alpha=c["alpha"], beta=c["beta"])
    if eos == "isothermal":
        args.update(cs=c["cs"], U_iso=c["U"])
    else:
        args.update(gamma=c["gamma"])
    t1 = time.time()
    s = SphB200(N, c["Kh"], eos, **args)
    s.upload(d["pos"], d["vel"], d["K"], 0.0)
    s.step(1, want_info=False)                     # cold start (no radius hints yet)
    s.synchronize()
    t2 = time.time()
    info = s.step(steps)
    t3 = time.time()
    try:
        tim = s.timings()
    except Exception:                      # small N: steps are replayed as a CUDA graph, which records no phase timers
        s.eval_state()
        tim = s.timings()
    p, v, K, t = s.download()
    ok = np.isfinite(p).all() and np.isfinite(v).all() and np.isfinite(info["stats"]).all()
    print(f"{name} {eos} {ict} N={N}: IC {t1 - t0:.1f}s, first step {1e3 * (t2 - t1):.0f} ms, then "
          f"{1e3 * (t3 - t2) / steps:.2f} ms/step = {N * steps / (t3 - t2):.3g} particle-steps/s; finite={ok}; "
          f"E drift {abs(info['stats'][-1, 4] / info['stats'][0, 4] - 1):.2e}; knn retries {tim['knn_retries']:.0f}; "
          f"phases {{knn {tim['knn_ms']:.2f}, walk {tim['gravity_ms']:.2f}, force {tim['force_ms']:.2f}, sort {tim['sort_ms']:.2f}, tree {tim['tree_ms']:.2f}}}",
          flush=True)
    s.close()
