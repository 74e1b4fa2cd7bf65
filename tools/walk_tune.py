"""Walk variants side by side (GPU box): python tools/walk_tune.py [N] [cfg ...]

Each cfg is a comma-separated list of ENV=VALUE settings ("-" = defaults); every cfg runs in its own process (the
library reads its switches once), evaluates the same Boss-Bodenheimer IC twice and prints the gravity phase time, the
visit count (SPH_B200_COUNT_VISITS=1 run, separate process) and the deviation of g / Phi from the first cfg.
Not part of the product path."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(N, tag):
    import astrophysical_sph_b200.iniconds as ic
    from astrophysical_sph_b200.libsph import SphB200

    cache = f"/tmp/walk_tune_ic_{N}.npz"
    if os.path.exists(cache):
        z = np.load(cache, allow_pickle=True)
        pos, vel, c = np.asfortranarray(z["pos"]), np.asfortranarray(z["vel"]), z["c"].item()
    else:
        d = ic.make_ic("isothermal", "boss_bodenheimer", N=N, T=10)
        pos, vel, c = d["pos"], d["vel"], d["constants"]
        np.savez(cache, pos=pos, vel=vel, c=np.array(c, dtype=object))
    s = SphB200(N, c["Kh"], "isothermal", m=c["m"], cs=c["cs"], G=c["G"], theta=c["theta"], alpha=c["alpha"], beta=c["beta"],
                U_iso=c["U"])
    s.upload(pos, vel, None, 0.0)
    ms, km = [], []
    for _ in range(4):
        s.eval_acc(pos, vel, None)
        ms.append(s.timings()["gravity_ms"])
        km.append(s.timings()["walk_kernel_ms"])
    tm = s.timings()
    g, phi = s.grav()
    s.close()
    ref = f"/tmp/walk_tune_ref_{N}.npz"
    dev = ""
    if os.path.exists(ref):
        z = np.load(ref)
        eg = np.linalg.norm(g - z["g"], axis=1) / np.linalg.norm(z["g"], axis=1)
        ep = np.abs(phi / z["phi"] - 1)
        dev = "  dev vs first cfg: g %.2e phi %.2e" % (eg.max(), ep.max())
    else:
        np.savez(ref, g=g, phi=phi)
    print("%-44s grav_ms %s  kernel_ms %s  visits/particle %.2f%s" % (tag, " ".join("%.3f" % x for x in ms), " ".join("%.3f" % x for x in km), tm["walk_visits"] / N, dev),
          flush=True)
    print("%-44s   sort %.3f tree %.3f knn %.3f density %.3f force %.3f finish %.3f total %.3f" % ("", tm["sort_ms"], tm["tree_ms"], tm["knn_ms"],
          tm["density_ms"], tm["force_ms"], tm["finish_ms"], tm["total_ms"]), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), sys.argv[3])
        sys.exit(0)
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    cfgs = sys.argv[2:] or ["SPH_B200_WALK_DFS=1", "-", "SPH_B200_WALK_T=0", "SPH_B200_WALK_T=4", "SPH_B200_WALK_T=8", "SPH_B200_WALK_T=16"]
    for f in (f"/tmp/walk_tune_ref_{N}.npz",):
        if os.path.exists(f):
            os.remove(f)
    for cfg in cfgs:
        for count in ((False,) if os.environ.get("WALK_TUNE_NOCOUNT") else (False, True)):
            env = dict(os.environ)
            if cfg != "-":
                for kv in cfg.split(","):
                    k, v = kv.split("=")
                    env[k] = v
            if count:
                env["SPH_B200_COUNT_VISITS"] = "1"
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(N), cfg + (" [count]" if count else "")],
                               env=env, capture_output=True, text=True, timeout=150)
            print(r.stdout.strip() or ("FAILED: " + r.stderr[-800:]), flush=True)
