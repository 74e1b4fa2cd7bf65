"""GPU bring-up script: runs the CUDA path next to the CPU oracle and prints / dumps every intermediate.

Usage (on a GPU box):  python tools/gpu_debug.py [N ...]   ->  gpurun_out/debug_<ic>_<N>.npz + stdout summary
Not part of the product path; the oracle is imported here only as the checker.
"""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import astrophysical_sph_b200.iniconds as ic  # noqa: E402
from astrophysical_sph_b200.libsph import SphB200  # noqa: E402
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def relerr(a, b, floor=0.0):
    a = np.asarray(a); b = np.asarray(b)
    den = np.maximum(np.abs(b), floor) if floor else np.abs(b)
    with np.errstate(divide="ignore", invalid="ignore"):
        e = np.abs(a - b) / den
    e[~np.isfinite(e)] = 0.0
    return float(e.max())


def vec_relerr(a, b):
    na = np.linalg.norm(a - b, axis=1); nb = np.linalg.norm(b, axis=1)
    return float((na / np.maximum(nb, 1e-300)).max())


def run_case(eos, ic_type, N, nthreads, dump=True, steps=0, **kw):
    print(f"=== {eos} {ic_type} N={N}", flush=True)
    d = ic.make_ic(eos, ic_type, N=N, **kw)
    c = d["constants"]
    pos, vel, K = d["pos"], d["vel"], d["K"]
    Kh = c["Kh"]
    args = dict(m=c["m"], G=c["G"], theta=c["theta"], alpha=c["alpha"], beta=c["beta"])
    if eos == "isothermal":
        args.update(cs=c["cs"], U_iso=c["U"])
    else:
        args.update(gamma=c["gamma"])
    s = SphB200(N, Kh, eos, **args)
    t0 = time.time()
    out = s.eval_acc(pos, vel, K)
    t1 = time.time()
    print("gpu eval_acc wall %.3fs  timings %s" % (t1 - t0, s.timings()), flush=True)
    out = s.eval_acc(pos, vel, K)
    print("gpu eval_acc (2nd) timings", {k: round(v, 3) for k, v in s.timings().items()}, flush=True)
    idx, r = s.neighbors()
    hy = s.hydro()
    g, phi = s.grav()
    tree = s.octree()
    res = dict(idx=idx, r=r, g=g, phi=phi, acc=out["acc"], tree=tree, **{"hy_" + k: v for k, v in hy.items()})
    ok = True
    if N <= 300000:
        t0 = time.time()
        eo = O.ISOTHERMAL if eos == "isothermal" else O.POLYTROPIC
        oh = O.hydro(pos, vel, c["m"], Kh, eos=eo, cs=c.get("cs", 0.0), Kent=K, gamma=c.get("gamma", 5 / 3),
                     alpha=c["alpha"], beta=c["beta"], nthreads=nthreads)
        l = np.abs(pos).max()
        og, ophi, st = O.gravity(l, c["m"], pos, c["theta"], oh["h"], nthreads=nthreads)
        otree = O.octree(l, c["m"], pos)
        print("oracle %.2fs; tree stats %s" % (time.time() - t0, st), flush=True)
        same_idx = np.array_equal(idx, oh["idx"])
        nbad = int((idx != oh["idx"]).any(axis=1).sum())
        print(f"idx identical: {same_idx}  rows differing: {nbad}")
        if nbad:
            rows = np.where((idx != oh["idx"]).any(axis=1))[0][:5]
            for rr in rows:
                print(" row", rr, "gpu", idx[rr][:8], "...", idx[rr][-4:], "ora", oh["idx"][rr][:8], "...", oh["idx"][rr][-4:])
                print("     r gpu", r[rr][-4:], "r ora", oh["r"][rr][-4:])
        print("r max abs diff rel:", relerr(r, oh["r"], 1e-300))
        print("h relerr", relerr(hy["h"], oh["h"]), " rho relerr", relerr(hy["rho"], oh["rho"]))
        print("ahyd vec relerr", vec_relerr(hy["ahyd"], oh["ahyd"]), " sum_vdw relerr",
              relerr(hy["sum_vdw"], oh["sum_vdw"], np.abs(oh["sum_vdw"]).max() * 1e-6), " mumax maxabs",
              float(np.abs(hy["mumax"] - oh["mumax"]).max()))
        if eos == "polytropic":
            print("cs_i relerr", relerr(hy["cs_i"], oh["cs_i"]), " dkdt relerr",
                  relerr(hy["dkdt"], oh["dkdt"], np.abs(oh["dkdt"]).max() * 1e-6))
        print("tree nodes gpu", tree.shape[0], "oracle", otree.shape[0])
        if tree.shape == otree.shape:
            for name, sl in (("Length", slice(0, 1)), ("centre", slice(1, 4)), ("lo", slice(4, 7)), ("hi", slice(7, 10)),
                             ("Mass", slice(10, 11)), ("rCOM", slice(11, 14)), ("count", slice(14, 15)),
                             ("depth", slice(15, 16))):
                dd = np.abs(tree[:, sl] - otree[:, sl]).max()
                print(f"   tree {name:7s} max abs diff {dd:.3e}  (scale {np.abs(otree[:, sl]).max():.3e})")
        print("g vec relerr", vec_relerr(g, og), " phi relerr", relerr(phi, ophi))
        oacc = oh["ahyd"] - c["G"] * og
        print("acc vec relerr", vec_relerr(out["acc"], oacc))
        res.update(o_idx=oh["idx"], o_r=oh["r"], o_rho=oh["rho"], o_h=oh["h"], o_ahyd=oh["ahyd"], o_g=og, o_phi=ophi,
                   o_tree=otree)
    if steps:
        s.upload(pos, vel, K, 0.0)
        t0 = time.time()
        info = s.step(steps)
        tg = time.time() - t0
        gp, gv, gK, gt = s.download()
        print(f"gpu {steps} steps wall {tg:.3f}s  dts {info['dts']}", flush=True)
        if N <= 20000:
            eo = O.ISOTHERMAL if eos == "isothermal" else O.POLYTROPIC
            t0 = time.time()
            oo = O.step(pos, vel, c["m"], Kh, c["G"], c["theta"], 0.0, steps, eos=eo, cs=c.get("cs", 0.0), Kent=K,
                        gamma=c.get("gamma", 5 / 3), alpha=c["alpha"], beta=c["beta"], U_iso=c.get("U", 0.0),
                        nthreads=nthreads)
            print(f"oracle {steps} steps {time.time() - t0:.2f}s dts {oo['dts']}")
            print("dt relerr", relerr(info["dts"], oo["dts"]), " t", gt, oo["t"])
            print("pos relerr(vec)", vec_relerr(gp, oo["pos"]), " vel vec relerr", vec_relerr(gv, oo["vel"]))
            print("stats |diff| per column", [float(np.abs(info["stats"][:, k] - oo["stats"][:, k]).max()) for k in range(10)])
            print("stats gpu row0", info["stats"][0]); print("stats ora row0", oo["stats"][0])
            if gK is not None:
                print("K relerr", relerr(gK, oo["K"]))
    if dump and N <= 20000:
        np.savez_compressed(os.path.join(OUT, f"debug_{eos}_{ic_type}_{N}.npz"), pos=pos, vel=vel, **res)
    s.close()
    return ok


if __name__ == "__main__":
    nthreads = O.max_threads()
    print("oracle threads", nthreads)
    cases = sys.argv[1:] or ["small"]
    for cs_ in cases:
        try:
            if cs_ == "small":
                run_case("isothermal", "gaussian_sphere", 5000, nthreads, steps=3, R=ic.R0)
                run_case("polytropic", "gaussian_sphere", 5000, nthreads, steps=3, R=ic.R0)
            elif cs_ == "plummer":
                run_case("polytropic", "sample_plummer_sphere", 100000, nthreads, steps=1)
            elif cs_ == "bb100k":
                run_case("isothermal", "boss_bodenheimer", 100000, nthreads, steps=1, T=10)
            elif cs_ == "bb1m":
                run_case("isothermal", "boss_bodenheimer", 1000000, nthreads, steps=2, T=10, dump=False)
        except Exception:
            traceback.print_exc()
