"""Host-side mirror of the reference's simulation drivers, on top of libsph_b200.so.

Mirrors isothermalSim / polytropeSim (F/isothermal_sim.jl, F/polytrope_sim.jl; F = julia_version/fastv1_kd&single_oc):
  getAcc(...)          one force evaluation with the reference's argument lists (:16 / :17)
  run_simulation(...)  read snapshot -> while t < tEnd {step, stats row, snapshot?} (:72-298 / :84-323)
The Julia toolchain is absent from the build image, so the host side above the C ABI is Python; the Julia `ccall`
shim a maintainer would add to the reference is julia/SphB200.jl.

Deliberate differences from the reference, all outside the hot path:
  * the seven N x K pass-through matrices of getAcc (dWx.., vij_.., mu) are reduced on the device to the two row
    reductions the caller consumes (sum_j v_ij.gradW_ij and max_j mu_ij, F/isothermal_sim.jl:160,165);
  * GLMakie windows / PNG output (F/isothermal_sim.jl:129-142, 228-270) are not provided; `showPlots` is accepted
    and ignored;
  * `snapshots/<ic>/{bin,graphs}` are created when missing instead of being required by the README.
"""
from __future__ import annotations

import time

import numpy as np

from . import snapshot_rw as SnapshotRW
from .libsph import SphB200

PLOT_N = {"isothermal": 1000, "polytropic": 10000}      # F/isothermal_sim.jl:122, F/polytrope_sim.jl:129


def _handle(eos, N, Kh, m, G, theta, alpha, beta, cs=0.0, gamma=5.0 / 3, U_iso=0.0, device=0):
    return SphB200(N, Kh, eos, m=m, cs=cs, gamma=gamma, G=G, theta=theta, alpha=alpha, beta=beta, U_iso=U_iso,
                   device=device)


def getAcc_isothermal(pos, vel, m, cs, G, theta, alpha, beta, Kh, handle=None):
    """isothermalSim.getAcc (F/isothermal_sim.jl:16-49) -> (acc, rho, h, sum_vdw, mumax, PHI)."""
    own = handle is None
    s = handle or _handle("isothermal", pos.shape[0], Kh, m, G, theta, alpha, beta, cs=cs)
    try:
        out = s.eval_acc(pos, vel)
        hy = s.hydro()
        return out["acc"], out["rho"], out["h"], hy["sum_vdw"], hy["mumax"], out["phi"]
    finally:
        if own:
            s.close()


def getAcc_polytropic(pos, vel, m, K, gamma, G, theta, alpha, beta, Kh, handle=None):
    """polytropeSim.getAcc (F/polytrope_sim.jl:17-51) -> (acc, rho, h, sum_vdw, mumax, cs_i, dK_sum, PHI)."""
    own = handle is None
    s = handle or _handle("polytropic", pos.shape[0], Kh, m, G, theta, alpha, beta, gamma=gamma)
    try:
        out = s.eval_acc(pos, vel, K)
        hy = s.hydro()
        return out["acc"], out["rho"], out["h"], hy["sum_vdw"], hy["mumax"], hy["cs_i"], hy["dkdt"], out["phi"]
    finally:
        if own:
            s.close()


def find_star_radius(rlin, rho_radial, threshold=1e-20):
    """F/polytrope_sim.jl:75-80."""
    assert len(rlin) == len(rho_radial), "rlin and rho_radial must be the same length"
    below = np.nonzero(np.asarray(rho_radial) < threshold)[0]
    return float(rlin[-1]) if below.size == 0 else float(rlin[below[0]])


def run_simulation(eos, ic_type, snapID, snapInterval, keepSnaps, showPlots=False, root=".", device=0,
                   max_steps=None, verbose=True):
    """run_simulation(ic_type, snapID, snapInterval, keepSnaps, showPlots) of either driver.

    Returns dict(steps, t, runtime_s, snapshots=[ids written]).  `max_steps` (not in the reference) bounds the
    loop for tests and benchmarks."""
    start = time.time()
    snap = SnapshotRW.read_snapshot(SnapshotRW.snapshot_path(snapID, ic_type, root))
    pos, vel, constants = snap["pos"], snap["vel"], dict(snap["constants"])
    poly = eos == "polytropic"
    iterID = int(constants["iterID"]); N = int(constants["N"]); Kh = int(constants["Kh"])
    t = float(constants["t"]); tEnd = float(constants["tEnd"]); R = float(constants["R"])
    m = float(constants["m"])
    K = np.asarray(snap["K"], dtype=np.float64) if poly else None     # F/polytrope_sim.jl:116-117
    if poly and K is None:
        raise ValueError("polytropic snapshot without a K column")
    s = _handle(eos, N, Kh, m, float(constants["G"]), float(constants["theta"]), float(constants["alpha"]),
                float(constants["beta"]), cs=float(constants.get("cs", 0.0)), gamma=float(constants.get("gamma", 5 / 3)),
                U_iso=float(constants.get("U", 0.0)), device=device)
    intervalCounter = snapInterval                                      # F/isothermal_sim.jl:108
    plotN = PLOT_N[eos]
    rlin = (np.linspace(-1, 1, plotN) * R) if not poly else np.linspace(0, 1.5 * R, plotN)   # :124 / poly :131
    stats_arr, _ = SnapshotRW.open_or_create_stats_mmap(
        SnapshotRW.os.path.join(root, "snapshots", ic_type, "stats"))
    if verbose:
        print("Starting simulation...")
    s.upload(pos, vel, K, t)
    written, steps = [], 0
    try:
        while t < tEnd and (max_steps is None or steps < max_steps):
            info = s.step(1)                                            # the whole loop body :155-212 on the device
            row = info["stats"][0]
            if poly and verbose:
                print("Virial Ratio: ", abs(row[2] / row[3]) if row[3] else float("inf"))   # F/polytrope_sim.jl:190
            SnapshotRW.update_stats_row(stats_arr, iterID, row)         # :192
            t = t + info["dts"][0]
            if verbose:
                print("Time: ", t)                                      # :213
            if keepSnaps * intervalCounter == snapInterval or t >= tEnd:    # :216
                rr = np.zeros((plotN, 3), order="F")
                rr[:, 0] = rlin + row[5]                                # samples on the x axis through the COM :218-219
                rr[:, 1] = row[6]; rr[:, 2] = row[7]
                rho_radial = s.density_at(rr)                           # HJL.density_plot :220
                intervalCounter = 0
                constants["iterID"] = iterID
                constants["t"] = t
                if poly:
                    R = find_star_radius(rlin, rho_radial, threshold=0.01 * rho_radial[0])   # poly :242
                    constants["N"] = N
                    constants["R"] = R
                    if verbose:
                        print(f"Saving snapshot with ID: {iterID}")
                p, v, Kd, _ = s.download()
                stats_arr.flush()                                       # Mmap.sync!
                SnapshotRW.write_snapshot(str(iterID), ic_type, p, v, K=Kd, constants=constants, rlin=rlin,
                                          rho_radial=rho_radial, root=root)
                written.append(iterID)
            iterID += 1
            intervalCounter += 1
            steps += 1
    finally:
        stats_arr.flush()
        s.close()
    runtime = time.time() - start
    if verbose:
        print(f"B200: octant-key tree for Pressure/AV + Octree for Smoothed Gravity. Runtime: {runtime} seconds")
    return dict(steps=steps, t=t, runtime_s=runtime, snapshots=written)
