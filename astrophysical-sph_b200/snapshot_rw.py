"""Snapshot CSV and stats-file I/O in the reference's on-disk layout.

Mirrors module SnapshotRW (F/SnapshotRW.jl, F = julia_version/fastv1_kd&single_oc): same function names, same
argument meaning, same files, so snapshots written here are read by the Julia `read_snapshot` (F/SnapshotRW.jl:123-159)
and vice versa.  PNG figure output (:101-107) is not provided (GLMakie is the reference's GUI layer, out of scope).

  snapshots/<ic_type>/bin/<snapID>snap.csv   header  type,x,y,z,vx,vy,vz,K,rlin,rho_radial,constants      (:37-49)
                                             N rows  particle,<x>,...,<vz>,<K or empty>,,,
                                             rows    rlin / rho_radial (';'-joined vectors, :57-83), constants (:86-97)
  snapshots/<ic_type>/stats                  100000 x 10 Float64, column-major, memory mapped                 (:171-184)

Floats are printed like Julia prints Float64 (shortest round-trip digits, `1.0`, `5.0e12`, `6.6743e-8`): the reader
re-types constants by text shape -- a value containing `e`, `E` or `.` becomes Float64, anything else Int
(F/SnapshotRW.jl:147) -- and the Julia drivers' Float64-typed signatures reject an Int where a Float64 is expected.
"""
from __future__ import annotations

import os

import numpy as np

NSTEPS = 100000   # F/SnapshotRW.jl:171
NFIELDS = 10      # F/SnapshotRW.jl:172

COLUMNS = ("type", "x", "y", "z", "vx", "vy", "vz", "K", "rlin", "rho_radial", "constants")


def julia_float_str(x) -> str:
    """Text Julia's `string(::Float64)` produces: fixed notation for 1e-4 <= |x| < 1e6, else `<mantissa>e<exp>`;
    always at least one fractional digit; shortest digits that round-trip."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Inf" if x > 0 else "-Inf"
    if x == 0.0:
        return "-0.0" if str(x).startswith("-") else "0.0"
    r = repr(x)
    sign = "-" if r.startswith("-") else ""
    r = r.lstrip("-")
    # decimal digits and exponent from the shortest repr
    if "e" in r:
        mant, ex = r.split("e")
        ex = int(ex)
    else:
        mant, ex = r, 0
    if "." in mant:
        ip, fp = mant.split(".")
    else:
        ip, fp = mant, ""
    digits = (ip + fp).lstrip("0")
    point = len(ip) + ex                      # value = 0.<ip fp> * 10^point  (before stripping zeros)
    lead = len(ip + fp) - len((ip + fp).lstrip("0"))
    point -= lead
    digits = digits.rstrip("0") or "0"
    e10 = point - 1                           # value = d.ddd * 10^e10
    if -4 <= e10 < 6:                         # Julia: 0.0001 -> "0.0001", 0.00001 -> "1.0e-5", 1e6 -> "1.0e6"
        if point <= 0:
            s = "0." + "0" * (-point) + digits
        elif point >= len(digits):
            s = digits + "0" * (point - len(digits)) + ".0"
        else:
            s = digits[:point] + "." + digits[point:]
    else:
        s = digits[0] + "." + (digits[1:] or "0") + "e" + str(e10)
    return sign + s


def _const_str(v) -> str:
    if isinstance(v, (bool, np.bool_)):
        return "true" if v else "false"
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    return julia_float_str(v)


def _ensure_dirs(root, ic_type):
    for sub in ("bin", "graphs"):     # the README asks the user to create these by hand (R/README.md:41-53)
        os.makedirs(os.path.join(root, "snapshots", ic_type, sub), exist_ok=True)


def snapshot_path(snapID, ic_type, root="."):
    return os.path.join(root, "snapshots", ic_type, "bin", f"{snapID}snap.csv")


def write_snapshot(snapID, ic_type, pos, vel, K=None, constants=None, rlin=None, rho_radial=None, fig1=None,
                   fig2=None, type="particle", root="."):
    """F/SnapshotRW.jl:22-109.  fig1 / fig2 are accepted for signature parity and ignored."""
    pos = np.asarray(pos, dtype=np.float64)
    vel = np.asarray(vel, dtype=np.float64)
    N = pos.shape[0]
    _ensure_dirs(root, ic_type)
    fmt = julia_float_str
    cols = [[fmt(v) for v in pos[:, k]] for k in range(3)] + [[fmt(v) for v in vel[:, k]] for k in range(3)]
    kcol = [fmt(v) for v in np.asarray(K, dtype=np.float64).reshape(-1)] if K is not None else [""] * N
    path = snapshot_path(snapID, ic_type, root)
    with open(path, "w", newline="") as f:
        f.write(",".join(COLUMNS) + "\n")
        f.writelines(f"{type},{x},{y},{z},{vx},{vy},{vz},{k},,,\n"
                     for x, y, z, vx, vy, vz, k in zip(*cols, kcol))
        if rlin is not None and len(rlin):
            f.write("rlin,,,,,,,," + ";".join(fmt(v) for v in rlin) + ",,\n")
        if rho_radial is not None and len(rho_radial):
            f.write("rho_radial,,,,,,,,," + ";".join(fmt(v) for v in rho_radial) + ",\n")
        if constants:
            f.write("constants,,,,,,,,,," + ";".join(f"{k}={_const_str(v)}" for k, v in constants.items()) + "\n")
    return path


def read_snapshot(filename):
    """F/SnapshotRW.jl:123-159 -> dict(pos, vel, K, rlin, rho_radial, constants); pos/vel are N x 3 Fortran-ordered
    float64 (Matrix{Float64} layout), K is None when the column is empty."""
    import pandas as pd

    df = pd.read_csv(filename, dtype={"type": str, "rlin": str, "rho_radial": str, "constants": str},
                     float_precision="round_trip")   # Julia parses Float64 text exactly; so must we
    part = df[df["type"] == "particle"]
    pos = np.asfortranarray(part[["x", "y", "z"]].to_numpy(dtype=np.float64))
    vel = np.asfortranarray(part[["vx", "vy", "vz"]].to_numpy(dtype=np.float64))
    K = None
    if "K" in part.columns and part["K"].notna().any():
        K = part["K"].to_numpy(dtype=np.float64)

    def vec(name):
        rows = df[df["type"] == name]
        return np.array([float(x) for x in rows.iloc[0][name].split(";")]) if len(rows) == 1 else np.zeros(0)

    constants = {}
    rows = df[df["type"] == "constants"]
    if len(rows) == 1:
        for pair in rows.iloc[0]["constants"].split(";"):
            k, v = pair.split("=")
            constants[k] = float(v) if any(ch in v for ch in "eE.") else int(v)    # :147
    return dict(pos=pos, vel=vel, K=K, rlin=vec("rlin"), rho_radial=vec("rho_radial"), constants=constants)


def open_or_create_stats_mmap(filename):
    """F/SnapshotRW.jl:174-184 -> (arr, io): 100000 x 10 float64, column-major, zero-initialised on creation."""
    isnew = not os.path.isfile(filename)
    if isnew:
        os.makedirs(os.path.dirname(filename) or ".", exist_ok=True)
        with open(filename, "wb") as f:
            f.write(b"\0" * (NSTEPS * NFIELDS * 8))
    arr = np.memmap(filename, dtype=np.float64, mode="r+", shape=(NSTEPS, NFIELDS), order="F")
    return arr, arr     # the memmap doubles as the io handle (flush() = Mmap.sync!, del = close)


def update_stats_row(arr, iterID, stats):
    """F/SnapshotRW.jl:191-195 (`update_stats_row!`); iterID is 1-based as in the reference."""
    assert 1 <= iterID <= arr.shape[0], "Iteration index out of bounds"
    assert len(stats) == arr.shape[1], "Mismatch in stats length"
    arr[iterID - 1, :] = stats


def get_stats_up_to(arr, iterID):
    """F/SnapshotRW.jl:203-205."""
    return np.array(arr[:iterID, :])
