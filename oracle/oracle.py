"""ctypes front-end of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY -- see the header of oracle/sph_oracle.cpp.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
PARITY UNPINNED: the reference has no golden vectors and cannot run here (no Julia).

All matrices follow the Julia layout: N x 3 arrays are passed as Fortran-ordered
float64 (three contiguous length-N columns); neighbour indices are N x K Int32,
1-based, column 1 = self (F/isothermal_hydroKDTree.jl:118-163).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

ISOTHERMAL, POLYTROPIC = 0, 1


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (g++ -O2 -ffp-contract=off)."""
    src = os.path.join(_HERE, "sph_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_last_error.restype = C.c_char_p
        _lib.oracle_dt.restype = C.c_double
        _lib.oracle_octree.restype = C.c_longlong
    return _lib


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _f(a):
    """float64, Fortran order (Julia column-major)."""
    return np.asfortranarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _chk(rc):
    if rc != 0:
        raise RuntimeError("oracle: " + lib().oracle_last_error().decode())


def _d(x):
    return C.c_double(float(x))


def knn(q, pos, K, nthreads=1):
    """getNeighbors core: (idx 1-based M x K int32, r M x K)."""
    q = _f(q); pos = _f(pos)
    M, N = q.shape[0], pos.shape[0]
    idx = np.zeros((M, K), dtype=np.int32, order="F")
    r = np.zeros((M, K), dtype=np.float64, order="F")
    _chk(lib().oracle_knn(M, _p(q), N, _p(pos), K, _p(idx), _p(r), nthreads))
    return idx, r


def hydro(pos, vel, m, K, eos=ISOTHERMAL, cs=0.0, Kent=None, gamma=5.0 / 3, alpha=1.0, beta=2.0, nthreads=1):
    """HJL.hydrodynamics reduced to what its caller consumes (dict)."""
    pos = _f(pos); vel = _f(vel)
    N = pos.shape[0]
    Kent_a = None if Kent is None else np.ascontiguousarray(Kent, dtype=np.float64)
    out = dict(
        idx=np.zeros((N, K), dtype=np.int32, order="F"),
        r=np.zeros((N, K), order="F"),
        ahyd=np.zeros((N, 3), order="F"),
        rho=np.zeros(N), h=np.zeros(N), sum_vdw=np.zeros(N), mumax=np.zeros(N),
        cs_i=np.zeros(N), dkdt=np.zeros(N),
    )
    _chk(lib().oracle_hydro(N, _p(pos), _p(vel), _d(m), eos, _d(cs), _p(Kent_a), _d(gamma), _d(alpha), _d(beta),
                            K, _p(out["idx"]), _p(out["r"]), _p(out["ahyd"]), _p(out["rho"]), _p(out["h"]),
                            _p(out["sum_vdw"]), _p(out["mumax"]), _p(out["cs_i"]), _p(out["dkdt"]), nthreads))
    return out


def gravity(l_domain, m, pos, theta, h, nthreads=1):
    """GJL.gravity: (g N x 3, PHI N, tree_stats[nodes, depth, leaf, mono, opened])."""
    pos = _f(pos); h = np.ascontiguousarray(h, dtype=np.float64)
    N = pos.shape[0]
    g = np.zeros((N, 3), order="F"); phi = np.zeros(N); st = np.zeros(5)
    _chk(lib().oracle_gravity(N, _d(l_domain), _d(m), _p(pos), _d(theta), _p(h), _p(g), _p(phi), _p(st), nthreads))
    return g, phi, st


def getacc(pos, vel, m, K, G, theta, eos=ISOTHERMAL, cs=0.0, Kent=None, gamma=5.0 / 3, alpha=1.0, beta=2.0,
           nthreads=1):
    pos = _f(pos); vel = _f(vel)
    N = pos.shape[0]
    Kent_a = None if Kent is None else np.ascontiguousarray(Kent, dtype=np.float64)
    out = dict(acc=np.zeros((N, 3), order="F"), rho=np.zeros(N), h=np.zeros(N), phi=np.zeros(N),
               sum_vdw=np.zeros(N), mumax=np.zeros(N), cs_i=np.zeros(N), dkdt=np.zeros(N))
    _chk(lib().oracle_getacc(N, _p(pos), _p(vel), _d(m), eos, _d(cs), _p(Kent_a), _d(gamma), _d(G), _d(theta),
                             _d(alpha), _d(beta), K, _p(out["acc"]), _p(out["rho"]), _p(out["h"]), _p(out["phi"]),
                             _p(out["sum_vdw"]), _p(out["mumax"]), _p(out["cs_i"]), _p(out["dkdt"]), nthreads))
    return out


def step(pos, vel, m, K, G, theta, t, nsteps, eos=ISOTHERMAL, cs=0.0, Kent=None, gamma=5.0 / 3, alpha=1.0,
         beta=2.0, U_iso=0.0, nthreads=1):
    """nsteps iterations of the simulation loop body. Returns dict(pos, vel, K, t, dts, stats)."""
    pos = _f(np.array(pos, copy=True)); vel = _f(np.array(vel, copy=True))
    N = pos.shape[0]
    Kent_a = None if Kent is None else np.array(Kent, dtype=np.float64, copy=True)
    tt = C.c_double(float(t))
    dts = np.zeros(nsteps); stats = np.zeros((nsteps, 10))
    _chk(lib().oracle_step(N, _p(pos), _p(vel), _p(Kent_a), _d(m), eos, _d(cs), _d(gamma), _d(G), _d(theta),
                           _d(alpha), _d(beta), K, _d(U_iso), C.byref(tt), nsteps, _p(dts), _p(stats), nthreads))
    return dict(pos=pos, vel=vel, K=Kent_a, t=tt.value, dts=dts, stats=stats)


def dt_from(vel, acc, rho, h, sum_vdw, mumax, m, eos=ISOTHERMAL, cs=0.0, cs_i=None, alpha=1.0, beta=2.0):
    vel = _f(vel); acc = _f(acc)
    N = vel.shape[0]
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (rho, h, sum_vdw, mumax)]
    ci = None if cs_i is None else np.ascontiguousarray(cs_i, dtype=np.float64)
    return float(lib().oracle_dt(N, _p(vel), _p(acc), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(ci), _d(m), eos,
                                 _d(cs), _d(alpha), _d(beta)))


def density_at(pts, pos, m, K, eos=ISOTHERMAL, nthreads=1):
    pts = _f(pts); pos = _f(pos)
    M, N = pts.shape[0], pos.shape[0]
    out = np.zeros(M)
    _chk(lib().oracle_density_at(M, _p(pts), N, _p(pos), _d(m), K, eos, _p(out), nthreads))
    return out


def octree(l_domain, m, pos):
    """Node table of the reference's BFS octree (for structural tests). Columns:
    Length, centre(3), lo(3), hi(3), Mass, rCOM(3), particle_count, depth."""
    pos = _f(pos)
    N = pos.shape[0]
    n = lib().oracle_octree(N, _d(l_domain), _d(m), _p(pos), None, C.c_longlong(0))
    if n < 0:
        _chk(-1)
    out = np.zeros((n, 16))
    lib().oracle_octree(N, _d(l_domain), _d(m), _p(pos), _p(out), C.c_longlong(n))
    return out
