"""ctypes binding of libsph_b200.so (include/sph_b200.h) -- the only way this package computes anything.

There is no CPU fallback: importing works anywhere (the library loads without a GPU so its symbols can be
checked), but creating a handle without an sm_100 device raises SphError(SPH_ERR_NO_DEVICE).

Array conventions are the reference's (Julia column-major): an N x 3 matrix is passed as a Fortran-ordered
float64 array; neighbour indices come back N x Kh int32, 1-based, column 1 = self
(F/isothermal_hydroKDTree.jl:118-163, F = julia_version/fastv1_kd&single_oc).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SPH_B200_LIB selects another build of the same sources (tuning variants, csrc/Makefile VARIANT=...)
LIB_PATH = os.environ.get("SPH_B200_LIB") or os.path.join(_HERE, "libsph_b200.so")
CSRC = os.path.join(_HERE, "csrc")

ISOTHERMAL, POLYTROPIC = 0, 1
FLAG_COUNT_VISITS = 1       # sph_params.flags: count the node visits of the tree walk (timings()["walk_visits"])
FLAG_SERIAL_PHASES = 2      # sph_params.flags: density / force before the walk on one stream: timings() of every phase alone
EOS_CODES = {"isothermal": ISOTHERMAL, "polytropic": POLYTROPIC}

SPH_OK = 0
SPH_ERR_INVALID, SPH_ERR_CUDA, SPH_ERR_NO_DEVICE, SPH_ERR_TREE_DEPTH = -1, -2, -3, -4
SPH_ERR_TREE_NODES, SPH_ERR_NCCL, SPH_ERR_STATE, SPH_ERR_NAN = -5, -6, -7, -8

# every symbol include/sph_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = (
    "sph_create", "sph_destroy", "sph_last_error", "sph_abi_version", "sph_launch_count", "sph_device_count", "sph_set_stream",
    "sph_synchronize", "sph_upload", "sph_download", "sph_eval_acc", "sph_eval_state", "sph_step",
    "sph_get_neighbors", "sph_get_hydro", "sph_get_grav", "sph_get_acc", "sph_get_octree", "sph_get_timings",
    "sph_get_dt", "sph_density_at", "sph_comm_unique_id", "sph_comm_init", "sph_measure_fp64_peak",
)


class SphParams(C.Structure):
    _fields_ = [("N", C.c_int64), ("Kh", C.c_int32), ("eos", C.c_int32), ("m", C.c_double), ("cs", C.c_double),
                ("gamma", C.c_double), ("G", C.c_double), ("theta", C.c_double), ("alpha", C.c_double),
                ("beta", C.c_double), ("U_iso", C.c_double), ("device", C.c_int32), ("flags", C.c_int32)]


class SphStepInfo(C.Structure):
    _fields_ = [("dt", C.c_double), ("stats", C.c_double * 10)]


class SphTimings(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("sort_ms", "tree_ms", "knn_ms", "density_ms", "force_ms", "gravity_ms",
                                          "finish_ms", "total_ms", "walk_visits", "knn_retries", "comm_ms",
                                          "walk_kernel_ms")]


class SphError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsph_b200 error {code}: {msg}")
        self.code = code


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libsph_b200.so in-tree with nvcc for sm_100a (csrc/Makefile)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "sph_b200.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        cmd = ["make", "-C", CSRC, "-j8"] + (["-B"] if force else [])
        subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    """Load the shared library (never builds implicitly on a box without nvcc; raises if it is missing)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SphError(SPH_ERR_NO_DEVICE, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                                              "g.build()'` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.sph_last_error.restype = C.c_char_p
        L.sph_last_error.argtypes = [C.c_void_p]
        L.sph_launch_count.restype = C.c_int64
        for name in ABI_SYMBOLS:
            getattr(L, name)
        _lib = L
    return _lib


def launch_count() -> int:
    return int(lib().sph_launch_count())


def device_count() -> int:
    return int(lib().sph_device_count())


def measure_fp64_peak(device: int = 0) -> float:
    """Measured FP64 FMA throughput of `device` in TFLOP/s (the library's own DFMA microbenchmark)."""
    v = C.c_double(0.0)
    rc = lib().sph_measure_fp64_peak(int(device), C.byref(v))
    if rc != 0:
        raise SphError(rc, lib().sph_last_error(None).decode())
    return v.value


def _f64(a, order="F"):
    return np.require(a, dtype=np.float64, requirements=["F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS", "ALIGNED"])


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class SphB200:
    """One handle = one B200.  Mirrors the data flow of run_simulation (F/isothermal_sim.jl:72-298)."""

    def __init__(self, N, Kh=50, eos="isothermal", m=1.0, cs=0.0, gamma=5.0 / 3, G=6.6743e-8, theta=0.576, alpha=1.0,
                 beta=2.0, U_iso=0.0, device=0, flags=0):
        self._h = C.c_void_p()
        self.N, self.Kh = int(N), int(Kh)
        self.eos = EOS_CODES[eos] if isinstance(eos, str) else int(eos)
        p = SphParams(self.N, self.Kh, self.eos, float(m), float(cs), float(gamma), float(G), float(theta),
                      float(alpha), float(beta), float(U_iso), int(device), int(flags))
        self.params = p
        rc = lib().sph_create(C.byref(p), C.byref(self._h))
        if rc != 0:
            raise SphError(rc, lib().sph_last_error(None).decode())

    # -- plumbing
    def _chk(self, rc):
        if rc != 0:
            raise SphError(rc, lib().sph_last_error(self._h).decode())

    def close(self):
        if self._h:
            lib().sph_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream_ptr):
        self._chk(lib().sph_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self._chk(lib().sph_synchronize(self._h))

    # -- state
    def upload(self, pos, vel, K=None, t=0.0):
        pos, vel = _f64(pos), _f64(vel)
        assert pos.shape == (self.N, 3) and vel.shape == (self.N, 3)
        Ka = None if K is None else _f64(np.asarray(K).reshape(-1), "C")
        self._chk(lib().sph_upload(self._h, _p(pos), _p(vel), _p(Ka), C.c_double(t)))

    def download(self):
        pos = np.zeros((self.N, 3), order="F")
        vel = np.zeros((self.N, 3), order="F")
        K = np.zeros(self.N) if self.eos == POLYTROPIC else None
        t = C.c_double(0.0)
        self._chk(lib().sph_download(self._h, _p(pos), _p(vel), _p(K), C.byref(t)))
        return pos, vel, K, t.value

    # -- hot path
    def eval_acc(self, pos, vel, K=None, out=None):
        """One getAcc on host arrays; returns dict(acc, rho, h, phi). `out` may hold preallocated arrays."""
        pos, vel = _f64(pos), _f64(vel)
        Ka = None if K is None else _f64(np.asarray(K).reshape(-1), "C")
        o = out or dict(acc=np.zeros((self.N, 3), order="F"), rho=np.zeros(self.N), h=np.zeros(self.N),
                        phi=np.zeros(self.N))
        self._chk(lib().sph_eval_acc(self._h, _p(pos), _p(vel), _p(Ka), _p(o["acc"]), _p(o["rho"]), _p(o["h"]),
                                     _p(o["phi"])))
        return o

    def eval_state(self):
        self._chk(lib().sph_eval_state(self._h))

    def step(self, nsteps=1, want_info=True):
        info = (SphStepInfo * nsteps)() if want_info else None
        self._chk(lib().sph_step(self._h, int(nsteps), info))
        if not want_info:
            return None
        return dict(dts=np.array([info[i].dt for i in range(nsteps)]),
                    stats=np.array([list(info[i].stats) for i in range(nsteps)]))

    # -- inspection
    def neighbors(self, want_r=True):
        idx = np.zeros((self.N, self.Kh), dtype=np.int32, order="F")
        r = np.zeros((self.N, self.Kh), order="F") if want_r else None
        self._chk(lib().sph_get_neighbors(self._h, _p(idx), _p(r)))
        return idx, r

    def hydro(self):
        o = dict(ahyd=np.zeros((self.N, 3), order="F"), rho=np.zeros(self.N), h=np.zeros(self.N),
                 sum_vdw=np.zeros(self.N), mumax=np.zeros(self.N), cs_i=np.zeros(self.N), dkdt=np.zeros(self.N))
        self._chk(lib().sph_get_hydro(self._h, _p(o["ahyd"]), _p(o["rho"]), _p(o["h"]), _p(o["sum_vdw"]),
                                      _p(o["mumax"]), _p(o["cs_i"]), _p(o["dkdt"])))
        return o

    def grav(self):
        g = np.zeros((self.N, 3), order="F")
        phi = np.zeros(self.N)
        self._chk(lib().sph_get_grav(self._h, _p(g), _p(phi)))
        return g, phi

    def acc(self):
        a = np.zeros((self.N, 3), order="F")
        self._chk(lib().sph_get_acc(self._h, _p(a)))
        return a

    def octree(self):
        n = C.c_int64(0)
        self._chk(lib().sph_get_octree(self._h, None, C.c_int64(0), C.byref(n)))
        out = np.zeros((n.value, 16))
        self._chk(lib().sph_get_octree(self._h, _p(out), C.c_int64(n.value), C.byref(n)))
        return out

    def timings(self):
        t = SphTimings()
        self._chk(lib().sph_get_timings(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in SphTimings._fields_}

    def dt(self):
        d = C.c_double(0.0)
        self._chk(lib().sph_get_dt(self._h, C.byref(d)))
        return d.value

    def density_at(self, pts):
        pts = _f64(pts)
        out = np.zeros(pts.shape[0])
        self._chk(lib().sph_density_at(self._h, _p(pts), C.c_int64(pts.shape[0]), _p(out)))
        return out

    # -- multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = lib().sph_comm_unique_id(buf)
        if rc != 0:
            raise SphError(rc, lib().sph_last_error(None).decode())
        return buf.raw

    def comm_init(self, nranks, rank, unique_id: bytes):
        assert len(unique_id) == 128
        self._chk(lib().sph_comm_init(self._h, int(nranks), int(rank), C.c_char_p(unique_id)))
