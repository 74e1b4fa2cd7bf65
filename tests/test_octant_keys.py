"""The two-word octant path keys of csrc/sph_internal.cuh (host-callable) against a plain-Python replay of addNodes!.

build_octree! (F/gravOctree_Single.jl:213-227) subdivides until every leaf holds one particle, at ANY depth; the GPU tree
is derived from keys that hold 21 levels per 63-bit word.  This pins, without a GPU: the octant digits of both words
(strict `>` against the parent centre, centre recurrence c -/+ L/2, :110-148), the cell geometry replayed from a key
(centre, bounds, half-width: bit-identical to the recurrence), and the common-leading-levels count that decides where two
particles part ways in the tree."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from conftest import ROOT

HARNESS = r"""
#include "sph_internal.cuh"
extern "C" {
void kt_key_words(double x, double y, double z, double l, unsigned long long *hi, unsigned long long *lo) {
    *hi = sph_octant_key_word(x, y, z, l, 0);
    *lo = sph_octant_key_word(x, y, z, l, 1);
}
int kt_common(unsigned long long a, unsigned long long b) { return sph_common_levels(a, b); }
void kt_cell(unsigned long long hi, unsigned long long lo, int depth, double l, double *out10) {
    SphCell g = sph_cell_of(hi, lo, depth, l);
    for (int a = 0; a < 3; ++a) { out10[a] = g.c[a]; out10[3 + a] = g.lo[a]; out10[6 + a] = g.hi[a]; }
    out10[9] = g.L;
}
}
"""


@pytest.fixture(scope="module")
def kt(tmp_path_factory):
    d = tmp_path_factory.mktemp("kt")
    src = d / "keys_harness.cpp"
    src.write_text(HARNESS)
    lib = d / "libkt.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++",
                           "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "astrophysical-sph_b200", "csrc"),
                           str(src), "-o", str(lib)])
    L = C.CDLL(str(lib))
    L.kt_common.restype = C.c_int
    L.kt_common.argtypes = [C.c_ulonglong, C.c_ulonglong]
    return L


def replay(p, l, levels=42):
    """addNodes! for one particle: octants and the geometry of the cell of every depth (F/gravOctree_Single.jl:110-148)."""
    c = [0.0, 0.0, 0.0]
    lo = [-l] * 3
    hi = [l] * 3
    L = l
    digits, cells = [], [(tuple(c), tuple(lo), tuple(hi), L)]
    for _ in range(levels):
        cl = L / 2
        o = 0
        for a in range(3):
            lc, rc = c[a] - cl, c[a] + cl
            mn, ctr, mx = lc - cl, lc + cl, rc + cl
            if p[a] - c[a] > 0:
                o |= 1 << a
                c[a], lo[a], hi[a] = rc, ctr, mx
            else:
                c[a], lo[a], hi[a] = lc, mn, ctr
        L = cl
        digits.append(o)
        cells.append((tuple(c), tuple(lo), tuple(hi), L))
    return digits, cells


def words(kt, p, l):
    hi, lo = C.c_ulonglong(0), C.c_ulonglong(0)
    kt.kt_key_words(C.c_double(p[0]), C.c_double(p[1]), C.c_double(p[2]), C.c_double(l), C.byref(hi), C.byref(lo))
    return hi.value, lo.value


def points():
    rng = np.random.default_rng(8)
    pts = [rng.standard_normal(3) for _ in range(40)]
    base = pts[3]
    pts += [base + s * rng.standard_normal(3) for s in (1e-5, 1e-7, 1e-9, 1e-11, 1e-12)]     # part ways at depths 18 .. 41
    pts += [np.array([0.0, 0.0, 0.0]), np.array([0.5, -0.25, 0.125]), np.array([-1.0, 1.0, 0.0])]   # on cell boundaries
    l = max(np.abs(q).max() for q in pts)
    return pts, l


def test_key_digits_and_cell_geometry_replay_addnodes(kt):
    pts, l = points()
    for p in pts:
        digits, cells = replay(p, l)
        hi, lo = words(kt, p, l)
        assert hi < 2**63 and lo < 2**63
        got = [(hi >> (3 * (20 - k))) & 7 for k in range(21)] + [(lo >> (3 * (20 - k))) & 7 for k in range(21)]
        assert got == digits
        out = (C.c_double * 10)()
        for depth in (0, 1, 7, 20, 21, 22, 30, 41, 42):
            kt.kt_cell(C.c_ulonglong(hi), C.c_ulonglong(lo), depth, C.c_double(l), out)
            c, blo, bhi, L = cells[depth]
            assert tuple(out[0:3]) == c and tuple(out[3:6]) == blo and tuple(out[6:9]) == bhi and out[9] == L


def test_common_levels_is_the_depth_where_two_particles_part(kt):
    pts, l = points()
    reps = [replay(p, l)[0] for p in pts]
    ws = [words(kt, p, l) for p in pts]
    deepest = 0
    for i in range(len(pts)):
        for j in range(i + 1, len(pts)):
            same = 0
            while same < 42 and reps[i][same] == reps[j][same]:
                same += 1
            c = kt.kt_common(ws[i][0], ws[j][0])
            if c == 21:
                c += kt.kt_common(ws[i][1], ws[j][1])
            assert c == same
            deepest = max(deepest, same)
    assert deepest > 30          # the set does exercise the second key word
