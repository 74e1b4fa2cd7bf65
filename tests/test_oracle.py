"""CPU tests of the oracle (oracle/sph_oracle.cpp) -- the checker the GPU parity tests rely on.

The reference has no tests, fixtures or golden vectors and cannot run here (no Julia):
parity is UNPINNED against the reference itself.  What pins the oracle instead:
  * an independent numpy/scipy restatement in the reference's matrix formulation (oracle/sph_numpy.py),
  * brute-force kNN and scipy's cKDTree (stand-ins for NearestNeighbors.jl's exact search),
  * direct-sum gravity in the theta -> 0 limit (the reference's own cross-check engine, B/adiabatic_forces.jl),
  * structural invariants that follow from the reference code,
  * the committed golden vectors (tests/golden, regression pin).
"""
import os

import numpy as np
import pytest
from conftest import GOLDEN, make_case, oracle_kwargs, vec_rel

from oracle import sph_numpy as NP


def test_knn_matches_brute_force_and_ckdtree(oracle):
    rng = np.random.default_rng(1)
    pos = np.asfortranarray(rng.standard_normal((700, 3)))
    K = 20
    idx, r = oracle.knn(pos, pos, K, nthreads=2)
    d2 = ((pos[:, None, :] - pos[None, :, :]) ** 2)
    d2 = (d2[..., 0] + d2[..., 1]) + d2[..., 2]
    order = np.argsort(d2, axis=1, kind="stable")[:, :K]
    assert np.array_equal(idx - 1, order)                       # F/isothermal_hydroKDTree.jl:131 sorted ascending
    assert np.array_equal(idx[:, 0] - 1, np.arange(700))        # column 1 = self
    assert np.array_equal(r, np.sqrt(np.take_along_axis(d2, order, axis=1)))
    from scipy.spatial import cKDTree

    _, idx2 = cKDTree(pos).query(pos, k=K)
    assert all(set(a) == set(b) for a, b in zip(idx - 1, idx2))


def test_knn_tie_break_is_by_index(oracle):
    # cubic lattice: many exact distance ties
    g = np.arange(6, dtype=float)
    pos = np.asfortranarray(np.array(np.meshgrid(g, g, g, indexing="ij")).reshape(3, -1).T)
    idx, r = oracle.knn(pos, pos, 12)
    for i in range(pos.shape[0]):
        rows = list(zip(r[i], idx[i]))
        assert rows == sorted(rows)


@pytest.mark.parametrize("eos", ["isothermal", "polytropic"])
def test_hydro_matches_numpy_twin(oracle, eos):
    pos, vel, K, c, _ = make_case(eos, "gaussian_sphere", 900, R=5.38552341e16)
    rng = np.random.default_rng(3)
    vel = np.asfortranarray(vel + 2e7 * rng.standard_normal(vel.shape))
    kw = oracle_kwargs(oracle, eos, c, K)
    a = oracle.hydro(pos, vel, c["m"], 50, **kw)
    kw.pop("eos")
    b = NP.hydrodynamics(np.asarray(pos), np.asarray(vel), c["m"], 50, eos=eos, **kw)
    assert np.array_equal(a["idx"] - 1, b["idx"])
    np.testing.assert_allclose(a["rho"], b["rho"], rtol=1e-12)
    np.testing.assert_allclose(a["h"], b["h"], rtol=0, atol=0)
    assert vec_rel(a["ahyd"], b["ahyd"]) < 1e-10
    np.testing.assert_allclose(a["sum_vdw"], b["sum_vdw"], rtol=1e-9, atol=1e-12 * np.abs(b["sum_vdw"]).max())
    np.testing.assert_allclose(a["mumax"], b["mumax"], rtol=0, atol=0)
    if eos == "polytropic":
        np.testing.assert_allclose(a["cs_i"], b["cs_i"], rtol=1e-13)
        np.testing.assert_allclose(a["dkdt"], b["dkdt"], rtol=1e-9, atol=1e-12 * np.abs(b["dkdt"]).max())


def test_hydro_invariants(oracle):
    pos, vel, K, c, _ = make_case("isothermal", "gaussian_sphere", 2000, R=5.38552341e16)
    a = oracle.hydro(pos, vel, c["m"], 50, cs=c["cs"])
    # K-th neighbour sits exactly at q = 2 (Appendix B-1); max_j mu_ij == 0 (B-2)
    assert np.array_equal(a["r"][:, -1] / a["h"], np.full(2000, 2.0))
    assert np.array_equal(a["mumax"], np.zeros(2000))
    # the pair scatter is exactly antisymmetric: total hydro force vanishes to round-off
    assert np.abs(a["ahyd"].sum(axis=0)).max() < 1e-12 * np.abs(a["ahyd"]).sum(axis=0).max()


def test_gravity_matches_numpy_twin_and_direct_sum(oracle):
    pos, vel, K, c, _ = make_case("isothermal", "gaussian_sphere", 400, R=1.0)
    m, theta = 1.0 / 400, 0.576
    h = oracle.hydro(pos, vel, m, 30, cs=1.0)["h"]
    l = np.abs(pos).max()
    g, phi, st = oracle.gravity(l, m, pos, theta, h)
    g2, phi2, nn, visits = NP.octree_gravity(l, m, np.asarray(pos), theta, h)
    assert st[0] == nn
    assert vec_rel(g, g2) < 1e-12
    np.testing.assert_allclose(phi, phi2, rtol=1e-12)
    # threaded walk (own leaf skipped instead of list surgery) agrees up to summation order
    g3, phi3, _ = oracle.gravity(l, m, pos, theta, h, nthreads=4)
    assert vec_rel(g3, g) < 1e-12
    # theta -> 0: every interaction goes through the softened pair kernel = direct sum
    g0, phi0, _ = oracle.gravity(l, m, pos, 1e-9, h)
    gd, phid = NP.direct_gravity(m, np.asarray(pos), h)
    assert vec_rel(g0, gd) < 1e-11
    np.testing.assert_allclose(phi0, phid, rtol=1e-11)
    # and the theta = 0.576 monopole walk is a decent approximation of it
    assert vec_rel(g, gd) < 0.05


def test_octree_structure(oracle):
    pos, *_ = make_case("isothermal", "gaussian_sphere", 3000, R=1.0)
    l = np.abs(pos).max()
    t = oracle.octree(l, 1.0, pos)
    depth, count, L = t[:, 15], t[:, 14], t[:, 0]
    assert np.all(np.diff(depth) >= 0)                       # breadth-first order (build_octree! :217-223)
    assert count[0] == 0 and (count[1:] >= 1).all()          # the root keeps particle_count = 0 (:94-104)
    assert (count == 1).sum() == 3000                        # one leaf per particle
    np.testing.assert_array_equal(L, l / 2.0 ** depth)
    np.testing.assert_allclose(t[:, 7:10] - t[:, 4:7], np.repeat(2 * L[:, None], 3, axis=1), rtol=1e-12)


def test_adaptive_dt_matches_numpy_twin(oracle):
    pos, vel, K, c, _ = make_case("isothermal", "gaussian_sphere", 800, R=5.38552341e16)
    rng = np.random.default_rng(5)
    vel = np.asfortranarray(1e6 * rng.standard_normal(vel.shape))
    out = oracle.getacc(pos, vel, c["m"], 50, c["G"], c["theta"], cs=c["cs"])
    dt = oracle.dt_from(vel, out["acc"], out["rho"], out["h"], out["sum_vdw"], out["mumax"], c["m"], cs=c["cs"])
    dt2 = NP.adaptive_dt(np.asarray(vel), np.asarray(out["acc"]), out, c["m"], np.full(800, c["cs"]), 1.0, 2.0)
    assert dt == pytest.approx(dt2, rel=1e-13)


@pytest.mark.parametrize("name,eos", [("gauss_iso_1024.npz", "isothermal"), ("gauss_poly_1024.npz", "polytropic")])
def test_oracle_reproduces_golden_vectors(oracle, name, eos):
    z = np.load(os.path.join(GOLDEN, name))
    m, cs, gamma, G, theta, alpha, beta, U, Kh = z["consts"]
    Kh = int(Kh)
    K = z["K"] if z["K"].size else None
    kw = dict(eos=oracle.ISOTHERMAL if eos == "isothermal" else oracle.POLYTROPIC, cs=cs, Kent=K, gamma=gamma,
              alpha=alpha, beta=beta)
    hy = oracle.hydro(z["pos"], z["vel"], m, Kh, nthreads=2, **kw)
    assert np.array_equal(hy["idx"], z["idx"])
    np.testing.assert_array_equal(hy["r"][:, -1], z["rK"])
    np.testing.assert_allclose(hy["rho"], z["rho"], rtol=1e-14)
    assert vec_rel(hy["ahyd"], z["ahyd"]) < 1e-12
    g, phi, st = oracle.gravity(np.abs(z["pos"]).max(), m, z["pos"], theta, hy["h"], nthreads=2)
    assert vec_rel(g, z["g"]) < 1e-12
    np.testing.assert_allclose(phi, z["phi"], rtol=1e-12)
    stp = oracle.step(z["pos"], z["vel"], m, Kh, G, theta, 0.0, len(z["dts"]), U_iso=U, **kw)
    np.testing.assert_allclose(stp["dts"], z["dts"], rtol=1e-12)
    assert vec_rel(stp["pos"], z["pos_end"]) < 1e-12


def test_kernel_normalisation_and_continuity():
    # int W dV = 1 and W, gradW continuous at q = 1, 2 (F/isothermal_hydroKDTree.jl:22-31, :57-65)
    h = 0.7
    q = np.linspace(0, 2, 200001)
    w = NP.W(np.array([h]), q[None, :])[0]
    integral = np.trapezoid(w * 4 * np.pi * (q * h) ** 2, q * h)
    assert integral == pytest.approx(1.0, rel=1e-8)
    for qq in (1.0, 2.0):
        lo = NP.W(np.array([h]), np.array([[qq - 1e-9]]))[0, 0]
        hi = NP.W(np.array([h]), np.array([[qq + 1e-9]]))[0, 0]
        assert lo == pytest.approx(hi, abs=1e-8)
    # gravity kernels continuous at q = 1, 2 (F/gravOctree_Single.jl:8-21)
    for qq in (1.0, 2.0):
        a = NP.grav_kernels(h * (qq - 1e-9), h)
        b = NP.grav_kernels(h * (qq + 1e-9), h)
        assert a[0] == pytest.approx(b[0], rel=1e-7) and a[1] == pytest.approx(b[1], rel=1e-7)
