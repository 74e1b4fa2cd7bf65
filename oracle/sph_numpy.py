"""numpy/scipy twin of the CPU oracle -- an INDEPENDENT restatement used to cross-check
oracle/sph_oracle.cpp (tests/test_oracle.py).  TEST INFRASTRUCTURE ONLY.

It follows the reference's *matrix* formulation (N x K arrays, as the Julia code builds
them) rather than the streaming loops of the C++ oracle, uses scipy's cKDTree for the
exact kNN that NearestNeighbors.jl provides in the reference, and a pure-Python BFS
octree for small N.  PARITY UNPINNED (no reference golden vectors exist).

F/ = /root/reference/julia_version/fastv1_kd&single_oc/
"""
from __future__ import annotations

from collections import deque

import numpy as np
from scipy.spatial import cKDTree


# ----------------------------------------------------------------------------- hydro
def get_neighbors(ri, rj, K):
    """F/isothermal_hydroKDTree.jl:118-163 (indices 0-based here)."""
    tree = cKDTree(rj)
    r, idx = tree.query(ri, k=K)
    d = ri[:, None, :] - rj[idx]           # getTreeDiffs :76-97
    h = r[:, -1] / 2                        # :151
    q = r / h[:, None]                      # :154
    return d[..., 0], d[..., 1], d[..., 2], r, h, q, idx


def W(h, q, poly=False):
    """F/isothermal_hydroKDTree.jl:5-35."""
    ct = 1 / (np.pi * h**3)
    m1 = q <= 1.0
    m2 = ~m1 if poly else ((q > 1.0) & (q <= 2.0))
    w = np.zeros_like(q)
    w1 = ct[:, None] * (1 - 3 / 2 * q**2 + 3 / 4 * q**3)
    w2 = ct[:, None] * 1 / 4 * (2 - q) ** 3
    w[m1] = w1[m1]
    w[m2] = w2[m2]
    return w


def gradW(dx, dy, dz, r, h, q, poly=False):
    """F/isothermal_hydroKDTree.jl:38-73."""
    ct = (1 / (np.pi * h**4))[:, None]
    hh = h[:, None]
    m1 = q <= 1.0
    m2 = ~m1 if poly else ((q > 1.0) & (q <= 2.0))
    dWdr = np.zeros_like(q)
    with np.errstate(divide="ignore", invalid="ignore"):
        a = ct * (9 / 4 * r / hh**2 - 3 / hh)
        b = ct * (-3 / 4 * (2 - q) ** 2) / r
    dWdr[m1] = a[m1]
    dWdr[m2] = b[m2]
    return dWdr * dx, dWdr * dy, dWdr * dz


def hydrodynamics(pos, vel, m, K, eos="isothermal", cs=0.0, Kent=None, gamma=5 / 3, alpha=1.0, beta=2.0):
    """HJL.hydrodynamics (iso :248-288, poly :251-292) + the evolve_K! sum (poly :296-316)."""
    poly = eos == "polytropic"
    N = pos.shape[0]
    dx, dy, dz, r, h, q, idx = get_neighbors(pos, pos, K)
    w = W(h, q, poly)
    rho = m * w.sum(axis=1)
    if poly:
        c = np.sqrt(gamma * Kent * rho ** (gamma - 1))
        P = Kent * rho**gamma
    else:
        c = np.full(N, cs)
        P = cs**2 * rho
    dWx, dWy, dWz = gradW(dx, dy, dz, r, h, q, poly)
    # getAV :196-216
    h_avg = (h[:, None] + h[idx]) / 2
    rho_avg = (rho[:, None] + rho[idx]) / 2
    vij = vel[:, None, :] - vel[idx]
    v_dot_r = vij[..., 0] * dx + vij[..., 1] * dy + vij[..., 2] * dz
    mu = np.minimum(h_avg * v_dot_r / (r**2 + 0.01 * h_avg**2), 0)
    Pi = (-alpha * c[:, None] * mu + beta * mu**2) / rho_avg
    # hydroCalculation :219-245
    if poly:
        ct = m * (((P / rho**2)[:, None] + (P / rho**2)[idx]) + Pi) / 2
    else:
        ct = m * ((P / rho**2)[:, None] + Pi / 2)
    a = np.zeros((N, 3))
    for comp, dWc in enumerate((dWx, dWy, dWz)):
        t = (ct * dWc)[:, 1:]
        a[:, comp] -= t.sum(axis=1)
        np.add.at(a[:, comp], idx[:, 1:].ravel(), t.ravel())
    vdw = vij[..., 0] * dWx + vij[..., 1] * dWy + vij[..., 2] * dWz
    s = (m * Pi * vdw / 2)[:, 1:]
    dkdt = s.sum(axis=1)
    np.add.at(dkdt, idx[:, 1:].ravel(), s.ravel())
    return dict(idx=idx, r=r, ahyd=a, rho=rho, h=h, sum_vdw=vdw.sum(axis=1), mumax=mu.max(axis=1), cs_i=c,
                dkdt=dkdt, P=P)


# ----------------------------------------------------------------------------- gravity
def grav_kernels(r, h):
    """GJL.Kernels F/gravOctree_Single.jl:5-29 -> (gPHI, PHI) scalars."""
    q = r / h
    if q <= 1:
        g = (1 / h**2) * (4 / 3 / h - 6 / 5 * (r**2 / h**3) + 1 / 2 * (r**3 / h**4))
        p = (1 / h) * (2 / 3 * q**2 - 3 / 10 * q**4 + 1 / 10 * q**5 - 7 / 5)
    elif q <= 2:
        g = ((1 / h**2) * (8 / 3 * q - 3 * q**2 + 6 / 5 * q**3 - 1 / 6 * q**4 - 1 / 15 * (1 / q**2))) / r
        p = (1 / h) * (4 / 3 * q**2 - q**3 + 3 / 10 * q**4 - 1 / 30 * q**5 - 8 / 5 + 1 / 15 / q)
    else:
        g = 1 / r**3
        p = -1 / r
    return g, p


def direct_gravity(m, pos, h):
    """theta -> 0 limit of GJL.gravity: every other particle through the leaf kernel
    (cf. B/adiabatic_forces.jl:78-136 for the same softened pair kernels)."""
    N = pos.shape[0]
    g = np.zeros((N, 3))
    phi = np.zeros(N)
    for i in range(N):
        d = pos[i] - pos
        r = np.sqrt((d**2).sum(axis=1))
        hij = (h[i] + h) / 2
        for j in range(N):
            if j == i:
                continue
            gp, pp = grav_kernels(r[j], hij[j])
            g[i] += m * gp * d[j]
            phi[i] += m * pp
    return g, phi - m * (7 / 5) / h


class _Node:
    __slots__ = ("L", "c", "lo", "hi", "M", "com", "parent", "count", "plist", "children", "leaf")


def octree_gravity(l, m, pos, theta, h, depth_cap=200):
    """GJL.gravity F/gravOctree_Single.jl:307-319 in plain Python (small N only)."""
    N = pos.shape[0]
    root = _Node()
    root.L, root.c = l, np.zeros(3)
    root.lo, root.hi = np.full(3, -l), np.full(3, l)
    root.M, root.com, root.parent, root.count = 0.0, np.zeros(3), -1, 0
    root.plist, root.children, root.leaf = list(range(N)), [], False
    nodes = [root]
    i = 0
    while i < len(nodes):                                  # build_octree! :213-227
        nd = nodes[i]
        if nd.count != 1:
            cl = nd.L / 2                                   # addNodes! :107-181
            lc, rc = nd.c - cl, nd.c + cl
            mn, ctr, mx = lc - cl, lc + cl, rc + cl
            buckets = [[] for _ in range(8)]
            for p in nd.plist:
                rel = pos[p] - nd.c
                buckets[4 * int(rel[2] > 0) + 2 * int(rel[1] > 0) + int(rel[0] > 0)].append(p)
            for ci in range(8):
                if not buckets[ci]:
                    continue
                b = np.array([ci & 1, (ci >> 1) & 1, (ci >> 2) & 1], dtype=bool)
                ch = _Node()
                ch.L = cl
                ch.c = np.where(b, rc, lc)
                ch.lo = np.where(b, ctr, mn)
                ch.hi = np.where(b, mx, ctr)
                ch.M = m * len(buckets[ci])
                ch.count = len(buckets[ci])
                ch.plist, ch.children, ch.parent, ch.leaf = buckets[ci], [], i, False
                ch.com = np.zeros(3)
                nodes.append(ch)
                nd.children.append(len(nodes) - 1)
            nd.plist = []
            if len(nodes) > 50 * N + 100:
                raise RuntimeError("octree runaway")
        i += 1
    leaves = []
    for i in range(len(nodes) - 1, -1, -1):                 # setCOMs! :183-211
        nd = nodes[i]
        if nd.count == 1:
            nd.leaf = True
            leaves.append(i)
            nd.com = pos[nd.plist[0]].copy()
        else:
            tm, ws = 0.0, np.zeros(3)
            for c in nd.children:
                tm += nodes[c].M
                ws = ws + nodes[c].M * nodes[c].com
            nd.M, nd.com = tm, ws / tm
    g = np.zeros((N, 3))
    phi = np.zeros(N)
    th2 = theta**2
    visits = np.zeros(3, dtype=np.int64)
    for lid in leaves:                                      # gravity_acc :280-304
        leaf = nodes[lid]
        pi = leaf.plist[0]
        par = nodes[leaf.parent]
        par.children.remove(lid)
        gi, ph = np.zeros(3), 0.0
        dq = deque(nodes[0].children)
        p, hi = pos[pi], h[pi]
        while dq:                                           # compute_g :239-278
            nd = nodes[dq.popleft()]
            d = p - nd.com
            d2 = d[0] ** 2 + d[1] ** 2 + d[2] ** 2
            s = nd.L * 2
            if nd.leaf:
                j = nd.plist[0]
                gp, pp = grav_kernels(np.sqrt(d2), (hi + h[j]) / 2)
                gi += nd.M * gp * d
                ph += nd.M * pp
                visits[0] += 1
                continue
            md = np.maximum(np.maximum(nd.lo - p, 0), p - nd.hi)
            md2 = md[0] ** 2 + md[1] ** 2 + md[2] ** 2
            with np.errstate(divide="ignore"):
                ok = (s**2 / d2 < th2) and (hi**2 / md2 < 0.25)
            if ok:
                dd = np.sqrt(d2)
                gi += nd.M / dd**3 * d
                ph += -nd.M / dd
                visits[1] += 1
            else:
                dq.extend(nd.children)
                visits[2] += 1
        g[pi], phi[pi] = gi, ph
        par.children.append(lid)
    return g, phi - m * (7 / 5) / h, len(nodes), visits


# ----------------------------------------------------------------------------- step
def adaptive_dt(vel, acc, hy, m, cs_vec, alpha, beta):
    """F/isothermal_sim.jl:158-166."""
    vel_r = np.sqrt((vel**2).sum(axis=1))
    a_r = np.sqrt((acc**2).sum(axis=1))
    with np.errstate(divide="ignore"):
        abs_div_v = np.abs(-(m * hy["sum_vdw"]) / hy["rho"])
        return 0.3 * min(
            (1 / abs_div_v).min(),
            (hy["h"] / vel_r).min(),
            np.sqrt(hy["h"] / a_r).min(),
            (hy["h"] / (cs_vec + 1.2 * (alpha * cs_vec + beta * hy["mumax"]))).min(),
        )
