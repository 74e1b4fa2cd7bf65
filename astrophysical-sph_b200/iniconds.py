"""Initial-condition generators -- host-side, offline.

Mirrors module INICONDS of the reference (F/iniconds.jl, F = julia_version/fastv1_kd&single_oc):
same distributions, same defaults (F/iniconds.jl:536-566), same constants rows (:655-690), written
through snapshot_rw.write_snapshot in the reference's CSV layout.  Random streams are numpy's
(default_rng(seed)); Julia's streams cannot be reproduced and the reference's own seeding is
ineffective (F/iniconds.jl:434-436), so parity is defined on identical snapshot *files*.

Deviations from reference defects are deliberate and listed in DESIGN.md:
  * ICs whose polytropic `K` the reference never defines (plummer, isothermal sphere, Bonnor-Ebert)
    get the gaussian_sphere recipe K = kB*T / (mu*mH*rho0^(gamma-1)), rho0 = SPH density at the COM.
  * bonnor_ebert_sphere / polytropic_sphere tabulate M(xi) once instead of a quadrature inside every
    bisection step (intractable at 16M particles).
"""
from __future__ import annotations

import numpy as np

R0 = 5.38552341e16      # "pc in [cm]" of the reference (F/iniconds.jl:532)
M0 = 1.9891e33          # solar mass [g] (:533)
G_CGS = 6.67430e-8
KB = 1.380649e-16       # (:572)
MH = 1.6735575e-24      # (:573)

DEFAULTS = dict(         # F/iniconds.jl:536-566
    N=10000, R=2.0 * R0, Kh=50, Kgr=20, t=0, tEnd=5e12, alpha=1.0, beta=2.0, G=G_CGS, theta=0.576,
    M=1 * M0, rho_c=150.0, xi_max=7.5, Omega_frac=0.5, gamma=5 / 3, mu=0.61, T=15_000_000, a=0.01,
    velocity_mode="virial", mach_number=1.0, alpha_vir=1.0, seed=42, spectrum="burgers",
    add_turbulence=False, turb_frac=0.1, n=3.0, axis=None, beta_rot=0.26, A=0.1,
)

IC_TYPES = ("sample_isothermal_sphere", "sample_plummer_sphere", "bonnor_ebert_sphere",
            "turbulent_molecular_cloud", "rotating_cloud", "polytropic_sphere", "gaussian_sphere",
            "boss_bodenheimer")


def _iso_dirs(rng, N):
    th = np.arccos(2 * rng.random(N) - 1)
    ph = 2 * np.pi * rng.random(N)
    return np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)


def _uniform_sphere(rng, N, R):
    """Rejection sampling in the cube (F/iniconds.jl:468-477, :205-214)."""
    out = np.empty((0, 3))
    while out.shape[0] < N:
        c = 2 * R * (rng.random((int(2.2 * (N - out.shape[0])) + 16, 3)) - 0.5)
        out = np.vstack([out, c[np.linalg.norm(c, axis=1) <= R]])
    return out[:N]


def sample_isothermal_sphere(N, R, cs, rng):
    """F/iniconds.jl:16-40."""
    radii = np.clip(np.abs(R / 3 * rng.standard_normal(N)), 0, R)
    ux, uy, uz = _iso_dirs(rng, N)
    return np.column_stack([radii * ux, radii * uy, radii * uz]), cs * rng.standard_normal((N, 3))


def sample_plummer_sphere(N, M, a, rng):
    """F/iniconds.jl:42-95: r = a (xi^(-2/3) - 1)^(-1/2); velocities by rejection."""
    r = a * (rng.random(N) ** (-2 / 3) - 1) ** (-0.5)
    ux, uy, uz = _iso_dirs(rng, N)
    pos = np.column_stack([r * ux, r * uy, r * uz])
    v_esc = np.sqrt(2 * G_CGS * M / np.sqrt(r**2 + a**2))
    v = np.zeros(N)
    todo = np.arange(N)
    while todo.size:
        x1, x2 = rng.random(todo.size), rng.random(todo.size)
        vv = x1**2 * v_esc[todo]
        g = vv**2 * (1 - vv**2 / v_esc[todo] ** 2) ** 3.5
        ok = 0.1 * x2 < g
        v[todo[ok]] = vv[ok]
        todo = todo[~ok]
    ux, uy, uz = _iso_dirs(rng, N)
    return pos, np.column_stack([v * ux, v * uy, v * uz])


def _lane_emden_iso(xi_max, n=4096):
    """Isothermal Lane-Emden psi(xi) and the mass integral I(xi) = int x^2 exp(-psi) dx, tabulated."""
    from scipy.integrate import solve_ivp

    def f(xi, y):
        return [y[1], -2 / xi * y[1] + np.exp(-y[0])]

    xs = np.linspace(1e-8, xi_max, n)
    sol = solve_ivp(f, (1e-8, xi_max), [0.0, 0.0], t_eval=xs, rtol=1e-8, atol=1e-8)
    integrand = xs**2 * np.exp(-sol.y[0])
    I = np.concatenate([[0.0], np.cumsum(0.5 * (integrand[1:] + integrand[:-1]) * np.diff(xs))])
    return xs, sol.y[0], I


def bonnor_ebert_sphere(N, cs, rho_c, xi_max, rng, velocity_mode="none", mach_number=1.0, alpha_vir=1.0):
    """F/iniconds.jl:98-194 with M(xi) tabulated once (inverse-CDF by interpolation)."""
    xs, _, I = _lane_emden_iso(xi_max)
    a = cs / np.sqrt(4 * np.pi * G_CGS * rho_c)
    Mtot = 4 * np.pi * a**3 * rho_c * I[-1]
    xi = np.interp(rng.random(N) * I[-1], I, xs)
    r = a * xi
    ux, uy, uz = _iso_dirs(rng, N)
    pos = np.column_stack([r * ux, r * uy, r * uz])
    vel = np.zeros((N, 3))
    if velocity_mode == "mach":
        vel = rng.standard_normal((N, 3)) * (mach_number * cs / np.sqrt(3))
        vel -= vel.mean(axis=0)
    elif velocity_mode == "virial":
        vel = rng.standard_normal((N, 3))
        m_part = Mtot / N
        cur = 0.5 * m_part * (vel**2).sum()
        R_eff = np.linalg.norm(pos, axis=1).max()
        want = 0.5 * alpha_vir * abs(-(3 / 5) * G_CGS * Mtot**2 / R_eff)
        vel *= np.sqrt(want / cur)
        vel -= vel.mean(axis=0)
    elif velocity_mode != "none":
        raise ValueError("velocity_mode must be none, mach, or virial")
    return pos, vel, Mtot


def turbulent_molecular_cloud(N, R, M, spectrum, cs, rng):
    """F/iniconds.jl:198-282.  The 32^3 'spectrum' is written directly in real space (no FFT), as in the
    reference; trilinear interpolation; zero mean; std(|v|) = cs."""
    rho_cloud = M / ((4 / 3) * np.pi * R**3)
    pos = _uniform_sphere(rng, N, R)
    gs = 32
    k1 = np.arange(1, gs + 1)
    ks = np.where(k1 <= gs // 2, k1, k1 - gs).astype(float)
    kx, ky, kz = np.meshgrid(ks, ks, ks, indexing="ij")
    kmag = np.sqrt(kx**2 + ky**2 + kz**2)
    power = -2.0 if spectrum == "burgers" else -11 / 3
    with np.errstate(divide="ignore"):
        amp = rng.standard_normal(kmag.shape) * np.where(kmag > 0, kmag, 1.0) ** power
    amp[kmag == 0] = 0.0
    phi = 2 * np.pi * rng.random(kmag.shape)
    d = rng.standard_normal(kmag.shape + (3,))
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    cube = (amp * np.cos(phi))[..., None] * d
    dx = 2 * R / gs
    f = (pos + R) / dx
    i0 = np.clip(np.floor(f).astype(int), 1, gs - 1)       # 1-based cell index, as the reference clamps
    w = f - i0
    i0 -= 1                                                # to 0-based storage
    vel = np.zeros((N, 3))
    for ox in (0, 1):
        for oy in (0, 1):
            for oz in (0, 1):
                wt = (w[:, 0] if ox else 1 - w[:, 0]) * (w[:, 1] if oy else 1 - w[:, 1]) * (w[:, 2] if oz else 1 - w[:, 2])
                vel += wt[:, None] * cube[i0[:, 0] + ox, i0[:, 1] + oy, i0[:, 2] + oz]
    vel -= vel.mean(axis=0)
    vel *= cs / np.std(np.linalg.norm(vel, axis=1), ddof=1)
    return pos, vel, np.full(N, rho_cloud)


def rotating_cloud(N, rng, Mtot=1.99e33, Rcloud=3e17, rho_c=1e-18, Omega_frac=0.5, add_turbulence=False, turb_frac=0.1):
    """F/iniconds.jl:285-340."""
    r0 = Rcloud / 3
    r = np.zeros(N)
    todo = np.arange(N)
    while todo.size:
        rr = Rcloud * rng.random(todo.size) ** (1 / 3)
        ok = rng.random(todo.size) < 1 / (1 + (rr / r0) ** 2) ** 2.5
        r[todo[ok]] = rr[ok]
        todo = todo[~ok]
    ux, uy, uz = _iso_dirs(rng, N)
    x, y, z = r * ux, r * uy, r * uz
    Rc = np.sqrt(x**2 + y**2)
    v_rot = Omega_frac * np.sqrt(G_CGS * Mtot * Rc / Rcloud**3)
    with np.errstate(invalid="ignore", divide="ignore"):
        vx, vy = -v_rot * y / Rc, v_rot * x / Rc
    vx[np.isnan(vx)] = 0.0
    vy[np.isnan(vy)] = 0.0
    vel = np.column_stack([vx, vy, np.zeros(N)])
    if add_turbulence:
        vel = vel + turb_frac * np.linalg.norm(vel, axis=1).mean() * (rng.standard_normal((N, 3)) / np.sqrt(3))
    return np.column_stack([x, y, z]), vel


def polytropic_sphere(N, n, K, rho_c, xi_max, rng):
    """F/iniconds.jl:342-415 with the mass profile tabulated once."""
    from scipy.integrate import solve_ivp

    def f(xi, y):
        return [y[1], -2 / xi * y[1] - np.sign(y[0]) * abs(y[0]) ** n]

    xs = np.linspace(1e-8, xi_max, 4096)
    sol = solve_ivp(f, (1e-8, xi_max), [1.0, 0.0], t_eval=xs, rtol=1e-8, atol=1e-10)
    th = np.clip(sol.y[0], 0, None)
    integrand = xs**2 * th**n
    I = np.concatenate([[0.0], np.cumsum(0.5 * (integrand[1:] + integrand[:-1]) * np.diff(xs))])
    a = np.sqrt((n + 1) * K / (4 * np.pi * G_CGS) * rho_c ** (1 / n - 1))
    Mtot = 4 * np.pi * a**3 * rho_c * I[-1]
    keep = np.concatenate([[True], np.diff(I) > 0])
    r = a * np.interp(rng.random(N) * I[-1], I[keep], xs[keep])
    ux, uy, uz = _iso_dirs(rng, N)
    return np.column_stack([r * ux, r * uy, r * uz]), np.zeros((N, 3)), Mtot


def gaussian_sphere(N, R, rng, axis=None, Omega_frac=0.0):
    """F/iniconds.jl:418-454."""
    pos = rng.standard_normal((N, 3)) * R
    pos -= pos.mean(axis=0)
    vel = np.zeros((N, 3))
    if axis is not None and Omega_frac != 0.0:
        ax = np.asarray(axis, dtype=float)
        ax = ax / np.linalg.norm(ax)
        vel = Omega_frac * np.cross(ax[None, :], pos)
    return pos, vel


def boss_bodenheimer(N, R, M, rng, A=0.1, beta_rot=0.26):
    """F/iniconds.jl:457-525: uniform sphere, m=2 azimuthal perturbation, solid-body rotation."""
    rho_cloud = M / ((4 / 3) * np.pi * R**3)
    pos = _uniform_sphere(rng, N, R)
    pos -= pos.mean(axis=0)
    phi = np.arctan2(pos[:, 1], pos[:, 0])
    ps = phi.copy()                                       # Newton: (ps + A sin 2ps)/2 = phi  (:484-497)
    active = np.ones(N, dtype=bool)
    for _ in range(50):
        fv = (ps + A * np.sin(2 * ps)) / 2 - phi
        fp = (1 + 2 * A * np.cos(2 * ps)) / 2
        new = ps - fv / fp
        conv = np.abs(new - ps) < 1e-12
        ps = np.where(active, new, ps)
        active &= ~conv
        if not active.any():
            break
    rxy = np.sqrt(pos[:, 0] ** 2 + pos[:, 1] ** 2)
    pos[:, 0], pos[:, 1] = rxy * np.cos(ps), rxy * np.sin(ps)
    I = 0.4 * M * R**2
    Egrav = -3 / 5 * G_CGS * M**2 / R
    Omega = np.sqrt(2 * beta_rot * abs(Egrav) / I)
    vel = np.column_stack([-Omega * pos[:, 1], Omega * pos[:, 0], np.zeros(N)])
    vel -= vel.mean(axis=0)
    return pos, vel, np.full(N, rho_cloud)


def sph_density_at_point(pt, pos, m, Kh, poly=True):
    """HJL.density_plot at one point (F/polytrope_hydroKDTree.jl:344-350) for the K recipe; offline, scipy."""
    from scipy.spatial import cKDTree

    r, _ = cKDTree(pos).query(np.asarray(pt, dtype=float)[None, :], k=Kh)
    r = r[0]
    h = r[-1] / 2
    q = r / h
    w = np.where(q <= 1, 1 - 1.5 * q**2 + 0.75 * q**3, 0.25 * (2 - q) ** 3) / (np.pi * h**3)
    return m * w.sum()


def make_ic(EOS, ic_type, **kwargs):
    """iniconds_setup (F/iniconds.jl:528-697) without the file write.

    Returns dict(pos, vel, K (or None), constants) -- `constants` has the reference's keys."""
    p = dict(DEFAULTS)
    p.update(kwargs)
    if ic_type not in IC_TYPES:
        raise ValueError(f"Invalid ic_type: {ic_type}")
    if EOS not in ("isothermal", "polytropic"):
        raise ValueError(f"Invalid EOS: {EOS}. Available options: 'isothermal' or 'polytropic'")
    rng = np.random.default_rng(p["seed"])
    N = int(p["N"])
    cs = float(np.sqrt(KB * p["T"] / (p["mu"] * MH)))           # :576
    m = p["M"] / N                                              # :577
    U = 3 / 2 * p["M"] * cs**2                                  # :578
    K = None
    gam = p["gamma"]

    def k_recipe(rho0):
        return np.full(N, KB * p["T"] / (p["mu"] * MH * rho0 ** (gam - 1)))

    if ic_type == "sample_isothermal_sphere":
        pos, vel = sample_isothermal_sphere(N, p["R"], cs, rng)
    elif ic_type == "sample_plummer_sphere":
        pos, vel = sample_plummer_sphere(N, p["M"], p["a"], rng)
    elif ic_type == "bonnor_ebert_sphere":
        pos, vel, _ = bonnor_ebert_sphere(N, cs, p["rho_c"], p["xi_max"], rng, p["velocity_mode"],
                                          p["mach_number"], p["alpha_vir"])
    elif ic_type == "turbulent_molecular_cloud":
        pos, vel, rho_vec = turbulent_molecular_cloud(N, p["R"], p["M"], p["spectrum"], cs, rng)
        K = cs**2 / gam * rho_vec ** (1 - gam)                  # :611
    elif ic_type == "rotating_cloud":
        pos, vel = rotating_cloud(N, rng, p["M"], p["R"], p["rho_c"], p["Omega_frac"], p["add_turbulence"],
                                  p["turb_frac"])
        K = k_recipe(p["rho_c"])                                # :624
    elif ic_type == "polytropic_sphere":
        if "K" not in p:
            raise ValueError("Missing required arguments for polytropic_sphere: ['K']")
        pos, vel, M_actual = polytropic_sphere(N, p["n"], p["K"], p["rho_c"], p["xi_max"], rng)
        K = np.full(N, float(p["K"]))
        m = M_actual / N
        p["M"] = M_actual
    elif ic_type == "gaussian_sphere":
        pos, vel = gaussian_sphere(N, p["R"], rng, p["axis"], p["Omega_frac"] if p["axis"] is not None else 0.0)
    else:  # boss_bodenheimer
        pos, vel, rho = boss_bodenheimer(N, p["R"], p["M"], rng, p["A"], p["beta_rot"])
        K = k_recipe(rho[0])                                    # :643
    if K is None and EOS == "polytropic":
        K = k_recipe(sph_density_at_point(pos.mean(axis=0), pos, m, int(p["Kh"])))   # :636-638
    r_com = pos.mean(axis=0)
    R_max = float(np.linalg.norm(pos - r_com, axis=1).max())   # :650-651
    constants = {"iterID": 1, "N": N, "Kh": int(p["Kh"]), "Kgr": int(p["Kgr"]), "t": float(p["t"]),
                 "tEnd": float(p["tEnd"]), "M": float(p["M"]), "R": R_max, "alpha": float(p["alpha"]),
                 "beta": float(p["beta"]), "G": float(p["G"]), "theta": float(p["theta"]), "m": float(m)}
    if EOS == "isothermal":
        constants.update({"cs": cs, "U": float(U)})
        K = None
    else:
        constants["gamma"] = float(gam)
    return dict(pos=np.asfortranarray(pos), vel=np.asfortranarray(vel), K=K, constants=constants)


def iniconds_setup(EOS, ic_type, root=".", **kwargs):
    """Generate and write `snapshots/<ic_type>/bin/1snap.csv` (F/iniconds.jl:672,691)."""
    from . import snapshot_rw

    ic = make_ic(EOS, ic_type, **kwargs)
    snapshot_rw.write_snapshot("1", ic_type, ic["pos"], ic["vel"], K=ic["K"], constants=ic["constants"], root=root)
    kind = "an isothermal" if EOS == "isothermal" else "a polytropic"
    print(f"Initial conditions for {kind} {ic_type} have been produced.")
    return ic
