"""Multi-GPU parity check, run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/mgpu_check.py [N]
Every rank evaluates the same initial condition with targets split by key range; rank 0 compares the assembled
result (and three steps) with the CPU oracle.  Exit code 0 = parity within the single-GPU tolerances.
bench.py calls run_check() before it times a multi-GPU run and prints the errors in its JSON line."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TOL = dict(rho=1e-9, h=0.0, ahyd=1e-9, dkdt=1e-9, sum_vdw=1e-9, g=1e-6, phi=1e-6, acc=1e-6, dt=1e-9, pos=1e-9, K=1e-9)


def run_check(dist, rank, world, device, N=20000):
    """All ranks call this.  Returns (errs, ok): the relative errors against the oracle (rank 0; None elsewhere) and
    the verdict, which is broadcast so that every rank agrees on it."""
    import torch

    import astrophysical_sph_b200.iniconds as ic
    from astrophysical_sph_b200 import parallel
    from astrophysical_sph_b200.libsph import SphB200

    d = ic.make_ic("polytropic", "gaussian_sphere", N=N, R=ic.R0)
    c = d["constants"]
    rng = np.random.default_rng(3)
    vel = np.asfortranarray(d["vel"] + 2e7 * rng.standard_normal(d["vel"].shape))
    s = SphB200(N, c["Kh"], "polytropic", m=c["m"], gamma=c["gamma"], G=c["G"], theta=c["theta"], alpha=c["alpha"],
                beta=c["beta"], device=device)
    parallel.init_handle_comm(s, dist)
    out = s.eval_acc(d["pos"], vel, d["K"])
    hy = s.hydro()
    g, phi = s.grav()
    s.upload(d["pos"], vel, d["K"], 0.0)
    info = s.step(3)
    p, v, Kend, t = s.download()
    s.close()
    errs, ok = None, True
    if rank == 0:
        from oracle import oracle as O

        nt = O.max_threads()
        kw = dict(eos=O.POLYTROPIC, Kent=d["K"], gamma=c["gamma"], alpha=c["alpha"], beta=c["beta"])
        oh = O.hydro(d["pos"], vel, c["m"], c["Kh"], nthreads=nt, **kw)
        og, ophi, _ = O.gravity(np.abs(d["pos"]).max(), c["m"], d["pos"], c["theta"], oh["h"], nthreads=nt)
        oo = O.step(d["pos"], vel, c["m"], c["Kh"], c["G"], c["theta"], 0.0, 3, nthreads=nt, **kw)

        def vrel(a, b):
            nb = np.linalg.norm(b, axis=1)
            return float((np.linalg.norm(a - b, axis=1) / np.maximum(nb, 1e-3 * np.median(nb))).max())

        errs = dict(rho=float(np.abs(hy["rho"] / oh["rho"] - 1).max()), h=float(np.abs(hy["h"] - oh["h"]).max()),
                    ahyd=vrel(hy["ahyd"], oh["ahyd"]), dkdt=float(np.abs(hy["dkdt"] - oh["dkdt"]).max() / np.abs(oh["dkdt"]).max()),
                    sum_vdw=float(np.abs(hy["sum_vdw"] - oh["sum_vdw"]).max() / np.abs(oh["sum_vdw"]).max()),
                    g=vrel(g, og), phi=float(np.abs(phi / ophi - 1).max()),
                    acc=vrel(out["acc"], oh["ahyd"] - c["G"] * og), dt=float(np.abs(info["dts"] / oo["dts"] - 1).max()),
                    pos=float(np.abs(p - oo["pos"]).max() / np.abs(oo["pos"]).max()), K=float(np.abs(Kend / oo["K"] - 1).max()))
        ok = all(errs[k] <= TOL[k] for k in TOL)
        errs = dict(errs, N=N, ranks=world, ok=bool(ok), checked="one getAcc + 3 steps of a polytropic Gaussian sphere against the oracle")
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{device}")
    dist.broadcast(flag, src=0)
    return errs, bool(flag.item())


def main():
    import torch
    import torch.distributed as dist

    N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    errs, ok = run_check(dist, rank, world, local, N)
    if rank == 0:
        print("mgpu_check world", world, "N", N, errs, flush=True)
        print("MGPU PARITY", "OK" if ok else "FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
