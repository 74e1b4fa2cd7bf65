"""Multi-GPU parity check, run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/mgpu_check.py [N]
Every rank evaluates the same initial condition with targets split by key range; rank 0 compares the assembled
result (and three steps) with the CPU oracle.  Exit code 0 = parity within the single-GPU tolerances."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import astrophysical_sph_b200.iniconds as ic  # noqa: E402
from astrophysical_sph_b200 import parallel  # noqa: E402
from astrophysical_sph_b200.libsph import SphB200  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = ic.make_ic("polytropic", "gaussian_sphere", N=N, R=ic.R0)
    c = d["constants"]
    rng = np.random.default_rng(3)
    vel = np.asfortranarray(d["vel"] + 2e7 * rng.standard_normal(d["vel"].shape))
    s = SphB200(N, c["Kh"], "polytropic", m=c["m"], gamma=c["gamma"], G=c["G"], theta=c["theta"], alpha=c["alpha"],
                beta=c["beta"], device=local)
    parallel.init_handle_comm(s, dist)
    out = s.eval_acc(d["pos"], vel, d["K"])
    hy = s.hydro()
    g, phi = s.grav()
    s.upload(d["pos"], vel, d["K"], 0.0)
    info = s.step(3)
    p, v, Kend, t = s.download()
    ok = True
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
                    g=vrel(g, og), phi=float(np.abs(phi / ophi - 1).max()),
                    acc=vrel(out["acc"], oh["ahyd"] - c["G"] * og), dt=float(np.abs(info["dts"] / oo["dts"] - 1).max()),
                    pos=float(np.abs(p - oo["pos"]).max() / np.abs(oo["pos"]).max()), K=float(np.abs(Kend / oo["K"] - 1).max()))
        print("mgpu_check world", world, "N", N, errs, flush=True)
        ok = (errs["rho"] < 1e-9 and errs["h"] == 0 and errs["ahyd"] < 1e-9 and errs["dkdt"] < 1e-9 and errs["g"] < 1e-6
              and errs["phi"] < 1e-6 and errs["acc"] < 1e-6 and errs["dt"] < 1e-9 and errs["pos"] < 1e-9 and errs["K"] < 1e-9)
        print("MGPU PARITY", "OK" if ok else "FAILED", flush=True)
    s.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
